"""GPU primitives (scan, stable radix sort, filters, transposition, integer
marginals) against numpy, through the C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _table(n_bins, nnz_target, seed, chrom_offset=None):
    rng = np.random.default_rng(seed)
    i = rng.integers(0, n_bins, nnz_target)
    j = rng.integers(0, n_bins, nnz_target)
    i, j = np.minimum(i, j), np.maximum(i, j)
    key = np.unique(i.astype(np.int64) * n_bins + j)
    b1, b2 = key // n_bins, key % n_bins
    cnt = rng.integers(0, 50, len(b1)).astype(np.int32)     # includes stored zeros
    ind = np.searchsorted(b1, np.arange(n_bins + 1)).astype(np.int64)
    if chrom_offset is None:
        chrom_offset = np.array([0, n_bins // 3, n_bins // 3 + 1, n_bins])
    return b1, b2, cnt, ind, np.asarray(chrom_offset)


@pytest.mark.parametrize("n", [0, 1, 5, 2048, 2049, 100_003, 5_000_000])
def test_exclusive_scan(n):
    from cooler_b200 import table as T
    T._use_device("cuda")
    rng = np.random.default_rng(n)
    x = rng.integers(0, 1000, n).astype(np.int32)
    out = T.exclusive_scan_i32(torch.from_numpy(x).cuda()).cpu().numpy()
    want = np.r_[0, np.cumsum(x.astype(np.int64))]
    np.testing.assert_array_equal(out, want)


@pytest.mark.parametrize("n,bits", [(1, 1), (31, 3), (4096, 8), (4097, 9), (200_000, 15),
                                    (1_000_003, 22), (300_000, 31)])
def test_sort_pairs_stable(n, bits):
    from cooler_b200 import table as T
    T._use_device("cuda")
    rng = np.random.default_rng(bits)
    keys = rng.integers(0, 2 ** bits if bits < 31 else 2 ** 31 - 1, n).astype(np.int32)
    vals = np.stack([np.arange(n, dtype=np.int32), rng.integers(0, 9, n).astype(np.int32)], 1)
    ko, vo = T.sort_pairs(torch.from_numpy(keys).cuda(), torch.from_numpy(vals).cuda(), bits)
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(ko.cpu().numpy(), keys[order])
    np.testing.assert_array_equal(vo[:n].cpu().numpy(), vals[order])


@pytest.mark.parametrize("n_bins,nnz,seed", [(50, 300, 0), (3000, 200_000, 1), (100_000, 150_000, 2),
                                             (7, 0, 3), (2500, 1_000_000, 4)])
@pytest.mark.parametrize("ndiag,keep", [(0, 0), (2, 0), (1, 1), (3, 2)])
def test_filter_transpose_marginals(n_bins, nnz, seed, ndiag, keep):
    from cooler_b200 import table as T
    b1, b2, cnt, ind, chrom_offset = _table(n_bins, nnz, seed)
    tab = T.PixelTable.from_arrays(ind, b2, cnt, chrom_offset)
    r, c, v = tab.raw.to_host()
    np.testing.assert_array_equal(r, b1), np.testing.assert_array_equal(c, b2), np.testing.assert_array_equal(v, cnt)

    chrom = np.repeat(np.arange(len(chrom_offset) - 1), np.diff(chrom_offset))
    cis = chrom[b1] == chrom[b2]
    m = np.ones(len(b1), bool)
    if ndiag:
        m &= np.abs(b1 - b2) >= ndiag
    if keep == 1:
        m &= cis
    elif keep == 2:
        m &= ~cis
    csr, tkey, tval = T.filter_copy(tab.raw, tab.row_lo, tab.row_hi, ndiag, keep, want_transpose=True)
    r, c, v = csr.to_host()
    np.testing.assert_array_equal(r, b1[m]), np.testing.assert_array_equal(c, b2[m])
    np.testing.assert_array_equal(v, cnt[m])
    np.testing.assert_array_equal(csr.indptr.cpu().numpy(), np.searchsorted(b1[m], np.arange(n_bins + 1)))

    csc = T.transpose_from(tkey, tval, n_bins)
    r, c, v = csc.to_host()          # rows of the CSC copy are bin2, 'other' is bin1
    order = np.lexsort((b1[m], b2[m]))
    np.testing.assert_array_equal(r, b2[m][order]), np.testing.assert_array_equal(c, b1[m][order])
    np.testing.assert_array_equal(v, cnt[m][order])

    for copy, rows in ((csr, b1[m]), (csc, b2[m][order])):
        vals = cnt[m] if copy is csr else cnt[m][order]
        nnz_m, sum_m = T.marginals_i64(copy)
        np.testing.assert_array_equal(nnz_m.cpu().numpy(),
                                      np.bincount(rows, weights=(vals != 0), minlength=n_bins).astype(np.int64))
        np.testing.assert_array_equal(sum_m.cpu().numpy(),
                                      np.bincount(rows, weights=vals, minlength=n_bins).astype(np.int64))


def test_from_arrays_int32_bin2_equals_int64():
    from cooler_b200 import table as T
    b1, b2, cnt, ind, chrom_offset = _table(3000, 200_000, 7)
    a = T.PixelTable.from_arrays(ind, b2, cnt, chrom_offset)
    b = T.PixelTable.from_arrays(ind, b2.astype(np.int32), cnt, chrom_offset)
    assert torch.equal(a.raw.px[: a.nnz], b.raw.px[: b.nnz])


def _rows_to_table(row_lens, n_cols, seed):
    """Upper-triangle-free synthetic table with prescribed row lengths (columns random, sorted)."""
    rng = np.random.default_rng(seed)
    n = len(row_lens)
    ind = np.r_[0, np.cumsum(row_lens)].astype(np.int64)
    b1 = np.repeat(np.arange(n, dtype=np.int64), row_lens)
    b2 = np.empty(ind[-1], dtype=np.int64)
    for r in np.flatnonzero(row_lens):
        k = row_lens[r]
        b2[ind[r]:ind[r + 1]] = np.sort(rng.choice(n_cols, size=k, replace=False)) if k <= n_cols // 2 \
            else np.sort(rng.permutation(n_cols)[:k])
    cnt = rng.integers(1, 100, ind[-1]).astype(np.int32)
    return b1, b2, cnt, ind


@pytest.mark.parametrize("name", ["tile_aligned", "ones", "around_chunk", "one_giant", "alternating_empty",
                                  "span_aligned", "mixed"])
def test_row_walk_adversarial_shapes(name):
    """Row lengths chosen to hit every boundary case of the warp-tile walk (rows ending exactly on
    chunk / tile / span boundaries, single-pixel rows, one row over many tiles, empty rows)."""
    import ctypes as C

    from cooler_b200 import _lib
    from cooler_b200 import table as T
    from cooler_b200.balance import IceProblem, IceState
    n_cols = 6000
    shapes = {
        "tile_aligned": [2048] * 7 + [0, 0] + [4096, 2048, 1, 2047],
        "ones": [1] * 5000,
        "around_chunk": [63, 64, 65, 1, 127, 128, 129, 0, 191, 192, 193] * 20,
        "one_giant": [0] * 10 + [5999] + [0] * 10 + [3] + [0] * 5,
        "alternating_empty": [0, 700] * 300,
        "span_aligned": [2048 * 6, 2048 * 6, 5, 2048 * 12 - 5 - 100, 100],
        "mixed": list(np.random.default_rng(5).integers(0, 3000, 400) * (np.random.default_rng(6).random(400) < 0.6)),
    }
    row_lens = np.asarray(shapes[name], dtype=np.int64)
    if name == "span_aligned":
        n_cols = 30000
    row_lens = np.minimum(row_lens, n_cols)
    n = max(len(row_lens), n_cols)
    row_lens = np.r_[row_lens, np.zeros(n - len(row_lens), dtype=np.int64)]
    b1, b2, cnt, ind = _rows_to_table(row_lens, n_cols, 3)
    chrom_offset = np.array([0, n // 2, n])
    tab = T.PixelTable.from_arrays(ind, b2, cnt, chrom_offset)
    # exact integer marginals of the raw copy
    nnz_m, sum_m = T.marginals_i64(tab.raw)
    np.testing.assert_array_equal(sum_m.cpu().numpy(), np.bincount(b1, weights=cnt, minlength=n).astype(np.int64))
    np.testing.assert_array_equal(nnz_m.cpu().numpy(), np.bincount(b1, minlength=n))
    # one fp64 sweep over both copies with a random bias, whole copies and strips (smem kernel)
    rng = np.random.default_rng(9)
    w = rng.lognormal(0, 0.5, n)
    want = np.bincount(b1, weights=w[b2] * cnt, minlength=n) + np.bincount(b2, weights=w[b1] * cnt, minlength=n)
    for strip_bins in (0, 2048):
        prob = IceProblem(tab, ignore_diags=0, strip_bins=strip_bins, band=False if strip_bins else None)
        state = IceState(prob, w, None, np.array([0, n]), 1e-5, 5)
        st = C.byref(state.st)
        _lib.call("cb200_ice_sweep", st, T._stream())
        _lib.call("cb200_ice_combine", st, T._stream())
        got = state.t["s"][:n].cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-9)


def test_segment_medians_are_bit_identical_to_numpy():
    """cb200_segment_medians (radix selection): np.median of every segment, ties, negatives,
    zeros, empty segments, odd / even counts, the positive-only rule of the MAD-max filter."""
    import torch
    from cooler_b200.table import segment_medians, _use_device
    dev = _use_device("cuda")
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.lognormal(0, 2, 100_001), -rng.lognormal(0, 1, 5000), np.zeros(777),
                        np.repeat(rng.random(50), 400), [1e-310, 1e300, -1e-310]])
    rng.shuffle(x)
    off = np.array([0, 0, 1, 2, 5, 1000, 1000, 40_000, len(x) - 1, len(x)])
    xd = torch.from_numpy(x).to(dev)
    got = segment_medians(xd, off)
    got_pos = segment_medians(xd, off, positive_only=True)
    for s, (lo, hi) in enumerate(zip(off[:-1], off[1:])):
        seg = x[lo:hi]
        want = np.median(seg) if len(seg) else np.nan
        pos = seg[seg > 0]
        want_pos = np.median(pos) if len(pos) else np.nan
        assert (np.isnan(want) and np.isnan(got[s])) or got[s] == want, (s, got[s], want)
        assert (np.isnan(want_pos) and np.isnan(got_pos[s])) or got_pos[s] == want_pos, (s, got_pos[s], want_pos)


def test_device_medians_give_the_same_filter_mask_as_numpy():
    """prefilter_bias with the medians selected on the device == the all-numpy statement-for-statement path."""
    import cooler_b200.balance as B
    from cooler_b200.table import _use_device
    dev = _use_device("cuda")
    rng = np.random.default_rng(5)
    sizes = [70_000, 50_001, 0, 3, 40_000]
    co = np.r_[0, np.cumsum(sizes)].astype(np.int64)
    n = int(co[-1])

    class P:
        pass
    p = P()
    p.table = P()
    p.table.n_bins, p.table.device = n, dev
    p.marg_sum = np.floor(rng.lognormal(7, 0.5, n))
    p.marg_sum[rng.random(n) < 0.03] = 0
    p.marg_sum[rng.random(n) < 0.01] *= 0.01
    p.marg_nnz = np.floor(p.marg_sum / 3)
    kw = dict(mad_max=5, min_nnz=10, min_count=0, blacklist=None, x0=None)
    a = B.prefilter_bias(p, co, **kw)
    old, B.DEVICE_MEDIAN_MIN_BINS = B.DEVICE_MEDIAN_MIN_BINS, 1 << 40
    try:
        b = B.prefilter_bias(p, co, **kw)
    finally:
        B.DEVICE_MEDIAN_MIN_BINS = old
    assert n >= old and (a == 0).sum() > 100
    np.testing.assert_array_equal(a, b)
