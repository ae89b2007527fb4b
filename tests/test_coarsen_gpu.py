"""Parity of the CUDA coarsening path with the reference's golden files
(toy.*.bg2 chains, tests/test_reduce.py:83-175), with reference outputs frozen in
tests/golden (GM12878 x2/x5/x10, odd.1 x4 & co., variable bins) and with the oracle
on larger random tables.  Bar: bit-exact pixels."""
import os

import numpy as np
import pytest

from tests import _golden

pytestmark = pytest.mark.gpu

CO = _golden.coarsen_golden()


def _cat(chunks):
    return np.stack([np.concatenate([c[k] for c in chunks]) for k in ("bin1_id", "bin2_id", "count")], 1) \
        if chunks else np.zeros((0, 3), dtype=np.int64)


def _mem(names, sizes, bc, bs, be, px, binsize, symm=True):
    from cooler_b200 import MemCooler
    return MemCooler.from_arrays(names, sizes, bc, bs, be, px[:, 0], px[:, 1], px[:, 2].astype(np.int32),
                                 binsize=binsize, symmetric_upper=symm)


@pytest.mark.parametrize("factor", [2, 5, 10])
@pytest.mark.parametrize("chunksize", [10_000, 777])
def test_coarsen_gm12878(factor, chunksize):
    from cooler_b200.reduce import CoolerCoarsener
    t = _golden.gm12878()
    px = np.stack([t["bin1_id"], t["bin2_id"], t["count"]], 1)
    clr = _mem(t["chromnames"], t["chromsizes"], t["bins_chrom"], t["bins_start"], t["bins_end"], px, 2_000_000)
    it = CoolerCoarsener(clr, factor, chunksize)
    chunks = list(it)
    got = _cat(chunks)
    np.testing.assert_array_equal(got[:, 0], CO[f"gm/x{factor}/bin1_id"])
    np.testing.assert_array_equal(got[:, 1], CO[f"gm/x{factor}/bin2_id"])
    np.testing.assert_array_equal(got[:, 2], CO[f"gm/x{factor}/count"])
    assert chunks[0]["bin1_id"].dtype == np.int64 and chunks[0]["count"].dtype == np.int32
    nb = it.new_bins
    np.testing.assert_array_equal(nb["start"].values, CO[f"gm/x{factor}/bins_start"])
    np.testing.assert_array_equal(nb["end"].values, CO[f"gm/x{factor}/bins_end"])
    np.testing.assert_array_equal(nb["chrom"].cat.codes.values, CO[f"gm/x{factor}/bins_chrom"])
    # chunks never split a coarse row and are globally sorted
    for a, b in zip(chunks[:-1], chunks[1:]):
        assert a["bin1_id"][-1] < b["bin1_id"][0]


@pytest.mark.parametrize("fam,r0,r1", [("symm.upper", 2, 4), ("asymm", 2, 4), ("asymm", 4, 8),
                                       ("asymm", 8, 16), ("asymm", 16, 32), ("asymm", 2, 32)])
def test_coarsen_toy_golden_files(fam, r0, r1):
    from cooler_b200.reduce import coarsen_cooler
    sizes = np.array([32, 32])
    bc, bs, be = _golden.uniform_bins(sizes, r0)
    clr = _mem(["chr1", "chr2"], sizes, bc, bs, be, CO[f"toy.{fam}/{r0}"], r0, symm=(fam == "symm.upper"))
    out = coarsen_cooler(clr, None, r1 // r0, chunksize=7)
    with out.open() as g:
        got = np.stack([g["pixels/bin1_id"][:], g["pixels/bin2_id"][:], g["pixels/count"][:]], 1)
    np.testing.assert_array_equal(got, CO[f"toy.{fam}/{r1}"])
    assert out.binsize == r1 and out.info["nbins"] == 2 * (32 // r1)
    assert out.storage_mode == clr.storage_mode


def test_zoomify_toy_chain():
    """tests/test_reduce.py:140-175: toy.asymm.2 -> [4, 8, 16, 32], each vs its golden."""
    from cooler_b200.reduce import zoomify_cooler
    sizes = np.array([32, 32])
    bc, bs, be = _golden.uniform_bins(sizes, 2)
    clr = _mem(["chr1", "chr2"], sizes, bc, bs, be, CO["toy.asymm/2"], 2, symm=False)
    out = zoomify_cooler(clr, None, [4, 8, 16, 32], chunksize=10)
    for res in (4, 8, 16, 32):
        with out[res].open() as g:
            got = np.stack([g["pixels/bin1_id"][:], g["pixels/bin2_id"][:], g["pixels/count"][:]], 1)
        np.testing.assert_array_equal(got, CO[f"toy.asymm/{res}"])
    with pytest.raises(ValueError):
        zoomify_cooler(clr, None, [4, 5, 32], chunksize=10)


@pytest.mark.parametrize("factor", [4, 3, 7, 200])
def test_coarsen_odd_partitions_correctly(factor):
    from cooler_b200.reduce import CoolerCoarsener
    sizes = CO["odd/chromsizes"]
    bc, bs, be = _golden.uniform_bins(sizes, 1)
    clr = _mem(["chr1", "chr2", "chr3"], sizes, bc, bs, be, CO["odd/1"], 1)
    got = _cat(list(CoolerCoarsener(clr, factor, 2)))
    want = CO["odd/4"] if factor == 4 else CO[f"odd/x{factor}"]
    np.testing.assert_array_equal(got, want)
    assert len(np.unique(got[:, :2], axis=0)) == len(got)


@pytest.mark.parametrize("factor", [2, 3])
def test_coarsen_variable_bins(factor):
    from cooler_b200.reduce import CoolerCoarsener
    bins = CO["var/bins"]
    clr = _mem(["chr1", "chr2"], np.array([32, 32]), bins[:, 0], bins[:, 1], bins[:, 2], CO["var/px"], None)
    it = CoolerCoarsener(clr, factor, 11)
    got = _cat(list(it))
    np.testing.assert_array_equal(got, CO[f"var/x{factor}"])
    nb = it.new_bins
    np.testing.assert_array_equal(
        np.stack([nb["chrom"].cat.codes.values, nb["start"].values, nb["end"].values], 1), CO[f"var/x{factor}/bins"])


def test_coarsen_missing_column_raises():
    from cooler_b200.reduce import coarsen_cooler
    sizes = np.array([32, 32])
    bc, bs, be = _golden.uniform_bins(sizes, 2)
    clr = _mem(["chr1", "chr2"], sizes, bc, bs, be, CO["toy.asymm/2"], 2, symm=False)
    with pytest.raises(ValueError):
        coarsen_cooler(clr, None, 2, 10, columns=["missing"])


@pytest.mark.parametrize("factor", [2, 5])
def test_coarsen_matches_oracle_large(factor):
    """~2e6 pixels, long and empty rows, a 1-bin chromosome."""
    from cooler_b200 import PixelTable
    from cooler_b200.reduce import coarsen_table, table_to_host
    from oracle import coarsen_numpy
    from tests.test_balance_gpu import _synthetic
    b1, b2, cnt, chrom_offset = _synthetic(21, [30000, 17001, 1, 9000], 0.0015)
    n = int(chrom_offset[-1])
    sizes = np.diff(chrom_offset) * 1000
    bc, bs, be = _golden.uniform_bins(sizes, 1000)
    want = coarsen_numpy.coarsen_pixels(dict(bin1_id=b1, bin2_id=b2, count=cnt), bc, bs, be, sizes, factor)
    tab = PixelTable.from_arrays(np.searchsorted(b1, np.arange(n + 1)), b2, cnt, chrom_offset)
    new, ids = coarsen_table(tab, factor)
    g1, g2, gc = table_to_host(new, ids)
    np.testing.assert_array_equal(g1, want[0]), np.testing.assert_array_equal(g2, want[1])
    np.testing.assert_array_equal(gc, want[2])
    assert int(gc.sum()) == int(cnt.sum())           # size-independent invariant: total count preserved
    # the coarse table is itself a valid input: its indptr matches its ids
    np.testing.assert_array_equal(new.raw.indptr.cpu().numpy(), np.searchsorted(g1, np.arange(new.n_bins + 1)))


@pytest.mark.parametrize("factor", [2, 3, 5, 16, 64])
def test_coarsen_many_small_chromosomes(factor):
    """400 chromosomes of 1 .. 300 bins (plus empty ones): blocks of the column-remap table hold no, one or
    several chromosome boundaries (the latter fall back to the bin_map gather); rounds 1 .. 6."""
    from cooler_b200 import PixelTable
    from cooler_b200.reduce import coarsen_table, table_to_host
    from oracle import coarsen_numpy
    rng = np.random.default_rng(factor)
    nb = rng.integers(1, 300, 400)
    nb[::7] = 1
    nb[5::50] = 0                                     # chromosomes without bins
    chrom_offset = np.r_[0, np.cumsum(nb)].astype(np.int64)
    n = int(chrom_offset[-1])
    m = 1_500_000
    i = rng.integers(0, n, m)
    j = np.where(rng.random(m) < 0.5, np.minimum(i + rng.integers(0, 40, m), n - 1), rng.integers(0, n, m))
    i, j = np.minimum(i, j), np.maximum(i, j)
    key = np.unique(i.astype(np.int64) * n + j)
    b1, b2 = key // n, key % n
    cnt = rng.integers(1, 100, len(b1)).astype(np.int32)
    sizes = np.maximum(nb, 0) * 1000
    bc, bs, be = _golden.uniform_bins(sizes, 1000)
    want = coarsen_numpy.coarsen_pixels(dict(bin1_id=b1, bin2_id=b2, count=cnt), bc, bs, be, sizes, factor)
    tab = PixelTable.from_arrays(np.searchsorted(b1, np.arange(n + 1)), b2, cnt, chrom_offset)
    new, ids = coarsen_table(tab, factor)
    g1, g2, gc = table_to_host(new, ids)
    np.testing.assert_array_equal(g1, want[0]), np.testing.assert_array_equal(g2, want[1])
    np.testing.assert_array_equal(gc, want[2])
    assert int(gc.sum()) == int(cnt.sum())


def _gm_part(k):
    t = _golden.gm12878()
    return _mem(t["chromnames"], t["chromsizes"], t["bins_chrom"], t["bins_start"], t["bins_end"],
                CO[f"merge/part{k}"], 2_000_000)


@pytest.mark.parametrize("mergebuf", [20_000_000, 5000])
def test_merge_matches_reference_golden(mergebuf):
    """3-way merge vs the unmodified reference's CoolerMerger (tests/golden/make_golden.py)."""
    from cooler_b200.reduce import CoolerMerger
    chunks = list(CoolerMerger([_gm_part(k) for k in range(3)], mergebuf))
    got = _cat(chunks)
    np.testing.assert_array_equal(got, CO["merge/out"])
    if mergebuf == 5000:
        assert len(chunks) > 3


def test_merge_coolers_semantics():
    """tests/test_reduce.py:34-75 re-targeted: doubled counts, dtype promotion, ValueErrors."""
    from cooler_b200.reduce import merge_coolers
    sizes = np.array([32, 32])
    bc, bs, be = _golden.uniform_bins(sizes, 2)
    toy2 = _mem(["chr1", "chr2"], sizes, bc, bs, be, CO["toy.symm.upper/2"], 2)
    out = merge_coolers(None, [toy2, toy2], mergebuf=3)
    with out.open() as g:
        got = np.stack([g["pixels/bin1_id"][:], g["pixels/bin2_id"][:], g["pixels/count"][:]], 1)
    want = CO["toy.symm.upper/2"].copy()
    want[:, 2] *= 2
    np.testing.assert_array_equal(got, want)
    bc4, bs4, be4 = _golden.uniform_bins(sizes, 4)
    toy4 = _mem(["chr1", "chr2"], sizes, bc4, bs4, be4, CO["toy.symm.upper/4"], 4)
    with pytest.raises(ValueError):
        merge_coolers(None, [toy2, toy4], mergebuf=3)                 # different resolution
    asym = _mem(["chr1", "chr2"], sizes, bc, bs, be, CO["toy.asymm/2"], 2, symm=False)
    with pytest.raises(ValueError):
        merge_coolers(None, [toy2, asym], mergebuf=3)                 # incompatible symmetry
    bins = CO["var/bins"]
    var = _mem(["chr1", "chr2"], sizes, bins[:, 0], bins[:, 1], bins[:, 2], CO["var/px"], None)
    with pytest.raises(ValueError):
        merge_coolers(None, [var, toy2], mergebuf=3)                  # incompatible bins
    with pytest.raises(ValueError):
        merge_coolers(None, [toy2, var], mergebuf=3)
    with pytest.raises(ValueError):
        merge_coolers(None, [toy2, toy2], mergebuf=3, columns=["missing"])


def test_merge_matches_oracle_large():
    from cooler_b200 import PixelTable
    from cooler_b200.reduce import merge_tables, table_to_host
    from oracle import coarsen_numpy
    from tests.test_balance_gpu import _synthetic
    tabs, pxs = [], []
    for seed in (41, 42, 43, 44):
        b1, b2, cnt, chrom_offset = _synthetic(seed, [30000, 17001, 1, 9000], 0.0008)
        n = int(chrom_offset[-1])
        pxs.append(dict(bin1_id=b1, bin2_id=b2, count=cnt))
        tabs.append(PixelTable.from_arrays(np.searchsorted(b1, np.arange(n + 1)), b2, cnt, chrom_offset))
    want = coarsen_numpy.merge_pixels(pxs)
    merged, ids = merge_tables(tabs)
    g1, g2, gc = table_to_host(merged, ids)
    np.testing.assert_array_equal(g1, want[0]), np.testing.assert_array_equal(g2, want[1])
    np.testing.assert_array_equal(gc, want[2])
    assert int(gc.sum()) == sum(int(p["count"].sum()) for p in pxs)


def _merge_part_coolers():
    from cooler_b200 import MemCooler
    from tests import _merge_cases
    gm = _golden.gm12878()
    out = []
    for sel, cols in _merge_cases.parts(gm):
        extra = {c: v for c, v in cols.items() if c != "count"}
        out.append(MemCooler.from_arrays([str(x) for x in gm["chromnames"]], gm["chromsizes"], gm["bins_chrom"],
                                         gm["bins_start"], gm["bins_end"], gm["bin1_id"][sel], gm["bin2_id"][sel],
                                         cols["count"], binsize=2_000_000, extra_pixel_columns=extra))
    return out


@pytest.mark.parametrize("mergebuf", [20_000_000, 5000])
@pytest.mark.parametrize("key", ["mean", "min", "max_last", "first", "cols_sum", "fmean", "mixed_sum"])
def test_general_merge_matches_reference_golden(key, mergebuf):
    """CoolerMerger with value columns / aggregators (mean, min, max, first, last; float64, wide int64 and
    mixed int32 / float32 columns) against the UNMODIFIED reference (pandas concat + groupby,
    _reduce.py:188-203; tests/golden/make_merge_golden.py): ids, integer results and dtypes exact,
    floating-point results to 1e-12."""
    from cooler_b200.reduce import CoolerMerger
    from tests import _merge_cases
    columns, agg = _merge_cases.CASES[key]
    chunks = list(CoolerMerger(_merge_part_coolers(), mergebuf, columns=columns, agg=agg))
    got = {k: np.concatenate([c[k] for c in chunks]) for k in chunks[0]}
    G = _golden._npz("merge_agg_golden.npz")
    assert list(got) == [k.split("/", 1)[1] for k in G if k.startswith(key + "/")]
    for k, v in got.items():
        want = G[f"{key}/{k}"]
        assert v.dtype == want.dtype, (k, v.dtype, want.dtype)
        if want.dtype.kind in "iu":
            np.testing.assert_array_equal(v, want)
        else:
            np.testing.assert_allclose(v, want, rtol=1e-12)
    if mergebuf == 5000:
        assert len(chunks) > 3


def test_merge_coolers_keeps_extra_columns_and_float_counts():
    """merge_coolers(None, ...) with columns=[count, fcount]: both land in the result with the promoted
    dtypes (merge_coolers: np.result_type over the inputs, _reduce.py:286-291)."""
    from cooler_b200.reduce import merge_coolers
    G = _golden._npz("merge_agg_golden.npz")
    out = merge_coolers(None, _merge_part_coolers(), mergebuf=10_000, columns=["count", "fcount", "big"])
    with out.open() as g:
        np.testing.assert_array_equal(g["pixels/bin2_id"][:], G["cols_sum/bin2_id"])
        np.testing.assert_array_equal(g["pixels/count"][:], G["cols_sum/count"])
        np.testing.assert_array_equal(g["pixels/big"][:], G["cols_sum/big"])
        np.testing.assert_allclose(g["pixels/fcount"][:], G["cols_sum/fcount"], rtol=1e-12)
    out = merge_coolers(None, _merge_part_coolers(), mergebuf=10_000, columns=["mixed"])
    with out.open() as g:
        assert g["pixels/mixed"][:].dtype == np.float64
        np.testing.assert_allclose(g["pixels/mixed"][:], G["mixed_sum/mixed"], rtol=1e-12)


# --- general value columns / aggregators (reference tests/test_reduce.py:100-118) ---------
@pytest.mark.parametrize("key", ["gm_mean_x2", "gm_min_x5", "gm_max_x3", "gm_first_x2", "gm_last_x7",
                                 "gm_cols_x2", "gm_fsum_x5", "gmf_sum_x2", "gmf_mean_x10"])
def test_general_aggregation_matches_reference_golden(key):
    """CoolerCoarsener with agg = mean / min / max / first / last, extra value columns (float64, wide
    int64) and a float32 count column, against the UNMODIFIED reference (pandas groupby): ids and
    integer results bit-exact, float64 results to 1e-12, float32 to 1e-6; dtypes as pandas gives them."""
    from cooler_b200 import MemCooler
    from cooler_b200.reduce import CoolerCoarsener
    from tests import _agg_cases
    tab, factor, columns, agg = _agg_cases.CASES[key]
    gm = _golden.gm12878()
    fcount, big = _agg_cases.derived(gm)
    count = gm["count"] if tab == "gm" else (gm["count"] * 0.5).astype(np.float32)
    clr = MemCooler.from_arrays([str(x) for x in gm["chromnames"]], gm["chromsizes"], gm["bins_chrom"],
                                gm["bins_start"], gm["bins_end"], gm["bin1_id"], gm["bin2_id"], count,
                                binsize=2_000_000, extra_pixel_columns={"fcount": fcount, "big": big})
    chunks = list(CoolerCoarsener(clr, factor, 5000, columns=columns, agg=agg))
    got = {k: np.concatenate([c[k] for c in chunks]) for k in chunks[0]}
    G = dict(np.load(os.path.join(_golden.GOLD, "coarsen_agg_golden.npz")))
    assert sorted(got) == sorted(k.split("/", 1)[1] for k in G if k.startswith(key + "/"))
    for k, v in got.items():
        want = G[f"{key}/{k}"]
        assert v.dtype == want.dtype, (k, v.dtype, want.dtype)
        if want.dtype.kind in "iu":
            np.testing.assert_array_equal(v, want)
        else:
            np.testing.assert_allclose(v, want, rtol=1e-6 if want.dtype == np.float32 else 1e-12)
    # chunks never split a coarse row and stay sorted / duplicate-free
    keyv = got["bin1_id"] * (got["bin2_id"].max() + 1) + got["bin2_id"]
    assert np.all(np.diff(keyv) > 0)


def test_coarsen_cooler_with_float_dtype_and_mean(tmp_path):
    """`--field count:dtype=float,agg=mean` (reference tests/test_cli_ops.py:58-65): the output column is
    float and holds the means."""
    from cooler_b200 import MemCooler, store
    from cooler_b200.cli import cli
    from click.testing import CliRunner
    gm = _golden.gm12878()
    clr = MemCooler.from_arrays([str(x) for x in gm["chromnames"]], gm["chromsizes"], gm["bins_chrom"],
                                gm["bins_start"], gm["bins_end"], gm["bin1_id"], gm["bin2_id"], gm["count"],
                                binsize=2_000_000)
    store.REGISTRY.clear()
    store.register("gm.cool", clr)
    out = str(tmp_path / "mean.cool")
    res = CliRunner().invoke(cli, ["coarsen", "gm.cool", "--factor", "2", "--field", "count:dtype=float,agg=mean",
                                   "-o", out])
    assert res.exit_code == 0, (res.output, res.exception)
    G = dict(np.load(os.path.join(_golden.GOLD, "coarsen_agg_golden.npz")))
    with store.open_cooler(out).open("r") as grp:
        cnt = np.asarray(grp["pixels"]["count"][:])
        assert cnt.dtype.kind == "f"
        np.testing.assert_allclose(cnt, G["gm_mean_x2/count"], rtol=1e-12)
        np.testing.assert_array_equal(np.asarray(grp["pixels"]["bin2_id"][:]), G["gm_mean_x2/bin2_id"])
