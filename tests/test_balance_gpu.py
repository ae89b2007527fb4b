"""Parity of the CUDA ICE path (through the C ABI) with the golden vectors of
the unmodified reference and with the oracle.  Bar (BASELINE.json north_star):
filter masks and iteration counts bit-exact, bias within 1e-9 relative."""
import logging
import warnings

import numpy as np
import pytest

from tests import _golden

pytestmark = pytest.mark.gpu

BAL, META = _golden.balance_golden()
RTOL = 1e-9


def _mem_cooler(t):
    from cooler_b200 import MemCooler
    return MemCooler.from_arrays(t["chromnames"], t["chromsizes"], t["bins_chrom"], t["bins_start"],
                                 t["bins_end"], t["bin1_id"], t["bin2_id"], t["count"],
                                 binsize=int(t["binsize"]),
                                 symmetric_upper=not bool((t["bin2_id"] < t["bin1_id"]).any()))


class _Count(logging.Handler):
    def __init__(self):
        super().__init__(level=logging.INFO)
        self.lines = []

    def emit(self, record):
        self.lines.append(record.getMessage())


def _check_stats(got, want, cis):
    for k in ("tol", "min_nnz", "min_count", "mad_max", "cis_only", "ignore_diags", "divisive_weights"):
        assert got[k] == want[k], k
    for k in ("scale", "var"):
        g = np.atleast_1d(np.asarray(got[k], dtype=float))
        w = np.array([np.nan if x is None else x for x in np.atleast_1d(want[k])], dtype=float)
        np.testing.assert_allclose(g, w, rtol=1e-9, atol=0, equal_nan=True, err_msg=k)
    np.testing.assert_array_equal(np.atleast_1d(got["converged"]), np.atleast_1d(want["converged"]))


@pytest.mark.parametrize("key", sorted(META))
def test_balance_matches_reference_golden(key):
    import cooler_b200
    src, case = key.split("/")
    clr = _mem_cooler(_golden.table_of(src))
    kw = dict(META[key]["kwargs"])
    if case == "x0":
        kw["x0"] = BAL["gm/x0_input"].copy()
    lg = logging.getLogger("cooler")
    h = _Count()
    lg.addHandler(h)
    old = lg.level
    lg.setLevel(logging.INFO)
    try:
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            bias, stats = cooler_b200.balance_cooler(clr, **kw)
    finally:
        lg.removeHandler(h)
        lg.setLevel(old)
    warned = any(issubclass(x.category, cooler_b200.ConvergenceWarning) for x in w)
    want = BAL[key]
    # masks bit-exact
    np.testing.assert_array_equal(np.isnan(bias), np.isnan(want))
    # iteration counts bit-exact ("variance is" log lines, as the reference logs them)
    n_var = sum(1 for m in h.lines if m.startswith("variance is"))
    assert n_var == sum(META[key]["passes"])
    assert warned == META[key]["warned"]
    # bias within 1e-9 relative
    np.testing.assert_allclose(bias, want, rtol=RTOL, atol=0, equal_nan=True)
    _check_stats(stats, META[key]["stats"], kw.get("cis_only", False))
    if case == "x0":
        assert bias is kw["x0"]          # the reference mutates and returns x0 (_balance.py:368-370)


FBAL, FMETA = _golden.balance_float_golden()


@pytest.mark.parametrize("strip_bins", [None, 256])
@pytest.mark.parametrize("key", sorted(FMETA))
def test_balance_float_counts_match_reference_golden(key, strip_bins):
    """Floating-point count columns (float64 / float32, stored zeros included) against the unmodified
    reference: masks and pass counts identical, bias / scale / var within 1e-9.  ``strip_bins`` 256 forces
    the column-strip kernel (value arrays attached per strip)."""
    import cooler_b200
    src, kind, _ = key.split("/")
    t = dict(_golden.table_of(src))
    t["count"] = _golden.float_counts(t["count"], kind)
    kw = dict(FMETA[key]["kwargs"])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        if strip_bins is None:                       # through the drop-in entry point
            bias, stats = cooler_b200.balance_cooler(_mem_cooler(t), **kw)
            passes = None
        else:
            chrom_offset, bin1_offset = _golden.offsets(t)
            bias, stats, info = cooler_b200.balance_arrays(bin1_offset, t["bin2_id"], t["count"], chrom_offset,
                                                           strip_bins=strip_bins, return_info=True, **kw)
            assert info["strip_shift"] == 8 and info["n_subcopies"] > 2
            passes = [p for p in info["passes"] if p]
    warned = any(issubclass(x.category, cooler_b200.ConvergenceWarning) for x in w)
    want = FBAL[key]
    np.testing.assert_array_equal(np.isnan(bias), np.isnan(want))
    if passes is not None:
        assert passes == [p for p in FMETA[key]["passes"] if p]
    assert warned == FMETA[key]["warned"]
    np.testing.assert_allclose(bias, want, rtol=RTOL, atol=0, equal_nan=True)
    _check_stats(stats, FMETA[key]["stats"], kw.get("cis_only", False))


def test_balance_float_counts_equal_integer_counts_when_integral():
    """The same integers stored as float64 give the integer path's result (different kernels, same sums)."""
    import cooler_b200
    t = _golden.gm12878()
    chrom_offset, bin1_offset = _golden.offsets(t)
    a, sa = cooler_b200.balance_arrays(bin1_offset, t["bin2_id"], t["count"], chrom_offset)
    b, sb = cooler_b200.balance_arrays(bin1_offset, t["bin2_id"], t["count"].astype(np.float64), chrom_offset)
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_allclose(a, b, rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(a, BAL["gm/defaults"], rtol=RTOL, equal_nan=True)
    assert sa["converged"] == sb["converged"]


def _synthetic(seed, n_per_chrom, density):
    rng = np.random.default_rng(seed)
    n = int(sum(n_per_chrom))
    chrom_offset = np.r_[0, np.cumsum(n_per_chrom)].astype(np.int64)
    b = rng.lognormal(0, 0.3, n)
    b[rng.random(n) < 0.02] *= 0.01
    nnz_target = int(density * n * n / 2)
    i = rng.integers(0, n, nnz_target)
    d = np.minimum((rng.pareto(0.7, nnz_target) * 3).astype(np.int64), n - 1)
    far = rng.random(nnz_target) < 0.3
    j = np.where(far, rng.integers(0, n, nnz_target), np.minimum(i + d, n - 1))
    i, j = np.minimum(i, j), np.maximum(i, j)
    key = np.unique(i.astype(np.int64) * n + j)
    b1, b2 = key // n, key % n
    cnt = (1 + rng.poisson(30 * b[b1] * b[b2] / (1 + np.abs(b1 - b2)) ** 0.5)).astype(np.int32)
    return b1, b2, cnt, chrom_offset


@pytest.mark.parametrize("mode", ["gw", "cis", "trans"])
@pytest.mark.parametrize("seed,n_per_chrom,density", [(11, [3000, 2500, 40, 1800, 1], 0.05),
                                                      (12, [20000, 15000, 9000], 0.004)])
def test_balance_matches_oracle_synthetic(mode, seed, n_per_chrom, density):
    """Sizes the oracle finishes in seconds; rows from empty to tens of thousands of pixels."""
    import cooler_b200
    from cooler_b200 import PixelTable
    from oracle import ice_numpy
    b1, b2, cnt, chrom_offset = _synthetic(seed, n_per_chrom, density)
    n = int(chrom_offset[-1])
    bin1_offset = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    kw = dict(cis_only=(mode == "cis"), trans_only=(mode == "trans"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want, wstats, wpasses, wwarned = ice_numpy.balance_arrays(
            dict(bin1_id=b1, bin2_id=b2, count=cnt), chrom_offset, bin1_offset,
            return_passes=True, **kw)
        tab = PixelTable.from_arrays(bin1_offset, b2, cnt, chrom_offset)
        bias, stats, info = cooler_b200.balance_table(tab, return_info=True, **kw)
    np.testing.assert_array_equal(np.isnan(bias), np.isnan(want))
    assert info["passes"] == wpasses
    np.testing.assert_allclose(bias, want, rtol=RTOL, atol=0, equal_nan=True)
    np.testing.assert_allclose(np.atleast_1d(stats["var"]), np.atleast_1d(wstats["var"]), rtol=1e-9, equal_nan=True)
    np.testing.assert_allclose(np.atleast_1d(stats["scale"]), np.atleast_1d(wstats["scale"]), rtol=1e-9, equal_nan=True)


@pytest.mark.parametrize("mode", ["gw", "cis", "trans"])
def test_balance_float_counts_match_oracle_default_strips(mode):
    """60 000 bins, float64 values: the automatic layout is column strips of 16 384 bins (the bias vector no
    longer fits L1); compared with the oracle, which is bit-identical to the reference on float columns."""
    import cooler_b200
    from oracle import ice_numpy
    b1, b2, cnt, chrom_offset = _synthetic(21, [30000, 20000, 10000], 0.008)
    rng = np.random.default_rng(3)
    val = cnt * rng.lognormal(0.0, 0.25, len(cnt))
    val[rng.random(len(val)) < 0.01] = 0.0
    n = int(chrom_offset[-1])
    bin1_offset = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    kw = dict(cis_only=(mode == "cis"), trans_only=(mode == "trans"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want, wstats, wpasses, _ = ice_numpy.balance_arrays(
            dict(bin1_id=b1, bin2_id=b2, count=val), chrom_offset, bin1_offset, return_passes=True, **kw)
        bias, stats, info = cooler_b200.balance_arrays(bin1_offset, b2, val, chrom_offset, return_info=True, **kw)
    assert not info["band"] and (info["strip_shift"] == 14 or mode == "trans")   # few trans pixels: whole copies
    np.testing.assert_array_equal(np.isnan(bias), np.isnan(want))
    assert info["passes"] == wpasses
    np.testing.assert_allclose(bias, want, rtol=RTOL, atol=0, equal_nan=True)
    np.testing.assert_allclose(np.atleast_1d(stats["var"]), np.atleast_1d(wstats["var"]), rtol=1e-9, equal_nan=True)


def test_balance_is_deterministic():
    import cooler_b200
    from cooler_b200 import PixelTable
    b1, b2, cnt, chrom_offset = _synthetic(5, [5000, 3000], 0.02)
    n = int(chrom_offset[-1])
    bin1_offset = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    outs = []
    for _ in range(2):
        tab = PixelTable.from_arrays(bin1_offset, b2, cnt, chrom_offset)
        outs.append(cooler_b200.balance_table(tab)[0])
    np.testing.assert_array_equal(outs[0], outs[1])       # bit-identical run to run


def test_store_writes_weight_column():
    import cooler_b200
    clr = _mem_cooler(_golden.gm12878())
    bias, stats = cooler_b200.balance_cooler(clr, store=True, store_name="weight")
    with clr.open("r") as grp:
        np.testing.assert_array_equal(grp["bins"]["weight"][:], bias)
        assert grp["bins"]["weight"].attrs["converged"]


def test_flat_marginals_property():
    """The reference's own test (tests/test_balance.py:13-44) re-targeted at the CUDA path."""
    import cooler_b200
    t = _golden.gm12878()
    clr = _mem_cooler(t)
    tol = 1e-2
    bias, stats = cooler_b200.balance_cooler(clr, ignore_diags=1, min_nnz=10, tol=tol)
    n = len(bias)
    m = np.zeros((n, n))
    m[t["bin1_id"], t["bin2_id"]] = t["count"]
    m = m + np.triu(m, 1).T
    np.fill_diagonal(m, 0)
    mask = np.isnan(bias)
    w = np.where(mask, 0, bias)
    marg = (m * w[:, None] * w[None, :]).sum(axis=0)[~mask]
    assert np.var(marg) < tol


@pytest.mark.parametrize("key,strip_bins", [("gm/defaults", 64), ("gm/cis_only", 128), ("gm/trans_only", 256),
                                            ("rnd3/gw", 32), ("rnd3/cis", 64), ("rnd3/trans", 16),
                                            ("gm/diag0", 512), ("gm/mad0_nnz0", 1024),
                                            ("sq22/gw", 64), ("sq22/cis", 32), ("sq21/trans", 16)])
def test_column_strips_match_reference_golden(key, strip_bins):
    """Cache-blocked layout (column strips of the gathered bias) gives the same answers."""
    import cooler_b200
    from cooler_b200 import PixelTable
    src, case = key.split("/")
    t = _golden.table_of(src)
    chrom_offset, bin1_offset = _golden.offsets(t)
    kw = dict(META[key]["kwargs"])
    tab = PixelTable.from_arrays(bin1_offset, t["bin2_id"], t["count"], chrom_offset)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bias, stats, info = cooler_b200.balance_table(tab, return_info=True, strip_bins=strip_bins, band=False, **kw)
    assert info["n_subcopies"] > 2 and info["strip_shift"] is not None and not info["band"]
    want = BAL[key]
    np.testing.assert_array_equal(np.isnan(bias), np.isnan(want))
    assert sum(info["passes"]) == sum(META[key]["passes"])
    np.testing.assert_allclose(bias, want, rtol=RTOL, atol=0, equal_nan=True)


def test_column_strips_large_synthetic_matches_unstripped():
    import cooler_b200
    from cooler_b200 import PixelTable
    b1, b2, cnt, chrom_offset = _synthetic(31, [40000, 30000, 20001], 0.003)
    n = int(chrom_offset[-1])
    bin1_offset = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    tab = PixelTable.from_arrays(bin1_offset, b2, cnt, chrom_offset)
    from cooler_b200.balance import IceProblem
    a, sa, ia = cooler_b200.balance_table(tab, return_info=True, strip_bins=0)
    b, sb, ib = cooler_b200.balance_table(tab, return_info=True, strip_bins=16384, band=False)
    assert ia["n_subcopies"] == 2 and ib["n_subcopies"] == 12
    assert ia["passes"] == ib["passes"]
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_allclose(a, b, rtol=1e-12, equal_nan=True)
    # the L1-gather kernel on the same strips is bit-identical to the shared-memory kernel
    prob = IceProblem(tab, strip_bins=16384, use_smem=False, band=False)
    c, sc, ic = cooler_b200.balance_table(tab, return_info=True, problem=prob)
    np.testing.assert_array_equal(b, c)


@pytest.mark.parametrize("key,strip_bins", [("gm/defaults", 64), ("gm/cis_only", 128), ("gm/trans_only", 32),
                                            ("rnd3/gw", 16), ("rnd2/cis", 16), ("rnd3/trans", 32)])
def test_band_layout_matches_reference_golden(key, strip_bins):
    """Diagonal-band strips + far remainder (the layout for very sparse, very large matrices)."""
    import cooler_b200
    from cooler_b200 import PixelTable
    src, case = key.split("/")
    t = _golden.table_of(src)
    chrom_offset, bin1_offset = _golden.offsets(t)
    kw = dict(META[key]["kwargs"])
    tab = PixelTable.from_arrays(bin1_offset, t["bin2_id"], t["count"], chrom_offset)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bias, stats, info = cooler_b200.balance_table(tab, return_info=True, strip_bins=strip_bins, band=True, **kw)
    assert info["band"] and info["n_subcopies"] > 2
    want = BAL[key]
    np.testing.assert_array_equal(np.isnan(bias), np.isnan(want))
    assert sum(info["passes"]) == sum(META[key]["passes"])
    np.testing.assert_allclose(bias, want, rtol=RTOL, atol=0, equal_nan=True)


def test_band_layout_large_synthetic_matches_unstripped():
    import cooler_b200
    from cooler_b200 import PixelTable
    b1, b2, cnt, chrom_offset = _synthetic(33, [50000, 30000, 20001], 0.002)
    n = int(chrom_offset[-1])
    bin1_offset = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    tab = PixelTable.from_arrays(bin1_offset, b2, cnt, chrom_offset)
    a, sa, ia = cooler_b200.balance_table(tab, return_info=True, strip_bins=0)
    b, sb, ib = cooler_b200.balance_table(tab, return_info=True, strip_bins=16384, band=True)
    assert ib["band"] and ib["n_subcopies"] > 4
    assert ia["passes"] == ib["passes"]
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_allclose(a, b, rtol=1e-12, equal_nan=True)


@pytest.mark.parametrize("key,strip_bins,col_shift", [
    ("gm/defaults", 64, 4), ("gm/cis_only", 128, 6), ("gm/trans_only", 32, 3),
    ("rnd3/gw", 16, 2), ("rnd2/cis", 16, 5), ("rnd3/trans", 32, 12), ("gm/diag0", 16, 1)])
def test_far_blocks_layout_matches_reference_golden(key, strip_bins, col_shift):
    """Band strips + far-field blocks (csrc/far.cuh: 2-D blocked far pixels, segmented-scan kernel)."""
    import cooler_b200
    from cooler_b200 import PixelTable
    src, case = key.split("/")
    t = _golden.table_of(src)
    chrom_offset, bin1_offset = _golden.offsets(t)
    kw = dict(META[key]["kwargs"])
    tab = PixelTable.from_arrays(bin1_offset, t["bin2_id"], t["count"], chrom_offset)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bias, stats, info = cooler_b200.balance_table(tab, return_info=True, strip_bins=strip_bins, band="blocks",
                                                      far_col_shift=col_shift, **kw)
    assert info["band"] == "blocks"
    if case != "cis_only":
        assert info["far_units"] > 0 and info["far_pixels"] > 0
    want = BAL[key]
    np.testing.assert_array_equal(np.isnan(bias), np.isnan(want))
    assert sum(info["passes"]) == sum(META[key]["passes"])
    np.testing.assert_allclose(bias, want, rtol=RTOL, atol=0, equal_nan=True)


@pytest.mark.parametrize("strip_bins,col_shift", [(1024, 8), (4096, 12), (256, 10), (16384, 12), (512, 5),
                                                  (1024, 9), (8192, 13), (16384, 13)])
def test_far_blocks_large_synthetic_matches_unstripped(strip_bins, col_shift):
    """Several 16384-row blocks, hundreds of column tiles, rows from empty to dense; every
    far-kernel variant agrees with the whole-copy sweep to rounding and is bit-reproducible."""
    import cooler_b200
    from cooler_b200 import PixelTable
    b1, b2, cnt, chrom_offset = _synthetic(35, [50000, 30000, 20001], 0.002)
    n = int(chrom_offset[-1])
    bin1_offset = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    tab = PixelTable.from_arrays(bin1_offset, b2, cnt, chrom_offset)
    a, sa, ia = cooler_b200.balance_table(tab, return_info=True, strip_bins=0)
    b, sb, ib = cooler_b200.balance_table(tab, return_info=True, strip_bins=strip_bins, band="blocks",
                                          far_col_shift=col_shift)
    c, sc, ic = cooler_b200.balance_table(tab, return_info=True, strip_bins=strip_bins, band="blocks",
                                          far_col_shift=col_shift)
    assert ib["band"] == "blocks" and ib["far_units"] > 4 and ib["far_pixels"] > 100000
    assert ia["passes"] == ib["passes"]
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_allclose(a, b, rtol=1e-12, equal_nan=True)
    np.testing.assert_array_equal(b, c)


@pytest.mark.parametrize("compact", [False, True])
@pytest.mark.parametrize("key,col_shift", [("gm/defaults", 4), ("gm/cis_only", 7), ("gm/trans_only", 5),
                                           ("rnd3/gw", 2), ("gm/diag0", 12), ("gm/mad0_nnz0", 6), ("rnd1/cis", 13),
                                           ("sq22/gw", 5), ("sq22/cis", 4)])
def test_all_blocks_layout_matches_reference_golden(key, col_shift, compact):
    """Every pixel (dense diagonal runs included) through the 2-D blocked kernel, with 8-byte and with
    compact 4-byte records (the GM12878 2 Mb counts reach thousands: every such pixel is split into
    pieces of <= 127, far beyond the 25 % growth at which the automatic choice keeps 8-byte records)."""
    import cooler_b200
    from cooler_b200 import PixelTable
    src, case = key.split("/")
    t = _golden.table_of(src)
    chrom_offset, bin1_offset = _golden.offsets(t)
    kw = dict(META[key]["kwargs"])
    tab = PixelTable.from_arrays(bin1_offset, t["bin2_id"], t["count"], chrom_offset)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import cooler_b200.table as tbl
        growth, tbl.FAR_COMPACT_MAX_GROWTH = tbl.FAR_COMPACT_MAX_GROWTH, 1e9
        try:
            bias, stats, info = cooler_b200.balance_table(tab, return_info=True, band="allblocks",
                                                          far_col_shift=col_shift, far_compact=compact, **kw)
        finally:
            tbl.FAR_COMPACT_MAX_GROWTH = growth
    assert info["band"] == "allblocks" and info["n_subcopies"] == 0 and info["far_units"] > 0
    assert info["far_rec_bytes"] == [4 if compact else 8] * 2
    assert info["far_records"] >= info["far_pixels"] and (compact or info["far_records"] == info["far_pixels"])
    want = BAL[key]
    np.testing.assert_array_equal(np.isnan(bias), np.isnan(want))
    assert sum(info["passes"]) == sum(META[key]["passes"])
    np.testing.assert_allclose(bias, want, rtol=RTOL, atol=0, equal_nan=True)


def test_all_blocks_large_synthetic_matches_unstripped():
    import cooler_b200
    from cooler_b200 import PixelTable
    b1, b2, cnt, chrom_offset = _synthetic(37, [50000, 30000, 20001], 0.002)
    n = int(chrom_offset[-1])
    bin1_offset = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    tab = PixelTable.from_arrays(bin1_offset, b2, cnt, chrom_offset)
    a, sa, ia = cooler_b200.balance_table(tab, return_info=True, strip_bins=0)
    for col_shift, compact in ((13, None), (13, False), (10, True)):   # the default 8192-column tiles and narrow ones
        b, sb, ib = cooler_b200.balance_table(tab, return_info=True, band="allblocks", far_col_shift=col_shift,
                                              far_compact=compact)
        assert ib["far_pixels"] == 2 * ia["nnz_eff"] and ia["passes"] == ib["passes"]
        assert ib["far_rec_bytes"] == [8 if compact is False else 4] * 2
        np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
        np.testing.assert_allclose(a, b, rtol=1e-12, equal_nan=True)


@pytest.mark.parametrize("mode", ["cis", "trans"])
def test_far_blocks_modes_match_oracle(mode):
    import cooler_b200
    from cooler_b200 import PixelTable
    from oracle import ice_numpy
    b1, b2, cnt, chrom_offset = _synthetic(36, [20000, 15000, 9000], 0.004)
    n = int(chrom_offset[-1])
    bin1_offset = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    kw = dict(cis_only=(mode == "cis"), trans_only=(mode == "trans"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want, wstats, wpasses, wwarned = ice_numpy.balance_arrays(
            dict(bin1_id=b1, bin2_id=b2, count=cnt), chrom_offset, bin1_offset, return_passes=True, **kw)
        tab = PixelTable.from_arrays(bin1_offset, b2, cnt, chrom_offset)
        bias, stats, info = cooler_b200.balance_table(tab, return_info=True, strip_bins=512, band="blocks",
                                                      far_col_shift=9, **kw)
    assert info["far_units"] > 0 and info["passes"] == wpasses
    np.testing.assert_array_equal(np.isnan(bias), np.isnan(want))
    np.testing.assert_allclose(bias, want, rtol=RTOL, atol=0, equal_nan=True)


@pytest.mark.parametrize("case", ["empty", "diag_only", "one_bin", "one_pixel", "all_masked"])
@pytest.mark.parametrize("mode", ["gw", "cis", "trans"])
def test_degenerate_tables_match_oracle(case, mode):
    """Edge cases of the reference: no pixels, everything filtered, a single bin."""
    import cooler_b200
    from cooler_b200 import PixelTable
    from oracle import ice_numpy
    if case == "empty":
        chrom_offset = np.array([0, 5, 9]); b1 = b2 = np.zeros(0, np.int64); cnt = np.zeros(0, np.int32)
    elif case == "diag_only":
        chrom_offset = np.array([0, 5, 9]); b1 = np.arange(9); b2 = b1.copy(); cnt = np.full(9, 7, np.int32)
    elif case == "one_bin":
        chrom_offset = np.array([0, 1]); b1 = np.array([0]); b2 = np.array([0]); cnt = np.array([3], np.int32)
    elif case == "one_pixel":
        chrom_offset = np.array([0, 4, 8]); b1 = np.array([1]); b2 = np.array([6]); cnt = np.array([5], np.int32)
    else:
        chrom_offset = np.array([0, 6, 12])
        i, j = np.triu_indices(12, 2)
        b1, b2, cnt = i, j, np.ones(len(i), np.int32)
    n = int(chrom_offset[-1])
    b1, b2 = b1.astype(np.int64), b2.astype(np.int64)
    bin1_offset = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    kw = dict(cis_only=(mode == "cis"), trans_only=(mode == "trans"))
    if case == "all_masked":
        kw["min_nnz"] = 1000
    elif case in ("one_pixel",):
        kw.update(min_nnz=0, mad_max=0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want, wstats, wpasses, _ = ice_numpy.balance_arrays(dict(bin1_id=b1, bin2_id=b2, count=cnt), chrom_offset,
                                                           bin1_offset, return_passes=True, max_iters=50, **kw)
        tab = PixelTable.from_arrays(bin1_offset, b2, cnt, chrom_offset)
        bias, stats, info = cooler_b200.balance_table(tab, return_info=True, max_iters=50, **kw)
    np.testing.assert_array_equal(np.isnan(bias), np.isnan(want))
    np.testing.assert_allclose(bias, want, rtol=RTOL, equal_nan=True)
    assert info["passes"] == wpasses
    np.testing.assert_allclose(np.atleast_1d(stats["var"]), np.atleast_1d(wstats["var"]), rtol=1e-9, equal_nan=True)
    np.testing.assert_allclose(np.atleast_1d(stats["scale"]), np.atleast_1d(wstats["scale"]), rtol=1e-9, equal_nan=True)
    np.testing.assert_array_equal(np.atleast_1d(stats["converged"]), np.atleast_1d(wstats["converged"]))


@pytest.mark.parametrize("key", ["gm/defaults", "gm/cis_only", "gm/trans_only", "gm/diag0", "rnd3/gw",
                                 "rnd3/cis", "rnd3/trans", "rnd1/cis", "sq22/gw", "sq22/cis", "sq21/trans"])
def test_streamed_build_matches_reference_golden(key):
    """balance_arrays: row-chunked build pipelined with the H2D copy (chunk CSR/CSC sub-copies)."""
    import cooler_b200
    src, case = key.split("/")
    t = _golden.table_of(src)
    chrom_offset, bin1_offset = _golden.offsets(t)
    kw = dict(META[key]["kwargs"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bias, stats, info = cooler_b200.balance_arrays(bin1_offset, t["bin2_id"], t["count"], chrom_offset,
                                                       return_info=True, **kw)
    assert info["n_subcopies"] > 2                        # several row chunks were used
    want = BAL[key]
    np.testing.assert_array_equal(np.isnan(bias), np.isnan(want))
    assert sum(info["passes"]) == sum(META[key]["passes"])
    np.testing.assert_allclose(bias, want, rtol=RTOL, atol=0, equal_nan=True)
    _check_stats(stats, META[key]["stats"], kw.get("cis_only", False))


@pytest.mark.parametrize("mode", ["gw", "cis", "trans"])
def test_streamed_build_large_synthetic_matches_table_path(mode):
    import cooler_b200
    from cooler_b200 import PixelTable
    import torch
    b1, b2, cnt, chrom_offset = _synthetic(51, [9000, 7000, 3, 5000], 0.02)
    n = int(chrom_offset[-1])
    bin1_offset = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    kw = dict(cis_only=(mode == "cis"), trans_only=(mode == "trans"))
    tab = PixelTable.from_arrays(bin1_offset, b2, cnt, chrom_offset)
    a, sa, ia = cooler_b200.balance_table(tab, return_info=True, **kw)
    pinned = [torch.from_numpy(x).pin_memory() for x in (bin1_offset, b2, cnt)]
    for b2_dtype in (torch.int64, torch.int32):
        b, sb, ib = cooler_b200.balance_arrays(pinned[0], pinned[1].to(b2_dtype).pin_memory(), pinned[2],
                                               chrom_offset, return_info=True, **kw)
        assert ib["n_subcopies"] >= 8 and ia["passes"] == ib["passes"] and ib["nnz_eff"] == ia["nnz_eff"]
        np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
        np.testing.assert_allclose(a, b, rtol=1e-12, equal_nan=True)


def test_auto_layout_uses_blocks_for_huge_sparse_tables():
    """From 2**20 bins on, sparse tables are swept by the 2-D blocked kernel automatically
    (balance.choose_layout); results equal the whole-copy sweep, and balance_arrays (the call
    balance_cooler makes) takes the same layout through the two-step build."""
    import cooler_b200
    from cooler_b200 import PixelTable
    from cooler_b200.balance import BLOCKS_MIN_BINS, choose_layout
    rng = np.random.default_rng(5)
    sizes = [600_000, 400_000, BLOCKS_MIN_BINS - 1_000_000 + 4097]
    chrom_offset = np.r_[0, np.cumsum(sizes)].astype(np.int64)
    n = int(chrom_offset[-1])
    assert n >= BLOCKS_MIN_BINS
    m = 24 * n
    i = rng.integers(0, n, m)
    d = np.where(rng.random(m) < 0.5, rng.integers(0, 200, m), rng.integers(0, n, m))
    j = np.minimum(i + d, n - 1)
    key = np.unique(i.astype(np.int64) * n + j)
    b1, b2 = key // n, key % n
    hidden = rng.lognormal(0, 0.3, n)
    cnt = (1 + rng.poisson(2 * hidden[b1] * hidden[b2])).astype(np.int32)
    bin1_offset = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    assert choose_layout(n, len(b1)) == (None, "allblocks")
    tab = PixelTable.from_arrays(bin1_offset, b2, cnt, chrom_offset)
    a, sa, ia = cooler_b200.balance_table(tab, return_info=True, strip_bins=0)
    b, sb, ib = cooler_b200.balance_table(tab, return_info=True)
    assert ia["band"] is False and ib["band"] == "allblocks" and ib["far_units"] > 0 and ib["n_subcopies"] == 0
    assert ia["passes"] == ib["passes"] and sa["converged"] and sb["converged"]
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_allclose(a, b, rtol=1e-12, equal_nan=True)
    c, sc = cooler_b200.balance_arrays(bin1_offset, b2, cnt, chrom_offset)
    np.testing.assert_array_equal(b, c)
    assert abs(sc["scale"] / sb["scale"] - 1) < 1e-12


def test_far_record_format_choice():
    """Compact 4-byte far-field records are the automatic choice when the counts allow it: counts
    above 127 are split (and must add up to the same sweep), a negative count or too many heavy
    counts keep the 8-byte records."""
    import cooler_b200
    from cooler_b200 import PixelTable
    b1, b2, cnt, chrom_offset = _synthetic(41, [9000, 6000, 3001], 0.01)
    n = int(chrom_offset[-1])
    bin1_offset = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    cnt = cnt.copy()
    heavy = np.flatnonzero(np.abs(b1 - b2) >= 2)[:: 301]       # +8 % records after splitting
    cnt[heavy] = 127 + (np.arange(len(heavy)) % 900) * 7            # 127 .. 6420: one to 51 pieces
    ref = None
    for name, c, want_bytes in (("split", cnt, 4), ("forced8", cnt, 8), ("negative", None, 8), ("heavy", None, 8)):
        cc = c
        kw = {}
        if name == "forced8":
            kw["far_compact"] = False
        if name == "negative":
            cc = cnt.copy()
            cc[heavy[0]] = -3
        if name == "heavy":
            cc = cnt * 200
        tab = PixelTable.from_arrays(bin1_offset, b2, cc, chrom_offset)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            a, sa, ia = cooler_b200.balance_table(tab, return_info=True, strip_bins=0, max_iters=30)
            b, sb, ib = cooler_b200.balance_table(tab, return_info=True, band="allblocks", far_col_shift=9,
                                                  max_iters=30, **kw)
        assert ib["far_rec_bytes"] == [want_bytes] * 2, (name, ib["far_rec_bytes"])
        assert ib["far_pixels"] == 2 * ia["nnz_eff"]
        assert (ib["far_records"] > ib["far_pixels"]) == (name == "split")
        assert ia["passes"] == ib["passes"]
        np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
        np.testing.assert_allclose(a, b, rtol=1e-11, equal_nan=True)
        if name in ("split", "forced8"):
            ref = a if ref is None else ref
            np.testing.assert_array_equal(a, ref)


@pytest.mark.parametrize("mode", ["gw", "cis", "trans"])
@pytest.mark.parametrize("shard", [None, (5000, 17000)])
def test_streamed_strip_build_is_bit_identical_to_table_path(mode, shard, monkeypatch):
    """Column-strip layout through balance_arrays: one H2D chunk per strip-aligned row range, the
    bin2-major copy of a chunk is one strip, the bin1-major pieces are concatenated per strip.  The
    sub-copies equal those of the two-step build, so the bias is identical to the last bit."""
    import cooler_b200
    from cooler_b200 import PixelTable, balance
    import torch
    monkeypatch.setattr(balance, "L1_RESIDENT_BINS", 1000)
    monkeypatch.setattr(balance, "DEFAULT_STRIP_BINS", 2048)
    monkeypatch.setattr(balance, "MIN_PIXELS_PER_ROW_STRIP", 0)
    b1, b2, cnt, chrom_offset = _synthetic(52, [9000, 7000, 3, 5000], 0.02)
    n = int(chrom_offset[-1])
    if shard is not None:
        keep = (b1 >= shard[0]) & (b1 < shard[1])
        b1, b2, cnt = b1[keep], b2[keep], cnt[keep]
    bin1_offset = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    kw = dict(cis_only=(mode == "cis"), trans_only=(mode == "trans"), row_range=shard)
    tab = PixelTable.from_arrays(bin1_offset, b2, cnt, chrom_offset, row_range=shard)
    kw2 = {k: v for k, v in kw.items() if k != "row_range"}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a, sa, ia = cooler_b200.balance_table(tab, return_info=True, max_iters=60, **kw2)
        pinned = [torch.from_numpy(x).pin_memory() for x in (bin1_offset, b2.astype(np.int32), cnt)]
        b, sb, ib = cooler_b200.balance_arrays(pinned[0], pinned[1], pinned[2], chrom_offset, return_info=True,
                                               max_iters=60, **kw)
    assert ia["strip_shift"] == 11 and ib["strip_shift"] == 11
    assert ib["n_subcopies"] == ia["n_subcopies"] and ia["passes"] == ib["passes"] and ib["nnz_eff"] == ia["nnz_eff"]
    np.testing.assert_array_equal(a, b)
