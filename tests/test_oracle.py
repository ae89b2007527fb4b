"""The oracle (oracle/*.py, a CPU restatement of the reference) against the
golden vectors frozen from the UNMODIFIED reference (tests/golden/make_golden.py).
Mirrors the reference's tests/test_balance.py and tests/test_reduce.py cases."""
import warnings

import numpy as np
import pytest

from oracle import coarsen_numpy, ice_numpy
from tests import _golden

BAL, META = _golden.balance_golden()


def _stats_equal(got, want):
    for k, w in want.items():
        g = got[k]
        if isinstance(w, list):
            g = [None if (isinstance(x, float) and np.isnan(x)) else x for x in np.asarray(g).tolist()]
            assert g == w, k
        elif isinstance(g, (float, np.floating)) and np.isnan(g):
            assert w is None, k
        else:
            assert g == w or bool(g) == w, (k, g, w)


@pytest.mark.parametrize("key", sorted(META))
def test_ice_oracle_bit_identical_to_reference(key):
    src, case = key.split("/")
    t = _golden.table_of(src)
    chrom_offset, bin1_offset = _golden.offsets(t)
    kw = dict(META[key]["kwargs"])
    if case == "x0":
        kw["x0"] = BAL["gm/x0_input"].copy()
    px = {k: t[k] for k in ("bin1_id", "bin2_id", "count")}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bias, stats, passes, warned = ice_numpy.balance_arrays(
            px, chrom_offset, bin1_offset, return_passes=True, **kw)
    assert passes == META[key]["passes"]
    assert warned == META[key]["warned"]
    np.testing.assert_array_equal(bias, BAL[key])          # bit-identical incl. NaN mask
    _stats_equal(stats, META[key]["stats"])


FBAL, FMETA = _golden.balance_float_golden()


@pytest.mark.parametrize("key", sorted(FMETA))
def test_ice_oracle_float_counts_bit_identical_to_reference(key):
    """Floating-point count columns (float64 and float32; _balance.py:25-26 copies the column's dtype)."""
    src, kind, _ = key.split("/")
    t = dict(_golden.table_of(src))
    t["count"] = _golden.float_counts(t["count"], kind)
    chrom_offset, bin1_offset = _golden.offsets(t)
    px = {k: t[k] for k in ("bin1_id", "bin2_id", "count")}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bias, stats, passes, warned = ice_numpy.balance_arrays(
            px, chrom_offset, bin1_offset, return_passes=True, **FMETA[key]["kwargs"])
    assert passes == FMETA[key]["passes"] and warned == FMETA[key]["warned"]
    np.testing.assert_array_equal(bias, FBAL[key])
    _stats_equal(stats, FMETA[key]["stats"])


@pytest.mark.parametrize("chunksize", [None, 5000, 977])
def test_ice_oracle_chunking_within_1e12(chunksize):
    t = _golden.gm12878()
    chrom_offset, bin1_offset = _golden.offsets(t)
    px = {k: t[k] for k in ("bin1_id", "bin2_id", "count")}
    bias, stats, passes, _ = ice_numpy.balance_arrays(px, chrom_offset, bin1_offset,
                                                      chunksize=chunksize, return_passes=True)
    assert passes == [14]
    np.testing.assert_allclose(bias, BAL["gm/defaults"], rtol=1e-12, equal_nan=True)


def test_reference_property_flat_marginals():
    """tests/test_balance.py:13-44 -- balanced marginals are flat."""
    t = _golden.gm12878()
    chrom_offset, bin1_offset = _golden.offsets(t)
    px = {k: t[k] for k in ("bin1_id", "bin2_id", "count")}
    tol = 1e-2
    bias, stats = ice_numpy.balance_arrays(px, chrom_offset, bin1_offset, ignore_diags=1,
                                           min_nnz=10, tol=tol)
    n = len(bias)
    m = np.zeros((n, n))
    m[t["bin1_id"], t["bin2_id"]] = t["count"]
    m = m + np.triu(m, 1).T
    np.fill_diagonal(m, 0)
    mask = np.isnan(bias)
    w = np.where(mask, 0, bias)
    marg = (m * w[:, None] * w[None, :]).sum(axis=0)[~mask]
    assert np.var(marg) < tol
    assert abs(marg.mean() - 1) < 1e-3


CO = _golden.coarsen_golden()


@pytest.mark.parametrize("factor", [2, 5, 10])
def test_coarsen_oracle_gm(factor):
    t = _golden.gm12878()
    px = {k: t[k] for k in ("bin1_id", "bin2_id", "count")}
    b1, b2, c, (nc, ns, ne) = coarsen_numpy.coarsen_pixels(
        px, t["bins_chrom"], t["bins_start"], t["bins_end"], t["chromsizes"], factor)
    np.testing.assert_array_equal(b1, CO[f"gm/x{factor}/bin1_id"])
    np.testing.assert_array_equal(b2, CO[f"gm/x{factor}/bin2_id"])
    np.testing.assert_array_equal(c, CO[f"gm/x{factor}/count"])
    assert c.dtype == np.int32
    np.testing.assert_array_equal(ns, CO[f"gm/x{factor}/bins_start"])
    np.testing.assert_array_equal(ne, CO[f"gm/x{factor}/bins_end"])


@pytest.mark.parametrize("fam,r0,r1", [("symm.upper", 2, 4), ("asymm", 2, 4), ("asymm", 4, 8),
                                       ("asymm", 8, 16), ("asymm", 16, 32), ("asymm", 2, 32)])
def test_coarsen_oracle_toy_chain(fam, r0, r1):
    a = CO[f"toy.{fam}/{r0}"]
    g = CO[f"toy.{fam}/{r1}"]
    sizes = np.array([32, 32])
    bc, bs, be = _golden.uniform_bins(sizes, r0)
    px = dict(bin1_id=a[:, 0], bin2_id=a[:, 1], count=a[:, 2].astype(np.int32))
    b1, b2, c, _ = coarsen_numpy.coarsen_pixels(px, bc, bs, be, sizes, r1 // r0)
    np.testing.assert_array_equal(np.stack([b1, b2, c], 1), g)


@pytest.mark.parametrize("factor", [4, 3, 7, 200])
def test_coarsen_oracle_odd(factor):
    a = CO["odd/1"]
    sizes = CO["odd/chromsizes"]
    bc, bs, be = _golden.uniform_bins(sizes, 1)
    px = dict(bin1_id=a[:, 0], bin2_id=a[:, 1], count=a[:, 2].astype(np.int32))
    b1, b2, c, (nc, ns, ne) = coarsen_numpy.coarsen_pixels(px, bc, bs, be, sizes, factor)
    g = CO["odd/4"] if factor == 4 else CO[f"odd/x{factor}"]
    np.testing.assert_array_equal(np.stack([b1, b2, c], 1), g)
    # no duplicate pixels (tests/test_reduce.py:134-137)
    assert len(np.unique(np.stack([b1, b2], 1), axis=0)) == len(b1)


@pytest.mark.parametrize("factor", [2, 3])
def test_coarsen_oracle_variable_bins(factor):
    bins, a = CO["var/bins"], CO["var/px"]
    px = dict(bin1_id=a[:, 0], bin2_id=a[:, 1], count=a[:, 2].astype(np.int32))
    b1, b2, c, (nc, ns, ne) = coarsen_numpy.coarsen_pixels(
        px, bins[:, 0], bins[:, 1], bins[:, 2], np.array([32, 32]), factor)
    np.testing.assert_array_equal(np.stack([b1, b2, c], 1), CO[f"var/x{factor}"])
    np.testing.assert_array_equal(np.stack([nc, ns, ne], 1), CO[f"var/x{factor}/bins"])


def test_merge_oracle_matches_reference_golden():
    parts = [dict(bin1_id=CO[f"merge/part{k}"][:, 0], bin2_id=CO[f"merge/part{k}"][:, 1],
                  count=CO[f"merge/part{k}"][:, 2].astype(np.int32)) for k in range(3)]
    b1, b2, c = coarsen_numpy.merge_pixels(parts)
    np.testing.assert_array_equal(np.stack([b1, b2, c], 1), CO["merge/out"])
    assert c.dtype == np.int32


@pytest.mark.parametrize("key", ["mean", "min", "max_last", "first", "cols_sum", "fmean", "mixed_sum"])
def test_general_merge_oracle_matches_reference_golden(key):
    """oracle.coarsen_numpy.merge_columns against the unmodified reference's CoolerMerger with value
    columns / aggregators (tests/golden/make_merge_golden.py): ids, integers and dtypes exact, floats to
    1e-12 (pandas sums groups with Kahan compensation)."""
    from tests import _merge_cases
    gm = _golden.gm12878()
    columns, agg = _merge_cases.CASES[key]
    tables = [dict(bin1_id=gm["bin1_id"][sel], bin2_id=gm["bin2_id"][sel], **cols)
              for sel, cols in _merge_cases.parts(gm)]
    got = coarsen_numpy.merge_columns(tables, columns, agg)
    G = _golden._npz("merge_agg_golden.npz")
    assert sorted(got) == sorted(k.split("/", 1)[1] for k in G if k.startswith(key + "/"))
    for k, v in got.items():
        want = G[f"{key}/{k}"]
        assert v.dtype == want.dtype, (k, v.dtype, want.dtype)
        if want.dtype.kind in "iu":
            np.testing.assert_array_equal(v, want)
        else:
            np.testing.assert_allclose(v, want, rtol=1e-12)


# --- the known answers SURVEY.md 8c records from its own probe of the unmodified reference ------------
_SURVEY_BALANCE = [   # kwargs, passes (sum), masked bins, scale, var, nansum(bias)
    ({}, 14, 118, 56.94014353567062, 3.7801217049722795e-06, 208.7776861403485),
    ({"trans_only": True}, 11, 118, 36.1473609947245, 5.314829440303912e-06, 286.2395194205529),
    ({"cis_only": True}, 295, 219, None, None, 297.58865027208265),
    ({"ignore_diags": 1, "tol": 1e-2}, 10, 120, None, 0.006900641970473413, 186.2982920213636),
    ({"mad_max": 0}, None, 105, 56.727150542796146, 3.543021571573216e-06, 214.07478676441906),
    ({"min_nnz": 0}, None, 107, 56.692392600713006, 6.064179624498289e-06, 233.53153166476125),
    ({"mad_max": 0, "min_nnz": 0}, 200, 0, 57.23676633284351, 0.005221833681348441, 2084.2446240263525),
    ({"min_count": 100}, 29, 1556, None, 8.003508570781497e-06, None),
]


@pytest.mark.parametrize("k", range(len(_SURVEY_BALANCE)))
def test_oracle_reproduces_survey_known_answers(k):
    """SURVEY.md 8c, table of known answers (the survey's own run of the unmodified reference on the
    GM12878 2 Mb fixture): pass counts, masked bins, scale, variance and the sum of the weights."""
    kw, passes, masked, scale, var, total = _SURVEY_BALANCE[k]
    t = _golden.gm12878()
    chrom_offset, bin1_offset = _golden.offsets(t)
    assert list(chrom_offset[:6]) == [0, 125, 247, 347, 443, 534] and chrom_offset[-1] == 1561
    px = {c: t[c] for c in ("bin1_id", "bin2_id", "count")}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bias, stats, n, warned = ice_numpy.balance_arrays(px, chrom_offset, bin1_offset, return_passes=True, **kw)
    if passes is not None:
        assert sum(n) == passes
    assert int(np.isnan(bias).sum()) == masked
    if kw.get("cis_only"):
        assert n[:5] == [15, 12, 11, 20, 21] and n[-2:] == [0, 0]          # chrY and chrM are empty
        np.testing.assert_allclose(stats["scale"][:2], [28.964751069515653, 30.062431951299605], rtol=1e-13)
        np.testing.assert_allclose(stats["var"][:3], [8.782106318499245e-06, 9.988300598962036e-06,
                                                      4.030316950816017e-06], rtol=1e-9)
    else:
        if scale is not None:
            assert abs(stats["scale"] / scale - 1) < 1e-13
        if var is not None:
            assert abs(stats["var"] / var - 1) < 1e-9
    if total is not None:
        assert abs(np.nansum(bias) / total - 1) < 1e-12
    assert warned == (kw == {"mad_max": 0, "min_nnz": 0})


@pytest.mark.parametrize("factor,n_new,nnz,diag,cmax,sha", [(2, 784, 30050, 742, 149, "9401b2d7389ad098"),
                                                            (5, 323, 19174, 312, 344, "a96dfa4bfa403a50"),
                                                            (10, 167, 10188, 165, 682, "dfaf335a324267bc")])
def test_coarsen_oracle_reproduces_survey_known_answers(factor, n_new, nnz, diag, cmax, sha):
    """SURVEY.md 8c, coarsening table: new bin count, output pixels, total, diagonal pixels, maximum and
    the sha1 of the (bin1, bin2, count) columns."""
    import hashlib
    t = _golden.gm12878()
    b1, b2, c, (nc, ns, ne) = coarsen_numpy.coarsen_pixels(
        {k: t[k] for k in ("bin1_id", "bin2_id", "count")}, t["bins_chrom"], t["bins_start"], t["bins_end"],
        t["chromsizes"], factor)
    assert len(nc) == n_new and len(b1) == nnz and int(c.sum()) == 100000
    assert int((b1 == b2).sum()) == diag and int(c.max()) == cmax
    h = hashlib.sha1(b1.astype("<i8").tobytes() + b2.astype("<i8").tobytes() + c.astype("<i4").tobytes())
    assert h.hexdigest()[:16] == sha
    if factor == 2:
        assert (int(ns[124 // 2]), int(ne[124 // 2])) == (248_000_000, 249_250_621)      # chr1's last x2 bin
        assert (int(ns[-1]), int(ne[-1])) == (0, 16_571)                                 # chrM stays one bin


def test_oracle_is_bit_identical_to_the_live_reference_on_random_tables():
    """300 balance runs + ~70 coarsenings on seeded random tables (symmetric-upper and square storage, integer
    and floating-point counts, all three modes, random filter settings, warm starts, blacklists) through the
    oracle and through the LIVE unmodified reference (oracle/refshim.py): bias bit-identical incl. the NaN
    pattern, pass counts, warnings, scale / var / converged equal; coarsened tables and bin tables equal.
    Runs in a subprocess (the reference needs stub modules); skipped where no copy of the reference exists."""
    import os
    import subprocess
    import sys

    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference sources not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tests", "_fuzz_oracle_vs_reference.py"), "150", "0"],
                         capture_output=True, text=True, timeout=900, cwd=root,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    assert res.stdout.strip().splitlines()[-1].startswith("ok: 300 balance runs")
