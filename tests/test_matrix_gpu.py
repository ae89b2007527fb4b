"""GPU: balanced matrix fetch (csrc/fetch.cu through the C ABI) against golden vectors from the
unmodified reference and against the oracle on larger random tables.  Bar: counts bit-exact,
balanced values bit-exact as well (same two roundings as `arr * np.outer(b1, b2)`)."""
import numpy as np
import pytest

from tests import _golden, _matrix_cases as MC

pytestmark = pytest.mark.gpu


def _clr(t, symmetric_upper=True, weights=None):
    from cooler_b200 import MemCooler
    clr = MemCooler.from_arrays([str(x) for x in t["chromnames"]], t["chromsizes"], t["bins_chrom"],
                                t["bins_start"], t["bins_end"], t["bin1_id"], t["bin2_id"], t["count"],
                                binsize=int(t["binsize"]) if "binsize" in t else 2_000_000,
                                symmetric_upper=symmetric_upper)
    if weights is not None:
        with clr.open("r+") as g:
            g["bins"].create_dataset("weight", data=weights)
    return clr


@pytest.mark.parametrize("case", MC.DENSE)
def test_dense_fetch_matches_reference_golden(case):
    import cooler_b200
    t, w, (i0, i1, j0, j1), kw = MC.inputs(case)
    clr = _clr(t, kw["symmetric_upper"], w)
    sel = cooler_b200.matrix(clr, balance=w is not None, divisive_weights=kw.get("divisive_weights", False))
    got = sel[i0:i1, j0:j1]
    want = MC.GOLD[case]
    assert got.shape == want.shape and got.dtype == want.dtype
    np.testing.assert_array_equal(got, want)


def test_region_fetch_and_missing_weights():
    import cooler_b200
    t, w, _, _ = MC.inputs("chr1_bal")
    clr = _clr(t, True, w)
    sel = cooler_b200.matrix(clr, balance=True)
    np.testing.assert_array_equal(sel.fetch("chr1"), MC.GOLD["chr1_bal"])
    np.testing.assert_array_equal(sel.fetch("chr5", "chr2:10,000,000-50,000,000"), MC.GOLD["rect_lower_bal"])
    raw = cooler_b200.matrix(clr, balance=False, table=sel.table)
    np.testing.assert_array_equal(raw.fetch("chr2:10,000,000-50,000,000", "chr5"), MC.GOLD["rect_raw"])
    assert sel.shape == (1561, 1561)
    with pytest.raises(ValueError, match="No column 'bins/nope'"):
        cooler_b200.matrix(clr, balance="nope")
    # weights handed over directly (e.g. the bias balance_cooler just returned)
    bias, _ = cooler_b200.balance_cooler(clr)
    direct = cooler_b200.matrix(clr, weights=bias, table=sel.table).fetch("chr1")
    np.testing.assert_allclose(direct, MC.GOLD["chr1_bal"], rtol=1e-9, equal_nan=True)


def test_4dn_weight_columns_are_divisive_by_default():
    """balance='KR' | 'VC' | 'VC_SQRT' (4DN / hic2cool columns) divide unless divisive_weights is
    given, exactly like Cooler.matrix (api.py:28, 367-368): the golden divisive window is reproduced."""
    import cooler_b200
    t, w, (i0, i1, j0, j1), kw = MC.inputs("divisive")
    clr = _clr(t, True, w)
    with clr.open("r+") as g:
        g["bins"].create_dataset("KR", data=w)
    want = MC.GOLD["divisive"]
    np.testing.assert_array_equal(cooler_b200.matrix(clr, balance="KR")[i0:i1, j0:j1], want)
    np.testing.assert_array_equal(clr.matrix(balance="KR")[i0:i1, j0:j1], want)
    mult = cooler_b200.matrix(clr, balance="KR", divisive_weights=False)[i0:i1, j0:j1]
    np.testing.assert_array_equal(mult, cooler_b200.matrix(clr, balance=True)[i0:i1, j0:j1])
    assert not np.array_equal(mult, want, equal_nan=True)


@pytest.mark.parametrize("case", MC.SPARSE)
def test_sparse_fetch_matches_reference_golden(case):
    import cooler_b200
    t, w, (i0, i1, j0, j1), kw = MC.inputs(case)
    m = cooler_b200.matrix(_clr(t, True, w), balance=w is not None, sparse=True)[i0:i1, j0:j1]
    o = np.lexsort((m.col, m.row))
    np.testing.assert_array_equal(m.row[o], MC.GOLD[case + "/row"])
    np.testing.assert_array_equal(m.col[o], MC.GOLD[case + "/col"])
    np.testing.assert_array_equal(m.data[o], MC.GOLD[case + "/data"])
    assert m.shape == tuple(MC.GOLD[case + "/shape"])


@pytest.mark.parametrize("case", MC.PIXELS)
def test_pixel_fetch_matches_reference_golden(case):
    import cooler_b200
    t, w, (i0, i1, j0, j1), kw = MC.inputs(case)
    df = cooler_b200.matrix(_clr(t, True, w), balance=w is not None, as_pixels=True)[i0:i1, j0:j1]
    for c in ("bin1_id", "bin2_id", "count") + (("balanced",) if w is not None else ()):
        np.testing.assert_array_equal(df[c].to_numpy(), MC.GOLD[f"{case}/{c}"])


def test_dense_fetch_matches_oracle_on_random_table():
    import cooler_b200
    from cooler_b200 import PixelTable
    from oracle import matrix_numpy
    rng = np.random.default_rng(5)
    n = 20000
    nnz = 600000
    i = rng.integers(0, n, nnz)
    j = np.minimum(i + (rng.pareto(0.8, nnz) * 4).astype(np.int64), n - 1)
    far = rng.random(nnz) < 0.2
    j = np.where(far, rng.integers(0, n, nnz), j)
    i, j = np.minimum(i, j), np.maximum(i, j)
    key = np.unique(i * n + j)
    b1, b2 = key // n, key % n
    cnt = rng.integers(0, 1000, len(key)).astype(np.int32)            # explicit zeros included
    w = rng.lognormal(0, 0.3, n)
    w[rng.random(n) < 0.05] = np.nan
    off = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    tab = PixelTable.from_arrays(off, b2, cnt, np.array([0, 9000, n]))
    for sym in (True, False):
        sel = cooler_b200.api.MatrixSelector(None, tab, w, fill_lower=sym)
        raw = cooler_b200.api.MatrixSelector(None, tab, None, fill_lower=sym)
        for (i0, i1, j0, j1) in [(0, 3000, 0, 3000), (5000, 5600, 4800, 9000), (9000, 20000, 100, 900),
                                 (19990, 20000, 0, 20000), (7, 8, 7, 8), (100, 100, 0, 50)]:
            want = matrix_numpy.fetch_dense(b1, b2, cnt, i0, i1, j0, j1, symmetric_upper=sym, weights=w)
            np.testing.assert_array_equal(sel[i0:i1, j0:j1], want)
            want = matrix_numpy.fetch_dense(b1, b2, cnt, i0, i1, j0, j1, symmetric_upper=sym)
            np.testing.assert_array_equal(raw[i0:i1, j0:j1], want)
        sp = cooler_b200.api.MatrixSelector(None, tab, w, sparse=True, fill_lower=sym)[4000:9500, 3000:9100]
        r, c, d = matrix_numpy.fetch_sparse(b1, b2, cnt, 4000, 9500, 3000, 9100, symmetric_upper=sym, weights=w)
        o = np.lexsort((sp.col, sp.row))
        np.testing.assert_array_equal(sp.row[o], r)
        np.testing.assert_array_equal(sp.col[o], c)
        np.testing.assert_array_equal(sp.data[o], d)


def test_cooler_objects_have_a_matrix_method():
    """clr.matrix(balance=...).fetch(...) on the store objects, table uploaded once."""
    import os

    from cooler_b200 import store
    clr = store.open_cooler(os.path.join(_golden.GOLD, "cool", "hg19.GM12878-MboI.matrix.2000kb.cool"))
    a = clr.matrix(balance=False).fetch("chr21")
    tab = clr._b200_table
    b = clr.matrix(balance=False, sparse=True).fetch("chr21")
    assert clr._b200_table is tab and a.shape == (25, 25) and np.array_equal(a, b.toarray())
    assert np.array_equal(a, a.T) and a.sum() > 0
    with pytest.raises(ValueError, match="No column 'bins/weight'"):
        clr.matrix(balance=True)


# --- `field=`: value columns other than an int32 count (api.py:326-329; cb200_fetch_*_f64) ----------
from tests import _matrix_field_cases as MF  # noqa: E402


def _field_clr(t, weights):
    from cooler_b200 import MemCooler
    clr = MemCooler.from_arrays([str(x) for x in t["chromnames"]], t["chromsizes"], t["bins_chrom"],
                                t["bins_start"], t["bins_end"], t["bin1_id"], t["bin2_id"], t["count"],
                                binsize=2_000_000, extra_pixel_columns=t["extra"] or None)
    if weights is not None:
        with clr.open("r+") as g:
            g["bins"].create_dataset("weight", data=weights)
    return clr


@pytest.mark.parametrize("key", sorted(MF.CASES))
def test_field_fetch_matches_reference_golden(key):
    """matrix(field=...) on float64 / wide int64 extra columns and on a float32 count column: dense, sparse
    and as_pixels, raw and balanced -- values and dtypes bit-identical to the unmodified reference."""
    import cooler_b200
    t, field, w, (i0, i1, j0, j1), kind, extra = MF.inputs(key)
    G = MF.gold()
    sel = cooler_b200.matrix(_field_clr(t, w), field=field, balance=w is not None, sparse=kind == "sparse",
                             as_pixels=kind == "pixels", **extra)
    got = sel[i0:i1, j0:j1]
    if kind == "dense":
        assert got.dtype == G[key].dtype and got.shape == G[key].shape
        np.testing.assert_array_equal(got, G[key])
    elif kind == "sparse":
        o = np.lexsort((got.col, got.row))
        np.testing.assert_array_equal(got.row[o], G[key + "/row"])
        np.testing.assert_array_equal(got.col[o], G[key + "/col"])
        assert got.data.dtype == G[key + "/data"].dtype
        np.testing.assert_array_equal(got.data[o], G[key + "/data"])
    else:
        cols = ["bin1_id", "bin2_id", field] + (["balanced"] if w is not None else [])
        assert list(got.columns) == cols
        for c in cols:
            assert got[c].to_numpy().dtype == G[f"{key}/{c}"].dtype
            np.testing.assert_array_equal(got[c].to_numpy(), G[f"{key}/{c}"])


def test_field_fetch_errors_and_table_reuse():
    import cooler_b200
    t, field, w, win, kind, extra = MF.inputs("f_dense_raw")
    clr = _field_clr(t, None)
    with pytest.raises(KeyError):
        cooler_b200.matrix(clr, field="nope", balance=False)
    a = clr.matrix(balance=False)[100:140, 90:130]                   # int32 counts: table kept on the store
    tab = clr._b200_table
    f = clr.matrix(field="fcount", balance=False)[100:140, 90:130]   # the kept table serves the structure
    assert clr._b200_table is tab and f.dtype == np.float64 and a.dtype == np.int32
    np.testing.assert_array_equal(f != 0, a != 0)
    np.testing.assert_array_equal(f, MF.gold()["f_dense_raw"])
