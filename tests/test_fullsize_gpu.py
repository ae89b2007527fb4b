"""BASELINE.json full-size configuration (synthetic hg38 100 kb, ~2e8 pixels, one GPU),
checked against the CPU oracle run with a process pool (a few seconds) and through size-independent properties:
flat marginals after balancing, idempotence (a warm start from the answer converges in one
pass), run-to-run bit reproducibility, layout independence (whole copies vs column strips),
and conservation of the total count + sortedness under coarsening."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    import cooler_b200
    from cooler_b200 import synth
    plan = synth.SynthPlan(100_000, 200_000_000, seed=0)
    data = synth.generate(plan)
    tab = cooler_b200.PixelTable.from_arrays(data["bin1_offset"], data["bin2_id"], data["count"],
                                             data["chrom_offset"])
    return plan, data, tab


def test_c2_balance_properties(c2):
    import cooler_b200
    plan, data, tab = c2
    assert 1.6e8 < data["nnz"] < 2.4e8 and plan.n_bins == 30895
    bias, stats, info = cooler_b200.balance_table(tab, return_info=True)
    assert stats["converged"] and 3 <= info["passes"][0] <= 200
    masked = np.isnan(bias)
    assert 0 < masked.sum() < 0.1 * len(bias)                 # the 1 % bad bins (and a few more) are filtered
    # reproducible bit for bit
    bias2, _ = cooler_b200.balance_table(tab)
    np.testing.assert_array_equal(bias, bias2)
    # idempotent: warm start from the answer is converged after one pass, marginals flat (mean 1, var < tol)
    x0 = bias.copy()
    b3, s3, i3 = cooler_b200.balance_table(tab, x0=x0, return_info=True)
    assert i3["passes"] == [1] and s3["converged"] and s3["var"] < 1e-5
    assert abs(s3["scale"] - 1.0) < 1e-2
    np.testing.assert_array_equal(np.isnan(b3), masked)
    np.testing.assert_allclose(b3[~masked], bias[~masked], rtol=2e-2)
    # layout independence: column strips give the same answer as whole copies
    b4, s4, i4 = cooler_b200.balance_table(tab, strip_bins=8192, band=False, return_info=True)
    assert i4["n_subcopies"] > 2 and i4["passes"] == info["passes"]
    np.testing.assert_array_equal(np.isnan(b4), masked)
    np.testing.assert_allclose(b4[~masked], bias[~masked], rtol=1e-11)


def test_c2_cis_and_trans_converge(c2):
    import cooler_b200
    plan, data, tab = c2
    b, s, i = cooler_b200.balance_table(tab, cis_only=True, return_info=True)
    assert len(i["passes"]) == 25 and np.all(s["converged"])
    b, s, i = cooler_b200.balance_table(tab, trans_only=True, return_info=True)
    assert s["converged"]


def test_c2_coarsen_invariants(c2):
    from cooler_b200.reduce import coarsen_table
    plan, data, tab = c2
    total = int(data["count"].to(torch.int64).sum())
    new, ids = coarsen_table(tab, 5)
    px = new.raw.px[: new.nnz]
    assert int(px[:, 1].to(torch.int64).sum()) == total        # counts conserved
    key = ids.to(torch.int64) * new.n_bins + px[:, 0].to(torch.int64)
    assert bool((key[1:] > key[:-1]).all())                    # strictly sorted by (bin1, bin2): no duplicates
    assert bool((px[:, 0] >= ids).all())                       # stays upper triangular
    ind = new.raw.indptr
    assert int(ind[-1]) == new.nnz and bool((ind[1:] >= ind[:-1]).all())
    # coarsening twice by 5 and 2 == once by 10
    a, ida = coarsen_table(new, 2)
    b, idb = coarsen_table(tab, 10)
    assert a.nnz == b.nnz
    assert torch.equal(a.raw.px[: a.nnz], b.raw.px[: b.nnz]) and torch.equal(ida, idb)


def _oracle_pool_balance(data, plan, **kw):
    """The CPU oracle (oracle/ice_numpy.py) run to convergence with a process pool over pixel spans
    (the reference's `-p N` structure), a few seconds at 2e8 pixels."""
    import multiprocessing as mp
    import os
    import warnings

    from oracle import ice_numpy
    off = data["bin1_offset"].numpy()
    n = plan.n_bins
    px = {"bin1_id": np.repeat(np.arange(n, dtype=np.int64), np.diff(off)),
          "bin2_id": data["bin2_id"].numpy(), "count": data["count"].numpy()}
    P = os.cpu_count() or 1
    chrom_ids = np.repeat(np.arange(len(plan.chrom_offset) - 1, dtype=np.int32), np.diff(plan.chrom_offset))
    ice_numpy.share(px, chrom_ids, n)              # inherited by the forked workers: only spans are pickled
    with mp.get_context("fork").Pool(P) as pool, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        chunk = max(1_000_000, len(px["bin2_id"]) // (4 * P))
        return ice_numpy.balance_arrays(px, plan.chrom_offset, off, chunksize=chunk, map=pool.imap,
                                        return_passes=True, **kw)


def test_c2_full_size_matches_oracle(c2):
    """BASELINE configs[1] at full size against the CPU oracle: mask and pass count exact, bias,
    scale and var to 1e-9 (through the default layout and through balance_arrays, the call
    balance_cooler makes)."""
    import cooler_b200
    plan, data, tab = c2
    want, wstats, wpasses, _ = _oracle_pool_balance(data, plan)
    bias, stats, info = cooler_b200.balance_table(tab, return_info=True)
    assert info["passes"] == wpasses
    np.testing.assert_array_equal(np.isnan(bias), np.isnan(want))
    np.testing.assert_allclose(bias, want, rtol=1e-9, equal_nan=True)
    assert abs(stats["scale"] / wstats["scale"] - 1) < 1e-9 and abs(stats["var"] / wstats["var"] - 1) < 1e-9
    b2, s2 = cooler_b200.balance_arrays(data["bin1_offset"], data["bin2_id"], data["count"], data["chrom_offset"])
    np.testing.assert_allclose(b2, want, rtol=1e-9, equal_nan=True)


@pytest.mark.parametrize("layout", ["strips", "allblocks"])
def test_default_geometry_layouts_match_oracle_on_a_huge_sparse_table(layout):
    """The layouts the 10 kb and 1 kb configurations get -- column strips of 16 384 bins and the 2-D
    blocked kernel with 4096-row blocks x 8192-column tiles -- at their DEFAULT geometry, on a table with
    2**20 + 4097 bins that the oracle still finishes in seconds: mask, pass count exact, bias to 1e-9."""
    import warnings

    import cooler_b200
    from cooler_b200 import PixelTable
    from cooler_b200.balance import BLOCKS_MIN_BINS, DEFAULT_STRIP_BINS, FAR_COL_SHIFT
    from oracle import ice_numpy
    rng = np.random.default_rng(5)
    sizes = [600_000, 400_000, BLOCKS_MIN_BINS - 1_000_000 + 4097]
    chrom_offset = np.r_[0, np.cumsum(sizes)].astype(np.int64)
    n = int(chrom_offset[-1])
    m = 24 * n
    i = rng.integers(0, n, m)
    d = np.where(rng.random(m) < 0.5, rng.integers(0, 200, m), rng.integers(0, n, m))
    j = np.minimum(i + d, n - 1)
    key = np.unique(i.astype(np.int64) * n + j)
    b1, b2 = key // n, key % n
    hidden = rng.lognormal(0, 0.3, n)
    cnt = (1 + rng.poisson(2 * hidden[b1] * hidden[b2])).astype(np.int32)
    bin1_offset = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want, wstats, wpasses, _ = ice_numpy.balance_arrays(dict(bin1_id=b1, bin2_id=b2, count=cnt), chrom_offset,
                                                           bin1_offset, return_passes=True)
    tab = PixelTable.from_arrays(bin1_offset, b2, cnt, chrom_offset)
    kw = dict(strip_bins=DEFAULT_STRIP_BINS, band=False) if layout == "strips" else dict(band="allblocks")
    bias, stats, info = cooler_b200.balance_table(tab, return_info=True, **kw)
    if layout == "strips":
        assert info["strip_shift"] == 14 and info["n_subcopies"] > 100
    else:
        assert info["band"] == "allblocks" and FAR_COL_SHIFT == 13 and info["far_units"] > 100
    assert info["passes"] == wpasses
    np.testing.assert_array_equal(np.isnan(bias), np.isnan(want))
    np.testing.assert_allclose(bias, want, rtol=1e-9, equal_nan=True)
    assert abs(stats["var"] / wstats["var"] - 1) < 1e-9
