"""CPU: the numpy restatement of the matrix fetch against golden vectors from the unmodified
reference (tests/golden/matrix_golden.npz)."""
import numpy as np
import pytest

from tests import _golden, _matrix_cases as MC


@pytest.mark.parametrize("case", MC.DENSE)
def test_oracle_dense_matches_reference(case):
    from oracle import matrix_numpy
    t, w, (i0, i1, j0, j1), kw = MC.inputs(case)
    got = matrix_numpy.fetch_dense(t["bin1_id"], t["bin2_id"], t["count"], i0, i1, j0, j1,
                                   symmetric_upper=kw["symmetric_upper"], weights=w,
                                   divisive_weights=kw.get("divisive_weights", False))
    want = MC.GOLD[case]
    assert got.shape == want.shape and got.dtype == want.dtype
    np.testing.assert_array_equal(got, want)                  # bit-exact, NaNs included


@pytest.mark.parametrize("case", MC.SPARSE)
def test_oracle_sparse_matches_reference(case):
    from oracle import matrix_numpy
    t, w, (i0, i1, j0, j1), kw = MC.inputs(case)
    r, c, d = matrix_numpy.fetch_sparse(t["bin1_id"], t["bin2_id"], t["count"], i0, i1, j0, j1, weights=w)
    np.testing.assert_array_equal(r, MC.GOLD[case + "/row"])
    np.testing.assert_array_equal(c, MC.GOLD[case + "/col"])
    np.testing.assert_array_equal(d, MC.GOLD[case + "/data"])


def test_region_parsing_matches_reference_extents():
    from cooler_b200 import MemCooler
    from cooler_b200.api import parse_region, region_to_extent
    t = _golden.gm12878()
    clr = MemCooler.from_arrays(t["chromnames"], t["chromsizes"], t["bins_chrom"], t["bins_start"],
                                t["bins_end"], t["bin1_id"], t["bin2_id"], t["count"], binsize=2_000_000)
    assert region_to_extent(clr, "chr2:10,000,000-50,000,000") == (130, 150)      # reference probe values
    assert region_to_extent(clr, "chr5") == (443, 534)
    assert region_to_extent(clr, ("chr1", 0, 30_000_000)) == (0, 15)
    assert region_to_extent(clr, "chr1:1.5M-3M") == (0, 2)
    assert parse_region("chr3:5kb-", {"chr3": 10_000}) == ("chr3", 5000, 10_000)
    for bad in ("chrNope", "chr1:10-5", "chr1:0-999,999,999,999", ":5-10"):
        with pytest.raises(ValueError):
            region_to_extent(clr, bad)
