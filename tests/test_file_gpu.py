"""File-backed entry points on the GPU: the pinned-ring streamed reader (table.read_cooler_device) and
balance_cooler on written .cool files, integer and floating-point count columns."""
import warnings

import numpy as np
import pytest

from tests.test_h5lite import _written_cool

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows", [None, (301, 1500)])
def test_streamed_reader_equals_host_reader(tmp_path, rows):
    """Pieces smaller than the table and a 2-slot ring: every slot is reused many times."""
    from cooler_b200 import store, table
    path, b1, b2, cnt = _written_cool(tmp_path, n_px=3_000_000)
    assert len(b2) > (1 << 20)
    clr = store.open_cooler(path)
    off_h, b2_h, cnt_h, co_h = table.read_cooler_arrays(clr, rows=rows)
    off_d, b2_d, cnt_d, co_d = table.read_cooler_device(store.open_cooler(path), rows=rows, piece_rows=1 << 17, slots=2,
                                                          min_pixels=0)
    assert b2_d.is_cuda and b2_d.dtype.__str__() == "torch.int32" and cnt_d.is_cuda
    np.testing.assert_array_equal(off_d, off_h), np.testing.assert_array_equal(co_d, co_h)
    np.testing.assert_array_equal(b2_d.cpu().numpy(), np.asarray(b2_h))
    np.testing.assert_array_equal(cnt_d.cpu().numpy(), np.asarray(cnt_h))
    # default geometry (one piece) gives the same columns
    _, b2_1, cnt_1, _ = table.read_cooler_device(store.open_cooler(path), rows=rows, min_pixels=0)
    assert bool((b2_1 == b2_d).all()) and bool((cnt_1 == cnt_d).all())


def test_balance_cooler_file_streams_and_matches_arrays(tmp_path):
    import cooler_b200
    from cooler_b200 import store
    path, b1, b2, cnt = _written_cool(tmp_path, n_px=3_000_000)
    n = 1801
    bin1_offset = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    chrom_offset = np.array([0, 900, 1500, 1501, 1801])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want, wstats = cooler_b200.balance_arrays(bin1_offset, b2, cnt, chrom_offset, ignore_diags=1)
        bias, stats = cooler_b200.balance_cooler(store.open_cooler(path), ignore_diags=1, store=True)
    # the streamed host build cuts the table into row chunks, the file path builds whole copies: the sums
    # are associated differently (1e-15), the masks and the converged state are the same
    np.testing.assert_array_equal(np.isnan(bias), np.isnan(want))
    np.testing.assert_allclose(bias, want, rtol=1e-12, atol=0, equal_nan=True)
    assert stats["converged"] == wstats["converged"]
    with store.open_cooler(path).open("r") as g:
        np.testing.assert_array_equal(g["bins/weight"][:], bias)


def test_balance_cooler_float_count_file_matches_oracle(tmp_path):
    """A .cool whose count column is float64 (e.g. written by coarsen --field count:dtype=float,agg=mean)."""
    import cooler_b200
    from cooler_b200 import store
    from oracle import ice_numpy
    path, b1, b2, val = _written_cool(tmp_path, n_px=3_000_000, float_counts=True)
    n = 1801
    bin1_offset = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    chrom_offset = np.array([0, 900, 1500, 1501, 1801])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want, wstats, wpasses, _ = ice_numpy.balance_arrays(dict(bin1_id=b1, bin2_id=b2, count=val), chrom_offset,
                                                            bin1_offset, return_passes=True, cis_only=True)
        bias, stats = cooler_b200.balance_cooler(store.open_cooler(path), cis_only=True)
    np.testing.assert_array_equal(np.isnan(bias), np.isnan(want))
    np.testing.assert_allclose(bias, want, rtol=1e-9, atol=0, equal_nan=True)
    # the variance of marginals that agree to 1e-16 is a difference of nearly equal numbers: at var ~ 1e-7 of
    # scale^2 its own relative error is ~1e-9 (the bar is on the bias; BASELINE north_star)
    np.testing.assert_allclose(stats["var"], wstats["var"], rtol=1e-6, equal_nan=True)
    np.testing.assert_allclose(stats["scale"], wstats["scale"], rtol=1e-9, equal_nan=True)
