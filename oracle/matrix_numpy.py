"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's 2-D matrix range query
with lower-triangle fill and balancing-weight application.

Follows open2c/cooler src/cooler/api.py:742-798 (weight application) and
core/_rangequery.py:153-368 (DirectRangeQuery2D / FillLowerRangeQuery2D semantics: the window of
the FULL matrix; for symmetric-upper storage the cells below the diagonal come from the mirror
images of stored pixels, the diagonal is not duplicated).  Pinned against golden vectors made by
the unmodified reference (tests/golden/make_matrix_golden.py -> matrix_golden.npz).
Only tests/ may import this module.
"""
import numpy as np


def fetch_pixels(bin1, bin2, count, i0, i1, j0, j1, fill_lower):
    """Window-relative (row, col, value) triples, direct part then mirrored part."""
    m = (bin1 >= i0) & (bin1 < i1) & (bin2 >= j0) & (bin2 < j1)
    row, col, val = bin1[m] - i0, bin2[m] - j0, count[m]
    if fill_lower:
        m2 = (bin1 >= j0) & (bin1 < j1) & (bin2 >= i0) & (bin2 < i1) & (bin2 > bin1)
        row = np.r_[row, bin2[m2] - i0]
        col = np.r_[col, bin1[m2] - j0]
        val = np.r_[val, count[m2]]
    return row, col, val


def fetch_dense(bin1, bin2, count, i0, i1, j0, j1, *, symmetric_upper=True, weights=None,
                divisive_weights=False, fill_lower=True):
    row, col, val = fetch_pixels(bin1, bin2, count, i0, i1, j0, j1, fill_lower and symmetric_upper)
    arr = np.zeros((i1 - i0, j1 - j0), dtype=count.dtype)
    arr[row, col] = val
    if weights is not None:                                           # api.py:788-796
        bias1 = weights[i0:i1]
        bias2 = bias1 if (i0, i1) == (j0, j1) else weights[j0:j1]
        if divisive_weights:
            bias1, bias2 = 1 / bias1, 1 / bias2
        arr = arr * np.outer(bias1, bias2)
    return arr


def fetch_sparse(bin1, bin2, count, i0, i1, j0, j1, *, symmetric_upper=True, weights=None,
                 divisive_weights=False, fill_lower=True):
    """(row, col, data) sorted by (row, col)."""
    row, col, val = fetch_pixels(bin1, bin2, count, i0, i1, j0, j1, fill_lower and symmetric_upper)
    data = val
    if weights is not None:                                           # api.py:773-780
        bias1, bias2 = weights[i0:i1], weights[j0:j1]
        if divisive_weights:
            bias1, bias2 = 1 / bias1, 1 / bias2
        data = bias1[row] * bias2[col] * val
    o = np.lexsort((col, row))
    return row[o], col[o], data[o]
