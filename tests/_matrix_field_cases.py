"""Cases of the matrix-fetch goldens on value columns other than an int32 count (`field=` of
Cooler.matrix, reference src/cooler/api.py:326-329, 724-726), shared by
tests/golden/make_matrix_field_golden.py (unmodified reference) and the tests.
Tables: "gm" = GM12878 2 Mb with the derived columns of tests/_agg_cases.derived (float64 ``fcount``,
wide int64 ``big``); "gmf" = the same table with a float32 ``count`` column (count * 0.5)."""

# key -> (table, field, weights key | None, (i0, i1, j0, j1), kind, extra matrix kwargs)
CASES = {
    "f_dense_raw": ("gm", "fcount", None, (100, 140, 90, 130), "dense", {}),
    "f_dense_bal": ("gm", "fcount", "gm/defaults", (0, 125, 0, 125), "dense", {}),
    "big_dense_raw": ("gm", "big", None, (300, 420, 350, 380), "dense", {}),
    "big_sparse_raw": ("gm", "big", None, (200, 260, 180, 240), "sparse", {}),
    "f_sparse_bal": ("gm", "fcount", "gm/defaults", (247, 347, 0, 15), "sparse", {}),
    "f_pixels_bal": ("gm", "fcount", "gm/defaults", (1401, 1426, 1401, 1426), "pixels", {}),
    "f32_dense_raw": ("gmf", "count", None, (0, 125, 0, 125), "dense", {}),
    "f32_dense_bal_div": ("gmf", "count", "gm/defaults", (0, 50, 20, 90), "dense", {"divisive_weights": True}),
    "f32_pixels_raw": ("gmf", "count", None, (10, 60, 30, 100), "pixels", {}),
}


def inputs(key):
    """(table dict incl. the value column as 'values', field, weights | None, window, kind, extra kwargs)."""
    import numpy as np

    from tests import _golden
    from tests._agg_cases import derived
    tab, field, wkey, win, kind, extra = CASES[key]
    t = dict(_golden.gm12878())
    fcount, big = derived(t)
    if tab == "gmf":
        t["count"] = (t["count"] * 0.5).astype(np.float32)
    t["extra"] = {"fcount": fcount, "big": big} if tab == "gm" else {}
    t["values"] = t["count"] if field == "count" else t["extra"][field]
    w = _golden.balance_golden()[0][wkey] if wkey is not None else None
    return t, field, w, win, kind, extra


def gold():
    from tests import _golden
    return _golden._npz("matrix_field_golden.npz")
