"""Cases of the general-aggregation coarsening goldens (shared by tests/golden/make_agg_golden.py,
which runs the unmodified reference, and tests/test_coarsen_gpu.py)."""
import numpy as np

CASES = {   # key -> (table, factor, columns, agg)
    "gm_mean_x2": ("gm", 2, ["count"], {"count": "mean"}),
    "gm_min_x5": ("gm", 5, ["count"], {"count": "min"}),
    "gm_max_x3": ("gm", 3, ["count"], {"count": "max"}),
    "gm_first_x2": ("gm", 2, ["count"], {"count": "first"}),
    "gm_last_x7": ("gm", 7, ["count"], {"count": "last"}),
    "gm_cols_x2": ("gm", 2, ["count", "fcount", "big"], {"fcount": "mean", "big": "max"}),
    "gm_fsum_x5": ("gm", 5, ["count", "fcount"], None),
    "gmf_sum_x2": ("gmf", 2, ["count"], None),
    "gmf_mean_x10": ("gmf", 10, ["count"], {"count": "mean"}),
}


def derived(gm):
    """(float64 column, wide int64 column) derived from the GM12878 counts."""
    rng = np.random.default_rng(7)
    n = len(gm["count"])
    fcount = gm["count"] * 0.37 + rng.random(n)
    big = gm["count"].astype(np.int64) * 3_000_000_000 + rng.integers(0, 1000, n)
    return fcount, big
