"""Cases of the general-merge goldens (shared by tests/golden/make_merge_golden.py, which runs the
unmodified reference's CoolerMerger, and the tests).  Three overlapping parts of the GM12878 2 Mb
table, each with an int32 ``count`` (part k scaled by k + 1), a float64 ``fcount``, an int64 ``big``
and -- part 1 only float32, the others int32 -- a mixed-dtype column ``mixed``."""
import numpy as np

CASES = {   # key -> (columns, agg)
    "mean": (["count"], {"count": "mean"}),
    "min": (["count"], {"count": "min"}),
    "max_last": (["count", "big"], {"count": "max", "big": "last"}),
    "first": (["count"], {"count": "first"}),
    "cols_sum": (["count", "fcount", "big"], None),
    "fmean": (["fcount", "count"], {"fcount": "mean"}),
    "mixed_sum": (["mixed"], None),
}


def parts(gm):
    """[(selection mask, {column: array})] of the three inputs."""
    rng = np.random.default_rng(11)
    n = len(gm["bin1_id"])
    out = []
    for k in range(3):
        sel = rng.random(n) < (0.5, 0.3, 0.8)[k]
        cnt = (gm["count"][sel] * (k + 1)).astype(np.int32)
        m = int(sel.sum())
        fcount = cnt * 0.37 + rng.random(m)
        big = cnt.astype(np.int64) * 3_000_000_000 + rng.integers(0, 1000, m)
        mixed = (cnt * 0.5).astype(np.float32) if k == 1 else cnt
        out.append((sel, {"count": cnt, "fcount": fcount, "big": big, "mixed": mixed}))
    return out
