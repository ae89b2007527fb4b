"""Quick check of the far-field kernel variants on a small table: all-blocks layout (as selected by
CB200_FAR_TMA / the default) against the whole-copy sweep.   CB200_FAR_TMA=1 python tools/check_far_tma.py"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cooler_b200  # noqa: E402

rng = np.random.default_rng(0)
n_per = [7000, 5000, 3, 3001]
chrom_offset = np.r_[0, np.cumsum(n_per)].astype(np.int64)
n = int(chrom_offset[-1])
m = 40 * n
i = rng.integers(0, n, m)
j = np.minimum(i + np.where(rng.random(m) < 0.5, rng.integers(0, 50, m), rng.integers(0, n, m)), n - 1)
key = np.unique(i.astype(np.int64) * n + j)
b1, b2 = key // n, key % n
cnt = (1 + rng.poisson(3, len(b1))).astype(np.int32)
off = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
tab = cooler_b200.PixelTable.from_arrays(off, b2, cnt, chrom_offset)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    a, _, ia = cooler_b200.balance_table(tab, return_info=True, strip_bins=0)
    for cs in (6, 9, 13):
        b, _, ib = cooler_b200.balance_table(tab, return_info=True, band="allblocks", far_col_shift=cs)
        ok = ia["passes"] == ib["passes"] and np.array_equal(np.isnan(a), np.isnan(b)) and \
            np.allclose(a, b, rtol=1e-12, atol=0, equal_nan=True)
        print("far_col_shift", cs, "tma", os.environ.get("CB200_FAR_TMA", "0"), "units", ib["far_units"],
              "OK" if ok else "FAIL", flush=True)
