"""Golden vectors for merging coolers with general value columns / aggregators, made by the
UNMODIFIED reference (CoolerMerger.__iter__, src/cooler/_reduce.py:170-208: pandas concat +
groupby(sort=True).aggregate(agg)) through oracle/refshim.py.

    python tests/golden/make_merge_golden.py     ->  tests/golden/merge_agg_golden.npz

Inputs: tests/_merge_cases.parts() of the GM12878 2 Mb table (tests/golden/gm12878_2mb.npz).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

from tests._merge_cases import CASES, parts  # noqa: E402


def main():
    assert refshim.available()
    gm = dict(np.load(os.path.join(HERE, "gm12878_2mb.npz")))
    names = []
    for k, (sel, cols) in enumerate(parts(gm)):
        nm = f"gm.mpart{k}.cool"
        extra = {c: v for c, v in cols.items() if c != "count"}
        refshim.register(nm, [str(x) for x in gm["chromnames"]], gm["chromsizes"], gm["bins_chrom"],
                         gm["bins_start"], gm["bins_end"], gm["bin1_id"][sel], gm["bin2_id"][sel], cols["count"],
                         2_000_000, extra_columns=extra)
        names.append(nm)
    out = {}
    for key, (columns, agg) in CASES.items():
        ref = None
        for mergebuf in (20_000_000, 5000):
            res = refshim.ref_merge(names, mergebuf=mergebuf, columns=columns, agg=agg)
            if ref is not None:                               # epochs only chunk the output
                for k in res:
                    assert np.array_equal(res[k], ref[k]), (key, k)
            ref = res
        for k, v in ref.items():
            out[f"{key}/{k}"] = v
        print(key, {k: (v.dtype, len(v)) for k, v in ref.items()})
    np.savez_compressed(os.path.join(HERE, "merge_agg_golden.npz"), **out)


if __name__ == "__main__":
    main()
