"""Golden vectors for coarsening with general value columns / aggregators, made by the UNMODIFIED
reference (CoolerCoarsener._aggregate, src/cooler/_reduce.py:582-620: pandas
groupby(sort=True).aggregate(agg)) through oracle/refshim.py.

    python tests/golden/make_agg_golden.py     ->  tests/golden/coarsen_agg_golden.npz

Inputs: the GM12878 2 Mb table (tests/golden/gm12878_2mb.npz) with two derived columns -- a
float64 column ``fcount`` and an int64 column ``big`` -- and the same table with a float32 count.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

from tests._agg_cases import CASES, derived  # noqa: E402


def main():
    assert refshim.available()
    gm = dict(np.load(os.path.join(HERE, "gm12878_2mb.npz")))
    fcount, big = derived(gm)
    args = ([str(x) for x in gm["chromnames"]], gm["chromsizes"], gm["bins_chrom"], gm["bins_start"],
            gm["bins_end"], gm["bin1_id"], gm["bin2_id"])
    refshim.register("gm.cool", *args, gm["count"], 2_000_000, extra_columns={"fcount": fcount, "big": big})
    refshim.register("gmf.cool", *args, (gm["count"] * 0.5).astype(np.float32), 2_000_000)
    out = {}
    for key, (tab, factor, columns, agg) in CASES.items():
        res, _ = refshim.ref_coarsen(tab + ".cool", factor, chunksize=5000, columns=columns, agg=agg)
        for k, v in res.items():
            out[f"{key}/{k}"] = v
        print(key, {k: (v.dtype, len(v)) for k, v in res.items()})
    np.savez_compressed(os.path.join(HERE, "coarsen_agg_golden.npz"), **out)


if __name__ == "__main__":
    main()
