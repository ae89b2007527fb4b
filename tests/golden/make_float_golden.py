"""Golden vectors for FLOATING-POINT count columns (the reference copies the column in its own dtype,
_balance.py:25-26): the UNMODIFIED reference balance_cooler run through oracle/refshim.py on the
committed tables with float64 / float32 values.  Build container only:

    python tests/golden/make_float_golden.py

The float columns are re-derived from the committed integer tables by ``tests/_golden.float_counts``
(seeded), so only the reference's answers are stored (balance_float_golden.npz / .json).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402
from tests import _golden  # noqa: E402
from tests.golden.make_golden import jsonable, table  # noqa: E402

CASES = {
    "gm/f64/defaults": {},
    "gm/f64/cis_only": {"cis_only": True},
    "gm/f64/trans_only": {"trans_only": True},
    "gm/f64/min_count": {"min_count": 150.5, "mad_max": 3},
    "gm/f64/diag0_nnz0": {"ignore_diags": 0, "min_nnz": 0},
    "gm/f64/iters3": {"max_iters": 3, "ignore_diags": 1},
    "gm/f32/defaults": {},
    "gm/f32/cis_only": {"cis_only": True},
    "rnd3/f64/defaults": {},
    "rnd3/f64/trans_only": {"trans_only": True},
    "rnd3/f64/cis_mad3": {"cis_only": True, "mad_max": 3, "min_nnz": 3, "ignore_diags": 1},
}


def main():
    assert refshim.available(), "reference not present"
    bal, meta = {}, {}
    for key, kw in CASES.items():
        src, kind, _ = key.split("/")
        t = dict(_golden.table_of(src))
        t["count"] = _golden.float_counts(t["count"], kind)
        name = f"{src}.{kind}.cool"
        refshim.register(name, *table(t))
        bias, stats, passes, warned = refshim.ref_balance(name, **kw)
        bal[key] = bias
        meta[key] = dict(kwargs=kw, stats=jsonable(stats), passes=passes, warned=warned)
        print(key, passes if len(passes) < 4 else (sum(passes), "total"), int(np.isnan(bias).sum()))
    np.savez_compressed(os.path.join(HERE, "balance_float_golden.npz"), **bal)
    with open(os.path.join(HERE, "balance_float_golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
