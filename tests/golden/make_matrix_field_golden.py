"""Golden vectors for the matrix fetch on value columns other than an int32 count (`field=`), made
by the UNMODIFIED reference (``cooler.Cooler.matrix(field=...)``, src/cooler/api.py:326-409, 665-800)
through oracle/refshim.py.

    python tests/golden/make_matrix_field_golden.py     ->  tests/golden/matrix_field_golden.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402
from tests import _golden  # noqa: E402
from tests._agg_cases import derived  # noqa: E402
from tests._matrix_field_cases import CASES  # noqa: E402


def main():
    cooler = refshim.import_reference()
    t = _golden.gm12878()
    fcount, big = derived(t)
    BAL, _ = _golden.balance_golden()
    out = {}
    for key, (tab, field, wkey, win, kind, extra) in CASES.items():
        count = t["count"] if tab == "gm" else (t["count"] * 0.5).astype(np.float32)
        g = refshim.register("gm_field.cool", [str(x) for x in t["chromnames"]], t["chromsizes"], t["bins_chrom"],
                             t["bins_start"], t["bins_end"], t["bin1_id"], t["bin2_id"], count, 2_000_000,
                             extra_columns={"fcount": fcount, "big": big} if tab == "gm" else None)
        if wkey is not None:
            g["bins"]["weight"] = BAL[wkey].copy()
        kw = dict(field=field, balance=wkey is not None, **extra)
        if kind == "sparse":
            kw["sparse"] = True
        elif kind == "pixels":
            kw.update(as_pixels=True, join=False)
        res = cooler.Cooler("gm_field.cool").matrix(**kw)[win[0]:win[1], win[2]:win[3]]
        if kind == "sparse":
            o = np.lexsort((res.col, res.row))
            out[key + "/row"], out[key + "/col"], out[key + "/data"] = res.row[o], res.col[o], res.data[o]
        elif kind == "pixels":
            for c in res.columns:
                out[key + "/" + c] = res[c].to_numpy()
        else:
            out[key] = np.asarray(res)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "matrix_field_golden.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)


if __name__ == "__main__":
    main()
