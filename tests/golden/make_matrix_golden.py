"""Golden vectors for the balanced matrix fetch, made by the UNMODIFIED
reference (``cooler.Cooler.matrix(...).fetch`` / ``[...]``, src/cooler/api.py:326-409, 665-800)
run through oracle/refshim.py in the build container.

    python tests/golden/make_matrix_golden.py     ->  tests/golden/matrix_golden.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402
from tests import _golden  # noqa: E402

CASES = [   # key, matrix kwargs, weights key, query ("fetch", r1, r2) or ("slice", i0, i1, j0, j1)
    ("chr1_bal", dict(balance=True), "gm/defaults", ("fetch", "chr1", None)),
    ("rect_raw", dict(balance=False), None, ("fetch", "chr2:10,000,000-50,000,000", "chr5")),
    ("rect_lower_bal", dict(balance=True), "gm/defaults", ("fetch", "chr5", "chr2:10,000,000-50,000,000")),
    ("diag_straddle_bal", dict(balance=True), "gm/defaults", ("slice", 100, 140, 90, 130)),
    ("diag_straddle_raw", dict(balance=False), None, ("slice", 100, 140, 90, 130)),
    ("overlap_tall_bal", dict(balance=True), "gm/defaults", ("slice", 300, 420, 350, 380)),
    ("cis_weights_bal", dict(balance=True), "gm/cis_only", ("fetch", "chr20", "chr21")),
    ("divisive", dict(balance=True, divisive_weights=True), "gm/defaults", ("slice", 0, 50, 20, 90)),
    ("whole_raw", dict(balance=False), None, ("slice", 0, 1561, 0, 1561)),
    ("sparse_bal", dict(balance=True, sparse=True), "gm/defaults", ("fetch", "chr3", "chr1:0-30,000,000")),
    ("sparse_raw_sym", dict(balance=False, sparse=True), None, ("slice", 200, 260, 180, 240)),
    ("pixels_bal", dict(balance=True, as_pixels=True, join=False), "gm/defaults", ("fetch", "chr21", None)),
    ("pixels_raw_rect", dict(balance=False, as_pixels=True, join=False), None, ("slice", 10, 60, 30, 100)),
    ("empty_window", dict(balance=True), "gm/defaults", ("slice", 5, 5, 7, 20)),
]


def main():
    cooler = refshim.import_reference()
    t = _golden.gm12878()
    BAL, _ = _golden.balance_golden()
    out = {}
    for key, kw, wkey, q in CASES:
        g = refshim.register("gm_matrix.cool", [str(x) for x in t["chromnames"]], t["chromsizes"],
                             t["bins_chrom"], t["bins_start"], t["bins_end"], t["bin1_id"], t["bin2_id"],
                             t["count"], 2_000_000)
        if wkey is not None:
            g["bins"]["weight"] = BAL[wkey].copy()
        clr = cooler.Cooler("gm_matrix.cool")
        sel = clr.matrix(**kw)
        res = sel.fetch(q[1], q[2]) if q[0] == "fetch" else sel[q[1]:q[2], q[3]:q[4]]
        if kw.get("sparse"):
            o = np.lexsort((res.col, res.row))
            out[key + "/row"], out[key + "/col"], out[key + "/data"] = res.row[o], res.col[o], res.data[o]
            out[key + "/shape"] = np.array(res.shape)
        elif kw.get("as_pixels"):
            for c in res.columns:
                out[key + "/" + c] = res[c].to_numpy()
        else:
            out[key] = np.asarray(res)
    # square (non-symmetric) storage: no lower-triangle fill
    CO = _golden.coarsen_golden()
    a = CO["toy.asymm/2"]
    bc, bs, be = _golden.uniform_bins(np.array([32, 32]), 2)
    refshim.register("toy_asymm_matrix.cool", ["chr1", "chr2"], np.array([32, 32]), bc, bs, be,
                     a[:, 0], a[:, 1], a[:, 2].astype(np.int32), 2, symmetric_upper=False)
    clr = cooler.Cooler("toy_asymm_matrix.cool")
    out["asymm_whole_raw"] = clr.matrix(balance=False)[:, :]
    out["asymm_rect_raw"] = clr.matrix(balance=False).fetch("chr2", "chr1:4-30")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "matrix_golden.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)


if __name__ == "__main__":
    main()
