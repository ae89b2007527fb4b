"""Shared description of the matrix-fetch golden cases (tests/golden/make_matrix_golden.py)."""
import os

import numpy as np

from tests import _golden

GOLD = dict(np.load(os.path.join(_golden.GOLD, "matrix_golden.npz"), allow_pickle=False))

# case -> (weights key | None, window, kwargs)
_GM = {
    "chr1_bal": ("gm/defaults", (0, 125, 0, 125), {}),
    "rect_raw": (None, (130, 150, 443, 534), {}),
    "rect_lower_bal": ("gm/defaults", (443, 534, 130, 150), {}),
    "diag_straddle_bal": ("gm/defaults", (100, 140, 90, 130), {}),
    "diag_straddle_raw": (None, (100, 140, 90, 130), {}),
    "overlap_tall_bal": ("gm/defaults", (300, 420, 350, 380), {}),
    "cis_weights_bal": ("gm/cis_only", (1369, 1401, 1401, 1426), {}),
    "divisive": ("gm/defaults", (0, 50, 20, 90), {"divisive_weights": True}),
    "whole_raw": (None, (0, 1561, 0, 1561), {}),
    "empty_window": ("gm/defaults", (5, 5, 7, 20), {}),
    "sparse_bal": ("gm/defaults", (247, 347, 0, 15), {}),
    "sparse_raw_sym": (None, (200, 260, 180, 240), {}),
    "pixels_bal": ("gm/defaults", (1401, 1426, 1401, 1426), {}),
    "pixels_raw_rect": (None, (10, 60, 30, 100), {}),
}
DENSE = ["chr1_bal", "rect_raw", "rect_lower_bal", "diag_straddle_bal", "diag_straddle_raw", "overlap_tall_bal",
         "cis_weights_bal", "divisive", "whole_raw", "empty_window", "asymm_whole_raw", "asymm_rect_raw"]
SPARSE = ["sparse_bal", "sparse_raw_sym"]
PIXELS = ["pixels_bal", "pixels_raw_rect"]


def toy_asymm():
    CO = _golden.coarsen_golden()
    a = CO["toy.asymm/2"]
    bc, bs, be = _golden.uniform_bins(np.array([32, 32]), 2)
    return {"chromnames": np.array(["chr1", "chr2"]), "chromsizes": np.array([32, 32]), "bins_chrom": bc,
            "bins_start": bs, "bins_end": be, "bin1_id": a[:, 0], "bin2_id": a[:, 1],
            "count": a[:, 2].astype(np.int32), "binsize": 2}


def inputs(case):
    """(table dict, weights | None, (i0, i1, j0, j1), kwargs incl. symmetric_upper)."""
    if case.startswith("asymm"):
        win = (0, 32, 0, 32) if case == "asymm_whole_raw" else (16, 32, 2, 15)
        return toy_asymm(), None, win, {"symmetric_upper": False}
    wkey, win, kw = _GM[case]
    BAL, _ = _golden.balance_golden()
    w = BAL[wkey] if wkey is not None else None
    return _golden.gm12878(), w, win, dict(kw, symmetric_upper=True)
