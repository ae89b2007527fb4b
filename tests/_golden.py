"""Loaders for the committed golden vectors (tests/golden/, made by make_golden.py
from the unmodified reference)."""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _npz(name):
    return dict(np.load(os.path.join(GOLD, name), allow_pickle=False))


def gm12878():
    return _npz("gm12878_2mb.npz")


def random_table(name):
    z = _npz("random_tables.npz")
    return {k.split("/", 1)[1]: v for k, v in z.items() if k.startswith(name + "/")}


def balance_golden():
    with open(os.path.join(GOLD, "balance_golden.json")) as f:
        meta = json.load(f)
    return _npz("balance_golden.npz"), meta


def coarsen_golden():
    return _npz("coarsen_golden.npz")


def table_of(key):
    """Input table for golden key prefix ('gm' or 'rnd1'...)."""
    return gm12878() if key == "gm" else random_table(key)


def offsets(t):
    n_chroms = len(t["chromsizes"])
    chrom_offset = np.searchsorted(t["bins_chrom"], np.arange(n_chroms + 1)).astype(np.int64)
    n_bins = len(t["bins_chrom"])
    bin1_offset = np.searchsorted(t["bin1_id"], np.arange(n_bins + 1)).astype(np.int64)
    return chrom_offset, bin1_offset


def nan_to_none(x):
    return None if (isinstance(x, float) and np.isnan(x)) else x


def uniform_bins(chromsizes, binsize):
    c, s, e = [], [], []
    for i, L in enumerate(chromsizes):
        st = np.arange(0, L, binsize, dtype=np.int64)
        c.append(np.full(len(st), i, np.int32)), s.append(st), e.append(np.minimum(st + binsize, L))
    return np.concatenate(c), np.concatenate(s), np.concatenate(e)
