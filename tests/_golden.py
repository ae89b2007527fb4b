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
    z = _npz("square_tables.npz" if name.startswith("sq") else "random_tables.npz")
    return {k.split("/", 1)[1]: v for k, v in z.items() if k.startswith(name + "/")}


def balance_golden():
    """Symmetric-upper cases (make_golden.py) + square-storage cases 'sq*/...' (make_square_golden.py)."""
    bal, meta = {}, {}
    for stem in ("balance_golden", "balance_square_golden"):
        with open(os.path.join(GOLD, stem + ".json")) as f:
            meta.update(json.load(f))
        bal.update(_npz(stem + ".npz"))
    return bal, meta


def balance_float_golden():
    with open(os.path.join(GOLD, "balance_float_golden.json")) as f:
        meta = json.load(f)
    return _npz("balance_float_golden.npz"), meta


def float_counts(count, kind):
    """Seeded floating-point value column derived from an integer count column: count x lognormal noise,
    ~2 % exact zeros (stored zeros must not count as nonzero, _balance.py:29-31); ``kind`` f64 | f32."""
    rng = np.random.default_rng(20240607)
    v = count.astype(np.float64) * rng.lognormal(0.0, 0.3, len(count))
    v[rng.random(len(count)) < 0.02] = 0.0
    return v.astype(np.float32) if kind == "f32" else v


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
