"""Run in a SUBPROCESS by tests/test_oracle.py: the numpy oracle (oracle/ice_numpy.py, oracle/coarsen_numpy.py)
against the LIVE unmodified reference (balance_cooler, CoolerCoarsener through oracle/refshim.py) on seeded
random tables -- symmetric-upper and square storage, integer and floating-point counts, all three modes, random
filter settings, warm starts and blacklists; coarsening by random factors.  Bias vectors must be bit-identical
(NaN pattern included), pass counts, warnings and statistics equal."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import coarsen_numpy, ice_numpy, refshim  # noqa: E402


def random_table(rng):
    n_chroms = int(rng.integers(1, 6))
    nb = [int(rng.integers(1, 40)) for _ in range(n_chroms)]
    binsize = 1000
    sizes = np.array([k * binsize - int(rng.integers(0, binsize // 2)) for k in nb], dtype=np.int64)
    bc = np.repeat(np.arange(n_chroms, dtype=np.int32), nb)
    bs = np.concatenate([np.arange(k) * binsize for k in nb]).astype(np.int64)
    be = np.minimum(bs + binsize, sizes[bc])
    n = len(bc)
    square = rng.random() < 0.35
    i, j = (np.indices((n, n)).reshape(2, -1) if square else np.triu_indices(n))
    d = np.abs(i - j)
    b = rng.lognormal(0, 0.4, n)
    b[rng.random(n) < 0.08] *= 0.02
    keep = rng.random(len(i)) < np.minimum(1.0, rng.uniform(0.1, 0.9) * (1 + 10.0 / (1 + d)))
    cnt = rng.poisson(30 * b[i] * b[j] / (1.0 + d) ** 0.7) * keep
    if rng.random() < 0.3:
        sel = keep                                         # stored zeros stay in the table
    else:
        sel = cnt > 0
    count = cnt[sel].astype(np.int32)
    if rng.random() < 0.3:
        count = (count * rng.lognormal(0, 0.2, len(count))).astype(np.float64 if rng.random() < 0.5 else np.float32)
    return dict(names=[f"c{k}" for k in range(n_chroms)], sizes=sizes, bc=bc, bs=bs, be=be,
                b1=i[sel].astype(np.int64), b2=j[sel].astype(np.int64), count=count, square=square, binsize=binsize)


def random_kwargs(rng, n):
    kw = {}
    mode = rng.integers(0, 3)
    if mode == 1:
        kw["cis_only"] = True
    elif mode == 2:
        kw["trans_only"] = True
    if rng.random() < 0.7:
        kw["ignore_diags"] = [0, 1, 2, 3, False][int(rng.integers(0, 5))]
    if rng.random() < 0.6:
        kw["mad_max"] = int(rng.integers(0, 6))
    if rng.random() < 0.6:
        kw["min_nnz"] = int(rng.integers(0, 8))
    if rng.random() < 0.3:
        kw["min_count"] = int(rng.integers(0, 40))
    if rng.random() < 0.5:
        kw["tol"] = float(10 ** rng.uniform(-7, -2))
    if rng.random() < 0.5:
        kw["max_iters"] = int(rng.integers(1, 60))
    if rng.random() < 0.3:
        kw["rescale_marginals"] = False
    if rng.random() < 0.25:
        kw["blacklist"] = sorted(set(int(x) for x in rng.integers(0, n, int(rng.integers(1, 5)))))
    if rng.random() < 0.2:
        x0 = rng.lognormal(0, 0.3, n)
        x0[rng.random(n) < 0.1] = np.nan
        kw["x0"] = x0
    return kw


def same(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def main(n_cases, seed):
    assert refshim.available()
    rng = np.random.default_rng(seed)
    n_bal = n_coarse = 0
    for case in range(n_cases):
        t = random_table(rng)
        n = len(t["bc"])
        name = f"fuzz{case}.cool"
        refshim.register(name, t["names"], t["sizes"], t["bc"], t["bs"], t["be"], t["b1"], t["b2"], t["count"],
                         t["binsize"], symmetric_upper=not t["square"])
        chrom_offset = np.searchsorted(t["bc"], np.arange(len(t["names"]) + 1)).astype(np.int64)
        bin1_offset = np.searchsorted(t["b1"], np.arange(n + 1)).astype(np.int64)
        px = {"bin1_id": t["b1"], "bin2_id": t["b2"], "count": t["count"]}
        for _ in range(2):
            kw = random_kwargs(rng, n)
            kw_ref = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
            kw_or = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                np.seterr(all="ignore")
                want, wstats, wpasses, wwarn = refshim.ref_balance(name, **kw_ref)
                got, gstats, gpasses, gwarn = ice_numpy.balance_arrays(px, chrom_offset, bin1_offset,
                                                                       return_passes=True, **kw_or)
            tag = (case, {k: v for k, v in kw.items() if k != "x0"}, t["square"], str(t["count"].dtype))
            if not kw.get("cis_only") and wpasses == []:   # refshim counts "variance is" log lines: none were
                wpasses = [0]                              # printed when the first marginal is empty (:98-102)
            assert gpasses == wpasses, ("passes", tag, gpasses, wpasses)
            assert gwarn == wwarn, ("warned", tag)
            assert same(got, want), ("bias", tag)
            for k in ("scale", "var"):
                assert same(gstats[k], wstats[k]), (k, tag, gstats[k], wstats[k])
            assert np.array_equal(np.atleast_1d(gstats["converged"]), np.atleast_1d(wstats["converged"])), tag
            n_bal += 1
        if t["count"].dtype.kind == "i" and not t["square"] and len(t["b1"]):
            factor = int(rng.integers(2, 9))
            ref, nb = refshim.ref_coarsen(name, factor, chunksize=int(rng.integers(5, 500)))
            b1, b2, c, (nc, ns, ne) = coarsen_numpy.coarsen_pixels(px, t["bc"], t["bs"], t["be"], t["sizes"], factor)
            assert np.array_equal(b1, ref["bin1_id"]) and np.array_equal(b2, ref["bin2_id"]), ("coarsen ids", case)
            assert np.array_equal(c, ref["count"]), ("coarsen counts", case, factor)
            assert np.array_equal(nc, nb["chrom"]) and np.array_equal(ns, nb["start"]) and np.array_equal(ne, nb["end"])
            n_coarse += 1
    print(f"ok: {n_bal} balance runs and {n_coarse} coarsenings identical")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 40, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
