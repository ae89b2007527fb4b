"""Generate the committed golden vectors under tests/golden/ by running the
UNMODIFIED reference (open2c/cooler at /root/reference) through
oracle/refshim.py.  Run in the build container only:

    python tests/golden/make_golden.py

Inputs come from the reference's own test fixtures
(/root/reference/tests/data/*.coo.txt, *.bg2, *.bed.gz, *.coo) -- the text
twins of the .cool files its tests use -- plus small seeded random tables.
The outputs (inputs + reference answers) are stored as .npz/.json so the tests
can run where /root/reference does not exist (the GPU box).
"""
from __future__ import annotations

import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402

DATA = "/root/reference/tests/data"


def load_gm12878():
    chroms, starts, ends = [], [], []
    with gzip.open(os.path.join(DATA, "hg19.bins.2000kb.bed.gz"), "rt") as f:
        for line in f:
            c, s, e = line.split()[:3]
            chroms.append(c), starts.append(int(s)), ends.append(int(e))
    names = list(dict.fromkeys(chroms))
    code = {n: i for i, n in enumerate(names)}
    bins_chrom = np.array([code[c] for c in chroms], dtype=np.int32)
    starts, ends = np.array(starts, np.int64), np.array(ends, np.int64)
    chromsizes = np.array([ends[bins_chrom == i].max() for i in range(len(names))], np.int64)
    px = np.loadtxt(os.path.join(DATA, "hg19.GM12878-MboI.matrix.2000kb.coo.txt"), dtype=np.int64)
    return dict(chromnames=np.array(names), chromsizes=chromsizes, bins_chrom=bins_chrom,
                bins_start=starts, bins_end=ends, bin1_id=px[:, 0], bin2_id=px[:, 1],
                count=px[:, 2].astype(np.int32), binsize=2_000_000)


def uniform_bins(chromsizes, binsize):
    c, s, e = [], [], []
    for i, L in enumerate(chromsizes):
        st = np.arange(0, L, binsize, dtype=np.int64)
        c.append(np.full(len(st), i, np.int32)), s.append(st), e.append(np.minimum(st + binsize, L))
    return np.concatenate(c), np.concatenate(s), np.concatenate(e)


def load_toy_bg2(name, binsize):
    """toy.*.bg2: chrom1 start1 end1 chrom2 start2 end2 count, 1-based starts
    (tests/test_cli_ingest.py:43-57); toy.chrom.sizes: chr1 32, chr2 32."""
    rows = []
    code = {"chr1": 0, "chr2": 1}
    per = 32 // binsize
    with open(os.path.join(DATA, name)) as f:
        for line in f:
            c1, s1, _, c2, s2, _, v = line.split()
            rows.append((code[c1] * per + (int(s1) - 1) // binsize,
                         code[c2] * per + (int(s2) - 1) // binsize, int(v)))
    a = np.array(rows, dtype=np.int64)
    order = np.lexsort((a[:, 1], a[:, 0]))
    return a[order]


def table(d):
    return (d["chromnames"].tolist(), d["chromsizes"], d["bins_chrom"], d["bins_start"],
            d["bins_end"], d["bin1_id"], d["bin2_id"], d["count"], d["binsize"])


def random_table(seed, nb_per_chrom, binsize, density, maxcount, symmetric=True, bad_frac=0.05):
    rng = np.random.default_rng(seed)
    chromsizes = np.array([n * binsize - rng.integers(0, binsize // 2) if n else 0
                           for n in nb_per_chrom], np.int64)
    chromsizes = np.maximum(chromsizes, 1)
    bc, bs, be = uniform_bins(chromsizes, binsize)
    n = len(bc)
    b = rng.lognormal(0, 0.4, n)
    b[rng.random(n) < bad_frac] *= 0.02
    i, j = np.triu_indices(n) if symmetric else np.indices((n, n)).reshape(2, -1)
    d = np.abs(i - j)
    lam = maxcount * b[i] * b[j] / (1.0 + d) ** 0.8
    keep = rng.random(len(i)) < np.minimum(1.0, density * (1 + 20.0 / (1 + d)))
    cnt = rng.poisson(lam) * keep
    sel = cnt > 0
    return dict(chromnames=np.array([f"c{k}" for k in range(len(nb_per_chrom))]),
                chromsizes=chromsizes, bins_chrom=bc, bins_start=bs, bins_end=be,
                bin1_id=i[sel].astype(np.int64), bin2_id=j[sel].astype(np.int64),
                count=cnt[sel].astype(np.int32), binsize=binsize)


def jsonable(stats):
    out = {}
    for k, v in stats.items():
        if isinstance(v, np.ndarray):
            out[k] = [None if (isinstance(x, float) and np.isnan(x)) else x for x in v.tolist()]
        elif isinstance(v, (np.floating, float)):
            out[k] = None if np.isnan(v) else float(v)
        elif isinstance(v, (np.bool_, bool)):
            out[k] = bool(v)
        elif isinstance(v, np.integer):
            out[k] = int(v)
        else:
            out[k] = v
    return out


BALANCE_CASES = {
    "defaults": {},
    "trans_only": {"trans_only": True},
    "cis_only": {"cis_only": True},
    "test_balance_cfg": {"ignore_diags": 1, "tol": 1e-2},
    "test_balance_cis": {"ignore_diags": 1, "tol": 1e-2, "cis_only": True},
    "test_balance_trans": {"ignore_diags": 1, "tol": 1e-2, "trans_only": True},
    "mad0": {"mad_max": 0},
    "nnz0": {"min_nnz": 0},
    "mad0_nnz0": {"mad_max": 0, "min_nnz": 0},
    "min_count100": {"min_count": 100},
    "diag0": {"ignore_diags": 0},
    "diagFalse_norescale": {"ignore_diags": False, "rescale_marginals": False},
    "diag5_iters3": {"ignore_diags": 5, "max_iters": 3},
    "blacklist": {"blacklist": [0, 4, 200, 201, 202, 1500]},
}


def main():
    assert refshim.available(), "reference not present"
    gm = load_gm12878()
    np.savez_compressed(os.path.join(HERE, "gm12878_2mb.npz"), **gm)
    refshim.register("gm.cool", *table(gm))

    bal, meta = {}, {}
    for key, kw in BALANCE_CASES.items():
        bias, stats, passes, warned = refshim.ref_balance("gm.cool", **kw)
        bal[f"gm/{key}"] = bias
        meta[f"gm/{key}"] = dict(kwargs=kw, stats=jsonable(stats), passes=passes, warned=warned)
        print(key, passes if len(passes) < 4 else (sum(passes), "total"), int(np.isnan(bias).sum()))
    # x0 warm start (reference mutates and returns x0; _balance.py:368-370)
    x0 = bal["gm/defaults"].copy() * 1.5
    bias, stats, passes, warned = refshim.ref_balance("gm.cool", x0=x0.copy())
    bal["gm/x0"] = bias
    bal["gm/x0_input"] = x0
    meta["gm/x0"] = dict(kwargs={}, stats=jsonable(stats), passes=passes, warned=warned)

    # small random tables (incl. a 1-bin chromosome and an empty-ish one), all three modes
    rnd_inputs = {}
    for seed, nb in [(1, [40, 25, 1, 30]), (2, [64, 3, 50]), (3, [200, 150, 10, 90, 2])]:
        t = random_table(seed, nb, 1000, 0.5 if seed < 3 else 0.25, 40)
        name = f"rnd{seed}"
        refshim.register(name + ".cool", *table(t))
        for k, v in t.items():
            rnd_inputs[f"{name}/{k}"] = v
        for mode, kw in [("gw", {}), ("cis", {"cis_only": True}), ("trans", {"trans_only": True}),
                         ("gw_mad3", {"mad_max": 3, "min_nnz": 3, "ignore_diags": 1})]:
            bias, stats, passes, warned = refshim.ref_balance(name + ".cool", **kw)
            bal[f"{name}/{mode}"] = bias
            meta[f"{name}/{mode}"] = dict(kwargs=kw, stats=jsonable(stats), passes=passes, warned=warned)
            print(name, mode, passes, int(np.isnan(bias).sum()))
    np.savez_compressed(os.path.join(HERE, "random_tables.npz"), **rnd_inputs)
    np.savez_compressed(os.path.join(HERE, "balance_golden.npz"), **bal)
    with open(os.path.join(HERE, "balance_golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)

    # ---------------- coarsen ----------------
    co = {}
    for factor in (2, 5, 10):
        out, nb = refshim.ref_coarsen("gm.cool", factor, chunksize=10000)
        for k in ("bin1_id", "bin2_id", "count"):
            co[f"gm/x{factor}/{k}"] = out[k]
        for k in ("chrom", "start", "end"):
            co[f"gm/x{factor}/bins_{k}"] = np.asarray(nb[k])
        print("gm coarsen", factor, len(out["bin1_id"]), int(out["count"].sum()))
    # chunksize independence
    out2, _ = refshim.ref_coarsen("gm.cool", 2, chunksize=777)
    assert all(np.array_equal(out2[k], co[f"gm/x2/{k}"]) for k in out2)

    # toy chains: goldens are the reference's own bg2 files
    toy_sizes = np.array([32, 32], np.int64)
    for fam, ress in (("symm.upper", [2, 4]), ("asymm", [2, 4, 8, 16, 32])):
        for r in ress:
            a = load_toy_bg2(f"toy.{fam}.{r}.bg2", r)
            co[f"toy.{fam}/{r}"] = a
        for r0, r1 in zip(ress[:-1], ress[1:]):
            a = co[f"toy.{fam}/{r0}"]
            bc, bs, be = uniform_bins(toy_sizes, r0)
            nm = f"toy.{fam}.{r0}.cool"
            refshim.register(nm, ["chr1", "chr2"], toy_sizes, bc, bs, be, a[:, 0], a[:, 1],
                             a[:, 2].astype(np.int32), r0, symmetric_upper=(fam == "symm.upper"))
            out, _ = refshim.ref_coarsen(nm, r1 // r0, chunksize=7)
            g = co[f"toy.{fam}/{r1}"]
            assert np.array_equal(out["bin1_id"], g[:, 0]) and np.array_equal(out["bin2_id"], g[:, 1]) \
                and np.array_equal(out["count"], g[:, 2]), (fam, r0, r1)
            print("toy", fam, r0, "->", r1, "reference == bg2 golden")

    # odd.1 x4 with chunksize 2 (tests/test_reduce.py:127-137); no text twin of odd.4 -> reference output
    sizes = np.array([51, 107, 33], np.int64)
    a = np.loadtxt(os.path.join(DATA, "odd.1.coo"), dtype=np.int64)
    bc, bs, be = uniform_bins(sizes, 1)
    refshim.register("odd.1.cool", ["chr1", "chr2", "chr3"], sizes, bc, bs, be,
                     a[:, 0], a[:, 1], a[:, 2].astype(np.int32), 1)
    co["odd/1"] = a
    co["odd/chromsizes"] = sizes
    out, nb = refshim.ref_coarsen("odd.1.cool", 4, chunksize=2)
    co["odd/4"] = np.stack([out["bin1_id"], out["bin2_id"], out["count"]], axis=1)
    co["odd/4/bins_start"], co["odd/4/bins_end"] = np.asarray(nb["start"]), np.asarray(nb["end"])
    for f in (3, 7, 200):
        out, nb = refshim.ref_coarsen("odd.1.cool", f, chunksize=50)
        co[f"odd/x{f}"] = np.stack([out["bin1_id"], out["bin2_id"], out["count"]], axis=1)
    print("odd x4", len(out["bin1_id"]))

    # variable-width bins (toy.bins.var.bed) with seeded random upper-triangular pixels, x2 and x3
    chroms, st, en = [], [], []
    with open(os.path.join(DATA, "toy.bins.var.bed")) as f:
        for line in f:
            c, s, e = line.split()[:3]
            chroms.append(0 if c == "chr1" else 1), st.append(int(s)), en.append(int(e))
    bc, bs, be = np.array(chroms, np.int32), np.array(st, np.int64), np.array(en, np.int64)
    n = len(bc)
    rng = np.random.default_rng(7)
    i, j = np.triu_indices(n)
    sel = rng.random(len(i)) < 0.6
    cnt = rng.integers(1, 50, sel.sum()).astype(np.int32)
    refshim.register("var.cool", ["chr1", "chr2"], toy_sizes, bc, bs, be, i[sel], j[sel], cnt, None)
    co["var/bins"] = np.stack([bc, bs, be], axis=1)
    co["var/px"] = np.stack([i[sel], j[sel], cnt], axis=1)
    for f in (2, 3):
        out, nb = refshim.ref_coarsen("var.cool", f, chunksize=11)
        co[f"var/x{f}"] = np.stack([out["bin1_id"], out["bin2_id"], out["count"]], axis=1)
        co[f"var/x{f}/bins"] = np.stack([np.asarray(nb["chrom"]), np.asarray(nb["start"]), np.asarray(nb["end"])], axis=1)
        print("var coarsen", f, len(out["bin1_id"]))
    # ---------------- merge (CoolerMerger, _reduce.py:132-208) ----------------
    rng = np.random.default_rng(11)
    parts = []
    for k in range(3):
        sel = rng.random(len(gm["bin1_id"])) < (0.5, 0.3, 0.8)[k]
        cnt = (gm["count"][sel] * (k + 1)).astype(np.int32)
        nm = f"gm.part{k}.cool"
        refshim.register(nm, gm["chromnames"].tolist(), gm["chromsizes"], gm["bins_chrom"], gm["bins_start"],
                         gm["bins_end"], gm["bin1_id"][sel], gm["bin2_id"][sel], cnt, 2_000_000)
        co[f"merge/part{k}"] = np.stack([gm["bin1_id"][sel], gm["bin2_id"][sel], cnt], axis=1)
        parts.append(nm)
    for mergebuf in (20_000_000, 5000):
        out = refshim.ref_merge(parts, mergebuf=mergebuf)
        m = np.stack([out["bin1_id"], out["bin2_id"], out["count"]], axis=1)
        if "merge/out" in co:
            assert np.array_equal(co["merge/out"], m)          # epochs only chunk the output
        co["merge/out"] = m
    print("merge 3-way", len(co["merge/out"]), int(co["merge/out"][:, 2].sum()))
    np.savez_compressed(os.path.join(HERE, "coarsen_golden.npz"), **co)


if __name__ == "__main__":
    main()
