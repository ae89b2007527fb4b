"""Run in a SUBPROCESS by tests/test_host_logic.py: pure host functions of cooler_b200 against their
counterparts in the LIVE unmodified reference on seeded random inputs -- zoom planners
(get_multiplier_sequence, preferred_sequence, _greedy_prune_partition), region parsing (util.parse_region,
region_to_extent), and the general-aggregation oracles (CoolerMerger / CoolerCoarsener with random
aggregators) that pin oracle.coarsen_numpy.merge_columns."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import coarsen_numpy, matrix_numpy, refshim  # noqa: E402

from cooler_b200 import api, reduce  # noqa: E402


def main(n_cases, seed):
    cooler = refshim.import_reference()
    import cooler._reduce as R
    import cooler.util as U
    rng = np.random.default_rng(seed)
    counts = {"multiplier": 0, "preferred": 0, "prune": 0, "region": 0, "merge": 0, "matrix": 0, "validate": 0,
              "binmap": 0, "extent": 0, "compat": 0}
    from cooler.create._ingest import BadInputError as RefBad
    from cooler.create._ingest import validate_pixels as ref_validate

    from cooler_b200.create import BadInputError as OurBad
    from cooler_b200.create import _validate_chunk
    nice = [1000, 2000, 5000, 10000, 25000, 50000, 100000, 250000, 500000, 1000000, 2500000, 5000000, 10000000]
    for case in range(n_cases):
        # ---- zoom planners -------------------------------------------------------------------------
        base = int(rng.choice([1, 2, 10, 100, 1000, 5000]))
        res = sorted({base * int(m) for m in rng.choice([1, 2, 3, 4, 5, 6, 8, 10, 20, 25, 50, 100], int(rng.integers(1, 7)))}
                     | ({int(x) for x in rng.choice(nice, 2)} if rng.random() < 0.3 else set()))
        bases = [base] + ([base * 3] if rng.random() < 0.3 else [])
        try:
            want = R.get_multiplier_sequence(list(res), list(bases))
            werr = None
        except ValueError as e:
            want, werr = None, str(e)[:30]
        try:
            got = reduce.get_multiplier_sequence(list(res), list(bases))
            gerr = None
        except ValueError as e:
            got, gerr = None, str(e)[:30]
        assert (werr is None) == (gerr is None), (res, bases, werr, gerr)
        if werr is None:
            for a, b in zip(got, want):
                assert list(a) == list(b), (res, bases, got, want)
        counts["multiplier"] += 1
        lo, hi = int(rng.integers(1, 5000)), int(rng.integers(1, 20_000_000))
        for style in ("nice", "binary"):
            assert list(reduce.preferred_sequence(lo, hi, style)) == list(R.preferred_sequence(lo, hi, style)), (lo, hi, style)
            counts["preferred"] += 1
        edges = np.unique(np.r_[0, np.sort(rng.integers(1, 5000, int(rng.integers(1, 60))))])
        maxlen = int(rng.integers(1, 800))
        assert list(reduce._prune_edges(edges, maxlen)) == list(R._greedy_prune_partition(edges, maxlen)), (edges, maxlen)
        counts["prune"] += 1
        # ---- region parsing ------------------------------------------------------------------------
        sizes = {"chr1": 1000, "chrX": 123_456_789, "c.2": 5}
        name = str(rng.choice(list(sizes) + ["nope", ""]))
        a, b = int(rng.integers(-5, 2000)), int(rng.integers(-5, 200_000_000))
        forms = [name, f"{name}:{a}-{b}", f"{name}:{a:,}-{b:,}", f"{name}:{a}-", f"{name}:{abs(a)}kb-{abs(b) // 1000}Mb",
                 f" {name} : {a} - {b} ", f"{name}:{a}", f"{name}:{a}-{b}-{a}", (name, a, b), (name, None, None), (name, a, None)]
        import pandas as pd
        cs = pd.Series(sizes)
        for reg in forms:
            try:
                want, werr = U.parse_region(reg, cs), None
            except Exception as e:                        # noqa: BLE001 - the reference raises several types
                want, werr = None, type(e).__name__
            try:
                got, gerr = api.parse_region(reg, cs), None
            except Exception as e:                        # noqa: BLE001
                got, gerr = None, type(e).__name__
            assert (werr is None) == (gerr is None), (reg, want, got, werr, gerr)
            if werr is None:
                assert tuple(got) == tuple(want), (reg, got, want)
            counts["region"] += 1
        # ---- general merge oracle vs CoolerMerger with random aggregators --------------------------------
        if case % 4 == 0:
            n_chroms = int(rng.integers(1, 4))
            nb = [int(rng.integers(2, 25)) for _ in range(n_chroms)]
            bc = np.repeat(np.arange(n_chroms, dtype=np.int32), nb)
            bs = np.concatenate([np.arange(k) * 10 for k in nb]).astype(np.int64)
            be = bs + 10
            n = len(bc)
            tables, names = [], []
            for k in range(int(rng.integers(2, 5))):
                i, j = np.triu_indices(n)
                sel = rng.random(len(i)) < rng.uniform(0.2, 0.8)
                m = int(sel.sum())
                cols = {"count": rng.integers(1, 50, m).astype(np.int32), "score": rng.random(m),
                        "w": rng.integers(-5, 5, m).astype(np.int64)}
                nm = f"fuzzmerge{case}_{k}.cool"
                refshim.register(nm, [f"c{q}" for q in range(n_chroms)], np.array(nb) * 10, bc, bs, be,
                                 i[sel], j[sel], cols["count"], 10, extra_columns={"score": cols["score"], "w": cols["w"]})
                tables.append(dict(bin1_id=i[sel].astype(np.int64), bin2_id=j[sel].astype(np.int64), **cols))
                names.append(nm)
            ops = ["sum", "min", "max", "first", "last", "mean"]
            columns = ["count", "score", "w"]
            agg = {c: str(rng.choice(ops)) for c in columns if rng.random() < 0.7}
            ref = refshim.ref_merge(names, mergebuf=int(rng.integers(20, 2000)), columns=columns, agg=dict(agg))
            got = coarsen_numpy.merge_columns(tables, columns, agg)
            assert sorted(got) == sorted(ref), (got.keys(), ref.keys())
            for c, v in got.items():
                w = ref[c]
                if w.dtype.kind in "iu":
                    assert v.dtype == w.dtype and np.array_equal(v, w), (case, c, agg)
                else:
                    assert v.dtype == w.dtype, (case, c, agg, v.dtype, w.dtype)
                    np.testing.assert_allclose(v, w, rtol=1e-12, err_msg=f"{case} {c} {agg}")
            counts["merge"] += 1
        # ---- coarsened bin table and the closed-form bin map vs the live CoolerCoarsener ------------------------
        if case % 2 == 0:
            variable = bool(rng.integers(0, 2))
            nch = int(rng.integers(1, 5))
            rows = []
            for c in range(nch):
                length = int(rng.integers(3, 500))
                edges_ = (np.unique(np.r_[0, rng.integers(1, length, int(rng.integers(0, 15))), length]) if variable
                          else np.unique(np.r_[np.arange(0, length, 9), length]))
                rows += [(c, int(a), int(b)) for a, b in zip(edges_[:-1], edges_[1:])]
            bt = np.array(rows, dtype=np.int64)
            bcx, bsx, bex = bt[:, 0].astype(np.int32), bt[:, 1], bt[:, 2]
            sizes_ = np.array([bex[bcx == c].max() for c in range(nch)], dtype=np.int64)
            nb_tot = len(bcx)
            diag = np.arange(nb_tot, dtype=np.int64)
            nm = f"fuzzbins{case}.cool"
            refshim.register(nm, [f"c{q}" for q in range(nch)], sizes_, bcx, bsx, bex, diag, diag,
                             np.ones(nb_tot, np.int32), None if variable else 9)
            # region -> bin extent (core/_rangequery.py:11-31) on the same bin table
            from cooler_b200 import MemCooler
            mine = MemCooler.from_arrays([f"c{q}" for q in range(nch)], sizes_, bcx, bsx, bex, diag, diag,
                                         np.ones(nb_tot, np.int32), binsize=None if variable else 9)
            theirs = cooler.Cooler(nm)
            for _q in range(6):
                c = int(rng.integers(0, nch))
                a_, b_ = sorted(int(x) for x in rng.integers(0, sizes_[c] + 1, 2))
                reg = (f"c{c}", a_, b_) if rng.random() < 0.5 else (f"c{c}:{a_}-{b_}" if rng.random() < 0.8 else f"c{c}")
                assert tuple(api.region_to_extent(mine, reg)) == tuple(int(x) for x in theirs.extent(reg)), (case, reg, variable)
                counts["extent"] += 1
            factor = int(rng.integers(2, 12))
            ref_px, ref_bins = refshim.ref_coarsen(nm, factor, chunksize=int(rng.integers(3, 200)))
            gc_, gs_, ge_ = reduce.coarsen_bins(bcx, bsx, bex, sizes_, factor)
            assert np.array_equal(gc_, ref_bins["chrom"]) and np.array_equal(gs_, ref_bins["start"]) \
                and np.array_equal(ge_, ref_bins["end"]), (case, factor, variable)
            off_ = np.searchsorted(bcx, np.arange(nch + 1))
            bmap, new_off = reduce.coarsen_bin_map(off_, factor)
            pooled = np.bincount(bmap, minlength=int(new_off[-1]))
            assert int(new_off[-1]) == len(gc_)
            # A variable-bin table whose COARSENED bins happen to look uniform (all but the last bin of every
            # chromosome equally wide) sends the reference down its fixed-bin formula `offset + start // binsize`
            # (_reduce.py:611-614), which mis-assigns pixels of such a table (ids can even exceed the new bin
            # table).  cooler_b200 pools `factor` consecutive bins, which is what the new bin table describes;
            # the comparison is made wherever the reference's formula applies.
            quirk = variable and coarsen_numpy._binsize_of(gc_, gs_, ge_) is not None
            if not quirk:
                assert np.array_equal(np.flatnonzero(pooled), ref_px["bin1_id"]), (case, factor, variable)
                assert np.array_equal(pooled[pooled > 0], ref_px["count"])
                assert np.array_equal(ref_px["bin1_id"], ref_px["bin2_id"])
                counts["binmap"] += 1
        # ---- CoolerMerger's compatibility checks (_reduce.py:150-166): same verdicts -------------------------
        if case % 3 == 0:
            from cooler_b200 import MemCooler as _MC

            def make(tag, nchr, bsz, lens, variable):
                rows_ = []
                for c in range(nchr):
                    e_ = (np.unique(np.r_[0, rng.integers(1, lens[c], 3), lens[c]]) if variable
                          else np.unique(np.r_[np.arange(0, lens[c], bsz), lens[c]]))
                    rows_ += [(c, int(a), int(b)) for a, b in zip(e_[:-1], e_[1:])]
                t_ = np.array(rows_, dtype=np.int64)
                nm_ = f"fuzzcompat{case}_{tag}.cool"
                nmz = [f"c{q}" for q in range(nchr)]
                z = np.zeros(1, np.int64)
                refshim.register(nm_, nmz, np.array(lens[:nchr]), t_[:, 0].astype(np.int32), t_[:, 1], t_[:, 2], z, z,
                                 np.ones(1, np.int32), None if variable else bsz)
                mine_ = _MC.from_arrays(nmz, np.array(lens[:nchr]), t_[:, 0].astype(np.int32), t_[:, 1], t_[:, 2], z, z,
                                        np.ones(1, np.int32), binsize=None if variable else bsz)
                return cooler.Cooler(nm_), mine_
            lens_a = [int(rng.integers(20, 90)) for _ in range(3)]
            spec_a = (int(rng.integers(1, 4)), int(rng.choice([5, 10])), lens_a, rng.random() < 0.25)
            spec_b = list(spec_a)
            r = rng.random()
            if r < 0.2:
                spec_b[0] = int(rng.integers(1, 4))
            elif r < 0.4:
                spec_b[1] = int(rng.choice([5, 10]))
            elif r < 0.6:
                spec_b[2] = [x + int(rng.integers(0, 2)) for x in lens_a]
            elif r < 0.7:
                spec_b[3] = not spec_a[3]
            ra, ma = make("a", *spec_a)
            rb, mb = make("b", *spec_b)
            try:
                R.CoolerMerger([ra, rb], mergebuf=100)
                werr = None
            except ValueError as e:
                werr = "ValueError"
            try:
                reduce.CoolerMerger([ma, mb], mergebuf=100)
                gerr = None
            except ValueError as e:
                gerr = "ValueError"
            assert werr == gerr, (case, spec_a, spec_b, werr, gerr)
            counts["compat"] += 1
        # ---- chunk validation (create/_ingest.py:401-446): same verdicts -----------------------------------
        nb_ = int(rng.integers(2, 30))
        m = int(rng.integers(0, 40))
        chunk = {"bin1_id": rng.integers(-1 if rng.random() < 0.1 else 0, nb_ + (1 if rng.random() < 0.1 else 0), m),
                 "bin2_id": rng.integers(0, nb_, m), "count": rng.integers(1, 5, m)}
        if rng.random() < 0.5:
            o = np.lexsort((chunk["bin2_id"], chunk["bin1_id"]))
            chunk = {k: v[o] for k, v in chunk.items()}
        flags = [bool(rng.integers(0, 2)) for _ in range(4)]
        import pandas as pd
        try:
            w = ref_validate(nb_, *flags)(pd.DataFrame(chunk))
            werr = None
        except RefBad as e:
            w, werr = None, str(e)[:22]
        try:
            g = _validate_chunk({k: v.copy() for k, v in chunk.items()}, nb_, *flags)
            gerr = None
        except OurBad as e:
            g, gerr = None, str(e)[:22]
        assert werr == gerr, (chunk, flags, werr, gerr)
        if werr is None:
            for k in chunk:
                assert np.array_equal(np.asarray(g[k]), w[k].to_numpy()), (k, flags)
        counts["validate"] += 1
        # ---- matrix oracle vs the live Cooler.matrix on random windows ----------------------------------------
        if case % 5 == 0:
            n = int(rng.integers(5, 60))
            square = rng.random() < 0.4
            i, j = (np.indices((n, n)).reshape(2, -1) if square else np.triu_indices(n))
            sel = rng.random(len(i)) < 0.4
            b1, b2 = i[sel].astype(np.int64), j[sel].astype(np.int64)
            cnt = rng.integers(0, 30, len(b1)).astype(np.int32)
            wts = rng.lognormal(0, 0.3, n)
            wts[rng.random(n) < 0.15] = np.nan
            grp = refshim.register(f"fuzzmat{case}.cool", ["a"], [n * 10], np.zeros(n, np.int32), np.arange(n) * 10,
                                   np.arange(1, n + 1) * 10, b1, b2, cnt, 10, symmetric_upper=not square)
            grp["bins"]["weight"] = wts.copy()
            clr = cooler.Cooler(f"fuzzmat{case}.cool")
            for _ in range(6):
                i0, i1 = sorted(int(x) for x in rng.integers(0, n + 1, 2))
                j0, j1 = sorted(int(x) for x in rng.integers(0, n + 1, 2))
                bal, div = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
                with np.errstate(all="ignore"):
                    want = clr.matrix(balance=bal, divisive_weights=div)[i0:i1, j0:j1]
                    got = matrix_numpy.fetch_dense(b1, b2, cnt, i0, i1, j0, j1, symmetric_upper=not square,
                                                   weights=wts if bal else None, divisive_weights=div)
                assert got.shape == want.shape and got.dtype == want.dtype, (case, got.dtype, want.dtype)
                assert np.array_equal(got, want, equal_nan=True), (case, i0, i1, j0, j1, bal, div, square)
                with np.errstate(all="ignore"):
                    sp = clr.matrix(balance=bal, sparse=True)[i0:i1, j0:j1]
                    r, c, d = matrix_numpy.fetch_sparse(b1, b2, cnt, i0, i1, j0, j1, symmetric_upper=not square,
                                                        weights=wts if bal else None)
                o = np.lexsort((sp.col, sp.row))
                assert np.array_equal(sp.row[o], r) and np.array_equal(sp.col[o], c)
                assert np.array_equal(sp.data[o], d, equal_nan=True), (case, "sparse")
                counts["matrix"] += 1
    print("ok:", counts)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
