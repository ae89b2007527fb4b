"""Run in a SUBPROCESS by tests/test_host_logic.py: the HOST glue of the coarsening / merging path --
reduce.CoolerCoarsener (fast and general value-column paths, chunking, dtypes, new bin tables),
coarsen_cooler(None), the in-memory zoomify_cooler chain (planning, level-to-level hand-over) and
CoolerMerger / merge_coolers(None) -- against the LIVE unmodified reference's CoolerCoarsener / CoolerMerger
chunk streams on seeded random tables.  The device functions (coarsen_table, coarsen_columns, merge_tables,
merge_columns, table_to_host, PixelTable.from_cooler) are replaced by numpy restatements of what the kernels
compute, so everything above the C ABI runs as it does on a GPU."""
import os
import sys
import warnings
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import coarsen_numpy, refshim  # noqa: E402


class FakeTensor:
    def __init__(self, a):
        self.a = np.asarray(a)

    def cpu(self):
        return self

    def numpy(self):
        return self.a


class FakeTable:
    """Stands in for a device-resident PixelTable: sorted (bin1, bin2, count) + chrom_offset."""

    def __init__(self, b1, b2, cnt, chrom_offset):
        self.b1, self.b2, self.cnt = np.asarray(b1, np.int64), np.asarray(b2, np.int64), np.asarray(cnt)
        self.chrom_offset = np.asarray(chrom_offset, np.int64)
        self.n_bins = int(self.chrom_offset[-1])
        self.device = None
        ind = np.searchsorted(self.b1, np.arange(self.n_bins + 1), side="left")
        self.raw = SimpleNamespace(indptr=FakeTensor(ind), nnz=len(self.b1))


def main(n_cases, seed):
    cooler = refshim.import_reference()
    import cooler_b200.reduce as R
    from cooler_b200 import MemCooler

    def from_cooler(clr, device="cuda", structure_only=False):
        return FakeTable(np.asarray(clr._load_dset("pixels/bin1_id")), np.asarray(clr._load_dset("pixels/bin2_id")),
                         np.asarray(clr._load_dset("pixels/count")), np.asarray(clr._load_dset("indexes/chrom_offset")))

    def group_sum(b1, b2, cols, agg):
        out = coarsen_numpy.merge_columns([dict(bin1_id=b1, bin2_id=b2, **cols)], list(cols), agg)
        return out.pop("bin1_id"), out.pop("bin2_id"), out

    def coarsen_table(table, factor):
        bmap, new_off = R.coarsen_bin_map(table.chrom_offset, factor)
        n1, n2 = bmap[table.b1].astype(np.int64), bmap[table.b2].astype(np.int64)
        b1, b2, out = group_sum(n1, n2, {"count": table.cnt.astype(np.int64)}, None)
        if len(out["count"]) and out["count"].max() > 2**31 - 1:
            raise OverflowError("coarsened count does not fit int32")
        new = FakeTable(b1, b2, out["count"].astype(np.int32), new_off)
        return new, b1

    def table_to_host(table, ids=None):
        return table.b1, table.b2, table.cnt.astype(np.int32)

    def coarsen_columns(bin1_offset, bin2_id, chrom_offset, factor, columns, agg, device="cuda"):
        bmap, new_off = R.coarsen_bin_map(chrom_offset, factor)
        b1 = np.repeat(np.arange(len(bin1_offset) - 1, dtype=np.int64), np.diff(bin1_offset))
        n1, n2 = bmap[b1].astype(np.int64), bmap[np.asarray(bin2_id)].astype(np.int64)
        g1, g2, out = group_sum(n1, n2, {k: np.asarray(v) for k, v in columns.items()}, agg)
        # the device returns accumulators: integers in int64, floats and means in float64
        acc = {k: (v.astype(np.float64) if (v.dtype.kind == "f" or agg.get(k, "sum") == "mean") else v.astype(np.int64))
               for k, v in out.items()}
        return g1, g2, acc, new_off

    def merge_tables(tables):
        b1 = np.concatenate([t.b1 for t in tables])
        b2 = np.concatenate([t.b2 for t in tables])
        cnt = np.concatenate([t.cnt.astype(np.int64) for t in tables])
        g1, g2, out = group_sum(b1, b2, {"count": cnt}, None)
        if len(out["count"]) and out["count"].max() > 2**31 - 1:
            raise OverflowError("merged count does not fit int32")
        return FakeTable(g1, g2, out["count"].astype(np.int32), tables[0].chrom_offset), g1

    def merge_columns(inputs, chrom_offset, columns, agg, device="cuda"):
        tables = []
        for off, b2, cols in inputs:
            b1 = np.repeat(np.arange(len(off) - 1, dtype=np.int64), np.diff(off))
            tables.append(dict(bin1_id=b1, bin2_id=np.asarray(b2, np.int64), **cols))
        out = coarsen_numpy.merge_columns(tables, list(columns), agg)
        return out.pop("bin1_id"), out.pop("bin2_id"), out

    R.PixelTable.from_cooler = staticmethod(from_cooler)
    R.coarsen_table, R.table_to_host, R.coarsen_columns = coarsen_table, table_to_host, coarsen_columns
    R.merge_tables, R.merge_columns = merge_tables, merge_columns
    rng = np.random.default_rng(seed)
    counts = {"coarsen": 0, "general": 0, "zoomify": 0, "merge": 0}
    ops = ["sum", "min", "max", "first", "last", "mean"]
    for case in range(n_cases):
        binsize = int(rng.choice([10, 1000]))
        nb = [int(rng.integers(1, 40)) for _ in range(int(rng.integers(1, 5)))]
        names = [f"c{k}" for k in range(len(nb))]
        sizes = np.array([k * binsize - int(rng.integers(0, binsize // 2)) for k in nb], dtype=np.int64)
        bc = np.repeat(np.arange(len(nb), dtype=np.int32), nb)
        bs = np.concatenate([np.arange(k) * binsize for k in nb]).astype(np.int64)
        be = np.minimum(bs + binsize, sizes[bc])
        n = len(bc)
        square = rng.random() < 0.25
        i, j = (np.indices((n, n)).reshape(2, -1) if square else np.triu_indices(n))
        sel = rng.random(len(i)) < rng.uniform(0.05, 0.7)
        b1, b2 = i[sel].astype(np.int64), j[sel].astype(np.int64)
        m = len(b1)
        cdt = rng.choice(["int32", "int32", "int16", "float64", "float32", "int64"])
        cnt = rng.integers(1, 200, m).astype(cdt) if cdt.startswith("int") else (rng.random(m) * 50).astype(cdt)
        extra = {"score": rng.random(m), "w": rng.integers(-9, 9, m).astype(np.int64)}
        name = f"fuzzred{case}.cool"
        refshim.register(name, names, sizes, bc, bs, be, b1, b2, cnt, binsize, symmetric_upper=not square,
                         extra_columns=extra)
        mine = MemCooler.from_arrays(names, sizes, bc, bs, be, b1, b2, cnt, binsize=binsize,
                                     symmetric_upper=not square, extra_pixel_columns=extra)
        # ---- CoolerCoarsener: count + sum (fast path when the count is an int32-class column) ---------------
        factor = int(rng.integers(2, 9))
        chunksize = int(rng.integers(1, 400))
        ref, ref_bins = refshim.ref_coarsen(name, factor, chunksize=chunksize)
        it = R.CoolerCoarsener(mine, factor, chunksize)
        chunks = list(it)
        got = {k: (np.concatenate([c[k] for c in chunks]) if chunks else np.zeros(0)) for k in ("bin1_id", "bin2_id", "count")}
        for a_, b_ in zip(chunks[:-1], chunks[1:]):
            assert a_["bin1_id"][-1] < b_["bin1_id"][0], ("a coarse row was split", case)
        if m:
            for k in ("bin1_id", "bin2_id"):
                assert np.array_equal(got[k], ref[k]), (case, k, factor)
            assert got["count"].dtype == ref["count"].dtype, (case, got["count"].dtype, ref["count"].dtype, cdt)
            np.testing.assert_allclose(got["count"].astype(float), ref["count"].astype(float),
                                       rtol=1e-6 if cdt == "float32" else 1e-12, err_msg=str(case))
        nbins = it.new_bins
        assert np.array_equal(np.asarray(nbins["chrom"].cat.codes), ref_bins["chrom"])
        assert np.array_equal(nbins["start"].values, ref_bins["start"]) and np.array_equal(nbins["end"].values, ref_bins["end"])
        counts["coarsen"] += 1
        # ---- general value columns / aggregators ------------------------------------------------------------
        if m and case % 2 == 0:
            columns = ["count"] + [c for c in ("score", "w") if rng.random() < 0.6]
            agg = {c: str(rng.choice(ops)) for c in columns if rng.random() < 0.7}
            ref, _ = refshim.ref_coarsen(name, factor, chunksize=chunksize, columns=columns, agg=dict(agg) or None)
            chunks = list(R.CoolerCoarsener(mine, factor, chunksize, columns=columns, agg=dict(agg) or None))
            got = {k: np.concatenate([c[k] for c in chunks]) for k in chunks[0]}
            assert list(got) == list(ref), (case, list(got), list(ref))
            for k, v in got.items():
                w_ = ref[k]
                assert v.dtype == w_.dtype, (case, k, agg, cdt, v.dtype, w_.dtype)
                if w_.dtype.kind in "iu":
                    assert np.array_equal(v, w_), (case, k, agg)
                else:
                    np.testing.assert_allclose(v, w_, rtol=1e-6 if w_.dtype == np.float32 else 1e-12, err_msg=f"{case} {k} {agg}")
            counts["general"] += 1
        # ---- in-memory zoomify chain: every level equals the reference's coarsening of the base ---------------
        if m and cdt == "int32" and case % 3 == 0:
            res = sorted({binsize * int(f) for f in rng.choice([2, 3, 4, 6, 8, 12], int(rng.integers(1, 4)))})
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                levels = R.zoomify_cooler(mine, None, res, chunksize)
            assert sorted(levels) == sorted([binsize] + res), (case, sorted(levels), res)
            for r_ in res:
                ref, _ = refshim.ref_coarsen(name, r_ // binsize, chunksize=10_000)
                for k in ("bin1_id", "bin2_id", "count"):
                    assert np.array_equal(np.asarray(levels[r_]._load_dset("pixels/" + k)), ref[k]), (case, r_, k)
                assert levels[r_].binsize == r_
            counts["zoomify"] += 1
        # ---- CoolerMerger on two or three row-subsets of the table ------------------------------------------
        if m and case % 2 == 1 and cdt.startswith("int") and cdt != "int64":
            parts_ref, parts_mine = [], []
            for k in range(int(rng.integers(2, 4))):
                s_ = rng.random(m) < 0.6
                nm = f"fuzzred{case}_p{k}.cool"
                refshim.register(nm, names, sizes, bc, bs, be, b1[s_], b2[s_], cnt[s_], binsize, symmetric_upper=not square)
                parts_ref.append(nm)
                parts_mine.append(MemCooler.from_arrays(names, sizes, bc, bs, be, b1[s_], b2[s_], cnt[s_], binsize=binsize,
                                                        symmetric_upper=not square))
            buf = int(rng.integers(5, 2000))
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref = refshim.ref_merge(parts_ref, mergebuf=buf) if any(len(np.asarray(p._load_dset("pixels/count"))) for p in parts_mine) else None
                chunks = list(R.CoolerMerger(parts_mine, buf))
            if ref is not None and chunks:
                got = {k: np.concatenate([c[k] for c in chunks]) for k in chunks[0]}
                for k in ("bin1_id", "bin2_id", "count"):
                    assert np.array_equal(got[k].astype(np.int64), ref[k].astype(np.int64)), (case, "merge", k)
                counts["merge"] += 1
    print("ok:", counts)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 60, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
