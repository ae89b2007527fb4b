"""Run in a SUBPROCESS by tests/test_matrix_oracle.py: the HOST side of the matrix fetch -- api.matrix (weight
column lookup, divisive defaults for KR / VC / VC_SQRT, `field=` routing and dtypes, errors) and the selector's
indexing (`[i0:i1, j0:j1]`, integers, negative and open slices, `.fetch(region[, region2])`) -- against the LIVE
unmodified reference's Cooler.matrix on seeded random coolers.  A subclass of MatrixSelector answers `_slice` with
the numpy restatement (oracle/matrix_numpy.py) instead of the kernels; everything else runs as on a GPU."""
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import matrix_numpy, refshim  # noqa: E402


def main(n_cases, seed):
    cooler = refshim.import_reference()
    import pandas as pd

    import cooler_b200.api as A
    from cooler_b200 import MemCooler

    class NumpySelector(A.MatrixSelector):
        def __init__(self, clr, table, weights, *, sparse=False, as_pixels=False, divisive_weights=False,
                     fill_lower=True, weight_name="weight", field="count", values=None):
            self.clr, self.table, self.sparse, self.as_pixels = clr, table, sparse, as_pixels
            self.divisive, self.field = bool(divisive_weights), field
            n = table.n_bins
            self.shape = (n, n)
            symmetric = clr.storage_mode == "symmetric-upper"
            self.fill_lower = bool(fill_lower and symmetric and not as_pixels)
            self.w = None if weights is None else np.asarray(weights, dtype=np.float64)
            self.vals = table.cnt if values is None else np.asarray(values)
            self.values = values

        def _slice(self, i0, i1, j0, j1):
            t = self.table
            if not (self.sparse or self.as_pixels):
                return matrix_numpy.fetch_dense(t.b1, t.b2, self.vals, i0, i1, j0, j1, symmetric_upper=self.fill_lower,
                                                weights=self.w, divisive_weights=self.divisive)
            r, c, v = matrix_numpy.fetch_pixels(t.b1, t.b2, self.vals, i0, i1, j0, j1, self.fill_lower)
            bal = None
            if self.w is not None:
                b1_, b2_ = self.w[i0:i1], self.w[j0:j1]
                if self.divisive:
                    b1_, b2_ = 1 / b1_, 1 / b2_
                bal = b1_[r] * b2_[c] * v
            if self.as_pixels:
                df = pd.DataFrame({"bin1_id": r + i0, "bin2_id": c + j0, self.field: v})
                if bal is not None:
                    df["balanced"] = bal
                return df
            from scipy.sparse import coo_matrix
            return coo_matrix((bal if bal is not None else v, (r, c)), shape=(i1 - i0, j1 - j0))

    def from_cooler(clr, device="cuda", structure_only=False):
        return SimpleNamespace(b1=np.asarray(clr._load_dset("pixels/bin1_id")), b2=np.asarray(clr._load_dset("pixels/bin2_id")),
                               cnt=np.asarray(clr._load_dset("pixels/count")), n_bins=int(clr.info["nbins"]), device=None,
                               raw=SimpleNamespace(nnz=len(np.asarray(clr._load_dset("pixels/count")))))

    A.MatrixSelector = NumpySelector
    A.PixelTable.from_cooler = staticmethod(from_cooler)
    rng = np.random.default_rng(seed)
    n_q = n_err = 0
    for case in range(n_cases):
        binsize = 100
        nb = [int(rng.integers(1, 25)) for _ in range(int(rng.integers(1, 4)))]
        names = [f"c{k}" for k in range(len(nb))]
        sizes = np.array([k * binsize - int(rng.integers(0, 50)) for k in nb], dtype=np.int64)
        bc = np.repeat(np.arange(len(nb), dtype=np.int32), nb)
        bs = np.concatenate([np.arange(k) * binsize for k in nb]).astype(np.int64)
        be = np.minimum(bs + binsize, sizes[bc])
        n = len(bc)
        square = rng.random() < 0.3
        i, j = (np.indices((n, n)).reshape(2, -1) if square else np.triu_indices(n))
        sel = rng.random(len(i)) < 0.4
        b1, b2 = i[sel].astype(np.int64), j[sel].astype(np.int64)
        m = len(b1)
        cdt = str(rng.choice(["int32", "float64", "float32"]))
        cnt = rng.integers(0, 90, m).astype(cdt) if cdt == "int32" else (rng.random(m) * 9).astype(cdt)
        extra = {"score": rng.random(m), "big": rng.integers(0, 2**40, m).astype(np.int64)}
        wcols = {}
        for wname in ("weight", "KR", "VC_SQRT", "mine"):
            if rng.random() < 0.6:
                w = rng.lognormal(0, 0.4, n)
                w[rng.random(n) < 0.15] = np.nan
                wcols[wname] = w
        name = f"fuzzapi{case}.cool"
        g = refshim.register(name, names, sizes, bc, bs, be, b1, b2, cnt, binsize, symmetric_upper=not square,
                             extra_columns=extra)
        mine = MemCooler.from_arrays(names, sizes, bc, bs, be, b1, b2, cnt, binsize=binsize,
                                     symmetric_upper=not square, extra_pixel_columns=extra)
        for k, w in wcols.items():
            g["bins"][k] = w.copy()
            mine.root["bins"].create_dataset(k, w.copy())
        theirs = cooler.Cooler(name)
        for _ in range(8):
            kw = {}
            r = rng.random()
            kw["balance"] = (False if r < 0.35 else True if r < 0.55 else str(rng.choice(["weight", "KR", "VC_SQRT", "mine", "nope"])))
            if rng.random() < 0.3:
                kw["divisive_weights"] = bool(rng.integers(0, 2))
            if rng.random() < 0.4:
                kw["field"] = str(rng.choice(["count", "score", "big", "absent"]))
            mode = rng.random()
            if mode < 0.3:
                kw["sparse"] = True
            elif mode < 0.5:
                kw.update(as_pixels=True, join=False)

            def index():
                # bounds inside [-n, n]: beyond them the reference neither clamps (Python slice semantics, which
                # cooler_b200 follows) nor checks -- it fails somewhere downstream
                q = rng.random()
                if q < 0.15:
                    return int(rng.integers(-n, n + 2))
                a, b = (int(x) for x in rng.integers(0, n + 1, 2))
                a, b = min(a, b), max(a, b)
                if rng.random() < 0.3:
                    a, b = a - n, (b - n if b < n else b)   # the same positions written from the end
                    if b <= 0 and b - a < 0:
                        a, b = b, a
                if q < 0.3:
                    return slice(None, b)
                if q < 0.45:
                    return slice(a, None)
                return slice(a, b)
            use_fetch = rng.random() < 0.3
            if use_fetch:
                c = int(rng.integers(0, len(names)))
                a, b = sorted(int(x) for x in rng.integers(0, sizes[c] + 1, 2))
                q1 = f"{names[c]}:{a}-{b}"
                q2 = names[int(rng.integers(0, len(names)))] if rng.random() < 0.5 else None
            else:
                key = (index(), index()) if rng.random() < 0.85 else index()
            outs = []
            for obj, fn in ((theirs, lambda k: theirs.matrix(**k)), (mine, lambda k: A.matrix(mine, **k))):
                try:
                    with np.errstate(all="ignore"):
                        s_ = fn(kw)
                        res = s_.fetch(q1, q2) if use_fetch else s_[key]
                    outs.append(("ok", res))
                except Exception as e:                    # noqa: BLE001
                    outs.append(("err", type(e).__name__))
            (ws, wv), (gs, gv) = outs
            assert ws == gs, (case, kw, None if use_fetch else key, wv, gv)
            if ws == "err":
                n_err += 1
                continue
            if kw.get("sparse"):
                o1, o2 = np.lexsort((wv.col, wv.row)), np.lexsort((gv.col, gv.row))
                assert wv.shape == gv.shape and np.array_equal(wv.row[o1], gv.row[o2]) and np.array_equal(wv.col[o1], gv.col[o2])
                assert wv.data.dtype == gv.data.dtype and np.array_equal(wv.data[o1], gv.data[o2], equal_nan=True), (case, kw)
            elif kw.get("as_pixels"):
                assert list(wv.columns) == list(gv.columns), (case, kw, list(wv.columns), list(gv.columns))
                for c_ in wv.columns:
                    x, y = wv[c_].to_numpy(), gv[c_].to_numpy()
                    assert x.dtype == y.dtype and np.array_equal(x, y, equal_nan=(x.dtype.kind == "f")), (case, kw, c_)
            else:
                assert wv.shape == gv.shape and wv.dtype == gv.dtype, (case, kw, wv.shape, gv.shape, wv.dtype, gv.dtype)
                assert np.array_equal(wv, gv, equal_nan=(wv.dtype.kind == "f")), (case, kw)
            n_q += 1
    print(f"ok: {n_q} matrix queries identical, {n_err} failing the same way")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 60, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
