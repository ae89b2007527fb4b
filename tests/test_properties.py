"""Property-based tests (hypothesis) of host-side logic: HDF5 writer round trips, the dump task
planner against the matrix oracle.  CPU only."""
import os

import numpy as np
import pandas as pd
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from cooler_b200 import h5lite, h5write

from ._h5check import check

DTYPES = ["i1", "i2", "i4", "i8", "u1", "u2", "u4", "u8", "f4", "f8", "S1", "S7"]


@st.composite
def columns(draw):
    out = []
    for k in range(draw(st.integers(1, 4))):
        dt = np.dtype(draw(st.sampled_from(DTYPES)))
        n = draw(st.integers(0, 3000))
        seed = draw(st.integers(0, 2**31 - 1))
        rng = np.random.default_rng(seed)
        if dt.kind in "iu":
            info = np.iinfo(dt)
            a = rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)
        elif dt.kind == "f":
            a = rng.standard_normal(n).astype(dt)
            if n:
                a[rng.integers(0, n)] = np.nan
        else:
            a = np.array(["".join(rng.choice(list("abcXYZ09"), dt.itemsize)) for _ in range(n)], dtype=dt)
            a = a.astype(dt)
        chunk = draw(st.sampled_from([None, 1, 7, 64, 1000]))
        comp = draw(st.sampled_from(["gzip", None]))
        out.append((f"col{k}", a, chunk, comp, draw(st.booleans())))
    return out


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(cols=columns(), nattr=st.integers(0, 5), seed=st.integers(0, 1000))
def test_h5write_roundtrip_random_columns(tmp_path, cols, nattr, seed):
    p = str(tmp_path / "r.h5")
    rng = np.random.default_rng(seed)
    pool = [("i", 7), ("f", 2.5), ("s", "text é"), ("b", True), ("n", np.int32(-3)),
            ("af", np.array([1.5, np.nan])), ("ai", np.arange(4)), ("ab", np.array([True, False, True]))]
    attrs = dict(pool[i] for i in rng.choice(len(pool), nattr, replace=False))
    with h5write.File(p, "w") as f:
        g = f.create_group("a/b")
        for name, a, chunk, comp, shuf in cols:
            g.create_dataset(name, data=a, chunks=(chunk,) if chunk else None, compression=comp, shuffle=shuf)
        if attrs:
            g.attrs.update(attrs)
    check(p)
    r = h5lite.File(p)["a/b"]
    assert sorted(r.keys()) == sorted(c[0] for c in cols)
    for name, a, *_ in cols:
        got = r[name][:]
        assert got.dtype == (a.dtype if a.dtype.itemsize else np.dtype("S1")), name
        assert np.array_equal(got, a, equal_nan=a.dtype.kind == "f"), name
    back = dict(r.attrs)
    assert set(back) == set(attrs)
    for k, v in attrs.items():
        if isinstance(v, np.ndarray):
            assert np.array_equal(back[k], v, equal_nan=v.dtype.kind == "f")
        else:
            assert back[k] == v
    # a second session appends another column and replaces an attribute; everything else survives
    with h5write.File(p, "r+") as f:
        f["a/b"].create_dataset("late", data=np.arange(10))
        f["a/b"].attrs["i"] = 8
    check(p)
    r = h5lite.File(p)["a/b"]
    assert r.attrs["i"] == 8 and np.array_equal(r["late"][:], np.arange(10))
    for name, a, *_ in cols:
        assert np.array_equal(r[name][:], a, equal_nan=a.dtype.kind == "f")
    os.remove(p)


@st.composite
def matrix_and_box(draw):
    n = draw(st.integers(4, 60))
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    dens = draw(st.sampled_from([0.05, 0.3, 0.9]))
    i, j = np.triu_indices(n)
    keep = rng.random(len(i)) < dens
    b1, b2 = i[keep].astype(np.int64), j[keep].astype(np.int64)
    cnt = rng.integers(1, 50, len(b1)).astype(np.int32)
    lo1 = draw(st.integers(0, n - 1))
    hi1 = draw(st.integers(lo1, n))
    lo2 = draw(st.integers(0, n - 1))
    hi2 = draw(st.integers(lo2, n))
    return n, b1, b2, cnt, (lo1, hi1, lo2, hi2), draw(st.sampled_from([1, 5, 40, 10_000]))


class _Sel:
    def __init__(self, b1, b2, cnt):
        self.b1, self.b2, self.cnt = b1, b2, cnt

    def _slice(self, i0, i1, j0, j1):
        m = (self.b1 >= i0) & (self.b1 < i1) & (self.b2 >= j0) & (self.b2 < j1)
        return pd.DataFrame({"bin1_id": self.b1[m], "bin2_id": self.b2[m], "count": self.cnt[m]})


@settings(max_examples=200, deadline=None)
@given(case=matrix_and_box(), fill=st.booleans())
def test_dump_tasks_cover_the_window_exactly_once(case, fill):
    """Whatever the box and the chunking: the planner's tasks emit every cell of the window of the
    FULL symmetric matrix exactly once (fill) / every stored pixel in the box exactly once (direct)."""
    from cooler_b200.dump import _run_task, plan_tasks
    from oracle import matrix_numpy
    n, b1, b2, cnt, box, chunksize = case
    off = np.searchsorted(b1, np.arange(n + 1)).astype(np.int64)
    sel = _Sel(b1, b2, cnt)
    frames = [_run_task(sel, t) for t in plan_tasks(box, off, chunksize, fill)]
    got = pd.concat(frames) if frames else pd.DataFrame({"bin1_id": [], "bin2_id": [], "count": []})
    i0, i1, j0, j1 = box
    row, col, val = matrix_numpy.fetch_pixels(b1, b2, cnt, i0, i1, j0, j1, fill)
    want = sorted(zip((row + i0).tolist(), (col + j0).tolist(), val.tolist()))
    have = sorted(zip(got["bin1_id"].astype(int).tolist(), got["bin2_id"].astype(int).tolist(),
                      got["count"].astype(int).tolist()))
    assert have == want
    if not fill:                                       # direct queries come out in table order
        assert have == list(zip(got["bin1_id"].astype(int).tolist(), got["bin2_id"].astype(int).tolist(),
                                got["count"].astype(int).tolist()))


@settings(max_examples=15, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(ops=st.lists(st.tuples(st.sampled_from(["add", "del", "attr", "group", "reopen"]), st.integers(0, 6),
                              st.integers(0, 2**31 - 1)), min_size=1, max_size=14),
       which=st.sampled_from(["toy.asymm.2.cool", "toy.symm.upper.2.mcool"]))
def test_random_edit_sessions_on_reference_files(tmp_path, ops, which):
    """Random sequences of in-place edits (new columns, deletions, attribute updates, new groups,
    several sessions) on copies of files written by libhdf5: the file stays structurally valid, the
    original columns never change, and the edits read back."""
    import shutil
    src = os.path.join(os.path.dirname(__file__), "golden", "cool", which)
    p = str(tmp_path / ("e_" + which))
    shutil.copy(src, p)
    os.chmod(p, 0o644)
    base = "/resolutions/4" if which.endswith(".mcool") else "/"
    ref = h5lite.File(src)[base]
    keep = {k: ref[k][:] for k in ("pixels/bin1_id", "pixels/bin2_id", "pixels/count", "bins/start", "indexes/bin1_offset")}
    model, attrs = {}, {}
    f = h5write.File(p, "r+")
    try:
        for op, k, seed in ops:
            g = f[base]["bins"]
            name = f"c{k}"
            rng = np.random.default_rng(seed)
            if op == "add" and name not in model:
                model[name] = rng.standard_normal(int(rng.integers(0, 300)))
                g.create_dataset(name, data=model[name])
            elif op == "del" and name in model:
                del g[name]
                del model[name]
            elif op == "attr" and model:
                tgt = sorted(model)[k % len(model)]
                val = [int(seed % 97), float(seed % 13) / 7, f"s{seed % 11}", bool(seed & 1)][k % 4]
                g[tgt].attrs[f"a{k % 3}"] = val
                attrs.setdefault(tgt, {})[f"a{k % 3}"] = val
            elif op == "group":
                f[base].require_group(f"extra/g{k}")
            elif op == "reopen":
                f.close()
                check(p)
                f = h5write.File(p, "r+")
            for gone in [t for t in attrs if t not in model]:
                del attrs[gone]
    finally:
        f.close()
    check(p)
    r = h5lite.File(p)[base]
    for k, v in keep.items():
        assert np.array_equal(r[k][:], v), k
    assert set(r["bins"].keys()) == {"chrom", "start", "end"} | set(model)
    for name, v in model.items():
        assert np.array_equal(r["bins"][name][:], v)
        assert dict(r["bins"][name].attrs) == attrs.get(name, {})
    os.remove(p)
