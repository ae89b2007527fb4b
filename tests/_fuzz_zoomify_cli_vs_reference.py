"""Run in a SUBPROCESS by tests/test_host_logic.py: the argument handling of `cooler zoomify` and `cooler coarsen`
(resolution strings -- explicit lists, N / B / 4DN / <k>N progressions --, --field specifications, -o) in
cooler_b200/cli.py against the LIVE unmodified reference CLIs (src/cooler/cli/zoomify.py, coarsen.py): both
commands' back ends (zoomify_cooler / coarsen_cooler) are replaced by recorders and the recorded calls compared.
One deliberate difference is not exercised: '<k>B' (the reference's second `endswith("n")` test makes its
binary-progression branch unreachable, so it raises; cooler_b200 runs the progression its help text promises)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402


def main(n_cases, seed):
    refshim.import_reference()
    from click.testing import CliRunner
    import cooler.cli.coarsen as ref_coarsen_mod
    import cooler.cli.zoomify as ref_zoom_mod
    from cooler.cli import cli as ref_cli

    import cooler_b200.cli as ours
    from cooler_b200 import MemCooler, store
    calls = {}

    def rec(tag):
        def f(*args, **kw):
            calls[tag] = (args, kw)
            return {} if tag.startswith("ours_zoom") else None
        return f

    ref_zoom_mod.zoomify_cooler = rec("ref_zoom")
    ref_coarsen_mod.coarsen_cooler = rec("ref_coarsen")
    ours.zoomify_cooler = rec("ours_zoom")
    ours.coarsen_cooler = rec("ours_coarsen")
    rng = np.random.default_rng(seed)
    n_runs = 0
    for case in range(n_cases):
        binsize = int(rng.choice([1, 10, 1000, 5000, 40000]))
        nb = [int(rng.integers(1, 3000)) for _ in range(int(rng.integers(1, 4)))]
        names = [f"c{k}" for k in range(len(nb))]
        sizes = np.array([k * binsize for k in nb], dtype=np.int64)
        bc = np.repeat(np.arange(len(nb), dtype=np.int32), nb)
        bs = np.concatenate([np.arange(k) * binsize for k in nb]).astype(np.int64)
        be = bs + binsize
        b1 = np.array([0], dtype=np.int64)
        name = f"fuzzzoom{case}.cool"
        refshim.register(name, names, sizes, bc, bs, be, b1, b1, np.array([1], np.int32), binsize)
        store.REGISTRY.clear()
        store.register(name, MemCooler.from_arrays(names, sizes, bc, bs, be, b1, b1, np.array([1], np.int32),
                                                   binsize=binsize))
        for _ in range(3):
            parts = []
            for _k in range(int(rng.integers(0, 4))):
                r = rng.random()
                if r < 0.4:
                    parts.append(str(binsize * int(rng.choice([2, 4, 5, 10, 50, 100]))))
                elif r < 0.55:
                    parts.append(str(rng.choice(["n", "N", " b", "B "])))
                elif r < 0.7:
                    parts.append(str(rng.choice(["4dn", "4DN"])))
                else:
                    parts.append(f"{binsize * int(rng.choice([1, 2, 5]))}{rng.choice(['n', 'N'])}")
            args = []
            if parts:
                args += ["-r", ",".join(parts)]
            fields = []
            if rng.random() < 0.4:
                fields.append(str(rng.choice(["count", "count:dtype=float", "count:agg=mean", "count:dtype=float32,agg=max"])))
                if rng.random() < 0.4:
                    fields.append(str(rng.choice(["score:agg=min", "score:dtype=int64"])))
            for f in fields:
                args += ["--field", f]
            if rng.random() < 0.5:
                args += ["-o", f"out{case}.mcool"]
            calls.clear()
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                a = CliRunner().invoke(ref_cli, ["zoomify", name, *args])
                b = CliRunner().invoke(ours.cli, ["zoomify", name, *args])
            assert (a.exit_code != 0) == (b.exit_code != 0), (args, a.exit_code, b.exit_code, a.exception, b.exception)
            if a.exit_code == 0:
                (ra, rk), (oa, ok) = calls["ref_zoom"], calls["ours_zoom"]
                assert [int(x) for x in ra[2]] == [int(x) for x in oa[2]], (args, ra[2], oa[2])      # resolutions
                assert int(ra[3]) == int(oa[3])                                                       # chunksize
                for k in ("columns", "dtypes", "agg"):
                    assert rk[k] == ok[k], (args, k, rk[k], ok[k])
            n_runs += 1
            # cooler coarsen
            cargs = ["-k", str(int(rng.integers(2, 9))), "-o", f"c{case}.cool"]
            for f in fields:
                cargs += ["--field", f]
            if rng.random() < 0.3:
                cargs += ["-c", str(int(rng.integers(1, 10**6)))]
            calls.clear()
            a = CliRunner().invoke(ref_cli, ["coarsen", name, *cargs])
            b = CliRunner().invoke(ours.cli, ["coarsen", name, *cargs])
            assert (a.exit_code != 0) == (b.exit_code != 0), (cargs, a.exception, b.exception)
            if a.exit_code == 0:
                (ra, rk), (oa, ok) = calls["ref_coarsen"], calls["ours_coarsen"]
                assert int(ra[2]) == int(oa[2]) and int(rk.get("chunksize", ra[3] if len(ra) > 3 else 0)) == \
                    int(ok.get("chunksize", oa[3] if len(oa) > 3 else 0)), (cargs, ra, oa, rk, ok)
                assert list(rk["columns"]) == list(ok["columns"]), (cargs, rk["columns"], ok["columns"])
                for k in ("dtypes", "agg"):
                    assert (rk[k] or None) == (ok[k] or None), (cargs, k, rk[k], ok[k])
            n_runs += 1
    print(f"ok: {n_runs} zoomify / coarsen command lines handled identically")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
