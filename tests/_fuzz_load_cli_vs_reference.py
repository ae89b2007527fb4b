"""Run in a SUBPROCESS by tests/test_load.py: `cooler load` of cooler_b200 against the LIVE unmodified reference CLI
(src/cooler/cli/load.py) on seeded random COO / BG2 text files and flag sets.  `create_from_unordered` is replaced
by a recorder on both sides; compared are the sanitized chunks it is handed (after a sort: the reference sorts each
chunk, cooler_b200 leaves that to the device), the output columns, their dtypes and the keyword arguments that
steer the checks (symmetric_upper, triucheck, mode, mergebuf)."""
import os
import sys
import tempfile

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402


def main(n_cases, seed):
    refshim.import_reference()
    from click.testing import CliRunner
    import cooler.cli.load as ref_mod
    from cooler.cli import cli as ref_cli

    import cooler_b200.cli as ours_cli
    import cooler_b200.create as ours_create
    calls = {}

    def recorder(tag):
        def f(cool_uri, bins, chunks, columns=None, dtypes=None, **kw):
            got = []
            for ch in chunks:
                df = ch if isinstance(ch, pd.DataFrame) else pd.DataFrame({k: np.asarray(v) for k, v in ch.items()})
                got.append(df)
            calls[tag] = dict(chunks=got, columns=list(columns), dtypes=dict(dtypes or {}), kw=kw, n_bins=len(bins))
        return f

    ref_mod.create_from_unordered = recorder("ref")
    ours_create.create_from_unordered = recorder("ours")          # cli.load imports it from .create at call time
    rng = np.random.default_rng(seed)
    tmp = tempfile.mkdtemp()
    sizes_path = os.path.join(tmp, "sizes")
    n_runs = n_fail = 0
    for case in range(n_cases):
        names = [f"c{k}" for k in range(int(rng.integers(1, 4)))]
        binsize = int(rng.choice([1, 5, 10]))
        lens = [int(rng.integers(binsize, 40 * binsize)) for _ in names]
        with open(sizes_path, "w") as f:
            for nme, ln in zip(names, lens):
                f.write(f"{nme}\t{ln}\n")
        nb = [-(-ln // binsize) for ln in lens]
        off = np.r_[0, np.cumsum(nb)]
        n_bins = int(off[-1])
        fmt = str(rng.choice(["coo", "bg2"]))
        one_based = bool(rng.integers(0, 2))
        m = int(rng.integers(1, 80))
        path = os.path.join(tmp, f"in{case}.txt")
        extra_col = rng.random() < 0.5
        broken = fmt == "bg2" and rng.random() < 0.15            # one anchor beyond its chromosome / negative
        with open(path, "w") as f:
            if rng.random() < 0.3:
                f.write("# a comment line\n")
            for _ in range(m):
                cnt = int(rng.integers(1, 50))
                sc = round(float(rng.random()), 4)
                if fmt == "coo":
                    a, b = (int(x) + one_based for x in rng.integers(0, n_bins, 2))
                    f.write(f"{a}\t{b}\t{cnt}" + (f"\t{sc}" if extra_col else "") + "\n")
                else:
                    pool = names + (["zz"] if rng.random() < 0.1 else [])
                    c1, c2 = str(rng.choice(pool)), str(rng.choice(pool))
                    l1 = lens[names.index(c1)] if c1 in names else 30
                    l2 = lens[names.index(c2)] if c2 in names else 30
                    s1 = int(rng.integers(0, l1)) + one_based
                    s2 = int(rng.integers(0, l2)) + one_based
                    if broken and rng.random() < 0.2:
                        s1 = l1 + 2 + one_based if rng.random() < 0.5 else one_based - 1
                    f.write(f"{c1}\t{s1}\t{s1 + 1}\t{c2}\t{s2}\t{s2 + 1}\t{cnt}" + (f"\t{sc}" if extra_col else "") + "\n")
        args = ["-f", fmt, f"{sizes_path}:{binsize}", path, os.path.join(tmp, "out.cool"),
                "--chunksize", str(int(rng.integers(1, 50)))]
        if one_based:
            args.append("--one-based")
        if rng.random() < 0.3:
            args.append("--no-symmetric-upper")
        if rng.random() < 0.3:
            args += ["--input-copy-status", str(rng.choice(["unique", "duplex"]))]
        if rng.random() < 0.25:
            args.append("--count-as-float")
        col = 3 if fmt == "coo" else 7
        r = rng.random()
        if r < 0.2:
            args += ["--field", f"count={col}:dtype=float"]
        elif r < 0.35:
            args += ["--field", f"count={col}"]
        elif r < 0.45:
            args += ["--field", "count:dtype=float32"]
        if extra_col and rng.random() < 0.6:
            args += ["--field", f"score={col + 1}" + (":dtype=float32" if rng.random() < 0.4 else "")]
        if rng.random() < 0.2:
            args.append("--append")
        if rng.random() < 0.2:
            args += ["--mergebuf", str(int(rng.integers(1, 1000)))]
        calls.clear()
        a = CliRunner().invoke(ref_cli, ["load", *args])
        b = CliRunner().invoke(ours_cli.cli, ["load", *args])
        assert (a.exit_code != 0) == (b.exit_code != 0), (args, a.exit_code, b.exit_code, a.exception, b.exception)
        n_runs += 1
        if a.exit_code != 0:
            n_fail += 1
            continue
        R, O = calls["ref"], calls["ours"]
        assert R["columns"] == O["columns"], (args, R["columns"], O["columns"])
        assert {k: np.dtype(v) for k, v in R["dtypes"].items()} == {k: np.dtype(v) for k, v in O["dtypes"].items()}, \
            (args, R["dtypes"], O["dtypes"])
        for k in ("symmetric_upper", "triucheck", "mode", "mergebuf"):
            assert R["kw"].get(k) == O["kw"].get(k), (args, k, R["kw"].get(k), O["kw"].get(k))
        assert len(R["chunks"]) == len(O["chunks"]), (args, len(R["chunks"]), len(O["chunks"]))
        value_cols = [c for c in R["columns"] if c not in ("bin1_id", "bin2_id")]
        for cr, co in zip(R["chunks"], O["chunks"]):
            keys = ["bin1_id", "bin2_id"] + value_cols
            cr = cr.sort_values(keys).reset_index(drop=True)
            co = co.sort_values(keys).reset_index(drop=True)
            assert len(cr) == len(co), (args, len(cr), len(co))
            for c in keys:
                x, y = cr[c].to_numpy(), co[c].to_numpy()
                assert np.array_equal(x.astype(np.float64), y.astype(np.float64)), (args, c, x[:5], y[:5])
                if c in value_cols:
                    assert x.dtype == y.dtype, (args, c, x.dtype, y.dtype)
    print(f"ok: {n_runs} load command lines agree ({n_fail} of them failing on both sides)")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 150, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
