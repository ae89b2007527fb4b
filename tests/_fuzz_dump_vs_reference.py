"""Run in a SUBPROCESS by tests/test_dump.py: `cooler dump` of cooler_b200 (task planning, lower-triangle fill,
joins, annotation, one-based shifts, text formatting -- cooler_b200/dump.py, with the numpy selector of
tests/test_dump.py standing in for the device range query) against the LIVE unmodified reference CLI
(src/cooler/cli/dump.py through oracle/refshim.py) on seeded random coolers and argument sets: byte-identical."""
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402


def main(n_cases, seed):
    refshim.import_reference()
    from click.testing import CliRunner
    from cooler.cli import cli as ref_cli

    from cooler_b200 import MemCooler
    from cooler_b200 import dump as D
    from tests.test_dump import _kwargs, _OracleSelector
    D.MatrixSelector = _OracleSelector
    rng = np.random.default_rng(seed)
    n_runs = 0
    for case in range(n_cases):
        binsize = int(rng.choice([10, 100, 2000]))
        names = [f"c{k}" for k in range(int(rng.integers(1, 4)))]
        nb = [int(rng.integers(1, 25)) for _ in names]
        sizes = np.array([k * binsize - int(rng.integers(0, binsize // 2)) for k in nb], dtype=np.int64)
        bc = np.repeat(np.arange(len(names), dtype=np.int32), nb)
        bs = np.concatenate([np.arange(k) * binsize for k in nb]).astype(np.int64)
        be = np.minimum(bs + binsize, sizes[bc])
        n = len(bc)
        square = rng.random() < 0.3
        i, j = (np.indices((n, n)).reshape(2, -1) if square else np.triu_indices(n))
        sel = rng.random(len(i)) < rng.uniform(0.1, 0.7)
        b1, b2 = i[sel].astype(np.int64), j[sel].astype(np.int64)
        cnt = rng.integers(1, 500, len(b1)).astype(np.int32)
        w = rng.lognormal(0, 0.5, n)
        w[rng.random(n) < 0.1] = np.nan
        name = f"fuzzdump{case}.cool"
        g = refshim.register(name, names, sizes, bc, bs, be, b1, b2, cnt, binsize, symmetric_upper=not square)
        has_w = rng.random() < 0.8
        if has_w:
            g["bins"]["weight"] = w.copy()
        ours = MemCooler.from_arrays(names, sizes, bc, bs, be, b1, b2, cnt, binsize=binsize, symmetric_upper=not square)
        if has_w:
            ours.root["bins"].create_dataset("weight", w.copy())

        def region():
            c = int(rng.integers(0, len(names)))
            if rng.random() < 0.3:
                return names[c]
            a, b = sorted(int(x) for x in rng.integers(0, sizes[c] + 1, 2))
            return f"{names[c]}:{a}-{b}"
        for _ in range(4):
            args = []
            if rng.random() < 0.8:
                args += ["-r", region()]
                if rng.random() < 0.6:
                    args += ["-r2", region()]
            for flag, p in (("-f", 0.5), ("-b", 0.5 if has_w else 0.1), ("--join", 0.4), ("-H", 0.4),
                            ("--one-based-ids", 0.3), ("--one-based-starts", 0.3)):
                if rng.random() < p:
                    args.append(flag)
            if rng.random() < 0.3 and has_w:
                args += ["--annotate", str(rng.choice(["weight", "weight,start", "end"]))]
            if rng.random() < 0.5:
                args += ["-k", str(int(rng.integers(1, 400)))]
            if rng.random() < 0.3:
                args += ["--na-rep", "NA"]
            if rng.random() < 0.3:
                args += ["--float-format", str(rng.choice([".10g", ".3f", "g"]))]
            res = CliRunner().invoke(ref_cli, ["dump", name, *args])
            buf = io.StringIO()
            try:
                D.dump(ours, buf, pixel_table=object(), **_kwargs(args))
                code = 0
            except SystemExit as e:
                code = int(e.code or 0)
            assert (res.exit_code != 0) == (code != 0), (case, args, res.exit_code, code, res.exception)
            if code == 0:
                assert buf.getvalue() == res.output, (case, args, square, buf.getvalue()[:300], res.output[:300])
            n_runs += 1
    print(f"ok: {n_runs} dump runs byte-identical")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 60, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
