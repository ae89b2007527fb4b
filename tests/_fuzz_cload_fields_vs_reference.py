"""Run in a SUBPROCESS by tests/test_load.py: `cooler cload pairs --field ...` of cooler_b200 against the LIVE
unmodified reference CLI (src/cooler/cli/cload.py:484-666) on seeded random pairs files.  The writers
(`create_cooler` in the reference, `create_from_unordered` here) are replaced by recorders and a numpy accumulator
stands in for the device one; compared are the per-chunk pixel tables (the reference's
``aggregate_records(sanitize_records(chunk))``), the output columns, dtypes and the symmetry flag."""
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
    import cooler.cli.cload as ref_mod
    from cooler.cli import cli as ref_cli

    import cooler_b200.cli as ours_cli
    import cooler_b200.create as ours_create
    from cooler_b200 import ingest
    from tests.test_load import _NumpyAccumulator
    ingest.UnorderedPixels = _NumpyAccumulator
    calls = {}

    def recorder(tag):
        def f(cool_uri, bins, chunks, columns=None, dtypes=None, **kw):
            got = [ch if isinstance(ch, pd.DataFrame) else pd.DataFrame({k: np.asarray(v) for k, v in ch.items()})
                   for ch in chunks]
            calls[tag] = dict(chunks=got, columns=list(columns), dtypes=dict(dtypes or {}), kw=kw)
        return f

    ref_mod.create_cooler = recorder("ref")
    ours_create.create_from_unordered = recorder("ours")

    def ours_plain(cool_uri, bins, pixels, columns=None, **kw):          # the count-only path writes directly
        calls["ours_plain"] = dict(px={k: np.asarray(v) for k, v in pixels.items()}, columns=list(columns), kw=kw)

    ours_create.create_cooler = ours_plain
    from oracle import ingest_numpy

    def oracle_bin_pairs(*a, device="cuda", **kw):
        try:
            return ingest_numpy.bin_pairs(*a, **kw)
        except ingest_numpy.BadInputError as e:
            raise ingest.BadInputError(str(e)) from e

    ingest.bin_pairs = oracle_bin_pairs
    rng = np.random.default_rng(seed)
    tmp = tempfile.mkdtemp()
    sizes_path = os.path.join(tmp, "sizes")
    n_runs = n_fail = 0
    for case in range(n_cases):
        names = [f"c{k}" for k in range(int(rng.integers(1, 4)))]
        binsize = int(rng.choice([1, 4, 10]))
        lens = [int(rng.integers(binsize, 30 * binsize)) for _ in names]
        with open(sizes_path, "w") as f:
            for nme, ln in zip(names, lens):
                f.write(f"{nme}\t{ln}\n")
        zero_based = bool(rng.integers(0, 2))
        path = os.path.join(tmp, f"p{case}.pairs")
        m = int(rng.integers(1, 120))
        broken = rng.random() < 0.1
        with open(path, "w") as f:
            for k in range(m):
                pool = names + (["zz"] if rng.random() < 0.1 else [])
                c1, c2 = str(rng.choice(pool)), str(rng.choice(pool))
                l1 = lens[names.index(c1)] if c1 in names else 20
                l2 = lens[names.index(c2)] if c2 in names else 20
                p1 = int(rng.integers(0, l1)) + (not zero_based)
                p2 = int(rng.integers(0, l2)) + (not zero_based)
                if broken and rng.random() < 0.1:
                    p1 = l1 + 3
                f.write(f"r{k}\t{c1}\t{p1}\t{c2}\t{p2}\t+\t-\t{round(float(rng.random()), 3)}\t{int(rng.integers(0, 60))}\n")
        args = [f"{sizes_path}:{binsize}", path, os.path.join(tmp, "o.cool"), "-c1", "2", "-p1", "3", "-c2", "4", "-p2", "5",
                "--chunksize", str(int(rng.integers(1, 60)))]
        if zero_based:
            args.append("--zero-based")
        if rng.random() < 0.25:
            args.append("--no-symmetric-upper")
        if rng.random() < 0.25:
            args += ["--input-copy-status", str(rng.choice(["unique", "duplex"]))]
        ops = ["sum", "min", "max", "first", "last", "mean"]
        r = rng.random()
        if r < 0.5:
            spec = "score=8"
        elif r < 0.7:
            spec = "count=9"
        else:
            spec = "mapq=9"
        props = []
        if rng.random() < 0.5:
            props.append(f"agg={rng.choice(ops)}")
        if rng.random() < 0.4:
            props.append(f"dtype={rng.choice(['float', 'float32', 'int64'])}")
        # field numbers ascending in the order given: pandas hands `names` to `usecols` in FILE order, so the
        # reference swaps the columns of `--field count=9 --field score=8` (cooler_b200 reads each field from the
        # column it names)
        if rng.random() < 0.3 and not spec.startswith("score"):
            args += ["--field", "score=8" + (f":agg={rng.choice(ops)}" if rng.random() < 0.5 else "")]
        args += ["--field", spec + (":" + ",".join(props) if props else "")]
        if rng.random() < 0.15:
            args += ["--field", "count:dtype=int64"] if not any(a.startswith("count=") for a in args) else []
        # ---- first without --field: one device-wide binning here, per-chunk tables merged by summing there ------
        plain = [x for x in args[:args.index("--field")]]
        if binsize > 0:
            calls.clear()
            a = CliRunner().invoke(ref_cli, ["cload", "pairs", *plain])
            b = CliRunner().invoke(ours_cli.cli, ["cload", "pairs", *plain])
            assert (a.exit_code != 0) == (b.exit_code != 0), (plain, a.exit_code, b.exit_code, a.exception, b.exception)
            n_runs += 1
            if a.exit_code == 0:
                R, O = calls["ref"], calls["ours_plain"]
                assert R["columns"] == O["columns"] == ["count"], (R["columns"], O["columns"])
                assert R["kw"].get("symmetric_upper") == O["kw"].get("symmetric_upper"), plain
                cat = pd.concat(R["chunks"], ignore_index=True)
                if len(cat):
                    want = cat.groupby(["bin1_id", "bin2_id"], sort=True)["count"].sum().reset_index()
                    for c in ("bin1_id", "bin2_id", "count"):
                        assert np.array_equal(want[c].to_numpy().astype(np.int64), O["px"][c].astype(np.int64)), (plain, c)
                else:
                    assert len(O["px"]["count"]) == 0
            else:
                n_fail += 1
        calls.clear()
        a = CliRunner().invoke(ref_cli, ["cload", "pairs", *args])
        b = CliRunner().invoke(ours_cli.cli, ["cload", "pairs", *args])
        assert (a.exit_code != 0) == (b.exit_code != 0), (args, a.exit_code, b.exit_code, a.exception, b.exception)
        n_runs += 1
        if a.exit_code != 0:
            n_fail += 1
            continue
        R, O = calls["ref"], calls["ours"]
        assert R["columns"] == O["columns"], (args, R["columns"], O["columns"])
        assert {k: np.dtype(v) for k, v in R["dtypes"].items()} == {k: np.dtype(v) for k, v in O["dtypes"].items()}, \
            (args, R["dtypes"], O["dtypes"])
        assert R["kw"].get("symmetric_upper") == O["kw"].get("symmetric_upper"), args
        assert len(R["chunks"]) == len(O["chunks"]), (args, len(R["chunks"]), len(O["chunks"]))
        for cr, co in zip(R["chunks"], O["chunks"]):
            cr = cr.sort_values(["bin1_id", "bin2_id"]).reset_index(drop=True)
            co = co.sort_values(["bin1_id", "bin2_id"]).reset_index(drop=True)
            assert len(cr) == len(co), (args, len(cr), len(co))
            for c in R["columns"] + ["bin1_id", "bin2_id"]:
                x, y = cr[c].to_numpy(), co[c].to_numpy()
                np.testing.assert_allclose(x.astype(np.float64), y.astype(np.float64), err_msg=f"{args} {c}",
                                           rtol=1e-6 if x.dtype == np.float32 else 1e-12)
                if c in R["columns"] and len(x):             # (an empty reference chunk has float64 / object columns)
                    assert x.dtype == y.dtype, (args, c, x.dtype, y.dtype)
    print(f"ok: {n_runs} cload pairs command lines (with and without --field) agree ({n_fail} of them failing on both sides)")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 150, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
