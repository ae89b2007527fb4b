"""Run in a SUBPROCESS by tests/test_host_logic.py: the option handling of `cooler balance` (cooler_b200/cli.py:
--check, --name, mode flags, --ignore-diags / --ignore-dist, bin filters, BED blacklists with and without a
header, --tol / --max-iters, the four convergence policies, --stdout formatting and the exit codes) against the
LIVE unmodified reference CLI (src/cooler/cli/balance.py through oracle/refshim.py) on seeded random coolers.
The numpy oracle stands in for the device balancer; every run uses --stdout or --check, i.e. nothing is written."""
import os
import sys
import tempfile
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ice_numpy, refshim  # noqa: E402


def numbers(text):
    """The weight column a run printed: one value or an empty string per bin; log lines dropped."""
    out = []
    for ln in text.split("\n"):
        t = ln.strip()
        if t == "":
            out.append("")
            continue
        try:
            float(t)
            out.append(t)
        except ValueError:
            pass                                               # a log line
    while out and out[-1] == "":
        out.pop()
    return out


def main(n_cases, seed):
    refshim.import_reference()
    from click.testing import CliRunner
    from cooler.cli import cli as ref_cli

    import cooler_b200.cli as ours_cli
    from cooler_b200 import MemCooler, store

    def oracle_balance_cooler(clr, *, group=None, chunksize=None, use_lock=None, **kw):
        chrom_offset = np.asarray(clr._load_dset("indexes/chrom_offset"))
        bin1_offset = np.asarray(clr._load_dset("indexes/bin1_offset"))
        px = {k: np.asarray(clr._load_dset("pixels/" + k)) for k in ("bin1_id", "bin2_id", "count")}
        return ice_numpy.balance_arrays(px, chrom_offset, bin1_offset, **kw)

    ours_cli.balance_cooler = oracle_balance_cooler
    rng = np.random.default_rng(seed)
    n_runs = n_fail = 0
    tmp = tempfile.mkdtemp()
    for case in range(n_cases):
        binsize = 1000
        names = [f"c{k}" for k in range(int(rng.integers(1, 4)))]
        nb = [int(rng.integers(3, 30)) for _ in names]
        sizes = np.array([k * binsize - int(rng.integers(0, 400)) for k in nb], dtype=np.int64)
        bc = np.repeat(np.arange(len(names), dtype=np.int32), nb)
        bs = np.concatenate([np.arange(k) * binsize for k in nb]).astype(np.int64)
        be = np.minimum(bs + binsize, sizes[bc])
        n = len(bc)
        i, j = np.triu_indices(n)
        b = rng.lognormal(0, 0.4, n)
        b[rng.random(n) < 0.08] *= 0.02
        cnt = rng.poisson(40 * b[i] * b[j] / (1.0 + np.abs(i - j)) ** 0.6) * (rng.random(len(i)) < 0.7)
        sel = cnt > 0
        b1, b2, cnt = i[sel].astype(np.int64), j[sel].astype(np.int64), cnt[sel].astype(np.int32)
        name = f"fuzzbal{case}.cool"
        g = refshim.register(name, names, sizes, bc, bs, be, b1, b2, cnt, binsize)
        ours = MemCooler.from_arrays(names, sizes, bc, bs, be, b1, b2, cnt, binsize=binsize)
        if rng.random() < 0.3:                                  # an existing weight column (for --check / --name)
            g["bins"]["weight"] = np.ones(n)
            ours.root["bins"].create_dataset("weight", np.ones(n))
        store.REGISTRY.clear()
        store.register(name, ours)
        for _ in range(3):
            args = []
            if rng.random() < 0.15:
                args.append("--check")
            else:
                args.append("--stdout")
            if rng.random() < 0.2:
                args += ["--name", str(rng.choice(["weight", "w2"]))]
            r = rng.random()
            if r < 0.25:
                args.append("--cis-only")
            elif r < 0.45:
                args.append("--trans-only")
            elif r < 0.5:
                args += ["--cis-only", "--trans-only"]
            if rng.random() < 0.5:
                args += ["--ignore-diags", str(int(rng.integers(0, 4)))]
            if rng.random() < 0.3:
                args += ["--ignore-dist", str(int(rng.integers(0, 4000)))]
            if rng.random() < 0.5:
                args += ["--mad-max", str(int(rng.integers(0, 6)))]
            if rng.random() < 0.5:
                args += ["--min-nnz", str(int(rng.integers(0, 8)))]
            if rng.random() < 0.3:
                args += ["--min-count", str(int(rng.integers(0, 60)))]
            if rng.random() < 0.4:
                args += ["--tol", f"{10 ** rng.uniform(-7, -2):.3g}"]
            if rng.random() < 0.5:
                args += ["--max-iters", str(int(rng.integers(1, 40)))]
            if rng.random() < 0.6:
                args += ["--convergence-policy", str(rng.choice(["store_final", "store_nan", "discard", "error"]))]
            if rng.random() < 0.35:
                path = os.path.join(tmp, f"bl{case}_{n_runs}.bed")
                with open(path, "w") as f:
                    if rng.random() < 0.4:
                        f.write("chrom\tstart\tend\n")
                    for _k in range(int(rng.integers(1, 4))):
                        c = int(rng.integers(0, len(names)))
                        hi = int(sizes[c]) + (500 if rng.random() < 0.1 else 0)      # sometimes out of bounds
                        a, bb = sorted(int(x) for x in rng.integers(0, hi + 1, 2))
                        f.write(f"{names[c] if rng.random() < 0.95 else 'nope'}\t{a}\t{bb}\n")
                args += ["--blacklist", path]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                np.seterr(all="ignore")
                want = CliRunner().invoke(ref_cli, ["balance", name, *args])
                got = CliRunner().invoke(ours_cli.cli, ["balance", name, *args])
            assert want.exit_code == got.exit_code, (case, args, want.exit_code, got.exit_code, want.exception, got.exception)
            if want.exit_code == 0 and "--stdout" in args:
                assert numbers(got.output) == numbers(want.output), (case, args, got.output[-400:], want.output[-400:])
            if "--check" in args:
                assert ("is balanced" in got.output) == ("is balanced" in want.output), (case, args)
            if want.exit_code != 0:
                n_fail += 1
            n_runs += 1
    print(f"ok: {n_runs} balance CLI runs agree ({n_fail} of them failing the same way)")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 60, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
