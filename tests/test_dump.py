"""`cooler dump`: byte-identical text against golden outputs of the unmodified
reference CLI (tests/golden/make_dump_golden.py).  CPU: task planning, joins and formatting with
the numpy oracle standing in for the device range query; GPU: the real thing."""
import hashlib
import io
import json
import os

import numpy as np
import pandas as pd
import pytest

from tests import _golden

with open(os.path.join(_golden.GOLD, "dump_golden.json")) as _f:
    CASES = json.load(_f)["cases"]


def _coolers():
    from cooler_b200 import MemCooler
    t = _golden.gm12878()
    BAL, _ = _golden.balance_golden()
    gm = MemCooler.from_arrays([str(x) for x in t["chromnames"]], t["chromsizes"], t["bins_chrom"], t["bins_start"],
                               t["bins_end"], t["bin1_id"], t["bin2_id"], t["count"], binsize=2_000_000)
    gm.root["bins"].create_dataset("weight", BAL["gm/defaults"].copy())
    a = _golden.coarsen_golden()["toy.asymm/2"]
    c, s, e = _golden.uniform_bins([32, 32], 2)
    toy = MemCooler.from_arrays(["chr1", "chr2"], [32, 32], c, s, e, a[:, 0], a[:, 1], a[:, 2].astype(np.int32),
                                binsize=2, symmetric_upper=False)
    gmf = MemCooler.from_arrays([str(x) for x in t["chromnames"]], t["chromsizes"], t["bins_chrom"], t["bins_start"],
                                t["bins_end"], t["bin1_id"], t["bin2_id"], (t["count"] * 0.5).astype(np.float32),
                                binsize=2_000_000)
    gmf.root["bins"].create_dataset("weight", BAL["gm/defaults"].copy())
    return {"gm": gm, "toy": toy, "gmf": gmf}


def _kwargs(args):
    """The subset of the CLI's option parsing the golden cases use."""
    kw, i = {}, 0
    flags = {"-b": "balanced", "--join": "join", "-H": "header", "-f": "fill_lower",
             "--one-based-ids": "one_based_ids", "--one-based-starts": "one_based_starts"}
    vals = {"-r": "range", "-r2": "range2", "-k": "chunksize", "--float-format": "float_format",
            "--na-rep": "na_rep", "-t": "table", "-c": "columns", "--annotate": "annotate"}
    while i < len(args):
        a = args[i]
        if a in flags:
            kw[flags[a]] = True
            i += 1
        else:
            v = args[i + 1]
            kw[vals[a]] = int(v) if a == "-k" else (tuple(v.split(",")) if a in ("-c", "--annotate") else v)
            i += 2
    return kw


def _compare(text, case):
    assert text.count("\n") == case["lines"], case["args"]
    if "output" in case:
        assert text == case["output"], case["args"]
    assert hashlib.sha1(text.encode()).hexdigest() == case["sha1"], case["args"]


class _OracleSelector:
    """Stands in for api.MatrixSelector(as_pixels=True) on the CPU: same _slice contract."""

    def __init__(self, clr, table, weights, as_pixels=True, **kw):
        self.b1 = np.asarray(clr._load_dset("pixels/bin1_id"))
        self.b2 = np.asarray(clr._load_dset("pixels/bin2_id"))
        self.cnt = np.asarray(clr._load_dset("pixels/count"))
        self.w = weights

    def _slice(self, i0, i1, j0, j1):
        m = (self.b1 >= i0) & (self.b1 < i1) & (self.b2 >= j0) & (self.b2 < j1)
        df = pd.DataFrame({"bin1_id": self.b1[m], "bin2_id": self.b2[m], "count": self.cnt[m]})
        if self.w is not None:
            df["balanced"] = self.w[df["bin1_id"].to_numpy()] * self.w[df["bin2_id"].to_numpy()] * df["count"].to_numpy()
        return df


@pytest.mark.parametrize("k", range(len(CASES)))
def test_dump_plumbing_matches_reference_cli(k, monkeypatch):
    from cooler_b200 import dump as D
    monkeypatch.setattr(D, "MatrixSelector", _OracleSelector)
    case = CASES[k]
    buf = io.StringIO()
    D.dump(_coolers()[case["cooler"]], buf, pixel_table=object(), **_kwargs(case["args"]))
    _compare(buf.getvalue(), case)


def test_plan_tasks_shapes():
    from cooler_b200.dump import plan_tasks
    off = np.arange(0, 1001, 10)                                # 100 rows of 10 pixels
    assert plan_tasks((0, 100, 0, 100), off, 250, False) == [((0, 100, 0, 100), s, False, False)
                                                             for s in [(0, 20), (20, 40), (40, 60), (60, 80), (80, 100)]]   # 2 + 1000 // 250 cuts
    t = plan_tasks((10, 40, 20, 60), off, 10_000, True)         # partial overlap: two stacked boxes
    assert [x[0] for x in t] == [(10, 20, 20, 60), (20, 40, 20, 60)] and all(x[2] and not x[3] for x in t)
    t = plan_tasks((30, 40, 10, 60), off, 10_000, True)         # rows nested in columns: transposed lower block first
    assert [(x[0], x[3]) for x in t] == [((10, 30, 30, 40), True), ((30, 40, 30, 60), False)]
    t = plan_tasks((50, 90, 0, 20), off, 10_000, True)          # entirely below the diagonal: transposed query
    assert [(x[0], x[3]) for x in t] == [((0, 20, 50, 90), True)]
    assert plan_tasks((5, 5, 0, 10), off, 100, True) == []


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(len(CASES)))
def test_dump_gpu_matches_reference_cli(k):
    from cooler_b200 import dump as D
    case = CASES[k]
    buf = io.StringIO()
    D.dump(_coolers()[case["cooler"]], buf, **_kwargs(case["args"]))
    _compare(buf.getvalue(), case)


@pytest.mark.gpu
def test_dump_cli_on_cool_file(tmp_path):
    """The click command on a real .cool file, output to a .gz."""
    import gzip

    from click.testing import CliRunner

    from cooler_b200.cli import cli
    src = os.path.join(_golden.GOLD, "cool", "hg19.GM12878-MboI.matrix.2000kb.cool")
    out = str(tmp_path / "d.tsv.gz")
    res = CliRunner().invoke(cli, ["dump", src, "-r", "chr20", "--one-based-ids", "-o", out])
    assert res.exit_code == 0, (res.output, res.exception)
    case = [c for c in CASES if c["args"] == ["-r", "chr20", "--one-based-ids"]][0]
    with gzip.open(out, "rt") as f:
        _compare(f.read(), case)
    res = CliRunner().invoke(cli, ["dump", src, "-b"])
    assert res.exit_code == 1                                    # "Balancing weights not found"


def test_dump_is_byte_identical_to_the_live_reference_cli_on_random_coolers():
    """240 `cooler dump` runs on seeded random coolers (symmetric-upper / square, NaN weights, missing weights)
    with random argument sets (-r / -r2 / -f / -b / --join / --annotate / one-based shifts / chunk sizes / formats)
    through cooler_b200/dump.py and through the live unmodified reference CLI: same exit status, byte-identical
    text (2 400 more runs with other seeds were made once by hand).  Subprocess; skipped without the reference."""
    import subprocess
    import sys

    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference sources not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tests", "_fuzz_dump_vs_reference.py"), "60", "0"],
                         capture_output=True, text=True, timeout=900, cwd=root,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    assert res.stdout.strip().splitlines()[-1] == "ok: 240 dump runs byte-identical"
