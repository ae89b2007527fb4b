"""Golden outputs of the UNMODIFIED reference's ``cooler dump`` (src/cooler/cli/dump.py) run through
oracle/refshim.py in the build container.

    python tests/golden/make_dump_golden.py     ->  tests/golden/dump_golden.json
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402
from tests import _golden  # noqa: E402

CASES = [   # (cooler, args)
    ("gm", ["-r", "chr21", "-b", "--join", "-H"]),
    ("gm", ["-r", "chr21:10,000,000-30,000,000", "-r2", "chr20", "-f", "-b"]),
    ("gm", ["-r", "chr21:10,000,000-30,000,000", "-r2", "chr20", "-b"]),
    ("gm", ["-r", "chr2:0-60,000,000", "-r2", "chr2:30,000,000-120,000,000", "-f", "-k", "300"]),
    ("gm", ["-r", "chr2:30,000,000-60,000,000", "-r2", "chr2:0-120,000,000", "-f", "-b", "-k", "200"]),
    ("gm", ["-r", "chr2:0-120,000,000", "-r2", "chr2:30,000,000-60,000,000", "-f"]),
    ("gm", ["-r", "chr2:30,000,000-120,000,000", "-r2", "chr2:0-60,000,000", "-f", "-H", "-k", "150"]),
    ("gm", ["-r", "chr1", "-f", "--join", "--one-based-starts", "--one-based-ids", "-k", "1000"]),
    ("gm", ["-r", "chr20", "--one-based-ids"]),
    ("gm", ["-r", "chrX", "--annotate", "weight", "-b", "--float-format", ".10g", "--na-rep", "NA"]),
    ("gm", ["-r", "chr19", "--annotate", "weight,start", "--join", "-H", "--one-based-starts"]),
    ("gm", ["-r", "chrM"]),
    ("gm", ["-r", "chrY", "-H", "-b"]),
    ("gm", ["-k", "5000"]),
    ("gm", ["-b", "-f", "-k", "20000"]),
    ("gm", ["-t", "bins", "-H"]),
    ("gm", ["-t", "bins", "-c", "chrom,start"]),
    ("gm", ["-t", "chroms"]),
    ("gm", ["-t", "chroms", "-H", "-c", "name"]),
    ("toy", ["-f", "-H"]),
    ("toy", ["-r", "chr1", "-r2", "chr2", "--join"]),
    ("toy", ["-r", "chr2:8-30", "-r2", "chr1:4-20"]),
    # a float32 count column (count * 0.5): values printed with the float format, balanced products in float64
    ("gmf", ["-r", "chr21", "-b", "-H"]),
    ("gmf", ["-r", "chr2:0-60,000,000", "-r2", "chr2:30,000,000-120,000,000", "-f", "-k", "300"]),
    ("gmf", ["-r", "chr20", "--join", "--float-format", ".10g"]),
]


def main():
    refshim.import_reference()
    from click.testing import CliRunner
    from cooler.cli import cli
    t = _golden.gm12878()
    BAL, _ = _golden.balance_golden()
    g = refshim.register("gm_dump.cool", [str(x) for x in t["chromnames"]], t["chromsizes"], t["bins_chrom"],
                         t["bins_start"], t["bins_end"], t["bin1_id"], t["bin2_id"], t["count"], 2_000_000)
    g["bins"]["weight"] = BAL["gm/defaults"].copy()
    g = refshim.register("gmf_dump.cool", [str(x) for x in t["chromnames"]], t["chromsizes"], t["bins_chrom"],
                         t["bins_start"], t["bins_end"], t["bin1_id"], t["bin2_id"],
                         (t["count"] * 0.5).astype("float32"), 2_000_000)
    g["bins"]["weight"] = BAL["gm/defaults"].copy()
    CO = _golden.coarsen_golden()
    a = CO["toy.asymm/2"]
    c, s, e = _golden.uniform_bins([32, 32], 2)
    refshim.register("toy_dump.cool", ["chr1", "chr2"], [32, 32], c, s, e, a[:, 0], a[:, 1], a[:, 2].astype("int32"), 2,
                     symmetric_upper=False)
    out = []
    for which, args in CASES:
        name = {"gm": "gm_dump.cool", "gmf": "gmf_dump.cool", "toy": "toy_dump.cool"}[which]
        res = CliRunner().invoke(cli, ["dump", name, *args])
        assert res.exit_code == 0, (args, res.exception)
        text = res.output
        rec = {"cooler": which, "args": args, "lines": text.count("\n"),
               "sha1": hashlib.sha1(text.encode()).hexdigest()}
        if len(text) <= 40_000:
            rec["output"] = text
        out.append(rec)
        print(which, args, rec["lines"], rec["sha1"][:12])
    with open(os.path.join(ROOT, "tests", "golden", "dump_golden.json"), "w") as f:
        json.dump({"cases": out}, f, indent=0)


if __name__ == "__main__":
    main()
