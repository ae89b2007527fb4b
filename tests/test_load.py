"""`cooler load` / create_from_unordered (COO and BG2 text in any order -> cooler):
the host sanitizers against the chunks of the unmodified reference's own sanitizers, and the whole command
against the reference's merged tables and its real .cool fixtures (tests/golden/make_load_golden.py).
CPU: a numpy stand-in replaces the device accumulator; GPU: the real thing."""
import os

import numpy as np
import pandas as pd
import pytest

from tests import _golden, _load_cases as LC

G = _golden._npz("load_golden.npz")
COOL = os.path.join(_golden.GOLD, "cool")


def _args(key, out):
    fmt, bins_spec, pixels, extra = LC.CASES[key]
    return ["load", "-f", fmt, os.path.join(LC.TEXT, bins_spec), os.path.join(LC.TEXT, pixels), out, *extra]


def _value_columns(key):
    return [k.split("/")[2] for k in G if k.startswith(key + "/merged/") and not k.endswith("_id")]


@pytest.mark.parametrize("key", sorted(LC.CASES))
def test_sanitizers_match_reference_chunks(key):
    """ingest.sanitize_pixels / sanitize_bg2 chunk by chunk against the reference's sanitize_pixels /
    sanitize_records(schema='bg2') (which also sort each chunk: compared after a sort)."""
    from cooler_b200 import ingest
    from cooler_b200.cli import parse_bins
    from cooler_b200.create import _bins_arrays
    fmt, bins_spec, pixels, extra = LC.CASES[key]
    chunksize = int(extra[extra.index("--chunksize") + 1])
    tril = None if "--no-symmetric-upper" in extra else ("drop" if "duplex" in extra else "reflect")
    _, bins = parse_bins(os.path.join(LC.TEXT, bins_spec))
    names, ids, start, end, _ = _bins_arrays(bins)
    lengths = np.asarray(end)[np.r_[ids[1:] != ids[:-1], True]]
    cols = {"bg2": ["chrom1", "start1", "end1", "chrom2", "start2", "end2", "count"],
            "coo": ["bin1_id", "bin2_id", "count"]}[fmt]
    if key == "coo_shuffled":
        cols.append("score")
    raw = pd.read_csv(os.path.join(LC.TEXT, pixels), sep="\t", header=None, names=cols,
                      dtype={"chrom1": str, "chrom2": str})
    lens = G[key + "/chunk_lengths"]
    o = 0
    for k, lo in enumerate(range(0, len(raw), chunksize)):
        chunk = {c: raw[c].to_numpy()[lo:lo + chunksize] for c in cols}
        if fmt == "bg2":
            got = ingest.sanitize_bg2(chunk, names, lengths, ids, start, end, is_one_based="--one-based" in extra,
                                      tril_action=tril)
        else:
            got = ingest.sanitize_pixels(chunk, is_one_based="--one-based" in extra, tril_action=tril)
        order = np.lexsort((got["bin2_id"], got["bin1_id"]))
        m = int(lens[k])
        assert len(order) == m
        for c in ["bin1_id", "bin2_id"] + _value_columns(key):
            np.testing.assert_array_equal(np.asarray(got[c])[order].astype(np.float64),
                                          G[f"{key}/chunks/{c}"][o:o + m].astype(np.float64), err_msg=c)
        o += m
    assert o == lens.sum()


class _NumpyAccumulator:
    """Stands in for ingest.UnorderedPixels on the CPU (same contract)."""

    def __init__(self, n_bins, device="cuda"):
        self.b1, self.b2, self.n_bins = [], [], n_bins

    def add(self, bin1, bin2, dupcheck=True):
        from cooler_b200.ingest import BadInputError
        if len(bin1) and (min(np.min(bin1), np.min(bin2)) < 0 or max(np.max(bin1), np.max(bin2)) >= self.n_bins):
            raise BadInputError("Found a bin ID that exceeds the declared number of bins.")
        if dupcheck and len(np.unique(np.asarray(bin1) * (1 << 32) + np.asarray(bin2))) != len(bin1):
            raise BadInputError("Found duplicate pixels")
        self.b1.append(np.asarray(bin1)), self.b2.append(np.asarray(bin2))

    def merge(self, columns, agg=None):
        from oracle import coarsen_numpy
        t = dict(bin1_id=np.concatenate(self.b1), bin2_id=np.concatenate(self.b2), **columns)
        out = coarsen_numpy.merge_columns([t], list(columns), agg)
        return out.pop("bin1_id"), out.pop("bin2_id"), out


def _check_output(key, out):
    from cooler_b200 import store
    from tests.test_h5write import check
    check(out)
    clr = store.open_cooler(out)
    for c in ["bin1_id", "bin2_id"] + _value_columns(key):
        got, want = np.asarray(clr._load_dset("pixels/" + c)), G[f"{key}/merged/{c}"]
        assert got.dtype == want.dtype, (c, got.dtype, want.dtype)
        if want.dtype.kind == "f":
            np.testing.assert_allclose(got, want, rtol=1e-12, err_msg=c)
        else:
            np.testing.assert_array_equal(got, want, err_msg=c)
    assert clr.info["nnz"] == len(G[key + "/merged/bin1_id"])
    assert clr.storage_mode == ("square" if "--no-symmetric-upper" in LC.CASES[key][3] else "symmetric-upper")
    # the reference's own .cool files made from the same text (tests/test_cli_ingest.py:201-208)
    ref = {"bg2_symm": os.path.join(COOL, "toy.symm.upper.2.mcool") + "::resolutions/2",
           "bg2_symm_ob": os.path.join(COOL, "toy.symm.upper.2.mcool") + "::resolutions/2",
           "bg2_asymm": os.path.join(COOL, "toy.asymm.2.cool")}.get(key)
    if ref is not None:
        r = store.open_cooler(ref)
        for c in ("bin1_id", "bin2_id", "count"):
            np.testing.assert_array_equal(np.asarray(clr._load_dset("pixels/" + c)), np.asarray(r._load_dset("pixels/" + c)))
        np.testing.assert_array_equal(np.asarray(clr._load_dset("indexes/bin1_offset")),
                                      np.asarray(r._load_dset("indexes/bin1_offset")))
        assert clr.chromnames == r.chromnames and clr.binsize == r.binsize


@pytest.mark.parametrize("key", sorted(LC.CASES))
def test_load_cli_plumbing_with_numpy_accumulator(key, tmp_path, monkeypatch):
    from click.testing import CliRunner

    from cooler_b200 import ingest
    from cooler_b200.cli import cli
    monkeypatch.setattr(ingest, "UnorderedPixels", _NumpyAccumulator)
    out = str(tmp_path / "out.cool")
    res = CliRunner().invoke(cli, _args(key, out))
    assert res.exit_code == 0, (res.output, res.exception)
    _check_output(key, out)


def _pairs_args(key, out):
    binsize, extra = LC.PAIRS_CASES[key]
    return ["cload", "pairs", os.path.join(LC.TEXT, f"toy.chrom.sizes:{binsize}"), os.path.join(LC.TEXT, "toy.pairs"),
            out, *LC.PAIRS_BASE, *extra]


def _check_pairs_output(key, out):
    from cooler_b200 import store
    from tests.test_h5write import check
    check(out)
    clr = store.open_cooler(out)
    cols = [k.split("/")[2] for k in G if k.startswith(key + "/merged/")]
    assert sorted(clr.pixels().dtypes.keys()) == sorted(cols)
    for c in cols:
        got, want = np.asarray(clr._load_dset("pixels/" + c)), G[f"{key}/merged/{c}"]
        assert got.dtype == want.dtype, (c, got.dtype, want.dtype)
        if want.dtype.kind == "f":
            np.testing.assert_allclose(got, want, rtol=1e-6 if want.dtype == np.float32 else 1e-12, err_msg=c)
        else:
            np.testing.assert_array_equal(got, want, err_msg=c)


@pytest.mark.parametrize("key", sorted(LC.PAIRS_CASES))
def test_cload_pairs_fields_plumbing_with_numpy_accumulator(key, tmp_path, monkeypatch):
    """`cooler cload pairs --field ...` (reference tests/test_cli_ingest.py:95-135): value columns aggregated
    per chunk with the requested function, chunk tables summed -- against the reference's own
    sanitize_records + aggregate_records run chunk by chunk (make_load_golden.py)."""
    from click.testing import CliRunner

    from cooler_b200 import ingest
    from cooler_b200.cli import cli
    monkeypatch.setattr(ingest, "UnorderedPixels", _NumpyAccumulator)
    out = str(tmp_path / "out.cool")
    res = CliRunner().invoke(cli, _pairs_args(key, out))
    assert res.exit_code == 0, (res.output, res.exception)
    _check_pairs_output(key, out)


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(LC.PAIRS_CASES))
def test_cload_pairs_fields_gpu(key, tmp_path):
    from click.testing import CliRunner

    from cooler_b200.cli import cli
    out = str(tmp_path / "out.cool")
    res = CliRunner().invoke(cli, _pairs_args(key, out))
    assert res.exit_code == 0, (res.output, res.exception)
    _check_pairs_output(key, out)


def _bad_inputs(tmp_path):
    dup = tmp_path / "dup.coo"
    dup.write_text("0\t1\t5\n3\t4\t1\n1\t0\t2\n")               # (0, 1) twice in one chunk after mirroring
    oob = tmp_path / "oob.coo"
    oob.write_text("0\t1\t5\n3\t64\t1\n")
    low = tmp_path / "low.coo"
    low.write_text("5\t1\t5\n")
    return str(dup), str(oob), str(low)


def _run_errors(tmp_path):
    from click.testing import CliRunner

    from cooler_b200.cli import cli
    from cooler_b200.create import BadInputError as CreateError
    from cooler_b200.ingest import BadInputError
    dup, oob, low = _bad_inputs(tmp_path)
    bins = os.path.join(LC.TEXT, "toy.chrom.sizes:1")
    out = str(tmp_path / "bad.cool")
    res = CliRunner().invoke(cli, ["load", "-f", "coo", bins, dup, out])
    assert res.exit_code != 0 and isinstance(res.exception, BadInputError) and "duplicate" in str(res.exception)
    res = CliRunner().invoke(cli, ["load", "-f", "coo", bins, dup, out, "--chunksize", "2"])
    assert res.exit_code == 0, res.exception                     # the two records sit in different chunks: summed
    res = CliRunner().invoke(cli, ["load", "-f", "coo", bins, oob, out])
    assert res.exit_code != 0 and isinstance(res.exception, CreateError) and "exceeds" in str(res.exception)
    # toy.pairs is one-based: read as zero-based its last position equals the chromosome length, which passes the
    # anchor check (`> chromsize`, create/_ingest.py:142) but lands on bin n_bins -- refused like in bin_pairs
    res = CliRunner().invoke(cli, _pairs_args("pairs_max_binsize8", out) + ["--zero-based"])
    assert res.exit_code != 0 and isinstance(res.exception, BadInputError) and "exceeds" in str(res.exception)
    res = CliRunner().invoke(cli, ["load", "-f", "coo", bins, low, out, "-N"])
    assert res.exit_code == 0, res.exception                     # square storage keeps the lower pixel
    return out


def test_load_cli_errors_with_numpy_accumulator(tmp_path, monkeypatch):
    from cooler_b200 import ingest, store
    monkeypatch.setattr(ingest, "UnorderedPixels", _NumpyAccumulator)
    out = _run_errors(tmp_path)
    clr = store.open_cooler(out)
    assert clr.storage_mode == "square" and list(np.asarray(clr._load_dset("pixels/bin1_id"))) == [5]


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(LC.CASES))
def test_load_cli_gpu(key, tmp_path):
    from click.testing import CliRunner

    from cooler_b200.cli import cli
    out = str(tmp_path / "out.cool")
    res = CliRunner().invoke(cli, _args(key, out))
    assert res.exit_code == 0, (res.output, res.exception)
    _check_output(key, out)


@pytest.mark.gpu
def test_load_cli_errors_gpu(tmp_path):
    _run_errors(tmp_path)


@pytest.mark.gpu
def test_unordered_pixels_matches_oracle_large():
    """The device accumulator on 2e6 random records in 7 chunks (duplicates only ACROSS chunks) with an int32
    and a float64 column against the numpy restatement of the merge."""
    from cooler_b200.ingest import BadInputError, UnorderedPixels
    from oracle import coarsen_numpy
    rng = np.random.default_rng(9)
    n_bins, tables = 50_000, []
    acc = UnorderedPixels(n_bins)
    for k in range(7):
        key = np.unique(rng.integers(0, n_bins * 40, 300_000))           # unique inside the chunk
        b1 = key // 40
        b2 = np.minimum(b1 + (key % 40) * (1 + (key % 7)), n_bins - 1)
        u = np.unique(b1 * n_bins + b2, return_index=True)[1]
        b1, b2 = b1[u], b2[u]
        p = rng.permutation(len(b1))
        b1, b2 = b1[p], b2[p]
        tables.append(dict(bin1_id=b1, bin2_id=b2, count=rng.integers(1, 100, len(b1)).astype(np.int32),
                           score=rng.random(len(b1))))
        acc.add(b1, b2)
    cols = {c: np.concatenate([t[c] for t in tables]) for c in ("count", "score")}
    g1, g2, vals = acc.merge(cols)
    want = coarsen_numpy.merge_columns(tables, ["count", "score"])
    np.testing.assert_array_equal(g1, want["bin1_id"]), np.testing.assert_array_equal(g2, want["bin2_id"])
    np.testing.assert_array_equal(vals["count"], want["count"])
    assert vals["count"].dtype == np.int32 and vals["score"].dtype == np.float64
    np.testing.assert_allclose(vals["score"], want["score"], rtol=1e-12)
    with pytest.raises(BadInputError, match="duplicate"):
        UnorderedPixels(n_bins).add(np.array([3, 9, 3]), np.array([7, 9, 7]))


def test_sanitizers_agree_with_the_live_reference_on_random_chunks():
    """~1 050 seeded random cases (fixed and variable bins, every tril_action, unknown chromosomes, anchors at and
    beyond the chromosome ends) through ingest.sanitize_records / sanitize_pixels and through the unmodified
    reference's functions -- and the pairs-binning oracle (oracle/ingest_numpy.py) against the reference's
    sanitize_records + aggregate_records: same tables or same errors.  Runs in a subprocess (the reference needs stub modules);
    skipped where no copy of the reference exists."""
    import subprocess
    import sys

    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference sources not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tests", "_fuzz_sanitize_vs_reference.py"), "300", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    assert res.stdout.strip().startswith("ok:")


def test_load_cli_agrees_with_the_live_reference_cli_on_random_text_files():
    """150 `cooler load` command lines on seeded random COO / BG2 files (records in any order, mirrored records,
    unknown chromosomes, comment lines, anchors outside their chromosome; --one-based, -N, --input-copy-status,
    --count-as-float, --field variants, --append, --mergebuf) through cooler_b200/cli.py and through the live
    unmodified reference CLI, `create_from_unordered` replaced by a recorder on both sides: same exit status, same
    sanitized chunks, output columns, dtypes and check flags (1 500 more command lines with other seeds were compared
    once by hand).  Subprocess; skipped without the reference."""
    import subprocess
    import sys

    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference sources not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tests", "_fuzz_load_cli_vs_reference.py"), "150", "0"],
                         capture_output=True, text=True, timeout=900, cwd=root,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    assert res.stdout.strip().splitlines()[-1].startswith("ok: 150 load command lines agree")


def test_cload_pairs_fields_agree_with_the_live_reference_cli_on_random_pairs_files():
    """150 `cooler cload pairs --field ...` command lines on seeded random pairs files (every aggregator, dtypes,
    a field called count, --zero-based, -N, --input-copy-status, unknown chromosomes, positions outside their
    chromosome) through cooler_b200/cli.py (numpy accumulator standing in for the device one) and through the live
    unmodified reference CLI, the writers replaced by recorders: same exit status, same per-chunk pixel tables,
    columns, dtypes (900 more command lines were compared once by hand).  Field numbers ascend in the order given --
    see the note on pandas' usecols in DESIGN.md 7.  Subprocess; skipped without the reference."""
    import subprocess
    import sys

    from oracle import refshim
    if not refshim.available():
        pytest.skip("reference sources not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tests", "_fuzz_cload_fields_vs_reference.py"), "150", "0"],
                         capture_output=True, text=True, timeout=900, cwd=root,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    assert res.stdout.strip().splitlines()[-1].startswith("ok: 300 cload pairs command lines")
