"""Pairs binning: oracle vs golden vectors of the unmodified reference (CPU) and the
CUDA path vs both (GPU).  Integer work: bit-exact."""
import os

import numpy as np
import pytest

from tests import _golden

G = dict(np.load(os.path.join(_golden.GOLD, "ingest_golden.npz"), allow_pickle=False))
CASES = {"toy4_zero_based": "reflect", "toy2_reflect": "reflect", "toy8_drop": "drop", "toy1_none": None, "gm_2mb": "reflect",
         "gm_100kb": "reflect", "gm_subset_500kb": "reflect", "gm_shuffled_1mb_reflect": "reflect",
         "gm_shuffled_1mb_drop": "drop"}


def _args(tag):
    return (G[f"{tag}/chrom1"].astype(np.int32), G[f"{tag}/pos1"].astype(np.int64),
            G[f"{tag}/chrom2"].astype(np.int32), G[f"{tag}/pos2"].astype(np.int64),
            G[f"{tag}/chromsizes"], int(G[f"{tag}/binsize"]))


def _kw(tag):
    return dict(tril_action=CASES[tag], is_one_based=bool(G[f"{tag}/one_based"]))


def _check(got, tag):
    for c in ("bin1_id", "bin2_id", "count"):
        np.testing.assert_array_equal(got[c], G[f"{tag}/out_{c}"], err_msg=f"{tag}/{c}")


@pytest.mark.parametrize("tag", sorted(CASES))
def test_oracle_matches_reference_golden(tag):
    from oracle import ingest_numpy
    _check(ingest_numpy.bin_pairs(*_args(tag), **_kw(tag)), tag)


def test_oracle_errors():
    from oracle import ingest_numpy
    sizes = np.array([100, 50])
    with pytest.raises(ingest_numpy.BadInputError):
        ingest_numpy.bin_pairs([0], [0], [1], [5], sizes, 10)                 # 1-based input holding a 0
    with pytest.raises(ingest_numpy.BadInputError):
        ingest_numpy.bin_pairs([0], [5], [1], [52], sizes, 10)                # beyond the chromosome
    with pytest.raises(ingest_numpy.BadInputError):
        ingest_numpy.bin_pairs([1], [5], [0], [5], sizes, 10, tril_action="raise")


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(CASES))
def test_gpu_bin_pairs_matches_reference_golden(tag):
    from cooler_b200 import ingest
    got = ingest.bin_pairs(*_args(tag), **_kw(tag))
    assert got["bin1_id"].dtype == np.int64 and got["count"].dtype == np.int32
    _check(got, tag)


@pytest.mark.gpu
def test_gpu_bin_pairs_errors_and_edges():
    from cooler_b200 import ingest
    sizes = np.array([100, 50])
    with pytest.raises(ingest.BadInputError, match="negative"):
        ingest.bin_pairs([0], [0], [1], [5], sizes, 10)
    with pytest.raises(ingest.BadInputError, match="exceeding"):
        ingest.bin_pairs([0], [5], [1], [52], sizes, 10)
    with pytest.raises(ingest.BadInputError, match="lower triangle"):
        ingest.bin_pairs([1], [5], [0], [5], sizes, 10, tril_action="raise")
    with pytest.raises(ingest.BadInputError):
        ingest.bin_pairs(np.array([0.5]), [5], [0], [5], sizes, 10)
    with pytest.raises(ValueError):
        ingest.bin_pairs([0], [5], [0], [5], sizes, 10, tril_action="bogus")
    out = ingest.bin_pairs([], [], [], [], sizes, 10)
    assert all(len(v) == 0 for v in out.values())
    # a 0-based position equal to the length of the LAST chromosome (a multiple of the bin size) passes
    # the anchor check (`> chromsize`, as in the reference) but is bin n_bins: bounds error, never a
    # silent drop (it used to collide with the dropped-record sentinel)
    with pytest.raises(ingest.BadInputError, match="exceeds the declared number of bins"):
        ingest.bin_pairs([0, 0], [5, 7], [0, 1], [9, 50], sizes, 10, is_one_based=False)
    # many dropped records (unplaced contigs) next to a few valid ones: cut off, not walked as one run
    nd = 300_000
    out = ingest.bin_pairs(np.r_[np.full(nd, -1), 0, 0], np.r_[np.ones(nd, int), 5, 7],
                           np.r_[np.zeros(nd, int), 0, 1], np.r_[np.ones(nd, int), 95, 3], sizes, 10)
    np.testing.assert_array_equal(out["bin1_id"], [0, 0])
    np.testing.assert_array_equal(out["bin2_id"], [9, 10])
    np.testing.assert_array_equal(out["count"], [1, 1])
    out = ingest.bin_pairs([-1, 0], [5, 7], [0, -1], [5, 9], sizes, 10)        # everything dropped
    assert all(len(v) == 0 for v in out.values())
    out = ingest.bin_pairs([0, 0, 1], [5, 7, 3], [0, 0, 0], [95, 99, 15], sizes, 10, is_one_based=False)
    np.testing.assert_array_equal(out["bin1_id"], [0, 1])
    np.testing.assert_array_equal(out["bin2_id"], [9, 10])
    np.testing.assert_array_equal(out["count"], [2, 1])


@pytest.mark.gpu
def test_gpu_bin_pairs_matches_oracle_large_random():
    from cooler_b200 import ingest
    from oracle import ingest_numpy
    rng = np.random.default_rng(9)
    sizes = np.array([5_000_000, 3_100_001, 999, 2_000_000])
    n = 3_000_000
    c1 = rng.integers(-1, 4, n).astype(np.int32)
    c2 = rng.integers(-1, 4, n).astype(np.int32)
    p1 = (rng.random(n) * sizes[np.maximum(c1, 0)]).astype(np.int64) + 1
    d = (rng.pareto(1.0, n) * 2000).astype(np.int64)
    p2 = np.where(c1 == c2, np.minimum(p1 + d, sizes[np.maximum(c2, 0)]),
                  (rng.random(n) * sizes[np.maximum(c2, 0)]).astype(np.int64) + 1)
    for binsize, tril in ((1000, "reflect"), (37, "drop"), (50_000, None)):
        want = ingest_numpy.bin_pairs(c1, p1, c2, p2, sizes, binsize, tril_action=tril)
        got = ingest.bin_pairs(c1, p1, c2, p2, sizes, binsize, tril_action=tril)
        for c in ("bin1_id", "bin2_id", "count"):
            np.testing.assert_array_equal(got[c], want[c])
        assert got["count"].sum() == want["count"].sum()


def _write_pairs(tmp_path, tag, n_max=200_000):
    """A .pairs text file + chromsizes file from a golden case (chromosome i is named 'c<i>';
    records on chromosomes the case marks as unknown (-1) are written as 'other')."""
    c1, p1, c2, p2, sizes, binsize = _args(tag)
    names = np.array([f"c{i}" for i in range(len(sizes))] + ["other"], dtype=object)
    path = str(tmp_path / f"{tag}.pairs")
    with open(path, "w") as f:
        f.write("## pairs format v1.0\n#columns: readID chr1 pos1 chr2 pos2 strand1 strand2\n")
        for k in range(len(c1)):
            f.write(f"r{k}\t{names[c1[k]]}\t{p1[k]}\t{names[c2[k]]}\t{p2[k]}\t+\t-\n")
        f.write("rX\tchrUn_zzz\t1\tc0\t1\t+\t-\n")            # contig missing from the bin table: dropped
    cs = str(tmp_path / f"{tag}.chrom.sizes")
    with open(cs, "w") as f:
        for n, L in zip(names, sizes):
            f.write(f"{n}\t{int(L)}\n")
    return path, cs, binsize


def _check_cload_output(out, tag):
    from cooler_b200 import store
    from ._h5check import check
    check(out)
    clr = store.open_cooler(out)
    for c in ("bin1_id", "bin2_id", "count"):
        np.testing.assert_array_equal(np.asarray(clr._load_dset("pixels/" + c)), G[f"{tag}/out_{c}"], err_msg=c)
    sizes, binsize = G[f"{tag}/chromsizes"], int(G[f"{tag}/binsize"])
    assert clr.chromnames == [f"c{i}" for i in range(len(sizes))] and list(clr.chromsizes.values) == list(sizes)
    assert clr.info["nbins"] == int(np.sum(-(-sizes // binsize))) and clr.info["sum"] == int(G[f"{tag}/out_count"].sum())
    assert clr.info["nnz"] == len(G[f"{tag}/out_count"]) and clr.storage_mode == "symmetric-upper"
    ind = np.asarray(clr._load_dset("indexes/bin1_offset"))
    np.testing.assert_array_equal(ind, np.searchsorted(G[f"{tag}/out_bin1_id"], np.arange(len(ind))))


@pytest.mark.parametrize("tag", ["gm_subset_500kb", "toy4_zero_based", "gm_shuffled_1mb_drop"])
def test_cload_pairs_cli_plumbing_with_oracle(tag, tmp_path, monkeypatch):
    """`cooler cload pairs BINS PAIRS OUT` (cli/cload.py:484-666): text parsing, contig names -> ids,
    flags, file output; the numpy oracle stands in for the device binning (GPU: next test)."""
    from click.testing import CliRunner

    from cooler_b200 import ingest
    from cooler_b200.cli import cli
    from oracle import ingest_numpy
    monkeypatch.setattr(ingest, "bin_pairs", lambda *a, device="cuda", **kw: ingest_numpy.bin_pairs(*a, **kw))
    path, cs, binsize = _write_pairs(tmp_path, tag)
    out = str(tmp_path / "out.cool")
    args = ["cload", "pairs", f"{cs}:{binsize}", path, out, "-c1", "2", "-p1", "3", "-c2", "4", "-p2", "5",
            "--chunksize", "1000", "--assembly", "toy"]
    if not bool(G[f"{tag}/one_based"]):
        args.append("--zero-based")
    if CASES[tag] == "drop":
        args += ["--input-copy-status", "duplex"]
    res = CliRunner().invoke(cli, args)
    assert res.exit_code == 0, (res.output, res.exception)
    _check_cload_output(out, tag)
    from cooler_b200 import store
    assert store.open_cooler(out).info["genome-assembly"] == "toy"
    res = CliRunner().invoke(cli, args[:-2] + ["-c1", "0"])
    assert res.exit_code != 0


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["gm_subset_500kb", "toy4_zero_based"])
def test_cload_pairs_cli_gpu(tag, tmp_path):
    from click.testing import CliRunner

    from cooler_b200.cli import cli
    path, cs, binsize = _write_pairs(tmp_path, tag)
    out = str(tmp_path / "out.cool")
    args = ["cload", "pairs", f"{cs}:{binsize}", path, out, "-c1", "2", "-p1", "3", "-c2", "4", "-p2", "5"]
    if not bool(G[f"{tag}/one_based"]):
        args.append("--zero-based")
    res = CliRunner().invoke(cli, args)
    assert res.exit_code == 0, (res.output, res.exception)
    _check_cload_output(out, tag)
