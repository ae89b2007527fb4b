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
