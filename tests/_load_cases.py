"""Cases of the `cooler load` goldens (tests/golden/make_load_golden.py runs the unmodified reference's
parser options + sanitizers + chunk checks; the tests run cooler_b200's).  Text inputs: the reference's own
toy fixtures (tests/golden/text/, copied from its tests/data) and a seeded shuffled COO file with mirrored
records, pixels split over several chunks and a floating-point extra field (written by ``write_synthetic``)."""
import os

import numpy as np

TEXT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "text")

# key -> (format, bins spec relative to TEXT, pixels file, extra CLI arguments)
CASES = {
    "bg2_symm": ("bg2", "toy.chrom.sizes:2", "toy.symm.upper.2.bg2", ["--chunksize", "10"]),
    "bg2_symm_ob": ("bg2", "toy.chrom.sizes:2", "toy.symm.upper.2.bg2",
                    ["--chunksize", "10", "--one-based", "--input-copy-status", "unique"]),
    "bg2_asymm": ("bg2", "toy.chrom.sizes:2", "toy.asymm.2.bg2",
                  ["--chunksize", "10", "--one-based", "--no-symmetric-upper"]),
    "bg2_float": ("bg2", "toy.chrom.sizes:2", "toy.symm.upper.2.bg2", ["--chunksize", "7", "--count-as-float"]),
    "bg2_field": ("bg2", "toy.chrom.sizes:2", "toy.symm.upper.2.bg2",
                  ["--chunksize", "10", "--field", "count=7:dtype=float"]),
    "coo_ob": ("coo", "toy.chrom.sizes:1", "toy.symm.upper.1.ob.coo", ["--chunksize", "10", "--one-based"]),
    "coo_zb": ("coo", "toy.chrom.sizes:1", "toy.symm.upper.1.zb.coo", ["--chunksize", "25"]),
    "coo_zb_duplex": ("coo", "toy.chrom.sizes:1", "toy.symm.upper.1.zb.coo",
                      ["--chunksize", "25", "--input-copy-status", "duplex"]),
    "coo_shuffled": ("coo", "toy.chrom.sizes:1", "synthetic.shuffled.coo",
                     ["--chunksize", "300", "--field", "count=3", "--field", "score=4:dtype=float64"]),
}


def write_synthetic(path, n_chunks=6, chunk=300, seed=3):
    """A COO file over the 64 bins of toy.chrom.sizes:1 whose records come in random order: pixels are
    split into 1-3 records that land in DIFFERENT chunks of ``chunk`` lines (a chunk must not hold a pixel
    twice: the duplicate check of create), about a third of the off-diagonal records are written mirrored
    (bin1 > bin2).  Columns: bin1, bin2, count, score.  Every chunk holds exactly ``chunk`` lines."""
    rng = np.random.default_rng(seed)
    n = 64
    i, j = np.triu_indices(n)
    order = rng.permutation(len(i))
    per_chunk = [[] for _ in range(n_chunks)]
    for a, b in zip(i[order], j[order]):
        free = [c for c in range(n_chunks) if len(per_chunk[c]) < chunk]
        if not free:
            break
        k = min(int(rng.integers(1, 4)), len(free))
        for c in rng.choice(free, size=k, replace=False):
            cnt, score = int(rng.integers(1, 40)), float(np.round(rng.random() * 10, 6))
            rec = (b, a, cnt, score) if (a != b and rng.random() < 0.33) else (a, b, cnt, score)
            per_chunk[c].append(rec)
    assert all(len(r) == chunk for r in per_chunk)
    with open(path, "w") as f:
        for recs in per_chunk:
            for q in rng.permutation(len(recs)):
                a, b, cnt, score = recs[q]
                f.write(f"{a}\t{b}\t{cnt}\t{score}\n")


# `cooler cload pairs --field ...` on the reference's toy.pairs (columns: readID chr1 pos1 chr2 pos2 strand1
# strand2 score), called like its tests/test_cli_ingest.py:_run_cload_pairs.  key -> (binsize, extra arguments)
PAIRS_BASE = ["-c1", "2", "-p1", "3", "-c2", "4", "-p2", "5", "--assembly", "toy", "--chunksize", "10"]
PAIRS_CASES = {
    "pairs_score": (2, ["--field", "score=8"]),
    "pairs_count_int": (2, ["--field", "count=8"]),
    "pairs_count_float": (2, ["--field", "count=8:dtype=float"]),
    "pairs_count_min": (2, ["--field", "count=8:agg=min,dtype=float"]),
    "pairs_score_mean_f32": (4, ["--field", "score=8:agg=mean,dtype=float32", "--field", "count:dtype=int64"]),
    "pairs_max_binsize8": (8, ["--field", "score=8:agg=max"]),
}
