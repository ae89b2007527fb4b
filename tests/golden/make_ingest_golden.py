"""Golden vectors for pairs binning made by the UNMODIFIED reference:
``aggregate_records(sort=True)(sanitize_records(bins, schema="pairs", decode_chroms=True, ...)(chunk))``
(src/cooler/create/_ingest.py) run through oracle/refshim.py in the build container.  The inputs are
committed next to the outputs as integer columns (chromosome names already decoded to ids in the
order of the chromsizes file; -1 = chromosome not requested).

    python tests/golden/make_ingest_golden.py    ->  tests/golden/ingest_golden.npz
"""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

DATA = "/root/reference/tests/data"


def main():
    cooler = refshim.import_reference()
    from cooler.create import aggregate_records, sanitize_records
    from cooler.util import binnify
    out = {}

    def run(tag, chunk, chromsizes, binsize, is_one_based=True, **kw):
        kw["is_one_based"] = is_one_based                # cli/cload.py:638 passes `not zero_based`
        bins = binnify(chromsizes, binsize)
        names = list(chromsizes.index)
        ids1 = pd.Categorical(chunk["chrom1"], names, ordered=True).codes.astype(np.int32)
        ids2 = pd.Categorical(chunk["chrom2"], names, ordered=True).codes.astype(np.int32)
        res = aggregate_records(sort=True)(
            sanitize_records(bins, schema="pairs", decode_chroms=True, **kw)(chunk.copy()))
        out[f"{tag}/chrom1"], out[f"{tag}/chrom2"] = ids1.astype(np.int8), ids2.astype(np.int8)
        out[f"{tag}/pos1"], out[f"{tag}/pos2"] = chunk["pos1"].to_numpy(np.int32), chunk["pos2"].to_numpy(np.int32)
        out[f"{tag}/chromsizes"] = chromsizes.to_numpy(np.int64)
        out[f"{tag}/binsize"] = np.array(binsize)
        out[f"{tag}/one_based"] = np.array(bool(is_one_based))
        for c in ("bin1_id", "bin2_id", "count"):
            out[f"{tag}/out_{c}"] = res[c].to_numpy()
        print(tag, len(chunk), "->", len(res), "pixels, sum", res["count"].sum())

    toy_sizes = pd.read_csv(f"{DATA}/toy.chrom.sizes", sep="\t", names=["name", "length"]).set_index("name")["length"]
    toy = pd.read_csv(f"{DATA}/toy.pairs", sep="\t", header=None,
                      names=["read_id", "chrom1", "pos1", "chrom2", "pos2", "strand1", "strand2", "value"])
    run("toy2_reflect", toy, toy_sizes, 2, tril_action="reflect")
    run("toy8_drop", toy, toy_sizes, 8, tril_action="drop")
    run("toy1_none", toy, toy_sizes, 1, tril_action=None)
    run("toy4_zero_based", toy.assign(pos1=toy.pos1 - 1, pos2=toy.pos2 - 1), toy_sizes, 4, is_one_based=False,
        tril_action="reflect")

    hg19 = pd.Series(
        [249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663, 146364022, 141213431,
         135534747, 135006516, 133851895, 115169878, 107349540, 102531392, 90354753, 81195210, 78077248,
         59128983, 63025520, 48129895, 51304566, 155270560, 59373566, 16571],
        index=[f"chr{i}" for i in list(range(1, 23)) + ["X", "Y", "M"]], name="length")
    gm = pd.read_csv(f"{DATA}/hg19.GM12878-MboI.pairs.subsample.sorted.txt.gz", sep="\t", header=None,
                     names=["chrom1", "pos1", "strand1", "chrom2", "pos2", "strand2"], nrows=8000)
    run("gm_2mb", gm, hg19, 2_000_000, tril_action="reflect")
    run("gm_100kb", gm, hg19, 100_000, tril_action="reflect")
    sub = hg19.iloc[[0, 2, 4, 22]]                      # only chr1, chr3, chr5, chrX requested: others dropped
    run("gm_subset_500kb", gm, sub, 500_000, tril_action="reflect")
    rng = np.random.default_rng(0)
    sh = gm.iloc[rng.permutation(len(gm))].reset_index(drop=True)      # unsorted, with lower-triangle records
    flip = rng.random(len(sh)) < 0.5
    sh2 = sh.copy()
    sh2.loc[flip, ["chrom1", "pos1", "chrom2", "pos2"]] = sh.loc[flip, ["chrom2", "pos2", "chrom1", "pos1"]].to_numpy()
    run("gm_shuffled_1mb_reflect", sh2, hg19, 1_000_000, tril_action="reflect")
    run("gm_shuffled_1mb_drop", sh2, hg19, 1_000_000, tril_action="drop")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ingest_golden.npz"), **out)


if __name__ == "__main__":
    main()
