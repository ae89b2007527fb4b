"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's pairs binning.

Follows open2c/cooler src/cooler/create/_ingest.py:68-214 (_sanitize_records with
decode_chroms=False, fixed-size bins) and :452-504 (aggregate_records(sort=True, count=True):
groupby([bin1_id, bin2_id]).size()).  Pinned against golden vectors made by the unmodified
reference (tests/golden/make_ingest_golden.py).  Only tests/ may import this module.
"""
import numpy as np


class BadInputError(ValueError):
    pass


def bin_pairs(chrom1_ids, pos1, chrom2_ids, pos2, chromsizes, binsize, *, is_one_based=True,
              tril_action="reflect", validate=True):
    c1 = np.asarray(chrom1_ids).astype(np.int64)
    c2 = np.asarray(chrom2_ids).astype(np.int64)
    a1 = np.asarray(pos1).astype(np.int64).copy()
    a2 = np.asarray(pos2).astype(np.int64).copy()
    chromsizes = np.asarray(chromsizes, dtype=np.int64)
    keep = ~((c1 < 0) | (c2 < 0))                                     # :101-107
    c1, c2, a1, a2 = c1[keep], c2[keep], a1[keep], a2[keep]
    if is_one_based:                                                  # :118-120
        a1 -= 1
        a2 -= 1
    if validate:                                                      # :123-147
        if np.any((a1 < 0) | (a2 < 0)):
            raise BadInputError("negative anchor")
        if np.any((a1 > chromsizes[c1]) | (a2 > chromsizes[c2])):
            raise BadInputError("anchor exceeds chromosome length")
    if tril_action is not None:                                       # :152-188
        tril = (c1 > c2) | ((c1 == c2) & (a1 > a2))
        if tril_action == "reflect":
            c1, c2 = np.where(tril, c2, c1), np.where(tril, c1, c2)
            a1, a2 = np.where(tril, a2, a1), np.where(tril, a1, a2)
        elif tril_action == "drop":
            c1, c2, a1, a2 = c1[~tril], c2[~tril], a1[~tril], a2[~tril]
        elif tril_action == "raise":
            if np.any(tril):
                raise BadInputError("lower triangle pairs")
        else:
            raise ValueError(f"Unknown tril_action value: '{tril_action}'")
    binoff = np.r_[0, np.cumsum(-(-chromsizes // binsize))]
    b1 = binoff[c1] + a1 // binsize                                   # :216-217
    b2 = binoff[c2] + a2 // binsize
    n_bins = int(binoff[-1]) + 2
    key, cnt = np.unique(b1 * n_bins + b2, return_counts=True)        # groupby(sort=True).size()
    return {"bin1_id": key // n_bins, "bin2_id": key % n_bins, "count": cnt.astype(np.int32)}
