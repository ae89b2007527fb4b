"""Run in a SUBPROCESS by tests/test_load.py (importing the reference installs stub modules that must not
leak into the test process): seeded random chunks through cooler_b200.ingest.sanitize_records /
sanitize_pixels and through the UNMODIFIED reference's sanitize_records / sanitize_pixels
(create/_ingest.py:68-400), fixed and variable bins, every tril_action, one- and zero-based, unknown
chromosomes, anchors at and beyond the chromosome ends -- equal tables (after a sort) or the same error."""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ingest_numpy, refshim  # noqa: E402

from cooler_b200 import ingest  # noqa: E402
from cooler_b200.create import _bins_arrays  # noqa: E402


def bins_table(rng, variable):
    names = [f"c{k}" for k in range(int(rng.integers(1, 5)))]
    rows = []
    for nm in names:
        length = int(rng.integers(5, 400))
        if variable:
            edges = np.unique(np.r_[0, rng.integers(1, length, int(rng.integers(0, 12))), length])
        else:
            edges = np.r_[np.arange(0, length, BINSIZE), length]
            edges = np.unique(edges)
        rows += [(nm, int(a), int(b)) for a, b in zip(edges[:-1], edges[1:])]
    bins = pd.DataFrame(rows, columns=["chrom", "start", "end"])
    bins["chrom"] = pd.Categorical(bins["chrom"], categories=names, ordered=True)
    return names, bins


BINSIZE = 7


def main(n_cases, seed):
    refshim.import_reference()
    from cooler.create._ingest import BadInputError as RefError
    from cooler.create._ingest import sanitize_pixels as ref_pixels
    from cooler.create._ingest import aggregate_records as ref_aggregate
    from cooler.create._ingest import sanitize_records as ref_records
    rng = np.random.default_rng(seed)
    n_err = n_ok = 0
    for case in range(n_cases):
        variable = bool(rng.integers(0, 2))
        names, bins = bins_table(rng, variable)
        _, ids, start, end, _ = _bins_arrays(bins)
        lengths = np.asarray(end)[np.r_[ids[1:] != ids[:-1], True]]
        tril = [None, "reflect", "drop", "raise"][int(rng.integers(0, 4))]
        one_based = bool(rng.integers(0, 2))
        m = int(rng.integers(0, 60))
        bad = rng.random() < 0.25                       # allow out-of-range anchors in a quarter of the cases
        pool = names + (["zz"] if rng.random() < 0.5 else [])
        c1 = rng.choice(pool, m)
        c2 = rng.choice(pool, m)
        size = {n: int(s) for n, s in zip(names, lengths)}

        def pos(c):
            hi = np.array([size.get(x, 50) for x in c]) + (3 if bad else 0)
            lo = -1 if bad else (1 if one_based else 0)
            return rng.integers(lo, np.maximum(hi, lo + 1) + (1 if one_based else 0))
        for schema, f in (("bg2", "start"), ("pairs", "pos")):
            chunk = pd.DataFrame({"chrom1": c1, f + "1": pos(c1), "chrom2": c2, f + "2": pos(c2),
                                  "v": rng.random(m)})
            if schema == "bg2":
                chunk["end1"], chunk["end2"] = chunk["start1"] + 1, chunk["start2"] + 1
            cols = {k: chunk[k].to_numpy() for k in chunk.columns}
            try:
                want = ref_records(bins, schema=schema, decode_chroms=True, is_one_based=one_based,
                                   tril_action=tril, sort=True, validate=True)(chunk.copy())
                werr = None
            except RefError as e:
                want, werr = None, str(e).split("\n")[0][:25]
            try:
                got = ingest.sanitize_records(cols, names, lengths, ids, start, end, anchor_field=f,
                                              is_one_based=one_based, tril_action=tril)
                gerr = None
            except ingest.BadInputError as e:
                got, gerr = None, str(e)[:25]
            assert (werr is None) == (gerr is None), (case, schema, werr, gerr)
            if werr is not None:
                assert werr[:20] == gerr[:20], (werr, gerr)
                n_err += 1
                continue
            o = np.lexsort((got["v"], got["bin2_id"], got["bin1_id"]))
            w = want.sort_values(["bin1_id", "bin2_id", "v"])
            assert len(o) == len(w), (case, schema)
            for k in ("bin1_id", "bin2_id", "v"):
                np.testing.assert_array_equal(np.asarray(got[k])[o], w[k].to_numpy(), err_msg=f"{case} {schema} {k}")
            n_ok += 1
        # the pairs-binning oracle (integer chromosome ids, fixed bins) against sanitize + aggregate_records
        if not variable and m:
            ids1 = rng.integers(-1, len(names), m)
            ids2 = rng.integers(-1, len(names), m)
            p1 = pos(np.array(names + ["zz"], dtype=object)[ids1])
            p2 = pos(np.array(names + ["zz"], dtype=object)[ids2])
            rec = pd.DataFrame({"chrom1": ids1, "pos1": p1, "chrom2": ids2, "pos2": p2})
            try:
                w = ref_aggregate(sort=True, count=True)(ref_records(
                    bins, schema="pairs", decode_chroms=False, is_one_based=one_based, tril_action=tril,
                    sort=False, validate=True)(rec.copy()))
                werr = None
            except RefError:
                w, werr = None, "bad"
            try:
                g = ingest_numpy.bin_pairs(ids1, p1, ids2, p2, lengths, BINSIZE, is_one_based=one_based,
                                           tril_action=tril)
                gerr = None
            except ingest_numpy.BadInputError:
                g, gerr = None, "bad"
            assert werr == gerr, (case, "pairs oracle", werr, gerr)
            if werr is None:
                for k in ("bin1_id", "bin2_id", "count"):
                    np.testing.assert_array_equal(g[k], w[k].to_numpy(), err_msg=f"{case} pairs oracle {k}")
                n_ok += 1
            else:
                n_err += 1
        # already-binned pixels
        n_bins = len(bins)
        px = pd.DataFrame({"bin1_id": rng.integers(one_based, n_bins + one_based, m),
                           "bin2_id": rng.integers(one_based, n_bins + one_based, m), "count": rng.integers(1, 9, m)})
        try:
            want = ref_pixels(bins, is_one_based=one_based, tril_action=tril, sort=True)(px.copy())
            werr = None
        except RefError as e:
            want, werr = None, str(e)[:25]
        try:
            got = ingest.sanitize_pixels({k: px[k].to_numpy() for k in px.columns}, is_one_based=one_based,
                                         tril_action=tril)
            gerr = None
        except ingest.BadInputError as e:
            got, gerr = None, str(e)[:25]
        assert werr == gerr, (case, werr, gerr)
        if werr is None:
            o = np.lexsort((got["count"], got["bin2_id"], got["bin1_id"]))
            w = want.sort_values(["bin1_id", "bin2_id", "count"])
            for k in ("bin1_id", "bin2_id", "count"):
                np.testing.assert_array_equal(np.asarray(got[k])[o], w[k].to_numpy(), err_msg=f"{case} px {k}")
            n_ok += 1
        else:
            n_err += 1
    print(f"ok: {n_ok} equal tables, {n_err} equal errors")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 300, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
