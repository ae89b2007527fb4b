"""``cooler dump`` on the GPU path (weight application / ``dump --balanced``).

Text export of a cooler's tables with the reference's options and byte-identical output
(reference src/cooler/cli/dump.py:19-304).  For the pixel table the 2-D range query
(``DirectRangeQuery2D`` / ``FillLowerRangeQuery2D``, core/_rangequery.py:405-565), its task
order and the balancing ``weight1 * weight2 * count`` run on the resident table through
``cooler_b200.api.MatrixSelector`` (csrc/fetch.cu); the host only plans the tasks, joins the
bin annotations and formats text.
"""
from __future__ import annotations

import gzip
import sys

import numpy as np

from .api import MatrixSelector, region_to_extent
from .table import PixelTable

__all__ = ["dump", "plan_tasks"]


def _comes_before(a0, a1, b0, b1, strict=False):          # core/_rangequery.py:52-57
    if a0 < b0:
        return a1 <= b0 if strict else a1 <= b1
    return False


def _contains(a0, a1, b0, b1):                            # core/_rangequery.py:60-67
    return a0 <= b0 and a1 >= b1


def _spans(bin1_offset, bbox, chunksize):
    """CSRReader.get_spans (core/_rangequery.py:182-193): row spans of about ``chunksize`` pixels."""
    i0, i1, j0, j1 = bbox
    if (i1 - i0 < 1) or (j1 - j0 < 1):
        return []
    seq = bin1_offset[i0:i1 + 1]
    lo, hi = int(seq[0]), int(seq[-1])
    num = 2 + (hi - lo) // int(chunksize)
    cuts = np.linspace(lo, hi, num, dtype=int)
    edges = i0 + np.unique(np.searchsorted(seq, cuts))
    return list(zip(edges[:-1].tolist(), edges[1:].tolist()))


def plan_tasks(bbox, bin1_offset, chunksize, fill_lower):
    """The reference's task list: [(bbox, row span, reflect, transpose)] in output order."""
    if not fill_lower:                                    # DirectRangeQuery2D, 443-446
        return [(bbox, span, False, False) for span in _spans(bin1_offset, bbox, chunksize)]
    i0, i1, j0, j1 = bbox                                 # FillLowerRangeQuery2D, 507-565
    use_t = i1 > j1
    if use_t:
        i0, i1, j0, j1 = j0, j1, i0, i1
    if i0 == j0 or _comes_before(i0, i1, j0, j1, strict=True):
        boxes, tr = [(i0, i1, j0, j1)], [use_t]
    elif _comes_before(i0, i1, j0, j1):
        boxes, tr = [(i0, j0, j0, j1), (j0, i1, j0, j1)], [use_t, use_t]
    elif _contains(j0, j1, i0, i1):
        boxes, tr = [(j0, i0, i0, i1), (i0, i1, i0, j1)], [not use_t, use_t]
    else:
        raise ValueError("This shouldn't happen.")
    tasks = []
    for box, t in zip(boxes, tr):
        tasks += [(box, span, True, t) for span in _spans(bin1_offset, box, chunksize)]
    return tasks


def _run_task(sel, task):
    """One CSRReader call (core/_rangequery.py:208-310) on the device + reflect / transpose."""
    (i0, i1, j0, j1), (s0, s1), reflect, transposed = task
    df = sel._slice(s0, s1, j0, j1)                       # rows [s0, s1) x columns [j0, j1), row-major
    if reflect and len(df):
        dup = ((df["bin1_id"] != df["bin2_id"]) & (df["bin2_id"] < i1)).to_numpy()
        if dup.any():
            extra = df[dup].rename(columns={"bin1_id": "bin2_id", "bin2_id": "bin1_id"})[list(df.columns)]
            import pandas as pd
            df = pd.concat([df, extra], axis=0, ignore_index=True)
    if transposed:
        df = df.rename(columns={"bin1_id": "bin2_id", "bin2_id": "bin1_id"})[list(df.columns)]
    return df


def _bins_table(clr):
    """``clr.bins()[:]`` with chromosome names (categorical), like ``cooler.Cooler.bins``."""
    import pandas as pd
    bins = clr.bins()[:]
    if "chrom" in bins.columns and bins["chrom"].dtype.kind in "iu":
        names = list(clr.chromsizes.index)
        bins = bins.assign(chrom=pd.Categorical.from_codes(bins["chrom"].to_numpy(), categories=names, ordered=True))
    first = [c for c in ("chrom", "start", "end") if c in bins.columns]
    return bins[first + [c for c in bins.columns if c not in first]]


def _take(bins_cols, ids, suffix):
    out = bins_cols.iloc[np.asarray(ids, dtype=np.int64)].reset_index(drop=True)
    return out.rename(columns=lambda x: x + suffix)


def dump(clr, out=None, *, table="pixels", columns=None, header=False, na_rep="", float_format="g",
         range=None, range2=None, fill_lower=False, balanced=False, join=False, annotate=None,
         one_based_ids=False, one_based_starts=False, chunksize=1_000_000, device="cuda", pixel_table=None):
    """Write ``table`` of ``clr`` as tab-separated text to ``out`` (path, '-' / None = stdout, or a
    file object).  Arguments as the options of ``cooler dump``."""
    import pandas as pd
    close = False
    if out is None or out == "-":
        f = sys.stdout
    elif isinstance(out, str):
        f = gzip.open(out, "wt") if out.endswith(".gz") else open(out, "w")
        close = True
    else:
        f = out
    try:
        if table == "chroms":
            ch = pd.DataFrame({"name": clr.chromnames, "length": clr.chromsizes.values})
            with clr.open("r") as g:
                for k in g["chroms"].keys():
                    if k not in ("name", "length"):
                        ch[k] = g["chroms"][k][:]
            chunks = ((ch[list(columns)] if columns is not None else ch),)
        elif table == "bins":
            bins = _bins_table(clr)
            chunks = ((bins[list(columns)] if columns is not None else bins),)
        else:
            bins = _bins_table(clr)
            n_bins = len(bins)
            if balanced and "weight" not in bins.columns:
                print("Balancing weights not found", file=sys.stderr)
                raise SystemExit(1)
            if annotate is not None:
                missing = [c for c in annotate if c not in bins.columns]
                if missing:
                    print(f"Column not found:\n {missing}")
                    raise SystemExit(1)
            if range:
                i0, i1 = region_to_extent(clr, range)
                j0, j1 = region_to_extent(clr, range2) if range2 is not None else (i0, i1)
                bbox = (i0, i1, j0, j1)
            else:
                bbox = (0, n_bins, 0, n_bins)
            if chunksize is None:
                chunksize = n_bins
            bin1_offset = np.asarray(clr._load_dset("indexes/bin1_offset"))
            tasks = plan_tasks(bbox, bin1_offset, chunksize,
                               bool(fill_lower) and clr.storage_mode == "symmetric-upper")
            # an int32-class count column lives in the table's records; a floating-point or wide one is
            # fetched as float64 values through the same queries (api.MatrixSelector(values=...))
            values = None
            dt = np.dtype(clr.pixels().dtypes["count"])
            if not (dt.kind in "iu" and dt.itemsize <= 4):
                values = np.asarray(clr._load_dset("pixels/count"))
            tab = pixel_table if pixel_table is not None else PixelTable.from_cooler(
                clr, device=device, structure_only=values is not None)
            weights = bins["weight"].to_numpy(np.float64) if balanced else None
            sel = MatrixSelector(clr, tab, weights, as_pixels=True, values=values)

            def gen():
                for task in tasks:
                    chunk = _run_task(sel, task)
                    if annotate is not None:              # make_annotator, cli/dump.py:19-52
                        extra = pd.concat([_take(bins[list(annotate)], chunk["bin1_id"], "1"),
                                           _take(bins[list(annotate)], chunk["bin2_id"], "2")], axis=1)
                    if join:
                        coords = bins[["chrom", "start", "end"]]
                        chunk = pd.concat([_take(coords, chunk["bin1_id"], "1"), _take(coords, chunk["bin2_id"], "2"),
                                           chunk.drop(columns=["bin1_id", "bin2_id"]).reset_index(drop=True)], axis=1)
                    if annotate is not None:
                        chunk = pd.concat([chunk.reset_index(drop=True), extra], axis=1)
                    if balanced or join or annotate:      # the reference shifts only inside the annotator
                        if one_based_ids:
                            for col in ("bin1_id", "bin2_id"):
                                if col in chunk.columns:
                                    chunk[col] += 1
                        if one_based_starts:
                            for col in ("start1", "start2"):
                                if col in chunk.columns:
                                    chunk[col] += 1
                    yield chunk
            chunks = gen()
        fmt = "%" + float_format if float_format is not None else None
        first = True
        for chunk in chunks:
            if first:
                if header:
                    chunk[0:0].to_csv(f, sep="\t", lineterminator="\n", index=False, header=True,
                                      float_format=fmt, na_rep=na_rep)
                first = False
            chunk.to_csv(f, sep="\t", lineterminator="\n", index=False, header=False, float_format=fmt,
                         na_rep=na_rep)
        f.flush()
    finally:
        if close:
            f.close()
