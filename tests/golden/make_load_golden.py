"""Golden vectors for `cooler load` (COO / BG2 text in any order -> cooler), made with the UNMODIFIED
reference's own building blocks run through oracle/refshim.py: the option handling of ``cli/load.py:196-360``
is followed literally, the chunks go through the reference's ``sanitize_pixels`` / ``sanitize_records(schema=
"bg2")`` (create/_ingest.py:68-400) and ``validate_pixels`` (:401-446) exactly as ``create_from_unordered`` ->
``create`` applies them to every chunk, value columns are cast to the output dtypes as the temporary coolers
store them, and the chunks are merged the way ``CoolerMerger`` merges the temporary coolers (``pd.concat`` +
``groupby([bin1_id, bin2_id], sort=True).sum()``, _reduce.py:188-203).  The HDF5 writes in between cannot run
here (no libhdf5), which is why the CLI itself is not invoked.

    python tests/golden/make_load_golden.py   ->  tests/golden/load_golden.npz

Stored per case: the sanitized chunks concatenated in input order (``<key>/chunks/<column>`` + the chunk
lengths) and the merged table (``<key>/merged/<column>``).
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

from tests._load_cases import CASES, PAIRS_BASE, PAIRS_CASES, TEXT  # noqa: E402


def reference_load(fmt, bins_spec, pixels, args):
    refshim.import_reference()
    from cooler.cli._util import parse_bins, parse_field_param
    from cooler.create._ingest import sanitize_pixels, sanitize_records, validate_pixels
    opt = {"chunksize": int(20e6), "one_based": False, "count_as_float": False, "no_symmetric_upper": False,
           "input_copy_status": "unique", "field": []}
    it = iter(args)
    for a in it:
        if a == "--chunksize":
            opt["chunksize"] = int(next(it))
        elif a == "--field":
            opt["field"].append(next(it))
        elif a == "--input-copy-status":
            opt["input_copy_status"] = next(it)
        else:
            opt[a.lstrip("-").replace("-", "_")] = True
    _, bins = parse_bins(os.path.join(TEXT, bins_spec))
    symmetric_upper = not opt["no_symmetric_upper"]
    tril_action = None
    if symmetric_upper:
        tril_action = "reflect" if opt["input_copy_status"] == "unique" else "drop"
    out_names = ["bin1_id", "bin2_id"]
    out_dtypes = {"bin1_id": np.int64, "bin2_id": np.int64, "count": np.int32}
    if fmt == "bg2":
        in_names = ["chrom1", "start1", "end1", "chrom2", "start2", "end2"]
        in_dtypes = {"chrom1": str, "start1": int, "end1": int, "chrom2": str, "start2": int, "end2": int,
                     "count": out_dtypes["count"]}
        in_numbers = {"chrom1": 0, "start1": 1, "end1": 2, "chrom2": 3, "start2": 4, "end2": 5, "count": 6}
        pipeline = sanitize_records(bins, schema="bg2", is_one_based=opt["one_based"], tril_action=tril_action,
                                    sort=True)
    else:
        in_names = ["bin1_id", "bin2_id"]
        in_dtypes = {"bin1_id": int, "bin2_id": int, "count": out_dtypes["count"]}
        in_numbers = {"bin1_id": 0, "bin2_id": 1, "count": 2}
        pipeline = sanitize_pixels(bins, is_one_based=opt["one_based"], tril_action=tril_action, sort=True)
    if opt["field"]:
        for arg in opt["field"]:
            name, colnum, dtype, _ = parse_field_param(arg, includes_agg=False)
            if colnum is None:
                assert name == "count" and dtype is not None
                in_names.append("count"), out_names.append("count")
                in_dtypes[name] = out_dtypes[name] = dtype
                continue
            if name not in in_names:
                in_names.append(name)
            if name not in out_names:
                out_names.append(name)
            in_numbers[name] = colnum
            if dtype is not None:
                in_dtypes[name] = out_dtypes[name] = dtype
    else:
        in_names.append("count"), out_names.append("count")
    if "count" in in_names and opt["count_as_float"]:
        in_dtypes["count"] = out_dtypes["count"] = np.float64
    reader = pd.read_csv(os.path.join(TEXT, pixels), sep="\t",
                         usecols=[in_numbers[n] for n in in_names], names=in_names, dtype=in_dtypes,
                         comment="#", iterator=True, chunksize=opt["chunksize"])
    check = validate_pixels(len(bins), boundscheck=True, triucheck=symmetric_upper, dupcheck=True,
                            ensure_sorted=False)
    value_cols = [c for c in out_names if c not in ("bin1_id", "bin2_id")]
    chunks = []
    for raw in reader:
        chunk = check(pipeline(raw))
        cols = {"bin1_id": chunk["bin1_id"].to_numpy(np.int64), "bin2_id": chunk["bin2_id"].to_numpy(np.int64)}
        for c in value_cols:                       # stored in the temporary cooler with its output dtype
            cols[c] = chunk[c].to_numpy().astype(out_dtypes.get(c, float))
        chunks.append(pd.DataFrame(cols))
    merged = (pd.concat(chunks, axis=0, ignore_index=True).groupby(["bin1_id", "bin2_id"], sort=True)
              .aggregate({c: "sum" for c in value_cols}).reset_index())
    return chunks, merged, {c: np.dtype(out_dtypes.get(c, float)) for c in value_cols}


def reference_cload_pairs(binsize, args):
    """cli/cload.py:512-666 followed literally up to ``create_cooler(..., ordered=False)``: the chunks
    ``aggregate_records(agg, count=True, sort=False)(sanitize_records(schema="pairs", ...)(chunk))`` are cast
    to the output dtypes (the temporary coolers of create_from_unordered) and merged by summing."""
    refshim.import_reference()
    from cooler.cli._util import parse_bins, parse_field_param
    from cooler.create._ingest import aggregate_records, sanitize_records
    _, bins = parse_bins(os.path.join(TEXT, f"toy.chrom.sizes:{binsize}"))
    opt = {"chunksize": 10, "zero_based": False, "field": []}
    nums = {}
    it = iter(args)
    for a in it:
        if a in ("-c1", "-p1", "-c2", "-p2"):
            nums[{"-c1": "chrom1", "-p1": "pos1", "-c2": "chrom2", "-p2": "pos2"}[a]] = int(next(it)) - 1
        elif a == "--chunksize":
            opt["chunksize"] = int(next(it))
        elif a == "--field":
            opt["field"].append(next(it))
        elif a == "--assembly":
            next(it)
        else:
            opt[a.lstrip("-").replace("-", "_")] = True
    in_names = ["chrom1", "pos1", "chrom2", "pos2"]
    in_dtypes = {"chrom1": str, "pos1": np.int64, "chrom2": str, "pos2": np.int64}
    in_numbers = dict(nums)
    out_names, out_dtypes, aggregations = [], {}, {}
    for arg in opt["field"]:
        name, colnum, dtype, agg = parse_field_param(arg)
        if colnum is None:
            assert agg is None and dtype is not None and name in {"bin1_id", "bin2_id", "count"}
            out_dtypes[name] = dtype
            continue
        if name not in in_names:
            in_names.append(name)
        if name not in out_names:
            out_names.append(name)
        in_numbers[name] = colnum
        if dtype is not None:
            in_dtypes[name] = out_dtypes[name] = dtype
        aggregations[name] = agg if agg is not None else "sum"
    if "count" not in out_names:
        out_names.append("count")
    order = sorted(in_names, key=in_numbers.get)          # field numbers ascending: names line up with usecols
    reader = pd.read_csv(os.path.join(TEXT, "toy.pairs"), sep="\t", usecols=[in_numbers[n] for n in order],
                         names=order, dtype=in_dtypes, iterator=True, chunksize=opt["chunksize"])
    sanitize = sanitize_records(bins, schema="pairs", decode_chroms=True, is_one_based=not opt["zero_based"],
                                tril_action="reflect", sort=True, validate=True)
    aggregate = aggregate_records(agg=dict(aggregations), count=True, sort=False)
    default = {"count": np.int32}
    chunks = []
    for raw in reader:
        chunk = aggregate(sanitize(raw))
        cols = {"bin1_id": chunk["bin1_id"].to_numpy(np.int64), "bin2_id": chunk["bin2_id"].to_numpy(np.int64)}
        for c in out_names:
            cols[c] = chunk[c].to_numpy().astype(out_dtypes.get(c, default.get(c, float)))
        chunks.append(pd.DataFrame(cols))
    merged = (pd.concat(chunks, axis=0, ignore_index=True).groupby(["bin1_id", "bin2_id"], sort=True)
              .aggregate({c: "sum" for c in out_names}).reset_index())
    return chunks, merged, {c: np.dtype(out_dtypes.get(c, default.get(c, float))) for c in out_names}


def main():
    assert refshim.available()
    out = {}
    for key, (binsize, extra) in PAIRS_CASES.items():
        chunks, merged, dtypes = reference_cload_pairs(binsize, PAIRS_BASE + extra)
        for c in merged.columns:
            v = merged[c].to_numpy()
            out[f"{key}/merged/{c}"] = v.astype(dtypes[c]) if c in dtypes else v
        print(key, sum(len(c) for c in chunks), len(merged), {c: str(out[f"{key}/merged/{c}"].dtype) for c in merged.columns},
              merged.iloc[0].to_dict())
    for key, (fmt, bins_spec, pixels, args) in CASES.items():
        chunks, merged, dtypes = reference_load(fmt, bins_spec, pixels, args)
        cat = pd.concat(chunks, axis=0, ignore_index=True)
        out[f"{key}/chunk_lengths"] = np.array([len(c) for c in chunks])
        for c in cat.columns:
            out[f"{key}/chunks/{c}"] = cat[c].to_numpy()
        for c in merged.columns:
            v = merged[c].to_numpy()
            out[f"{key}/merged/{c}"] = v.astype(dtypes[c]) if c in dtypes else v      # create casts to dtypes
        print(key, len(cat), len(merged), {c: str(merged[c].dtype) for c in merged.columns})
    np.savez_compressed(os.path.join(HERE, "load_golden.npz"), **out)


if __name__ == "__main__":
    main()
