"""Write a cooler to a ``.cool`` / ``.mcool`` file without h5py.

``create_cooler`` follows ``cooler.create.create`` (reference src/cooler/create/_create.py:470-713)
for the part of it the coarsening path uses: same positional / keyword arguments, same on-disk
schema (docs/schema_v3.rst: groups ``chroms``, ``bins``, ``pixels``, ``indexes`` + info
attributes), same chunk validation (``_validate_pixels``, create/_ingest.py:401-432) and the
same derived attributes (``nnz``, ``sum``, ``bin-size`` / ``bin-type``).  The HDF5 bytes are
produced by ``cooler_b200.h5write`` (shuffle + gzip-6 chunked columns like the reference's
``h5opts`` default).  When the real ``cooler`` package + h5py are importable the façade uses
them instead (``reduce.coarsen_cooler``); this module is what makes file output work in an
image that has neither.
"""
from __future__ import annotations

import json
import os
from datetime import datetime

import numpy as np

from . import h5lite, h5write
from .store import parse_cooler_uri

__all__ = ["create_cooler", "create_from_unordered", "copy_cooler", "BadInputError", "MAGIC", "MAGIC_MCOOL"]

MAGIC = "HDF5::Cooler"                    # create/_constants.py:3-5
MAGIC_MCOOL = "HDF5::MCOOL"
URL = "https://github.com/open2c/cooler"
FORMAT_VERSION = 3
FORMAT_VERSION_MCOOL = 2
_PIXEL_DTYPES = {"bin1_id": np.int64, "bin2_id": np.int64, "count": np.int32}


class BadInputError(ValueError):          # create/_ingest.py:38
    pass


def _validate_chunk(chunk, n_bins, boundscheck, triucheck, dupcheck, ensure_sorted):
    """create/_ingest.py:401-432, same messages."""
    b1, b2 = np.asarray(chunk["bin1_id"]), np.asarray(chunk["bin2_id"])
    if boundscheck:
        if np.any((b1 < 0) | (b2 < 0)):
            raise BadInputError("Found bin ID < 0")
        if np.any((b1 >= n_bins) | (b2 >= n_bins)):
            raise BadInputError("Found a bin ID that exceeds the declared number of bins. "
                                "Check whether your bin table is correct.")
    if triucheck and np.any(b1 > b2):
        raise BadInputError("Found bin1_id greater than bin2_id")
    d1 = np.diff(b1)
    in_order = bool(np.all((d1 > 0) | ((d1 == 0) & (b2[1:] >= b2[:-1])))) if len(b1) > 1 else True
    if dupcheck and len(b1) > 1:
        if in_order:                                   # lexsorted chunk (the usual case): duplicates are adjacent
            dup = (d1 == 0) & (b2[1:] == b2[:-1])
            k = np.flatnonzero(dup)[:5] + 1
        else:
            order = np.lexsort((b2, b1))
            s1, s2 = b1[order], b2[order]
            dup = (s1[1:] == s1[:-1]) & (s2[1:] == s2[:-1])
            k = order[1:][dup][:5]
        if dup.any():
            rows = "\n".join(f"{i}\t{b1[i]}\t{b2[i]}" for i in k)
            raise BadInputError("Found duplicate pixels:\n" + rows)
    if ensure_sorted and in_order:
        return chunk
    if ensure_sorted:
        order = np.lexsort((b2, b1))
        chunk = {k: np.asarray(v)[order] for k, v in chunk.items()}
    return chunk


def _bins_arrays(bins):
    """(chromnames, chrom ids, start, end, extra columns) from a bins DataFrame."""
    for col in ("chrom", "start", "end"):
        if col not in bins.columns:
            raise ValueError(f"Missing column from bin table: '{col}'.")
    chrom = bins["chrom"].astype(object).to_numpy()
    names, first = np.unique(chrom, return_index=True)
    names = [str(x) for x in names[np.argsort(first)]]               # order of first appearance
    idmap = {n: i for i, n in enumerate(names)}
    ids = np.fromiter((idmap[str(c)] for c in chrom), dtype=np.int32, count=len(chrom))
    if len(ids) > 1 and np.any(np.diff(ids) < 0):
        raise ValueError("Bin table is not sorted by chromosome")
    extra = {c: bins[c].to_numpy() for c in bins.columns if c not in ("chrom", "start", "end")}
    return names, ids, bins["start"].to_numpy(), bins["end"].to_numpy(), extra


def _binsize(ids, start, end):
    """util.get_binsize (util.py:393-411): one width shared by all but the last bin of each chromosome."""
    if len(ids) == 0:
        return None
    w = (np.asarray(end) - np.asarray(start))
    last = np.r_[ids[1:] != ids[:-1], True]
    sizes = np.unique(w[~last])
    return int(sizes[0]) if len(sizes) == 1 else None


def refuse_truncating_source(dst_path, sources, mode):
    """``mode="w"`` truncates the destination.  When the destination IS one of the source files
    (``zoomify x.mcool::resolutions/1000`` without ``-o``, or ``-o`` equal to the input) that would
    destroy the user's data before it has been read -- the sources are memory-mapped.  h5py refuses
    to truncate a file it has open (the reference therefore fails with an OSError); so do we."""
    if mode != "w":
        return
    dst = os.path.realpath(dst_path)
    for src in sources:
        name = getattr(src, "filename", None)
        if isinstance(name, str) and os.path.exists(name) and os.path.realpath(name) == dst:
            raise OSError(f"Unable to create file '{dst_path}': it is the input file '{name}' "
                          "(unable to truncate a file which is already open); choose another output path")


def _open_target(file_path, group_path, mode):
    """The file and the (emptied) destination group, as create() prepares them (_create.py:611-623)."""
    if mode in ("r+", "a") and not (os.path.exists(file_path) and os.path.getsize(file_path) > 0):
        if mode == "r+":
            raise FileNotFoundError(file_path)
        mode = "w"
    f = h5write.File(file_path, "w" if mode == "w" else "r+")
    if group_path == "/":
        for name in ("chroms", "bins", "pixels", "indexes"):
            if name in f:
                del f[name]
        return f, f
    if group_path in f:
        del f[group_path]
    return f, f.create_group(group_path)


def create_cooler(cool_uri, bins, pixels, columns=None, dtypes=None, metadata=None, assembly=None,
                  symmetric_upper=True, mode=None, h5opts=None, boundscheck=True, triucheck=True,
                  dupcheck=True, ensure_sorted=False, lock=None, append=False, **kwargs):
    """Create a cooler at ``cool_uri`` from a bins DataFrame and an iterable of pixel chunks.

    ``pixels`` is a DataFrame, a dict of arrays, or an iterable of either (the ContactBinner
    protocol, create/_ingest.py:507-528).  Returns ``(nnz, sum)``.
    """
    import pandas as pd
    if kwargs:
        raise TypeError(f"unsupported arguments: {sorted(kwargs)}")
    if not isinstance(bins, pd.DataFrame):
        raise ValueError("Second positional argument must be a pandas DataFrame. "
                         "Note that the `chromsizes` argument is now deprecated: "
                         "see documentation for `create`.")
    file_path, group_path = parse_cooler_uri(cool_uri)
    file_path = os.path.realpath(file_path)
    if mode is None:
        mode = "a" if append else "w"
    opts = {"compression": "gzip", "compression_opts": 6, "shuffle": True}
    opts.update(h5opts or {})
    if columns is None:
        columns = ["bin1_id", "bin2_id", "count"]
    else:
        columns = list(columns)
        for col in ("bin2_id", "bin1_id"):
            if col not in columns:
                columns.insert(0, col)
    dt = dict(_PIXEL_DTYPES)
    dt.update(dict(dtypes or {}))
    for col in columns:
        dt.setdefault(col, float)
    if isinstance(pixels, (pd.DataFrame, dict)):
        for col in columns:
            if col not in pixels:
                kind = "Standard" if col in _PIXEL_DTYPES else "User"
                raise ValueError(f"{kind} column not found in input: '{col}'")
        pixels = (pixels,)

    chromnames, chrom_ids, start, end, extra = _bins_arrays(bins)
    n_chroms, n_bins = len(chromnames), len(chrom_ids)
    last = np.r_[chrom_ids[1:] != chrom_ids[:-1], True] if n_bins else np.zeros(0, bool)
    lengths = np.asarray(end)[last]
    binsize = _binsize(chrom_ids, start, end)
    if not symmetric_upper:
        triucheck = False

    f, h5 = _open_target(file_path, group_path, mode)
    try:
        grp = h5.create_group("chroms")                                        # write_chroms
        grp.create_dataset("name", data=np.array(chromnames, dtype="S"), **opts)
        grp.create_dataset("length", data=np.asarray(lengths, dtype=np.int32), **opts)
        grp = h5.create_group("bins")                                          # write_bins
        enum = {n: i for i, n in enumerate(chromnames)}
        try:
            grp.create_dataset("chrom", data=chrom_ids, enum=enum, **opts)
        except h5write.H5FormatError:                                          # too many contigs for an enum header
            grp.create_dataset("chrom", data=chrom_ids, attrs={"enum_path": "/chroms/name"}, **opts)
        grp.create_dataset("start", data=np.asarray(start, dtype=np.int32), **opts)
        grp.create_dataset("end", data=np.asarray(end, dtype=np.int32), **opts)
        for k, v in extra.items():
            grp.create_dataset(k, data=v, **opts)
        grp = h5.create_group("pixels")                                        # prepare_pixels / write_pixels
        streams = {c: grp.create_dataset(c, dtype=dt[c], **opts) for c in columns}
        nnz, total = 0, 0
        row_counts = np.zeros(n_bins + 1, dtype=np.int64)
        prev_last = -1
        for chunk in pixels:
            if isinstance(chunk, pd.DataFrame):
                chunk = {k: v.values for k, v in chunk.items()}
            if boundscheck or triucheck or dupcheck or ensure_sorted:
                chunk = _validate_chunk(chunk, n_bins, boundscheck, triucheck, dupcheck, ensure_sorted)
            b1 = np.asarray(chunk["bin1_id"])
            n = len(b1)
            if n == 0:
                continue
            if b1[0] < prev_last or np.any(b1[1:] < b1[:-1]):
                raise BadInputError("Pixel chunks must be sorted by bin1_id (use ensure_sorted=True)")
            prev_last = int(b1[-1])
            row_counts[1:] += np.bincount(b1, minlength=n_bins)[:n_bins]
            for c in columns:
                streams[c].append(np.asarray(chunk[c]))
            nnz += n
            if "count" in chunk:
                total += np.asarray(chunk["count"]).sum()
        for s in streams.values():
            s.close()
        grp = h5.create_group("indexes")                                       # write_indexes
        chrom_offset = np.searchsorted(chrom_ids, np.arange(n_chroms + 1)).astype(np.int64)
        grp.create_dataset("chrom_offset", data=chrom_offset, **opts)
        grp.create_dataset("bin1_offset", data=np.cumsum(row_counts), **opts)
        total = total.item() if isinstance(total, np.generic) else total
        info = {                                                               # _create.py:700-713, write_info
            "bin-type": "fixed" if binsize is not None else "variable",
            "bin-size": binsize if binsize is not None else "null",
            "storage-mode": "symmetric-upper" if symmetric_upper else "square",
            "nchroms": n_chroms, "nbins": n_bins, "sum": total, "nnz": nnz,
            "genome-assembly": assembly if assembly is not None else "unknown",
            "metadata": json.dumps(metadata if metadata is not None else {}),
            "creation-date": datetime.now().isoformat(),
            "generated-by": "cooler_b200",
            "format": MAGIC, "format-version": FORMAT_VERSION, "format-url": URL,
        }
        h5.attrs.update(info)
    finally:
        f.close()
    return nnz, total


def create_from_unordered(cool_uri, bins, chunks, columns=None, dtypes=None, mode=None, mergebuf=20_000_000,
                          delete_temp=True, temp_dir=None, max_merge=200, device="cuda", **kwargs):
    """Create a cooler from pixel chunks given in ANY order (``cooler.create.create_from_unordered``,
    create/_create.py:716-812; the back end of ``cooler load``).

    The reference writes every chunk, sorted, to a temporary cooler (``create(...)`` with its chunk checks)
    and k-way merges the temporary files with ``CoolerMerger`` (every value column summed).  Here every
    chunk gets the same checks -- bounds and the upper-triangle rule on the host, duplicates inside a chunk by
    a sort + run count on the device --, is parked on the GPU, and one device-wide stable sort + reduce-by-key
    merges all of them (``ingest.UnorderedPixels``).  As in the reference, value columns are cast to their
    output ``dtypes`` chunk by chunk (the temporary coolers store them that way) before they are summed.
    ``mergebuf`` sets the chunking of the output; ``delete_temp``, ``temp_dir`` and ``max_merge`` are
    accepted and ignored (nothing is spilled to disk)."""
    import pandas as pd

    from .ingest import UnorderedPixels
    value_cols = [c for c in (["count"] if columns is None else list(columns)) if c not in ("bin1_id", "bin2_id")]
    dt = dict(_PIXEL_DTYPES)
    dt.update(dict(dtypes or {}))
    for c in value_cols:
        dt.setdefault(c, float)
    boundscheck = kwargs.pop("boundscheck", True)
    triucheck = kwargs.pop("triucheck", True) and kwargs.get("symmetric_upper", True)
    dupcheck = kwargs.pop("dupcheck", True)
    kwargs.pop("ensure_sorted", None)
    n_bins = len(bins)
    acc = UnorderedPixels(n_bins, device=device)
    parts = {c: [] for c in value_cols}
    for chunk in chunks:
        if isinstance(chunk, pd.DataFrame):
            chunk = {k: v.values for k, v in chunk.items()}
        for col in ("bin1_id", "bin2_id", *value_cols):
            if col not in chunk:
                kind = "Standard" if col in _PIXEL_DTYPES else "User"
                raise ValueError(f"{kind} column not found in input: '{col}'")
        _validate_chunk(chunk, n_bins, boundscheck, triucheck, False, False)
        acc.add(np.asarray(chunk["bin1_id"]), np.asarray(chunk["bin2_id"]), dupcheck=dupcheck)
        for c in value_cols:
            parts[c].append(np.asarray(chunk[c]).astype(dt[c], copy=False))
    cat = {c: (np.concatenate(parts[c]) if parts[c] else np.zeros(0, dtype=dt[c])) for c in value_cols}
    b1, b2, vals = acc.merge(cat)

    def merged():
        step = max(int(mergebuf), 1)
        for lo in range(0, len(b1), step):
            out = {"bin1_id": b1[lo:lo + step], "bin2_id": b2[lo:lo + step]}
            out.update({c: vals[c][lo:lo + step] for c in value_cols})
            yield out

    return create_cooler(cool_uri, bins, merged(), columns=value_cols, dtypes=dt, mode=mode, boundscheck=False,
                         triucheck=False, dupcheck=False, ensure_sorted=False, **kwargs)


def copy_cooler(src, dst_uri, columns=("count",), mode="w"):
    """Copy one cooler (chroms, bins, pixels, indexes, attributes) into ``dst_uri``.

    The base-resolution copy of ``zoomify_cooler`` (_reduce.py:838-854) and, with
    ``columns=None`` (every pixel column), of ``legacy_zoomify`` (909-914).  ``src`` is a
    Cooler-like object: columns of a file read by h5lite are copied chunk by chunk without
    re-encoding (like ``h5py.Group.copy``); other sources are read and encoded.
    """
    file_path, group_path = parse_cooler_uri(dst_uri)
    opts = {"compression": "gzip", "compression_opts": 6, "shuffle": True}
    refuse_truncating_source(file_path, [src], mode)
    f, h5 = _open_target(os.path.realpath(file_path), group_path, mode)
    try:
        with src.open("r") as g:
            for grp_name in ("chroms", "bins", "pixels", "indexes"):
                out = h5.create_group(grp_name)
                names = list(g[grp_name].keys())
                if grp_name == "pixels" and columns is not None:
                    names = [c for c in names if c in ("bin1_id", "bin2_id", *columns)]
                for k in names:
                    d = g[grp_name][k]
                    if isinstance(d, h5lite.Dataset):                    # file source: raw chunk copy (h5py's src.copy)
                        out.copy_dataset(d, k)
                        continue
                    enum = getattr(d, "enum", None)
                    if enum is None and grp_name == "bins" and k == "chrom" and np.asarray(d[:]).dtype.kind in "iu":
                        enum = {n: i for i, n in enumerate(src.chromnames)}      # write_bins: chrom ids as an enum
                    attrs = dict(getattr(d, "attrs", {}) or {})
                    out.create_dataset(k, data=np.asarray(d[:]), enum=enum, attrs=attrs or None, **opts)
            attrs = dict(getattr(g, "attrs", {}) or {})
        attrs = {k: v for k, v in attrs.items() if v is not None or k == "bin-size"}
        h5.attrs.update(attrs)
    finally:
        f.close()
