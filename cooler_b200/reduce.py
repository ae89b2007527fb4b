"""Coarsening / zoomify on B200 -- mirrors open2c/cooler v0.10.4
``src/cooler/_reduce.py`` (``CoolerCoarsener``, ``coarsen_cooler``,
``zoomify_cooler`` and the resolution planners).

The reference re-bins every pixel chunk by joining it with the bin table and
running a pandas ``groupby([bin1_id, bin2_id], sort=True).sum()``
(_reduce.py:582-620).  Here one level is: the fine rows of a coarse row are
merged pairwise (merge path, the bin remap fused into the first round), then
one reduce-by-key (csrc/coarsen.cu), all on the device-resident table; the
result is again a ``PixelTable`` so a zoomify chain (and per-level balancing)
never leaves HBM.  With ``group=`` (one rank per GPU) every rank coarsens the
stored pixels of its own row range, cut at coarse-bin boundaries so that no
exchange is needed; rank 0 gathers the pieces in rank order and writes.
The bin id map is the closed form of the reference's coordinate arithmetic:
``new_id = new_chrom_offset[c] + (old_id - old_chrom_offset[c]) // factor``.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .store import MemCooler, parse_cooler_uri
from .table import (PixelTable, TableCopy, _ptr, _stream, _use_device, exclusive_scan_i32,
                    indptr_from_sorted, key_bits_for, n_tiles, sort_pairs, swap_key_other)

HIGLASS_TILE_DIM = 256      # _reduce.py:28

__all__ = ["CoolerCoarsener", "coarsen_cooler", "zoomify_cooler", "coarsen_table",
           "CoolerMerger", "merge_coolers", "merge_tables",
           "get_multiplier_sequence", "preferred_sequence", "ZOOMS_4DN", "get_quadtree_depth",
           "legacy_zoomify"]

# Resolutions of the 4DN multires standard (_reduce.py:30-44).
ZOOMS_4DN = [1000, 2000, 5000, 10000, 25000, 50000, 100000, 250000, 500000,
             1000000, 2500000, 5000000, 10000000]

_I32, _I64 = torch.int32, torch.int64


# --------------------------------------------------------------------------
# host-side planning (pure integer logic)
# --------------------------------------------------------------------------
def coarsen_chrom_offset(chrom_offset, factor):
    """New ``chrom_offset``: each chromosome keeps ceil(n_bins_chrom / factor) bins."""
    nb = np.diff(np.asarray(chrom_offset, dtype=np.int64))
    return np.r_[0, np.cumsum(-(-nb // factor))].astype(np.int64)


def coarsen_bin_map(chrom_offset, factor):
    """int32 map old bin id -> new bin id (closed form of _reduce.py:597-614)."""
    chrom_offset = np.asarray(chrom_offset, dtype=np.int64)
    new_off = coarsen_chrom_offset(chrom_offset, factor)
    nb = np.diff(chrom_offset)
    old = np.arange(chrom_offset[-1], dtype=np.int64)
    start = np.repeat(chrom_offset[:-1], nb)
    return (np.repeat(new_off[:-1], nb) + (old - start) // factor).astype(np.int32), new_off


def coarsen_bins(bins_chrom, bins_start, bins_end, chromsizes, factor):
    """New bin table (semantics of ``CoolerCoarsener.coarsen_bins``, _reduce.py:563-580):
    per chromosome every ``factor``-th bin start; the end is the end of the last
    pooled bin, or the chromosome length for a trailing partial group."""
    bins_chrom = np.asarray(bins_chrom)
    bins_start = np.asarray(bins_start, dtype=np.int64)
    bins_end = np.asarray(bins_end, dtype=np.int64)
    n_chroms = len(chromsizes)
    off = np.searchsorted(bins_chrom, np.arange(n_chroms + 1))
    cs, ss, es = [], [], []
    for c in range(n_chroms):
        lo, hi = off[c], off[c + 1]
        if hi == lo:
            continue
        s = bins_start[lo:hi:factor]
        e = bins_end[lo + factor - 1:hi:factor]
        if len(e) < len(s):
            e = np.append(e, int(chromsizes[c]))
        cs.append(np.full(len(s), c, dtype=np.int32))
        ss.append(s)
        es.append(e)
    if not cs:
        z = np.zeros(0, dtype=np.int64)
        return z.astype(np.int32), z, z
    return np.concatenate(cs), np.concatenate(ss), np.concatenate(es)


def _prune_edges(edges, maxlen):
    """Thin a partition of 0..nnz to pieces of about ``maxlen``
    (role of ``_greedy_prune_partition``, _reduce.py:310-320)."""
    edges = np.asarray(edges, dtype=np.int64)
    total = int(edges[-1])
    cuts = list(range(0, total, max(int(maxlen), 1))) + [total]
    idx = np.unique(np.searchsorted(edges, cuts))
    return edges[idx]


def get_quadtree_depth(chromsizes, base_binsize, bins_per_tile):
    """Zoom levels needed to coarsen down to a single tile (_reduce.py:323-347)."""
    tile_bp = bins_per_tile * base_binsize
    n_tiles_ = math.ceil(sum(chromsizes) / tile_bp)
    return math.ceil(np.log2(n_tiles_))


def preferred_sequence(start, stop, style="nice"):
    """1-2-5 ("nice") or doubling ("binary") progression from ``start`` up to
    ``stop`` inclusive (_reduce.py:382-440)."""
    if start > stop:
        return []
    if style not in ("nice", "binary"):
        raise ValueError(f"Expected style value of 'binary' or 'nice'; got '{style}'.")
    seq, cur = [], int(start)
    if style == "binary":
        while cur <= stop:
            seq.append(cur)
            cur *= 2
        return seq
    decade = int(start)
    while True:
        for m in (1, 2, 5):
            v = decade * m
            if v > stop:
                return seq
            seq.append(v)
        decade *= 10


def get_multiplier_sequence(resolutions, bases=None):
    """For each target resolution pick the largest smaller resolution that divides it
    (_reduce.py:443-498).  Returns (resn, pred, mult) arrays; pred == -1 marks a base."""
    bases = {min(resolutions)} if bases is None else set(bases)
    resn = np.array(sorted(bases.union(resolutions)))
    pred = np.full(len(resn), -1, dtype=int)
    mult = np.full(len(resn), -1, dtype=int)
    for i in range(len(resn) - 1, -1, -1):
        for p in range(i - 1, -1, -1):
            if resn[i] % resn[p] == 0:
                pred[i], mult[i] = p, resn[i] // resn[p]
                break
    for i, p in enumerate(pred):
        if p == -1 and resn[i] not in bases:
            raise ValueError(
                f"Resolution {resn[i]} cannot be derived from the base resolutions: {bases}.")
    return resn, pred, mult


def align_rows_to_factor(chrom_offset, factor, row):
    """Largest row <= ``row`` that starts a coarse bin (a multiple of ``factor`` counted from
    its chromosome's first bin).  Shards cut at such rows coarsen independently: no coarse
    row is split between ranks (the reference's chunk edges obey the same rule,
    _reduce.py:550-561)."""
    chrom_offset = np.asarray(chrom_offset, dtype=np.int64)
    row = int(row)
    if row >= chrom_offset[-1]:
        return int(chrom_offset[-1])
    c = int(np.searchsorted(chrom_offset, row, side="right")) - 1
    return int(chrom_offset[c] + (row - chrom_offset[c]) // factor * factor)


# --------------------------------------------------------------------------
# device path
# --------------------------------------------------------------------------
def _coarse_maps(table: PixelTable, factor: int):
    """Device bin maps of one coarsening step: (bin_map, group_first, n_new, new chrom_offset, number of
    merge rounds, largest number of fine rows per coarse row)."""
    factor = int(factor)
    assert factor > 1
    dev = _use_device(table.device)
    csr = table.raw
    new_off = coarsen_chrom_offset(table.chrom_offset, factor)
    n_new = int(new_off[-1])
    if csr.row_base != 0 or csr.n_rows != table.n_bins:
        raise ValueError("coarsen_table needs the whole-row-range raw copy of the table")
    # old bin -> new bin and the first old bin of every new bin, on the device (closed forms)
    n_chroms = len(new_off) - 1
    offs = torch.from_numpy(np.stack([np.asarray(table.chrom_offset, dtype=np.int64), new_off])).to(dev)
    bin_map_d = torch.empty(max(table.n_bins, 1), dtype=_I32, device=dev)
    group_first_d = torch.empty(n_new + 1, dtype=_I32, device=dev)
    _lib.call("cb200_coarsen_maps", _ptr(offs[0]), _ptr(offs[1]), n_chroms, factor, table.n_bins, n_new,
              _ptr(bin_map_d), _ptr(group_first_d), _stream())
    nb = np.diff(table.chrom_offset)
    max_group = int(min(factor, nb.max())) if len(nb) and nb.max() > 0 else 1
    rounds = max(1, int(max_group - 1).bit_length())
    # column remap of round 0 (cb200_remap): block table + division by multiplication when it applies
    remap = _lib.Remap()
    remap.bin_map = bin_map_d.data_ptr()
    if table.n_bins > 0 and n_new * factor < 2**31:
        shift = max(4, int((table.n_bins - 1) >> 11).bit_length())        # <= 2048 blocks: the table stays in L1
        n_blocks = ((table.n_bins - 1) >> shift) + 1
        blocks_d = torch.empty((n_blocks, 4), dtype=_I32, device=dev)
        _lib.call("cb200_coarsen_block_table", _ptr(offs[0]), _ptr(offs[1]), n_chroms, factor, table.n_bins, shift,
                  _ptr(blocks_d), _stream())
        lg = int(factor - 1).bit_length()
        remap.blocks, remap.block_shift = blocks_d.data_ptr(), shift
        remap.magic, remap.magic_shift = -(-(1 << (31 + lg)) // factor), lg - 1
        remap._keep = (bin_map_d, blocks_d)
    return bin_map_d, group_first_d, n_new, new_off, rounds, max_group, remap


def _merge_rounds(csr, maps, n_rounds, keys=None):
    """Rounds 0 .. n_rounds-1 of the pairwise merges (no reduction); returns the buffer holding the result
    (the input copy's own pixels when n_rounds == 0).  ``keys``: filled with the coarse row of every
    element by the last round run here."""
    bin_map_d, group_first_d, n_new, _, _, max_group, remap = maps
    dev = csr.px.device
    n = csr.nnz
    claim = torch.zeros(1, dtype=_I64, device=dev)
    bufs = [torch.empty((n + 2, 2), dtype=_I32, device=dev) for _ in range(min(n_rounds, 2))]
    src = csr.px
    for r in range(n_rounds):
        dst = bufs[r % 2]
        lists = -(-max_group // (1 << r))
        _lib.call("cb200_coarsen_merge_round", _ptr(src), _ptr(dst), _ptr(csr.indptr), _ptr(group_first_d),
                  n_new, r, -(-lists // 2), C.byref(remap) if r == 0 else None,
                  _ptr(keys) if (keys is not None and r == n_rounds - 1) else C.c_void_p(0), _ptr(claim), _stream())
        src = dst
    return src


def _merge_coarse_rows(table: PixelTable, factor: int):
    """The merge part of K5 without the reduction (general aggregators): returns (keys = coarse row of every
    pixel, vals = {coarse column, payload} sorted by (coarse row, coarse column) with equal pairs adjacent in
    pixel order, n_new, new chrom_offset).  The payload is whatever the table's raw copy carries as "count"."""
    maps = _coarse_maps(table, factor)
    csr = table.raw
    keys = torch.empty(max(csr.nnz, 1), dtype=_I32, device=csr.px.device)
    vals = _merge_rounds(csr, maps, maps[4], keys=keys)
    return keys, vals, maps[2], maps[3]


def coarsen_table(table: PixelTable, factor: int):
    """Coarsen a device-resident table by ``factor``.  Returns (new PixelTable, bin1' ids).

    The new table's raw copy is sorted by (bin1', bin2') and duplicate-free, with
    counts summed (int64 accumulation; OverflowError if a sum leaves int32).

    No global sort: the fine rows of a coarse row are already sorted by the (monotone) coarse
    column, so the coarse row is their multi-way merge -- ceil(log2(rows per coarse row)) rounds of
    pairwise merge-path merges; the LAST round sums adjacent equal (bin1', bin2') on the fly
    (``cb200_coarsen_merge_reduce``) and writes every reduced row at its input offset, a scan of the
    row lengths is the new row pointer, and ``cb200_coarsen_compact`` moves the rows into place."""
    dev = table.device
    csr = table.raw
    n = csr.nnz
    maps = _coarse_maps(table, factor)
    bin_map_d, group_first_d, n_new, new_off, rounds, _, remap = maps
    src = _merge_rounds(csr, maps, rounds - 1)
    tmp = torch.empty((n + 2, 2), dtype=_I32, device=dev)
    row_count = torch.zeros(max(n_new, 1), dtype=_I32, device=dev)
    overflow = torch.zeros(1, dtype=_I32, device=dev)
    claim = torch.zeros(1, dtype=_I64, device=dev)
    _lib.call("cb200_coarsen_merge_reduce", _ptr(src), _ptr(tmp), _ptr(csr.indptr), _ptr(group_first_d), n_new,
              rounds - 1, C.byref(remap) if rounds == 1 else None, _ptr(row_count), _ptr(overflow),
              _ptr(claim), _stream())
    del src
    indptr = exclusive_scan_i32(row_count[:n_new])            # the new row pointer
    n_out = int(indptr[-1].item())
    k2 = torch.empty(max(n_out, 1), dtype=_I32, device=dev)
    v2 = torch.empty((n_out + 2, 2), dtype=_I32, device=dev)
    _lib.call("cb200_coarsen_compact", _ptr(tmp), _ptr(csr.indptr), _ptr(group_first_d), _ptr(indptr), n_new,
              _ptr(v2), _ptr(k2), _stream())
    if int(overflow.item()):
        raise OverflowError("coarsened count does not fit int32")
    new_raw = TableCopy(v2, indptr, n_out, n_new)
    return PixelTable(new_raw, new_off), k2[:n_out]


AGG_OPS = {"sum": 0, "min": 1, "max": 2, "first": 3, "last": 4, "mean": 5}
_SRC_DTYPES = {torch.int32: 0, torch.int64: 1, torch.float32: 2, torch.float64: 3}


def coarsen_columns(bin1_offset, bin2_id, chrom_offset, factor, columns, agg, device="cuda"):
    """Coarsening with arbitrary value columns and aggregators -- the general form of
    ``CoolerCoarsener._aggregate`` (_reduce.py:582-620: ``groupby([bin1_id, bin2_id], sort=True)
    .aggregate(agg)``).  ``columns``: {name: host array, any numeric dtype}; ``agg``: {name: "sum" |
    "min" | "max" | "first" | "last" | "mean"}.  The pixels are merged once carrying their own
    position as payload; every column is then gathered through that order and reduced per
    (bin1', bin2') run in pixel order, integers in int64, floats in float64 (mean always float64).
    Returns (bin1' int64, bin2' int64, {name: numpy array}, new chrom_offset)."""
    dev = _use_device(device)
    n = len(bin2_id)
    if n >= 2**31 - 1:
        raise NotImplementedError("the general aggregation path indexes pixels with int32")
    idx = torch.arange(n, dtype=_I32)
    table = PixelTable.from_arrays(bin1_offset, bin2_id, idx, chrom_offset, device=dev)
    keys, vals, n_new, new_off = _merge_coarse_rows(table, factor)
    b1, b2, out = _aggregate_runs(keys, vals, n, columns, agg, dev)
    return (b1, b2, out, new_off)


def _reduce_runs(keys, vals, n, swap, dev):
    """Reduce-by-key over adjacent equal (key, val.other) pairs; returns (key', val', n_out)."""
    nt = n_tiles(n)
    tile_count = torch.zeros(max(nt, 1), dtype=_I32, device=dev)
    _lib.call("cb200_runs_count", _ptr(keys), _ptr(vals), n, _ptr(tile_count), _stream())
    tile_off = exclusive_scan_i32(tile_count[:nt])
    n_out = int(tile_off[-1].item())
    k2 = torch.empty(max(n_out, 1), dtype=_I32, device=dev)
    v2 = torch.empty((n_out + 2, 2), dtype=_I32, device=dev)
    overflow = torch.zeros(1, dtype=_I32, device=dev)
    _lib.call("cb200_runs_write", _ptr(keys), _ptr(vals), n, _ptr(tile_off), int(swap), _ptr(k2), _ptr(v2),
              C.c_void_p(0), _ptr(overflow), _stream())
    return k2, v2, n_out, overflow


def _concat_sorted(tables):
    """Concatenate device-resident tables over identical bins in the given order and bring the pixels
    into (bin1, bin2) order by two stable radix sorts (by bin2, then by bin1): pixels of one (bin1, bin2)
    stay in table order.  Returns (keys = bin1, vals = {bin2, count}, n)."""
    n_bins = tables[0].n_bins
    dev = _use_device(tables[0].device)
    ident = torch.arange(n_bins, dtype=_I32, device=dev)
    keys, vals = [], []
    for t in tables:
        if t.n_bins != n_bins:
            raise ValueError("tables must share one bin segmentation")
        csr = t.raw
        key = torch.empty(max(csr.nnz, 1), dtype=_I32, device=dev)
        val = torch.empty((csr.nnz + 2, 2), dtype=_I32, device=dev)
        cs = csr.c_struct()
        _lib.call("cb200_coarsen_remap", C.byref(cs), csr.n_rows, _ptr(ident), _ptr(key), _ptr(val), _stream())
        keys.append(key[: csr.nnz])
        vals.append(val[: csr.nnz])
    n = int(sum(k.numel() for k in keys))
    key = torch.cat(keys)
    val = torch.cat(vals + [torch.zeros((2, 2), dtype=_I32, device=dev)])
    del keys, vals
    bits = key_bits_for(n_bins)
    ks, vs = sort_pairs(key, val, bits)                       # by bin2
    del key, val
    k2, v2 = swap_key_other(ks, vs, n)                         # key = bin1, val = {bin2, count}
    del ks, vs
    k3, v3 = sort_pairs(k2, v2, bits)                          # stable by bin1 -> (bin1, bin2) order
    return k3, v3, n


def merge_tables(tables):
    """Sum several device-resident tables over identical bins (the k-way merge of
    ``CoolerMerger``, _reduce.py:132-208): concatenate, stable radix sort by bin2 then by
    bin1, reduce-by-key.  Returns (PixelTable, bin1 ids)."""
    n_bins = tables[0].n_bins
    dev = _use_device(tables[0].device)
    k3, v3, n = _concat_sorted(tables)
    k4, v4, n_out, overflow = _reduce_runs(k3, v3, n, 0, dev)
    del k3, v3
    if int(overflow.item()):
        raise OverflowError("merged count does not fit int32")
    indptr = indptr_from_sorted(k4, n_out, n_bins)
    return PixelTable(TableCopy(v4, indptr, n_out, n_bins), tables[0].chrom_offset), k4[:n_out]


def _aggregate_runs(keys, vals, n, columns, agg, dev):
    """Reduce every (key, vals.other) run of a sorted stream whose payload ``vals.count`` is the pixel's
    position in the value columns: one ``cb200_runs_aggregate`` per column, in stream order inside a run.
    Returns (bin1 int64, bin2 int64, {name: numpy array; integers in int64, floats / means in float64})."""
    nt = n_tiles(n)
    tile_count = torch.zeros(max(nt, 1), dtype=_I32, device=dev)
    _lib.call("cb200_runs_count", _ptr(keys), _ptr(vals), n, _ptr(tile_count), _stream())
    tile_off = exclusive_scan_i32(tile_count[:nt])
    n_out = int(tile_off[-1].item())
    b1 = torch.empty(max(n_out, 1), dtype=_I32, device=dev)
    b2 = torch.empty(max(n_out, 1), dtype=_I32, device=dev)
    out = {}
    first = True
    for name, col in columns.items():
        op = agg.get(name, "sum")
        if callable(op) or op not in AGG_OPS:
            raise NotImplementedError(f"aggregator {op!r} is not supported by the CUDA path "
                                      f"(supported: {sorted(AGG_OPS)})")
        a = np.ascontiguousarray(col)
        t = torch.from_numpy(a if a.flags.writeable else a.copy())      # read-only views (npz, memmap) are copied
        if t.dtype not in _SRC_DTYPES:
            t = t.to(torch.float64 if t.dtype.is_floating_point else torch.int64)
        src = t.to(dev)
        is_float = src.dtype.is_floating_point or op == "mean"
        res = torch.empty(max(n_out, 1), dtype=torch.float64 if is_float else _I64, device=dev)
        _lib.call("cb200_runs_aggregate", _ptr(keys), _ptr(vals), n, _ptr(tile_off), _ptr(src),
                  _SRC_DTYPES[src.dtype], AGG_OPS[op], _ptr(b1) if first else C.c_void_p(0),
                  _ptr(b2) if first else C.c_void_p(0), _ptr(res) if not is_float else C.c_void_p(0),
                  _ptr(res) if is_float else C.c_void_p(0), _stream())
        out[name] = res[:n_out].cpu().numpy()
        first = False
    return b1[:n_out].to(_I64).cpu().numpy(), b2[:n_out].to(_I64).cpu().numpy(), out


def _settle_dtypes(vals, in_dtypes, agg):
    """Column dtypes as pandas' groupby gives them: sum / min / max / first / last keep the input dtype
    (an integer sum that leaves it stays int64), the mean is float64 (float32 for a float32 column)."""
    for c, v in vals.items():
        want = np.dtype(in_dtypes[c])
        if agg.get(c, "sum") == "mean":
            if want == np.float32:
                vals[c] = v.astype(np.float32)
            continue
        if want.kind in "iu" and len(v) and (v.min() < np.iinfo(want).min or v.max() > np.iinfo(want).max):
            continue
        vals[c] = v.astype(want, copy=False)
    return vals


def merge_columns(inputs, chrom_offset, columns, agg, device="cuda"):
    """Merge pixel tables over identical bins with arbitrary value columns and aggregators -- the
    general form of ``CoolerMerger.__iter__`` (_reduce.py:188-203: ``pd.concat`` of the inputs in cooler
    order, ``groupby([bin1_id, bin2_id], sort=True).aggregate(agg)``).  ``inputs``: one
    ``(bin1_offset, bin2_id, {name: host array})`` per cooler.  The tables are merged once carrying the
    pixel's position in the concatenated columns as payload (stable sorts keep the cooler order inside a
    group); every column is then reduced per (bin1, bin2) run.  Numeric dtypes promote like
    ``pd.concat``.  Returns (bin1 int64, bin2 int64, {name: numpy array})."""
    dev = _use_device(device)
    sizes = [len(b2) for _, b2, _ in inputs]
    n = int(sum(sizes))
    if n >= 2**31 - 1:
        raise NotImplementedError("the general aggregation path indexes pixels with int32")
    tables, base = [], 0
    for (off, b2, _), m in zip(inputs, sizes):
        idx = torch.arange(base, base + m, dtype=_I32)
        tables.append(PixelTable.from_arrays(off, b2, idx, chrom_offset, device=dev))
        base += m
    cat, in_dt = {}, {}
    for name in columns:
        arrs = [np.asarray(cols[name]) for _, _, cols in inputs]
        in_dt[name] = np.result_type(*[a.dtype for a in arrs])
        cat[name] = np.concatenate([a.astype(in_dt[name], copy=False) for a in arrs]) if arrs else np.zeros(0)
    if n == 0:
        empty = np.zeros(0, dtype=np.int64)
        return empty, empty, {c: cat[c][:0] for c in columns}
    keys, vals, n = _concat_sorted(tables)
    del tables
    b1, b2, out = _aggregate_runs(keys, vals, n, cat, agg, dev)
    return b1, b2, _settle_dtypes(out, in_dt, agg)


def table_to_host(table: PixelTable, bin1_ids=None):
    """(bin1_id int64, bin2_id int64, count int32) host arrays of a table's raw copy."""
    raw = table.raw
    n = raw.nnz
    dev = raw.device
    if bin1_ids is None:
        ind = raw.indptr
        bin1_ids = torch.repeat_interleave(torch.arange(raw.n_rows, device=dev, dtype=_I32),
                                           (ind[1:] - ind[:-1]))
    b1 = torch.empty(n, dtype=_I64, device=dev)
    b2 = torch.empty(n, dtype=_I64, device=dev)
    cnt = torch.empty(n, dtype=_I32, device=dev)
    _lib.call("cb200_unpack_pixels", _ptr(bin1_ids), _ptr(raw.px), n, _ptr(b1), _ptr(b2), _ptr(cnt), _stream())
    return b1.cpu().numpy(), b2.cpu().numpy(), cnt.cpu().numpy()


def _check_columns(clr, columns, agg):
    columns = ["count"] if columns is None else list(columns)
    input_dtypes = clr.pixels().dtypes
    for col in columns:
        if col not in input_dtypes:
            raise ValueError(
                f"Pixel value column '{col}' not found in input '{clr.filename}'.")
    for col, op in (agg or {}).items():
        if callable(op) or op not in AGG_OPS:
            raise NotImplementedError(f"aggregator {op!r} for column '{col}' is not supported by the CUDA "
                                      f"coarsener (supported: {sorted(AGG_OPS)})")
    return columns


def shard_rows_for_coarsening(clr, group, align):
    """Row range [lo, hi) of this rank for sharded coarsening: about nnz / world stored pixels
    (table.shard_rows_by_pixels), both ends moved down to a boundary of ``align`` pooled bins inside
    their chromosome (``align_rows_to_factor``), so that every coarse row -- of this level and of the
    levels chained after it, when ``align`` is their common multiple -- lives on exactly one rank."""
    import torch.distributed as dist
    from .table import shard_rows_by_pixels
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    off = np.asarray(clr._load_dset("indexes/bin1_offset"))
    chrom_offset = np.asarray(clr._load_dset("indexes/chrom_offset"))
    lo, hi = shard_rows_by_pixels(off, rank, world)
    lo = align_rows_to_factor(chrom_offset, align, lo) if rank > 0 else 0
    hi = align_rows_to_factor(chrom_offset, align, hi) if rank + 1 < world else len(off) - 1
    return lo, max(hi, lo)


def gather_pixels_to_rank0(arrays, group):
    """Concatenate equally long 1-D numpy arrays of every rank in rank order on group rank 0 (the
    others get empty arrays).  Padded ``dist.gather`` on the device for NCCL, on the host for gloo."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    n = len(arrays[0])
    sizes = [torch.zeros(1, dtype=_I64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n], dtype=_I64, device=dev), group=group)
    sizes = [int(x.item()) for x in sizes]
    dst = dist.get_global_rank(group, 0) if group is not None else 0
    out = []
    for a in arrays:
        a = np.ascontiguousarray(a)
        pad = torch.zeros(max(max(sizes), 1), dtype=torch.from_numpy(a[:0].copy()).dtype, device=dev)
        pad[:n] = torch.from_numpy(a).to(dev)
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, bufs, dst=dst, group=group)
        if rank == 0:
            out.append(torch.cat([b[:k] for b, k in zip(bufs, sizes)]).cpu().numpy())
        else:
            out.append(a[:0])
    return out


class CoolerCoarsener:
    """GPU replacement of ``cooler._reduce.CoolerCoarsener`` (_reduce.py:501-642).

    A ContactBinner-style iterable (create/_ingest.py:507-528): yields dict chunks
    ``{bin1_id, bin2_id, count}`` lexsorted by (bin1_id, bin2_id), globally sorted
    and duplicate-free across chunks, of about ``chunksize`` input pixels each and
    never splitting a coarse row.  ``new_bins`` is the coarsened bin table.
    """

    def __init__(self, source, factor, chunksize, columns=None, agg=None, batchsize=1, map=map,
                 device="cuda", table=None, group=None, align=None):
        assert isinstance(factor, int) and factor > 1
        self.group = group
        self.clr = source
        self.factor = factor
        self.chunksize = int(chunksize)
        self.columns = _check_columns(source, columns, agg)
        self.agg = {col: "sum" for col in self.columns}           # _reduce.py:529-531
        if agg is not None:
            self.agg.update(agg)
        dt = source.pixels().dtypes
        # the fast path carries an int32 count through the merge; anything else (other columns, other
        # aggregators, floating-point or wide counts) goes through the general index-payload path
        self._fast = (self.columns == ["count"] and self.agg["count"] == "sum"
                      and np.dtype(dt["count"]).kind in "iu" and np.dtype(dt["count"]).itemsize <= 4)
        self.old_binsize = source.binsize
        self.new_binsize = None if self.old_binsize is None else self.old_binsize * factor
        old_bins = source.bins()[:]
        chromsizes = source.chromsizes
        c, s, e = coarsen_bins(np.asarray(source._load_dset("bins/chrom")), old_bins["start"].values,
                               old_bins["end"].values, chromsizes.values, factor)
        self.new_bins_arrays = (c, s, e)
        self._chromnames = list(chromsizes.index)
        self._device = device
        self._align = int(align or factor)
        if table is None and self._fast and group is not None:
            # this rank's row range of the source, cut at boundaries of `align` pooled bins
            from .table import read_cooler_arrays
            rows = shard_rows_for_coarsening(source, group, int(align or factor))
            off, b2, cnt, co = read_cooler_arrays(source, rows=rows)
            table = PixelTable.from_arrays(off, b2, cnt, co, device=device, row_range=rows)
        self._table = table if (table is not None or not self._fast) else PixelTable.from_cooler(source, device=device)
        self.result_table = None

    @property
    def new_bins(self):
        import pandas as pd
        c, s, e = self.new_bins_arrays
        chrom = pd.Categorical.from_codes(c, categories=self._chromnames, ordered=True)
        return pd.DataFrame({"chrom": chrom, "start": s, "end": e})

    def __iter__(self):
        if not self._fast:
            yield from self._iter_general()
            return
        new_table, b1ids = coarsen_table(self._table, self.factor)
        self.result_table = new_table
        b1, b2, cnt = table_to_host(new_table, b1ids)
        count_dtype = self.clr.pixels().dtypes["count"]
        cnt = cnt.astype(count_dtype, copy=False)
        if self.group is not None:
            # shards hold disjoint, ascending coarse row ranges: the rank-ordered concatenation IS the
            # coarsened table (bit-identical to the single-GPU result); rank 0 alone yields chunks
            b1, b2, cnt = gather_pixels_to_rank0([b1, b2, cnt], self.group)
            ind = np.searchsorted(b1, np.arange(new_table.n_bins + 1), side="left")
        else:
            ind = new_table.raw.indptr.cpu().numpy()
        edges = _prune_edges(ind, self.chunksize)
        for lo, hi in zip(edges[:-1], edges[1:]):
            if hi > lo:
                yield {"bin1_id": b1[lo:hi], "bin2_id": b2[lo:hi], "count": cnt[lo:hi]}

    def _iter_general(self):
        """Any numeric value columns / aggregators (``coarsen_columns``).  Column dtypes follow pandas'
        groupby: sum / min / max / first / last keep the input dtype (an integer sum that leaves it
        widens to int64), mean is float64; ``create`` then casts to the requested ``dtypes``."""
        clr = self.clr
        chrom_offset = np.asarray(clr._load_dset("indexes/chrom_offset"))
        bin1_offset = np.asarray(clr._load_dset("indexes/bin1_offset"))
        in_dt = clr.pixels().dtypes
        sl = slice(None)
        if self.group is not None:
            # this rank's rows, cut at coarse-bin boundaries: every (bin1', bin2') run lives on one rank
            rows = shard_rows_for_coarsening(clr, self.group, self._align)
            p0, p1 = int(bin1_offset[rows[0]]), int(bin1_offset[rows[1]])
            sl = slice(p0, p1)
            bin1_offset = np.clip(bin1_offset, p0, p1) - p0
        with clr.open("r") as grp:
            bin2 = np.asarray(grp["pixels"]["bin2_id"][sl])
            cols = {c: np.asarray(grp["pixels"][c][sl]) for c in self.columns}
        b1, b2, vals, new_off = coarsen_columns(bin1_offset, bin2, chrom_offset, self.factor, cols, self.agg,
                                                device=self._device)
        if self.group is not None:
            # rank-ordered concatenation on rank 0 (accumulator dtypes: the casts below see the whole column)
            got = gather_pixels_to_rank0([b1, b2] + [vals[c] for c in self.columns], self.group)
            b1, b2 = got[0], got[1]
            vals = dict(zip(self.columns, got[2:]))
        for c in self.columns:
            want = np.dtype(in_dt[c])
            v = vals[c]
            if self.agg[c] == "mean":
                if want == np.float32:
                    vals[c] = v.astype(np.float32)             # pandas: the mean of a float32 column is float32
                continue
            if want.kind in "iu" and len(v) and (v.min() < np.iinfo(want).min or v.max() > np.iinfo(want).max):
                continue                                       # overflowing integer sum: stays int64
            vals[c] = v.astype(want, copy=False)
        self.result_table = None                               # nothing to chain: the next level re-reads its source
        ind = np.searchsorted(b1, np.arange(int(new_off[-1]) + 1), side="left")
        edges = _prune_edges(ind, self.chunksize)
        for lo, hi in zip(edges[:-1], edges[1:]):
            if hi > lo:
                chunk = {"bin1_id": b1[lo:hi], "bin2_id": b2[lo:hi]}
                chunk.update({c: vals[c][lo:hi] for c in self.columns})
                yield chunk


def _as_cooler(obj):
    """A Cooler-like object for a URI / path: the real ``cooler.Cooler`` when the package (h5py)
    is importable, else the built-in HDF5 reader (``store.H5Cooler``); registered in-memory
    stores and Cooler-like objects pass through."""
    from .store import open_cooler
    return open_cooler(obj)


def _bins_frame(clr):
    """``clr.bins()[["chrom", "start", "end"]][:]`` with ``chrom`` as names (a categorical), whatever
    the store returns (the built-in stores keep the enum codes of ``bins/chrom``)."""
    import pandas as pd
    bins = clr.bins()[["chrom", "start", "end"]][:]
    if bins["chrom"].dtype.kind in "iu":
        names = list(clr.chromsizes.index)
        bins = bins.assign(chrom=pd.Categorical.from_codes(bins["chrom"].to_numpy(), categories=names, ordered=True))
    return bins


def _create(output_uri, bins, iterator, **kwargs):
    """``cooler.create.create`` when available (h5py), else the built-in writer (create.py)."""
    try:
        from cooler.create import create
    except ImportError:
        from .create import create_cooler as create
    return create(output_uri, bins, iterator, **kwargs)


def _coarsen(clr, output_uri, factor, chunksize, nproc, columns, dtypes, agg, device, table, kwargs,
             group=None, align=None):
    """Shared body of coarsen_cooler / zoomify_cooler: returns (MemCooler or None, device table).
    With ``group`` every rank coarsens its row shard; group rank 0 receives the pixels and writes (or
    holds the MemCooler), the other ranks get a MemCooler without pixels (bins / metadata only)."""
    factor = int(factor)
    rank0 = True
    if group is not None:
        import torch.distributed as dist
        rank0 = dist.get_rank(group) == 0
    columns = _check_columns(clr, columns, agg)
    dtypes = dict(dtypes or {})
    input_dtypes = clr.pixels().dtypes
    for col in columns:
        dtypes.setdefault(col, input_dtypes[col])
    iterator = CoolerCoarsener(clr, factor, chunksize, columns=columns, agg=agg, batchsize=nproc,
                               device=device, table=table, group=group, align=align)
    symmetric_upper = clr.storage_mode == "symmetric-upper"
    if output_uri is None or not rank0:
        chunks = list(iterator)
        cat = {k: (np.concatenate([c[k] for c in chunks]) if chunks else np.zeros(0, dtype=np.int64))
               for k in ("bin1_id", "bin2_id", "count")}        # like create(): only 'count' is kept
        c, s, e = iterator.new_bins_arrays
        out = MemCooler.from_arrays(iterator._chromnames, clr.chromsizes.values, c, s, e,
                                    cat["bin1_id"], cat["bin2_id"],
                                    cat["count"].astype(dtypes["count"], copy=False),
                                    binsize=iterator.new_binsize, symmetric_upper=symmetric_upper)
        out.device_table = iterator.result_table
        return (out if output_uri is None else None), iterator.result_table
    kwargs = dict(kwargs)
    kwargs.setdefault("append", True)
    if kwargs.get("mode") == "w":                  # `coarsen x.cool -o x.cool`: never truncate the input
        from .create import refuse_truncating_source
        refuse_truncating_source(parse_cooler_uri(output_uri)[0], [clr], "w")
    _create(output_uri, iterator.new_bins, iterator, dtypes=dtypes, symmetric_upper=symmetric_upper,
            **kwargs)
    return None, iterator.result_table


def coarsen_cooler(base_uri, output_uri, factor, chunksize, nproc=1, columns=None, dtypes=None,
                   agg=None, device="cuda", group=None, **kwargs):
    """Coarsen a cooler by an integer factor on the GPU.

    Same arguments as ``cooler.coarsen_cooler`` (_reduce.py:645-749); ``nproc`` is
    accepted and ignored (no process pool).  ``base_uri`` may be a URI / path (read by the
    ``cooler`` package when it is importable, else by the built-in HDF5 reader) or a
    Cooler-like object.  With ``output_uri=None`` the result is returned as a ``MemCooler``;
    otherwise it is written like the reference does it -- ``create(..., append=True,
    symmetric_upper=<propagated>)`` -- through ``cooler.create`` (h5py) when available, else
    through the built-in writer (``cooler_b200.create`` / ``h5write``).

    ``group``: a ``torch.distributed`` process group (one rank per GPU).  Where the reference maps
    pixel chunks over a process pool (``nproc``, _reduce.py:726-744), every rank here reads and
    coarsens the stored pixels of its own bin1 row range, cut at coarse-bin boundaries so that no
    coarse row is split (no exchange step); group rank 0 gathers the pieces in rank order --
    bit-identical to the single-GPU result -- and writes / returns them (the other ranks return a
    MemCooler without pixels, or None).  Call it on all ranks.
    """
    clr = _as_cooler(base_uri)
    table = kwargs.pop("table", None)
    out, _ = _coarsen(clr, output_uri, factor, chunksize, nproc, columns, dtypes, agg, device, table, kwargs,
                      group=group)
    if group is not None and output_uri is not None:
        import torch.distributed as dist
        dist.barrier(group)                           # the file is complete when any rank returns
    return out


def zoomify_cooler(base_uris, outfile, resolutions, chunksize, nproc=1, columns=None, dtypes=None,
                   agg=None, device="cuda", group=None, **kwargs):
    """Multi-resolution coarsening chain (``cooler.zoomify_cooler``, _reduce.py:752-877).

    Each target resolution is derived from the largest already-available
    resolution that divides it.  Levels are chained on the device: level k's output
    table feeds level k+1 without a host round trip (the reference re-reads each
    predecessor from the file).  With ``outfile=None`` returns ``{resolution: MemCooler}``
    (the base included); otherwise writes the MCOOL v2 layout ``resolutions/<r>`` into
    ``outfile`` (base levels copied first, _reduce.py:838-854; root attributes ``format`` /
    ``format-version`` last, 873-877) and returns None.

    ``group`` (one rank per GPU): every rank keeps ITS row shard of every level on its device -- the
    base shards are cut at boundaries of the least common multiple of all zoom factors, so every
    level of the chain stays exchange-free -- and group rank 0 gathers and writes each level.
    """
    rank0 = True
    if group is not None:
        import torch.distributed as dist
        rank0 = dist.get_rank(group) == 0
    if not isinstance(base_uris, (list, tuple)):
        base_uris = [base_uris]
    bases = {}
    for b in base_uris:
        clr = _as_cooler(b)
        if clr.binsize is None:
            raise ValueError("Cannot zoomify a cooler with variable-size bins")   # cli/zoomify.py
        bases[int(clr.binsize)] = clr
    resn, pred, mult = get_multiplier_sequence(list(resolutions), list(bases))
    # zoom factor of every level relative to ITS base, and their common multiple per base
    to_base, base_of = {}, {}
    for i, res in enumerate(resn):
        if pred[i] == -1:
            to_base[int(res)], base_of[int(res)] = 1, int(res)
        else:
            prev = int(resn[pred[i]])
            to_base[int(res)], base_of[int(res)] = to_base[prev] * int(mult[i]), base_of[prev]
    align = {b: math.lcm(*[f for r, f in to_base.items() if base_of[r] == b]) for b in bases}
    if outfile is not None:
        from . import create as _cr
        _cr.refuse_truncating_source(outfile, bases.values(), "w")      # on every rank: all raise or none
    if outfile is not None and rank0:
        cols = tuple(columns) if columns is not None else ("count",)
        first = True
        for res, clr in bases.items():
            _cr.copy_cooler(clr, f"{outfile}::resolutions/{res}", columns=cols, mode="w" if first else "r+")
            first = False
    if group is not None and outfile is not None:
        dist.barrier(group)
    out = {}
    tables = {}
    for i, res in enumerate(resn):
        res = int(res)
        if pred[i] == -1:
            out[res] = bases[res]
            continue
        prev = int(resn[pred[i]])
        dest = None if outfile is None else f"{outfile}::resolutions/{res}"
        kw = dict(kwargs)
        if outfile is not None:
            kw["mode"] = "r+"
        # a level made from a base reads shards aligned for the whole chain; later levels inherit them
        lvl, tables[res] = _coarsen(out[prev], dest, int(mult[i]), chunksize, nproc, columns, dtypes, agg,
                                    device, tables.get(prev), kw, group=group,
                                    align=max(1, align[base_of[res]] // to_base[prev]))
        if group is not None and outfile is not None:
            dist.barrier(group)                       # rank 0 has written the level every rank opens next
        out[res] = lvl if outfile is None else _as_cooler(dest)
    if outfile is None:
        return out
    if rank0:
        from . import create as _cr
        from . import h5write
        with h5write.File(outfile, "r+") as fw:
            fw.attrs.update({"format": "HDF5::MCOOL", "format-version": _cr.FORMAT_VERSION_MCOOL})
    if group is not None:
        dist.barrier(group)
    return None


def legacy_zoomify(input_uri, outfile, nproc, chunksize, lock=None, device="cuda"):
    """Quad-tree tiling in the legacy MCOOL layout ``::0, ::1, ...`` (_reduce.py:880-944).

    The base matrix is copied to level ``n_zooms`` and coarsened by 2 down to level 0; the
    levels are chained on the device.  Returns ``(n_zooms, {level: binsize})`` and writes the
    root attributes ``max-zoom``, ``max-zooms`` and one ``<level> -> binsize`` entry per level.
    """
    from collections import OrderedDict

    from . import create as _cr
    from . import h5write
    clr = _as_cooler(input_uri)
    _cr.refuse_truncating_source(outfile, [clr], "w")
    n_zooms = get_quadtree_depth(clr.chromsizes, clr.binsize, HIGLASS_TILE_DIM)
    factor = 2
    zoom_levels = OrderedDict()
    level = str(n_zooms)
    binsize = clr.binsize
    _cr.copy_cooler(clr, f"{outfile}::{level}", columns=None, mode="w")
    zoom_levels[level] = binsize
    src, table = _as_cooler(f"{outfile}::{level}"), None
    for i in range(n_zooms - 1, -1, -1):
        binsize *= factor
        dest = f"{outfile}::{i}"
        _, table = _coarsen(src, dest, factor, chunksize, nproc, None, None, None, device, table, {"mode": "r+"})
        src = _as_cooler(dest)
        zoom_levels[str(i)] = binsize
    with h5write.File(outfile, "r+") as fw:
        fw.attrs.update({"max-zoom": n_zooms})
        fw.attrs["max-zooms"] = n_zooms
        fw.attrs.update({k: int(v) for k, v in zoom_levels.items()})
    return n_zooms, zoom_levels


class CoolerMerger:
    """GPU replacement of ``cooler._reduce.CoolerMerger`` (_reduce.py:132-208): a
    ContactBinner-style iterable over the merged, lexsorted, duplicate-free pixels of
    several coolers with identical axes, in chunks of about ``mergebuf`` pixels."""

    def __init__(self, coolers, mergebuf, columns=None, agg=None, device="cuda"):
        self.coolers = list(coolers)
        self.mergebuf = int(mergebuf)
        self.columns = ["count"] if columns is None else list(columns)
        self.agg = {col: "sum" for col in self.columns}            # _reduce.py:146-149
        if agg is not None:
            self.agg.update(agg)
        for col, op in self.agg.items():
            if callable(op) or op not in AGG_OPS:
                raise NotImplementedError(f"aggregator {op!r} for column '{col}' is not supported by the CUDA "
                                          f"merger (supported: {sorted(AGG_OPS)})")
        # the fast path sums int32 counts inside the sort / reduce-by-key kernels; anything else (other
        # columns, other aggregators, floating-point or wide counts) carries pixel positions instead
        dts = [c.pixels().dtypes for c in self.coolers]
        self._fast = (list(self.agg) == ["count"] and self.agg["count"] == "sum" and all(
            "count" in d and np.dtype(d["count"]).kind in "iu" and np.dtype(d["count"]).itemsize <= 4 for d in dts))
        binsize = self.coolers[0].binsize                          # compatibility checks: 150-166
        if binsize is not None:
            if len({c.binsize for c in self.coolers}) > 1:
                raise ValueError("Coolers must have the same resolution")
            chromsizes = self.coolers[0].chromsizes
            for c in self.coolers[1:]:
                cs = c.chromsizes
                if len(cs) != len(chromsizes) or not np.all(cs.values == chromsizes.values) \
                        or list(cs.index) != list(chromsizes.index):
                    raise ValueError("Coolers must have the same chromosomes")
        else:
            bins = self.coolers[0].bins()[["chrom", "start", "end"]][:]
            for c in self.coolers[1:]:
                if c.binsize is not None:
                    raise ValueError("Coolers must have same bin structure")
                b2 = c.bins()[["chrom", "start", "end"]][:]
                if len(b2) != len(bins) or not np.all(b2.values == bins.values):
                    raise ValueError("Coolers must have same bin structure")
        self.device = device
        self.result_table = None

    def __iter__(self):
        if not self._fast:
            yield from self._iter_general()
            return
        tables = [PixelTable.from_cooler(c, device=self.device) for c in self.coolers]
        merged, ids = merge_tables(tables)
        self.result_table = merged
        b1, b2, cnt = table_to_host(merged, ids)
        ind = merged.raw.indptr.cpu().numpy()
        edges = _prune_edges(ind, self.mergebuf)
        for lo, hi in zip(edges[:-1], edges[1:]):
            if hi > lo:
                yield {"bin1_id": b1[lo:hi], "bin2_id": b2[lo:hi], "count": cnt[lo:hi]}

    def _iter_general(self):
        """Any numeric value columns / aggregators (``merge_columns``); the chunks carry the columns of
        ``agg`` in its order, as ``groupby(...).aggregate(agg)`` returns them (_reduce.py:196-203)."""
        names = list(self.agg)
        inputs = []
        for c in self.coolers:
            dt = c.pixels().dtypes
            for col in names:
                if col not in dt:
                    raise ValueError(f"Pixel value column '{col}' not found in input '{c.filename}'.")
            with c.open("r") as grp:
                cols = {col: np.asarray(grp["pixels"][col][:]) for col in names}
                inputs.append((np.asarray(c._load_dset("indexes/bin1_offset")),
                               np.asarray(grp["pixels"]["bin2_id"][:]), cols))
        chrom_offset = np.asarray(self.coolers[0]._load_dset("indexes/chrom_offset"))
        b1, b2, vals = merge_columns(inputs, chrom_offset, names, self.agg, device=self.device)
        self.result_table = None
        ind = np.searchsorted(b1, np.arange(int(chrom_offset[-1]) + 1), side="left")
        edges = _prune_edges(ind, self.mergebuf)
        for lo, hi in zip(edges[:-1], edges[1:]):
            if hi > lo:
                chunk = {"bin1_id": b1[lo:hi], "bin2_id": b2[lo:hi]}
                chunk.update({col: vals[col][lo:hi] for col in names})
                yield chunk


def merge_coolers(output_uri, input_uris, mergebuf, columns=None, dtypes=None, agg=None,
                  device="cuda", **kwargs):
    """Merge coolers with identical axes on the GPU (``cooler.merge_coolers``,
    _reduce.py:211-307).  ``input_uris``: URIs / paths or Cooler-like objects.
    ``output_uri=None`` returns a ``MemCooler``; otherwise the result is written like the
    reference does it (``create(output_uri, bins, iterator, columns=..., dtypes=..., assembly=...)``)."""
    clrs = [_as_cooler(u) for u in input_uris]
    is_symm = [c.storage_mode == "symmetric-upper" for c in clrs]
    if all(is_symm):
        symmetric_upper = True
    elif not any(is_symm):
        symmetric_upper = False
    else:
        raise ValueError("Cannot merge symmetric and non-symmetric coolers.")
    columns = ["count"] if columns is None else list(columns)
    dtype_map = {col: [] for col in columns}
    for c in clrs:
        pd_ = c.pixels().dtypes
        for col in columns:
            if col not in pd_:
                raise ValueError(f"Pixel value column '{col}' not found in input '{c.filename}'.")
            dtype_map[col].append(pd_[col])
    dtypes = dict(dtypes or {})
    for col in columns:
        dtypes.setdefault(col, np.result_type(*dtype_map[col]))
    iterator = CoolerMerger(clrs, mergebuf=mergebuf, columns=columns, agg=agg, device=device)
    if output_uri is None:
        chunks = list(iterator)
        keys = ["bin1_id", "bin2_id", *columns]
        cat = {k: (np.concatenate([c[k] for c in chunks]) if chunks else np.zeros(0, dtype=np.int64)) for k in keys}
        c0 = clrs[0]
        bins = c0.bins()[:]
        extra = {col: cat[col].astype(dtypes[col], copy=False) for col in columns if col != "count"}
        # the in-memory store always has a count column: zeros when 'count' was not among the merged columns
        count = cat["count"].astype(dtypes["count"], copy=False) if "count" in cat \
            else np.zeros(len(cat["bin1_id"]), dtype=np.int32)
        out = MemCooler.from_arrays(list(c0.chromsizes.index), c0.chromsizes.values,
                                    np.asarray(c0._load_dset("bins/chrom")), bins["start"].values,
                                    bins["end"].values, cat["bin1_id"], cat["bin2_id"], count,
                                    binsize=c0.binsize, symmetric_upper=symmetric_upper,
                                    extra_pixel_columns=extra or None)
        out.device_table = iterator.result_table
        return out
    _create(output_uri, _bins_frame(clrs[0]), iterator, columns=columns, dtypes=dtypes,
            assembly=clrs[0].info.get("genome-assembly", None), symmetric_upper=symmetric_upper, **kwargs)
    return None
