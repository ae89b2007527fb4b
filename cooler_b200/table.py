"""Device-resident pixel table: the HBM layout of the hot path.

A *copy* of the table is ``px[nnz]`` of 8-byte records ``{other bin, count}``
ordered by row, plus ``indptr[n_rows+1]`` (int64) and ``tile_row[n_tiles]``
(row of the first pixel of every warp tile).  The bin1-major copy ("CSR")
stores ``other = bin2``; the bin2-major copy ("CSC"), built once by the
on-device stable radix sort, stores ``other = bin1``.

torch is used for device memory and streams only; every kernel is in
``libcooler_b200.so`` (include/cooler_b200.h).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_I32 = torch.int32
_I64 = torch.int64


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _use_device(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.NativeLibraryError("cooler_b200 runs on CUDA devices only (no CPU fallback)")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    torch.cuda.set_device(idx)
    _lib.call("cb200_set_device", idx)
    return torch.device("cuda", idx)


def to_host(t):
    """Device tensor -> numpy through a pinned staging buffer (torch caches pinned blocks, so repeated
    calls reuse them): 25 MB of per-bin results come back in ~0.5 ms instead of ~10 ms pageable."""
    if t.device.type != "cuda" or t.numel() < 65536:
        return t.cpu().numpy()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return h.numpy()


def _as_tensor(a):
    """Host array -> tensor; numpy dtypes torch lacks (uint32 / uint64) are widened to int64 first."""
    if isinstance(a, torch.Tensor):
        return a
    a = np.ascontiguousarray(a)
    if a.dtype in (np.uint32, np.uint64):
        if a.size and int(a.max()) > 2**63 - 1:
            raise OverflowError("count column holds values outside int64")
        a = a.astype(np.int64)
    elif a.dtype == np.uint16:
        a = a.astype(np.int32)
    return torch.from_numpy(a)


def counts_as_int32(count_t):
    """The count column as an int32 tensor.  Columns of a wider or unsigned integer type are range
    checked first: a value outside int32 would wrap silently and corrupt the marginals (the
    reference computes in whatever dtype the column has, _balance.py:25-26)."""
    if count_t.dtype == _I32:
        return count_t
    if count_t.dtype.is_floating_point:
        raise NotImplementedError(
            "this cooler_b200 path takes integer count columns only (balancing handles floating-point "
            "columns: balance_cooler / balance_arrays)")
    if count_t.numel():
        lo, hi = int(count_t.min()), int(count_t.max())
        if lo < -2**31 or hi > 2**31 - 1:
            raise OverflowError(f"count column ({count_t.dtype}) holds values outside int32 "
                                f"[{lo}, {hi}]: not supported by the CUDA path")
    return count_t.to(_I32)


def is_float_column(count):
    """True for a floating-point count column (numpy array or tensor)."""
    if isinstance(count, torch.Tensor):
        return count.dtype.is_floating_point
    return np.issubdtype(np.asarray(count).dtype, np.floating)


def attach_values(copy, values):
    """Floating-point count columns: ``copy`` was built with the pixel's position in the stored table in
    the count field; sets ``copy.val`` = the float64 values in the copy's own order and relabels the
    field with the position inside the copy (cb200_attach_values)."""
    val = torch.empty(max(copy.nnz, 1), dtype=torch.float64, device=copy.px.device)
    _lib.call("cb200_attach_values", _ptr(copy.px), copy.nnz, _ptr(values), int(values.numel()), _ptr(val), _stream())
    copy.val = val
    return copy


def values_nonzero(values):
    """float64 1.0 / 0.0 per value: ``data[data != 0] = 1`` (_binarize, _balance.py:29-31)."""
    out = torch.empty_like(values)
    _lib.call("cb200_values_nonzero", _ptr(values), int(values.numel()), _ptr(out), _stream())
    return out


def warp_tile():
    return _lib.load().cb200_warp_tile()


def n_tiles(nnz):
    wt = warp_tile()
    return (int(nnz) + wt - 1) // wt


def _new_px(nnz, device):
    # +2 records: pad so that the 128-bit load of the last pair stays in bounds
    return torch.empty((int(nnz) + 2, 2), dtype=_I32, device=device)


class TableCopy:
    """One row-ordered copy of the pixel table on the device."""

    def __init__(self, px, indptr, nnz, n_rows, tile_row=None, row_base=0):
        # rows [row_base, row_base + n_rows) of the matrix; indptr / tile_row are local to that range
        self.px, self.indptr, self.nnz, self.n_rows = px, indptr, int(nnz), int(n_rows)
        self.row_base = int(row_base)
        if tile_row is None:
            tile_row = torch.empty(max(n_tiles(nnz), 1), dtype=_I32, device=px.device)
            _lib.call("cb200_tile_rows", _ptr(indptr), self.n_rows, self.nnz, _ptr(tile_row), _stream())
        self.tile_row = tile_row

    @property
    def device(self):
        return self.px.device

    def c_struct(self):
        return _lib.Copy(self.px.data_ptr(), self.indptr.data_ptr(), self.tile_row.data_ptr(), self.nnz)

    def nbytes(self):
        return self.nnz * 8 + self.indptr.numel() * 8 + self.tile_row.numel() * 4

    # -- debugging / tests ---------------------------------------------------
    def tight(self):
        """Same pixels, row range shrunk to [first non-empty row, last non-empty row]."""
        if self.nnz == 0:
            return self
        edge = torch.tensor([0, self.nnz], dtype=_I64, device=self.indptr.device)
        pos = torch.searchsorted(self.indptr, edge, right=False).cpu().numpy()
        first_r = torch.searchsorted(self.indptr, edge[:1], right=True).item() - 1
        last_r = int(pos[1]) - 1                      # last row r with indptr[r] < nnz
        if first_r == 0 and last_r == self.n_rows - 1:
            return self
        sub = TableCopy(self.px, self.indptr[first_r:last_r + 2], self.nnz, last_r - first_r + 1,
                        row_base=self.row_base + first_r)
        for a in ("strip", "val"):
            if hasattr(self, a):
                setattr(sub, a, getattr(self, a))
        return sub

    def to_host(self):
        """(row ids, other ids, counts) as numpy int64/int64/int32."""
        ind = self.indptr.cpu().numpy()
        px = self.px[: self.nnz].cpu().numpy()
        rows = self.row_base + np.repeat(np.arange(self.n_rows, dtype=np.int64), np.diff(ind))
        return rows, px[:, 0].astype(np.int64), px[:, 1].copy()


def exclusive_scan_i32(x):
    """int32[n] -> int64[n+1] exclusive prefix sums (device)."""
    n = x.numel()
    out = torch.empty(n + 1, dtype=_I64, device=x.device)
    ws_bytes = _lib.load().cb200_scan_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    _lib.call("cb200_exclusive_scan_i32", _ptr(x), _ptr(out), n, _ptr(ws), ws_bytes, _stream())
    return out


def sort_pairs(keys, vals, key_bits, key_shift=0):
    """Stable sort of (int32 key, int32x2 value) by key bits [key_shift, key_shift + key_bits)."""
    n = keys.numel()
    dev = keys.device
    keys_out = torch.empty(n, dtype=_I32, device=dev)
    vals_out = _new_px(n, dev)
    if n == 0:
        return keys_out, vals_out
    passes = (max(key_bits, 1) + 7) // 8
    keys_tmp = torch.empty(n if passes > 1 else 1, dtype=_I32, device=dev)
    vals_tmp = torch.empty((n if passes > 1 else 1, 2), dtype=_I32, device=dev)
    ws_bytes = _lib.load().cb200_sort_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.call("cb200_sort_pairs", _ptr(keys), _ptr(vals), _ptr(keys_out), _ptr(vals_out),
              _ptr(keys_tmp), _ptr(vals_tmp), n, int(key_shift), int(key_bits), _ptr(ws), ws_bytes, _stream())
    return keys_out, vals_out


def indptr_from_sorted(keys, n, n_rows, shift=0, key_base=0):
    out = torch.empty(n_rows + 1, dtype=_I64, device=keys.device)
    _lib.call("cb200_indptr_from_sorted", _ptr(keys), int(n), int(n_rows), int(key_base), int(shift),
              _ptr(out), _stream())
    return out


def key_bits_for(n_rows):
    return max(1, int(n_rows - 1).bit_length()) if n_rows > 1 else 1


def filter_copy(copy, row_lo, row_hi, ndiag, keep, *, want_rows=False, want_transpose=False,
                band_shift=0):
    """Stable compaction of ``copy`` by the static filter (see cb200_filter_count).

    Returns (new_copy, tkey, tval); the new copy's indptr is rebuilt from the
    kept pixels' row ids.  tkey/tval (if requested) feed the transposition sort.
    """
    dev = copy.device
    nt = n_tiles(copy.nnz)
    tile_count = torch.zeros(max(nt, 1), dtype=_I32, device=dev)
    _lib.call("cb200_filter_count", _ptr(copy.px), _ptr(copy.indptr), _ptr(copy.tile_row),
              copy.n_rows, copy.nnz, _ptr(row_lo), _ptr(row_hi), int(ndiag), int(keep), int(band_shift),
              copy.row_base, _ptr(tile_count), _stream())
    tile_off = exclusive_scan_i32(tile_count[:nt] if nt else tile_count[:0])
    kept = int(tile_off[-1].item())                      # one sync per build step
    px_out = _new_px(kept, dev)
    rows = torch.empty(max(kept, 1), dtype=_I32, device=dev)
    tkey = torch.empty(max(kept, 1), dtype=_I32, device=dev) if want_transpose else None
    tval = _new_px(kept, dev) if want_transpose else None
    _lib.call("cb200_filter_write", _ptr(copy.px), _ptr(copy.indptr), _ptr(copy.tile_row),
              copy.n_rows, copy.nnz, _ptr(row_lo), _ptr(row_hi), int(ndiag), int(keep), int(band_shift),
              copy.row_base, _ptr(tile_off), _ptr(px_out), _ptr(rows), _ptr(tkey), _ptr(tval), _stream())
    indptr = indptr_from_sorted(rows, kept, copy.n_rows, key_base=copy.row_base)
    new = TableCopy(px_out, indptr, kept, copy.n_rows, row_base=copy.row_base)
    if want_rows:
        new.rows = rows
    return new, (tkey[:kept] if tkey is not None else None), tval


def earlier_chrom_pixels(copy, row_lo, row_hi):
    """The stored pixels (r, c) of ``copy`` whose column lies in a chromosome BEFORE r's own
    (``c < row_lo[r]``) as ``(px {c, count}, rows int32, n)``, or None when there is none -- the case
    of every symmetric-upper table, which then costs one counting pass.  Input of ``cb200_cis_touch``
    (cis_only on square-storage tables, see include/cooler_b200.h)."""
    dev = copy.device
    nt = n_tiles(copy.nnz)
    if nt == 0:
        return None
    tile_count = torch.zeros(nt, dtype=_I32, device=dev)
    args = (_ptr(copy.px), _ptr(copy.indptr), _ptr(copy.tile_row), copy.n_rows, copy.nnz, _ptr(row_lo),
            _ptr(row_hi), 0, _lib.KEEP_EARLIER_CHROM, 0, copy.row_base)
    _lib.call("cb200_filter_count", *args, _ptr(tile_count), _stream())
    tile_off = exclusive_scan_i32(tile_count)
    kept = int(tile_off[-1].item())
    if kept == 0:
        return None
    px_out = _new_px(kept, dev)
    rows = torch.empty(kept, dtype=_I32, device=dev)
    _lib.call("cb200_filter_write", *args, _ptr(tile_off), _ptr(px_out), _ptr(rows), _ptr(None), _ptr(None),
              _stream())
    return px_out, rows, kept


def transpose_from(tkey, tval, n_rows, return_keys=False):
    """(key = other bin, val = {row, count}) in row-major order -> the other-major copy."""
    n = tkey.numel()
    keys_sorted, vals_sorted = sort_pairs(tkey, tval, key_bits_for(n_rows))
    indptr = indptr_from_sorted(keys_sorted, n, n_rows)
    copy = TableCopy(vals_sorted, indptr, n, n_rows)
    return (copy, keys_sorted) if return_keys else copy


def swap_key_other(keys, vals, n):
    """(key, {a, v}) -> (a, {key, v})."""
    dev = keys.device
    keys_out = torch.empty(max(n, 1), dtype=_I32, device=dev)
    vals_out = _new_px(n, dev)
    _lib.call("cb200_swap_key_other", _ptr(keys), _ptr(vals), int(n), _ptr(keys_out), _ptr(vals_out), _stream())
    return keys_out[:n], vals_out


def strip_partition(key_other, val_rowcount, n, n_rows, shift):
    """Split a row-ordered stream (key = gathered bin, val = {row, count}) into column
    strips ``key >> shift``; every strip becomes a TableCopy of its own (views into one
    buffer, strip starts 16-byte aligned).  Row order inside a strip is preserved
    (stable radix pass), so a strip is a valid row-major copy."""
    dev = key_other.device
    n_strips = ((max(n_rows, 1) - 1) >> shift) + 1
    if n == 0:
        return []
    ks, vs = sort_pairs(key_other, val_rowcount, key_bits_for(n_strips), key_shift=shift)
    start = indptr_from_sorted(ks, n, n_strips, shift).cpu().numpy()
    counts = np.diff(start)
    pad_before = np.r_[0, np.cumsum(counts & 1)[:-1]].astype(np.int64)
    aligned = start[:-1] + pad_before
    total = int(n + n_strips + 2)
    px_out = torch.empty((total, 2), dtype=_I32, device=dev)
    rows_out = torch.empty(total, dtype=_I32, device=dev)
    pad_d = torch.from_numpy(pad_before).to(dev)
    _lib.call("cb200_strip_finish", _ptr(ks), _ptr(vs), int(n), int(shift), _ptr(pad_d), _ptr(px_out),
              _ptr(rows_out), _stream())
    del ks, vs
    live = np.flatnonzero(counts > 0)
    # row range of every strip = [first row id, last row id] of its (row-sorted) elements
    idx = np.concatenate([aligned[live], aligned[live] + counts[live] - 1])
    ends = rows_out[torch.from_numpy(idx).to(dev)].cpu().numpy()
    first_rows, last_rows = ends[: len(live)], ends[len(live):]
    subs = []
    for j, s_ in enumerate(live):
        c, a = int(counts[s_]), int(aligned[s_])
        r0, r1 = int(first_rows[j]), int(last_rows[j])
        indptr = indptr_from_sorted(rows_out[a:a + c], c, r1 - r0 + 1, key_base=r0)
        sub = TableCopy(px_out[a:a + c + 2], indptr, c, r1 - r0 + 1, row_base=r0)
        sub.strip = int(s_)
        subs.append(sub)
    return subs


def far_rows():
    """Rows per far-field block of the loaded library (1 << CB200_FAR_ROW_SHIFT, csrc/far.cuh)."""
    return 1 << _lib.far_row_shift()


class FarCopy:
    """Far-field pixels of one copy in block order (cb200_far_copy in the header)."""

    def __init__(self, px, ptr, nnz, rb0, n_rb, j0, n_j, col_shift, warps, tile_counts):
        self.px, self.ptr, self.nnz = px, ptr, int(nnz)
        self.rb0, self.n_rb, self.j0, self.n_j = int(rb0), int(n_rb), int(j0), int(n_j)
        self.col_shift, self.warps = int(col_shift), int(warps)
        self.tile_counts = tile_counts            # host int64 [n_rb, n_j]: records per (row block, column tile)

    def nbytes(self):
        return self.px.numel() * 4 + self.ptr.numel() * 8


FAR_COMPACT_MAX_GROWTH = 1.25   # compact records are used while splitting counts > 127 adds < 25 % records


def far_blocks(key_other, val_rowcount, n, n_bins, col_shift, warps=16, compact=None):
    """Row-ordered stream (key = gathered bin, val = {row, count}) -> FarCopy.

    Two stable radix sorts bring the records into (row >> RS, other >> col_shift, row,
    other) order: first on the column-tile bits of the gathered bin, then -- after swapping
    key and other -- on the row-block bits of the row.

    ``compact``: True = 4-byte records {row:12 | column:13 | count:7} (counts above 127 become
    several records of the same cell), False = 8-byte records, None = compact when the geometry
    allows it (<= 4096-row blocks, <= 8192-column tiles), no count is negative and splitting adds
    less than 25 % records.  The choice made is ``FarCopy.rec_bytes``."""
    dev = key_other.device
    n = int(n)
    if n == 0:
        return None
    FAR_ROW_SHIFT = _lib.far_row_shift()
    n_j_all = ((max(n_bins, 1) - 1) >> col_shift) + 1
    n_rb_all = ((max(n_bins, 1) - 1) >> FAR_ROW_SHIFT) + 1
    ks, vs = sort_pairs(key_other[:n], val_rowcount, key_bits_for(n_j_all), key_shift=col_shift)
    kr, vr = swap_key_other(ks, vs, n)            # key = row, val = {other, count}
    del ks, vs
    if n_rb_all > 1:
        kr, vr = sort_pairs(kr, vr, key_bits_for(n_rb_all), key_shift=FAR_ROW_SHIFT)
    n_px = n
    can_compact = FAR_ROW_SHIFT <= 12 and col_shift <= 13
    if compact and not can_compact:
        raise ValueError("compact far-field records need <= 4096-row blocks and <= 8192-column tiles")
    if compact is None:
        compact = can_compact
    if compact:
        pieces = torch.empty(n, dtype=_I32, device=dev)
        flags = torch.empty(2, dtype=_I32, device=dev)
        _lib.call("cb200_far_pieces", _ptr(vr), n, _ptr(pieces), _ptr(flags), _stream())
        big, neg = (int(x) for x in flags.cpu().numpy())
        if neg:
            compact = False                       # negative counts do not fit the 7-bit field
        elif big:
            off = exclusive_scan_i32(pieces)
            n2 = int(off[-1].item())
            if n2 > FAR_COMPACT_MAX_GROWTH * n:
                compact = False
            else:
                kr2 = torch.empty(n2, dtype=_I32, device=dev)
                vr2 = _new_px(n2, dev)
                _lib.call("cb200_far_expand", _ptr(kr), _ptr(vr), n, _ptr(off), _ptr(kr2), _ptr(vr2), _stream())
                kr, vr, n = kr2, vr2, n2
            del off
        del pieces, flags
    ends = kr[torch.tensor([0, n - 1], device=dev)].cpu().numpy()
    rb0, rb1 = int(ends[0]) >> FAR_ROW_SHIFT, int(ends[1]) >> FAR_ROW_SHIFT
    n_rb, j0, n_j = rb1 - rb0 + 1, 0, n_j_all
    n_slots = n_rb * n_j * warps
    if n_slots >= 2**31 - 1:
        raise ValueError("far-field slot table does not fit int32")
    rec = torch.empty(n, dtype=_I32, device=dev) if compact else _new_px(n, dev)
    slot = torch.empty(n, dtype=_I32, device=dev)
    _lib.call("cb200_far_finish4" if compact else "cb200_far_finish", _ptr(kr), _ptr(vr), n, int(col_shift),
              rb0, j0, n_j, int(warps), _ptr(rec), _ptr(slot), _stream())
    del kr, vr
    ptr_rec = indptr_from_sorted(slot, n, n_slots)            # first record of every slot
    # lane-interleaved chunks: every slot is padded to whole 64-record chunks
    nch = torch.empty(n_slots, dtype=_I32, device=dev)
    _lib.call("cb200_far_slot_chunks", _ptr(ptr_rec), n_slots, _ptr(nch), _stream())
    cptr = exclusive_scan_i32(nch)                            # first chunk of every slot
    del nch
    total_chunks = int(cptr[-1].item())
    if compact:
        px = torch.empty(total_chunks * 64 + 4, dtype=_I32, device=dev)
        _lib.call("cb200_far_interleave4", _ptr(rec), _ptr(slot), _ptr(ptr_rec), _ptr(cptr), n, n_slots,
                  total_chunks, _ptr(px), _stream())
    else:
        px = torch.empty((total_chunks * 64 + 2, 2), dtype=_I32, device=dev)
        _lib.call("cb200_far_interleave", _ptr(rec), _ptr(slot), _ptr(ptr_rec), _ptr(cptr), n, total_chunks,
                  _ptr(px), _stream())
    del rec, slot
    tile_start = ptr_rec[::warps].cpu().numpy()               # [n_rb * n_j + 1]
    counts = np.diff(tile_start).reshape(n_rb, n_j)
    fc = FarCopy(px, cptr, n, rb0, n_rb, j0, n_j, col_shift, warps, counts)
    fc.n_chunks = total_chunks
    fc.rec_bytes = 4 if compact else 8
    fc.n_pixels = n_px                                        # before counts > 127 were split
    return fc


def plan_far_units(copies_counts, target):
    """Split the non-empty (row block, column tile) pairs of every far copy into units.

    ``copies_counts``: list of (rb0, j0, counts[n_rb, n_j]).  A unit is a run of consecutive
    non-empty tiles of ONE row block holding about ``target`` records (a persistent CTA of the
    far kernel processes one unit at a time).  Returns (units int32 [n_units, 4] = copy, rb,
    t0, t1; tiles int32 [n_tiles] = column tile ids; rb_unit: list of int32 [n_rb + 1] per copy).
    Units are ordered by (copy, row block, t0) -- the order in which the per-bin kernels add
    their partial sums."""
    target = max(int(target), 1)
    units, tiles, rb_units = [], [], []
    n_units = n_tiles_ = 0
    for c, (rb0, j0, counts) in enumerate(copies_counts):
        n_rb = counts.shape[0]
        rows, cols = np.nonzero(counts)                       # row-major: the tile list of every row block in order
        per_row = np.bincount(rows, minlength=n_rb)           # non-empty tiles per row block
        tstart = np.r_[0, np.cumsum(per_row)]
        totals = counts.sum(axis=1)
        parts = np.maximum(1, np.rint(totals / target).astype(np.int64))
        parts[per_row == 0] = 0
        n_row_units = np.where(parts <= 1, parts, 0)          # rows split below are counted there
        split_units = {}
        for k in np.flatnonzero(parts > 1):                   # only the (few) heavy row blocks need a search
            js = cols[tstart[k]:tstart[k + 1]]
            cum = np.cumsum(counts[k, js])
            p = int(parts[k])
            edges = np.searchsorted(cum, cum[-1] * np.arange(1, p) / p, side="left") + 1
            edges = np.unique(np.r_[0, edges, len(js)])
            split_units[int(k)] = edges
            n_row_units[k] = len(edges) - 1
        first_unit = n_units + np.r_[0, np.cumsum(n_row_units)]
        u = np.zeros((int(n_row_units.sum()), 4), dtype=np.int32)
        single = np.flatnonzero(parts == 1)
        pos = (first_unit[single] - n_units).astype(np.int64)
        u[pos, 0], u[pos, 1] = c, rb0 + single
        u[pos, 2], u[pos, 3] = n_tiles_ + tstart[single], n_tiles_ + tstart[single + 1]
        for k, edges in split_units.items():
            p0 = int(first_unit[k] - n_units)
            m = len(edges) - 1
            u[p0:p0 + m, 0], u[p0:p0 + m, 1] = c, rb0 + k
            u[p0:p0 + m, 2] = n_tiles_ + tstart[k] + edges[:-1]
            u[p0:p0 + m, 3] = n_tiles_ + tstart[k] + edges[1:]
        units.append(u)
        tiles.append((cols + j0).astype(np.int32))
        rb_units.append(first_unit.astype(np.int32))
        n_units += len(u)
        n_tiles_ += len(cols)
    units = np.concatenate(units) if units else np.zeros((0, 4), dtype=np.int32)
    tiles = np.concatenate(tiles) if tiles else np.zeros(0, dtype=np.int32)
    return units.reshape(-1, 4), tiles, rb_units


def segment_medians(x, seg_offset, positive_only=False):
    """Exact medians of the segments ``x[seg_offset[s]:seg_offset[s+1]]`` of a float64 DEVICE vector
    (``cb200_segment_medians``: radix selection, bit-identical to ``np.median`` of the valid
    elements -- positive ones when ``positive_only``).  Returns a numpy array, NaN for segments
    without valid elements (as ``np.median`` of an empty array)."""
    dev = x.device
    seg_offset = np.asarray(seg_offset, dtype=np.int64)
    S = len(seg_offset) - 1
    if S <= 0:
        return np.zeros(0)
    off_d = torch.from_numpy(seg_offset).to(dev)
    mid = torch.empty(2 * S, dtype=torch.float64, device=dev)
    cnt = torch.empty(S, dtype=_I64, device=dev)
    ws_bytes = _lib.load().cb200_median_workspace_bytes(S)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.call("cb200_segment_medians", _ptr(x), _ptr(off_d), S, int(np.diff(seg_offset).max()), int(bool(positive_only)),
              _ptr(mid), _ptr(cnt), _ptr(ws), ws_bytes, _stream())
    mid = mid.cpu().numpy().reshape(S, 2)
    cnt = cnt.cpu().numpy()
    out = np.full(S, np.nan)
    for s in np.flatnonzero(cnt > 0):
        out[s] = np.median(mid[s])                 # the same mean-of-the-two-middle-elements numpy takes
    return out


def marginals_i64(copy):
    """(nnz_marg, sum_marg) int64 row marginals of one copy."""
    dev = copy.device
    n = copy.n_rows
    nnz_out = torch.empty(n, dtype=_I64, device=dev)
    sum_out = torch.empty(n, dtype=_I64, device=dev)
    scratch = torch.empty((2 * n_tiles(copy.nnz) + n) * 2 + 2, dtype=_I64, device=dev)
    cs = copy.c_struct()
    _lib.call("cb200_marginals_i64", C.byref(cs), n, _ptr(nnz_out), _ptr(sum_out), _ptr(scratch), _stream())
    return nnz_out, sum_out


class PixelTable:
    """The uploaded, unfiltered bin1-major copy plus the genome segmentation.

    Built once per cooler ("h5py reads the pixel table once"); balancing with
    different options and coarsening both start from it without touching the
    host again.  ``row_range`` marks a shard: this table holds the stored pixels
    of rows [row_range[0], row_range[1]) of an n_bins x n_bins matrix.
    """

    def __init__(self, raw, chrom_offset, row_range=None, fval=None):
        self.raw = raw
        self.fval = fval          # float64 values of a floating-point count column (raw's count field = position), else None
        self.n_bins = raw.n_rows
        self.chrom_offset = np.asarray(chrom_offset, dtype=np.int64)
        self.row_range = row_range or (0, self.n_bins)
        nb = np.diff(self.chrom_offset)
        lo = np.repeat(self.chrom_offset[:-1], nb).astype(np.int32)
        hi = np.repeat(self.chrom_offset[1:], nb).astype(np.int32)
        self.row_lo = torch.from_numpy(lo).to(raw.device)
        self.row_hi = torch.from_numpy(hi).to(raw.device)

    @property
    def device(self):
        return self.raw.device

    @property
    def nnz(self):
        return self.raw.nnz

    @classmethod
    def from_arrays(cls, bin1_offset, bin2_id, count, chrom_offset, device="cuda", row_range=None,
                    float_values=False):
        """Upload (host arrays or tensors; pinned memory is copied asynchronously).

        ``bin1_offset`` is the cooler's ``indexes/bin1_offset`` (the CSR indptr;
        docs/schema_v3.rst) -- bin1_id itself is never shipped to the device.

        ``float_values`` (balancing only): a floating-point count column is kept as float64
        (``table.fval``) and the records' count field holds the pixel's position, so the filters,
        the transposition and the strip partition carry the values along as indices.
        """
        device = _use_device(device)
        n_bins = len(bin1_offset) - 1
        if n_bins >= 2**31 - 1:
            raise ValueError("n_bins must fit int32")
        nnz = len(bin2_id)

        def dev(a, dtype):
            t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
            if t.dtype != dtype:
                t = t.to(dtype)
            return t.to(device, non_blocking=True)

        fval = None
        if float_values and is_float_column(count):
            if nnz >= 2**31 - 1:
                raise ValueError("a table with floating-point counts must hold fewer than 2^31 pixels per GPU")
            fval = dev(_as_tensor(count), torch.float64)
            count_d = torch.arange(nnz, dtype=_I32, device=device)
        else:
            count_d = dev(counts_as_int32(_as_tensor(count)), _I32)
        indptr = dev(bin1_offset, _I64)
        bin2_t = bin2_id if isinstance(bin2_id, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(bin2_id))
        narrow = bin2_t.dtype == _I32                 # reader already narrowed the ids: 8 B/pixel over PCIe
        bin2_d = dev(bin2_t, _I32 if narrow else _I64)
        px = _new_px(nnz, device)
        _lib.call("cb200_pack_pixels_i32" if narrow else "cb200_pack_pixels", _ptr(bin2_d), _ptr(count_d),
                  nnz, _ptr(px), _stream())
        raw = TableCopy(px, indptr, nnz, n_bins)
        return cls(raw, chrom_offset, row_range=row_range, fval=fval)

    @classmethod
    def from_cooler(cls, clr, device="cuda", structure_only=False):
        """Read the pixel table of a ``cooler.Cooler``-like object once (see read_cooler_device).
        ``structure_only``: rows and columns only, every count 0 -- for consumers that bring their own
        value column (``api.matrix(field=...)``)."""
        if structure_only:
            bin1_offset = np.asarray(clr._load_dset("indexes/bin1_offset"))
            chrom_offset = np.asarray(clr._load_dset("indexes/chrom_offset"))
            with clr.open("r") as grp:
                bin2 = np.asarray(grp["pixels"]["bin2_id"][:])
            return cls.from_arrays(bin1_offset, bin2, torch.zeros(len(bin2), dtype=_I32), chrom_offset,
                                   device=device)
        bin1_offset, bin2, count, chrom_offset = read_cooler_device(clr, device=device)
        return cls.from_arrays(bin1_offset, bin2, count, chrom_offset, device=device)


def shard_rows_by_pixels(bin1_offset, rank, world):
    """Contiguous bin1 row range [lo, hi) of ``rank``: boundaries at
    ``searchsorted(bin1_offset, g * nnz / world)`` -- about nnz / world stored pixels
    per rank, rows never split.  The reference's counterpart is the chunk partition of
    ``_balance.py:351-357`` handed to ``Pool.imap_unordered`` (cli/balance.py:237-243)."""
    off = np.asarray(bin1_offset)
    n_bins, nnz = len(off) - 1, int(off[-1])
    cuts = np.searchsorted(off, nnz * np.arange(world + 1, dtype=np.float64) / world, side="left")
    cuts = np.minimum(cuts, n_bins)
    cuts[0], cuts[-1] = 0, n_bins
    cuts = np.maximum.accumulate(cuts)
    return int(cuts[rank]), int(cuts[rank + 1])


_PINNED_RING: dict = {}


def _pinned_ring(slots, rows, dtypes):
    """``slots`` pinned staging buffers per column, allocated once per process (page-locking memory costs
    about as much as decoding into it)."""
    key = (slots, rows, tuple(str(d) for d in dtypes))
    if key not in _PINNED_RING:
        _PINNED_RING.clear()
        _PINNED_RING[key] = [[torch.empty(rows, dtype=d, pin_memory=True) for d in dtypes] for _ in range(slots)]
    return _PINNED_RING[key]


_NP2TORCH = {"int32": torch.int32, "int64": torch.int64, "float32": torch.float32, "float64": torch.float64,
             "int16": torch.int16, "uint8": torch.uint8, "int8": torch.int8}


def read_cooler_device(clr, rows=None, device="cuda", piece_rows=1 << 23, slots=3, min_pixels=1 << 20):
    """The pixel columns of a Cooler-like object straight into DEVICE memory: returns (bin1_offset host
    int64 -- local to ``rows`` as in ``read_cooler_arrays`` --, bin2_id device int32 / int64, count device,
    chrom_offset host).

    For files read by ``h5lite`` the stored chunks are inflated by a pool of host threads directly into a
    small ring of pinned staging buffers (bin2_id assembled as int32 from the low byte planes) and every
    finished piece of ``piece_rows`` pixels crosses PCIe on a copy stream while later pieces are still
    being inflated: the table is read ONCE, through pinned memory, and the host never holds a whole
    column.  For other stores (h5py, MemCooler) and tables below ``min_pixels`` the columns come back as HOST arrays
    (``read_cooler_arrays``) and the caller's build uploads them.  The staging ring is shared by all calls of
    the process: one call at a time."""
    from . import h5lite
    device = _use_device(device)
    chrom_offset = np.asarray(clr._load_dset("indexes/chrom_offset"))
    bin1_offset = np.asarray(clr._load_dset("indexes/bin1_offset"))
    n_bins = len(bin1_offset) - 1
    p0, p1 = 0, int(bin1_offset[-1])
    if rows is not None:
        p0, p1 = int(bin1_offset[rows[0]]), int(bin1_offset[rows[1]])
        bin1_offset = np.clip(bin1_offset, p0, p1) - p0
    with clr.open("r") as grp:
        d2, dc = grp["pixels"]["bin2_id"], grp["pixels"]["count"]
        fast = all(isinstance(d, h5lite.Dataset) and d.is_chunked_1d() and str(np.dtype(d.dtype)) in _NP2TORCH
                   and np.dtype(d.dtype).byteorder in "<=|" for d in (d2, dc))      # native little-endian columns
        if not fast or p1 - p0 < min_pixels:
            _, bin2, count, _ = read_cooler_arrays(clr, rows=rows)     # host arrays: the caller uploads (pipelined build)
            return bin1_offset, bin2, count, chrom_offset
        narrow = n_bins < 2**31 - 1 and h5lite._narrowable(d2.dtype, np.int32)
        t2 = torch.int32 if narrow else _NP2TORCH[str(np.dtype(d2.dtype))]
        tc = _NP2TORCH[str(np.dtype(dc.dtype))]
        csz = max(d2.chunk_rows(), dc.chunk_rows())
        piece = max(csz, piece_rows // csz * csz)          # pieces start on chunk boundaries of the file
        start = p0 // csz * csz
        pieces = [(max(a, p0), min(a + piece, p1)) for a in range(start, p1, piece)]
        ring = _pinned_ring(slots, piece, (t2, tc))
        events = [None] * slots
        bin2_d = torch.empty(p1 - p0, dtype=t2, device=device)
        count_d = torch.empty(p1 - p0, dtype=tc, device=device)
        copy_stream = torch.cuda.Stream(device=device)
        for k, (a, b) in enumerate(pieces):
            slot = k % slots
            if events[slot] is not None:
                events[slot].synchronize()                 # the slot's previous upload has left the buffer
            # blocking, all host cores, no interpreter lock held; the upload of the previous piece runs meanwhile
            h5lite.read_rows_into([(d2, a, b, ring[slot][0].numpy()[: b - a]), (dc, a, b, ring[slot][1].numpy()[: b - a])])
            with torch.cuda.stream(copy_stream):
                bin2_d[a - p0:b - p0].copy_(ring[slot][0][: b - a], non_blocking=True)
                count_d[a - p0:b - p0].copy_(ring[slot][1][: b - a], non_blocking=True)
                events[slot] = torch.cuda.Event()
                events[slot].record(copy_stream)
        torch.cuda.current_stream().wait_stream(copy_stream)
        for ev in events:                                  # the ring is reused by the next call
            if ev is not None:
                ev.synchronize()
    return bin1_offset, bin2_d, count_d, chrom_offset


def read_cooler_arrays(clr, rows=None):
    """(bin1_offset, bin2_id, count, chrom_offset) host arrays of a Cooler-like object.

    Protocol used (reference: src/cooler/parallel.py:289-304, _balance.py:141-142):
    ``clr.open('r')`` -> group with ``pixels/bin2_id``, ``pixels/count``;
    ``clr._load_dset('indexes/chrom_offset' | 'indexes/bin1_offset')``.  bin1_id is never read.

    ``rows=(lo, hi)`` reads only the stored pixels of bin1 rows [lo, hi) -- the slice
    ``pixels[bin1_offset[lo]:bin1_offset[hi]]`` a multi-GPU rank owns -- and returns LOCAL offsets
    (``bin1_offset`` clipped to the slice and shifted; rows outside it are empty)."""
    chrom_offset = np.asarray(clr._load_dset("indexes/chrom_offset"))
    bin1_offset = np.asarray(clr._load_dset("indexes/bin1_offset"))
    n_bins = len(bin1_offset) - 1
    sl = slice(None)
    if rows is not None:
        p0, p1 = int(bin1_offset[rows[0]]), int(bin1_offset[rows[1]])
        sl = slice(p0, p1)
        bin1_offset = np.clip(bin1_offset, p0, p1) - p0
    with clr.open("r") as grp:
        d2 = grp["pixels"]["bin2_id"]
        if hasattr(d2, "astype") and hasattr(d2, "id") and n_bins < 2**31 - 1:
            bin2 = d2.astype(np.int32)[sl]             # h5py converts while reading: no int64 staging
        else:
            bin2 = d2[sl]
        count = grp["pixels"]["count"][sl]
    return bin1_offset, bin2, count, chrom_offset
