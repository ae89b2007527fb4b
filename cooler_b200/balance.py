"""ICE matrix balancing on B200 -- drop-in for ``cooler.balance_cooler``.

Mirrors open2c/cooler v0.10.4 ``src/cooler/_balance.py`` (names, arguments,
return values, warnings).  The chunk map-reduce of the reference
(``parallel.split(...).pipe(...).reduce(add)``) and the three
``_balance_*`` loops are replaced by CUDA kernels in libcooler_b200.so; the
O(n_bins) filter arithmetic (min_nnz, min_count, MAD-max, blacklist;
_balance.py:375-413) stays on the host in numpy, statement for statement, so
that the filter masks are bit-exact.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
import warnings

import numpy as np
import torch

from . import _lib
from .table import (PixelTable, TableCopy, _as_tensor, _new_px, _ptr, _stream, _use_device, counts_as_int32, to_host,  # noqa: F401
                    attach_values, earlier_chrom_pixels, is_float_column, values_nonzero, far_blocks, far_rows, filter_copy, marginals_i64, n_tiles, plan_far_units, read_cooler_arrays, read_cooler_device,
                    segment_medians, shard_rows_by_pixels,
                    strip_partition, swap_key_other, transpose_from)

__all__ = ["balance_cooler", "balance_table", "balance_arrays", "ConvergenceWarning", "IceProblem"]

logger = logging.getLogger("cooler")   # same logger name as the reference (_logging.py)

_F64, _I32 = torch.float64, torch.int32
UPD_BINS = 256                          # smallest number of bins per block of the per-bin kernels (csrc/ice.cu: 256 threads)
UPD_BLOCKS_TARGET = 1184                # about 8 blocks per SM: every block takes a ticket on ONE counter per launch


class ConvergenceWarning(UserWarning):
    """Same role as cooler._balance.ConvergenceWarning (_balance.py:21)."""


def mad(data, axis=None):
    """util.mad (src/cooler/util.py:513-514)."""
    return np.median(np.abs(data - np.median(data, axis)), axis)


def _allreduce(t, group):
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


_COPY_STREAMS: dict = {}


def _copy_stream(dev):
    key = str(dev)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _COPY_STREAMS[key]


class _TableInfo:
    """What balance_table needs to know about a table that was consumed while streaming."""

    def __init__(self, n_bins, chrom_offset, device, row_range):
        self.n_bins, self.chrom_offset, self.device, self.row_range = n_bins, chrom_offset, device, row_range
        nb = np.diff(chrom_offset)
        self.row_lo = torch.from_numpy(np.repeat(chrom_offset[:-1], nb).astype(np.int32)).to(device)
        self.row_hi = torch.from_numpy(np.repeat(chrom_offset[1:], nb).astype(np.int32)).to(device)


class _PeerVectors:
    """Double-buffered per-bin vectors in NVLink peer-mapped (symmetric) memory: halves 0 / 1 hold
    the local partial sums s of alternating passes, halves 2 / 3 the slice-wise reduced vector of
    the two-stage all-reduce (long vectors, see TWO_STAGE_BYTES)."""

    def __init__(self, local, handle, ptrs, stride, world):
        self.local, self.handle, self.ptrs, self.stride, self.world = local, handle, ptrs, stride, world
        self.rank = handle.rank
        self.tick = 0     # which half is written next; never reset, so a half is only reused after
                          # a later barrier proved every peer is done reading it

    _cache: dict = {}

    @classmethod
    def create(cls, n, group, dev):
        """Allocate + rendezvous once per (group, size, device); later calls reuse the buffers
        (the rendezvous is a collective that costs ~0.1 s)."""
        import torch.distributed as dist
        key = (id(group), int(n), str(dev))
        if key in cls._cache:
            return cls._cache[key]
        pv = cls._create(n, group, dev)
        cls._cache[key] = pv
        return pv

    @classmethod
    def _create(cls, n, group, dev):
        import torch.distributed as dist
        try:
            import torch.distributed._symmetric_memory as symm_mem
            if dist.get_backend(group) != "nccl":
                return None
            stride = -(-max(n, 1) // 64) * 64
            local = symm_mem.empty(4 * stride, dtype=_F64, device=dev)
            local.zero_()
            handle = symm_mem.rendezvous(local, group)
            ptrs = torch.tensor(list(handle.buffer_ptrs), dtype=torch.int64, device=dev)
            return cls(local, handle, ptrs, stride, handle.world_size)
        except Exception as e:  # pragma: no cover - depends on the platform
            logger.warning(f"peer-memory reduction unavailable ({e!r}); using NCCL all-reduce")
            return None


def merge_row_pieces(pieces, strip):
    """Concatenate row-ordered sub-copies that cover ascending, disjoint row ranges (the pieces of
    one column strip coming from successive row chunks) into ONE sub-copy: pixel arrays back to back
    (device-to-device copies), row pointers shifted by the pixels before them, rows between two
    pieces empty."""
    if len(pieces) == 1:
        pieces[0].strip = strip
        return pieces[0]
    dev = pieces[0].px.device
    total = sum(p.nnz for p in pieces)
    base = pieces[0].row_base
    n_rows = pieces[-1].row_base + pieces[-1].n_rows - base
    px = _new_px(total, dev)
    indptr = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
    o, a_prev_end = 0, 0
    for p in pieces:
        a = p.row_base - base
        assert a >= a_prev_end - 1 and p.indptr.numel() == p.n_rows + 1
        if a > a_prev_end:
            indptr[a_prev_end:a].fill_(o)               # empty rows between two pieces
        px[o:o + p.nnz].copy_(p.px[: p.nnz])
        torch.add(p.indptr, o, out=indptr[a:a + p.n_rows + 1])
        o += p.nnz
        a_prev_end = a + p.n_rows + 1
    sub = TableCopy(px, indptr, total, n_rows, row_base=base)
    sub.strip = strip
    return sub


class IceProblem:
    """Filtered CSR/CSC copies of one table for one (mode, ignore_diags) choice.

    Stage A applies the reference's ``base_filters`` (_balance.py:360-364):
    ``_zero_trans`` for cis_only, ``_zero_diags``; the prefilter marginals
    (min_nnz pass and raw marginal, _balance.py:375-393) are taken on it.
    For trans_only stage B additionally removes cis pixels (``_zero_cis`` is
    only piped inside ``_balance_transonly``, _balance.py:227).
    """

    def __init__(self, table: PixelTable, *, cis_only=False, trans_only=False, ignore_diags=2,
                 group=None, strip_bins=None, use_smem=True, band=None, peer_reduce=None,
                 far_col_shift=None, far_compact=None):
        self.table = table
        self.use_smem = use_smem
        self.far_copies = []
        self.far_col_shift = int(far_col_shift or FAR_COL_SHIFT)
        self.far_warps = 16
        self.peer_reduce = True if peer_reduce is None else bool(peer_reduce)
        self.group = group
        self.mode = "cis" if cis_only else ("trans" if trans_only else "gw")
        n_bins = table.n_bins
        fval = getattr(table, "fval", None)          # floating-point count column: records carry positions
        self.float_counts = fval is not None
        if self.float_counts and band not in (None, False):
            raise NotImplementedError("the band / 2-D blocked layouts hold integer counts only")
        ndiag = int(ignore_diags) if ignore_diags else 0
        keep = _lib.KEEP_CIS if cis_only else _lib.KEEP_ALL
        csr, tkey, tval = filter_copy(table.raw, table.row_lo, table.row_hi, ndiag, keep,
                                      want_transpose=True)
        self.earlier = []                            # cis_only, square storage: see cis_touch_flags
        if cis_only:
            e = earlier_chrom_pixels(table.raw, table.row_lo, table.row_hi)
            if e is not None:
                self.earlier.append(e)
        csc, ckeys = transpose_from(tkey, tval, n_bins, return_keys=True)
        if fval is None:
            # prefilter marginals (exact integers)
            nnz_r, sum_r = marginals_i64(csr)
            nnz_c, sum_c = marginals_i64(csc)
            marg = torch.stack([nnz_r + nnz_c, sum_r + sum_c])
            del nnz_r, sum_r, nnz_c, sum_c
        else:
            marg = float_marginals([csr, csc], fval, n_bins)
        _allreduce(marg, group)
        marg = to_host(marg)
        self.marg_nnz = marg[0].astype(np.float64)
        self.marg_sum = marg[1].astype(np.float64)
        cval = None
        if trans_only:
            tkey = tval = ckeys = None
            csr, tkey, tval = filter_copy(csr, table.row_lo, table.row_hi, 0, _lib.KEEP_TRANS,
                                          want_transpose=True)
            csc, ckeys, cval = filter_copy(csc, table.row_lo, table.row_hi, 0, _lib.KEEP_TRANS,
                                           want_transpose=True)
        self._nnz_eff = csr.nnz
        # layout of the sweep: whole copies, column strips, or diagonal band strips + far rest
        self.strip_shift, self.band = choose_layout(n_bins, csr.nnz, strip_bins, band)
        if self.float_counts:
            self.band = False                         # whole copies or column strips
        sh = self.strip_shift
        if band is None and self.band == "allblocks":
            # automatic choice: the block build keeps ~5 transient 12-byte copies of the pixel stream
            # alive (sort in / out / scratch of both orientations); fall back to whole copies when
            # that does not fit next to the table (a whole 1 kb matrix on ONE GPU)
            free = torch.cuda.mem_get_info(table.device)[0]
            if 60 * csr.nnz > free:
                self.band = False
        if band == "allblocks" or self.band == "allblocks":   # every pixel through the 2-D blocked kernel (csrc/far.cuh)
            self.band, self.strip_shift = "allblocks", None
            if cval is None:
                ckeys, cval = swap_key_other(ckeys, csc.px, csc.nnz)
            n1, n2 = csr.nnz, csc.nnz
            self.csr = self.csc = csr = csc = None
            # both orientations hold the same counts, so they agree on compact (4-byte) records
            self.far_copies = [far_blocks(tkey[:n1], tval, n1, n_bins, self.far_col_shift, compact=far_compact)]
            tkey = tval = None
            compact = self.far_copies[0].rec_bytes == 4 if self.far_copies[0] is not None else far_compact
            self.far_copies.append(far_blocks(ckeys[:n2], cval, n2, n_bins, self.far_col_shift, compact=compact))
            self.subs = []
        elif sh is None:
            if fval is not None:
                attach_values(csr, fval), attach_values(csc, fval)
            self.subs = [csr.tight(), csc.tight()]
            self.csr, self.csc = csr, csc
        elif not self.band:
            subs = strip_partition(tkey[: csr.nnz], tval, csr.nnz, n_bins, sh)
            tkey = tval = None
            if cval is None:                      # csc came from the sort: (key = row, val = {other, count})
                ckeys, cval = swap_key_other(ckeys, csc.px, csc.nnz)
            subs += strip_partition(ckeys[: csc.nnz], cval, csc.nnz, n_bins, sh)
            if fval is not None:
                for sub in subs:
                    attach_values(sub, fval)
            self.subs = subs
            self.csr = self.csc = None            # the strips replace the whole copies
        else:
            tkey = tval = ckeys = cval = None
            subs = []
            for copy in (csr, csc):
                near, nkey, nval = filter_copy(copy, table.row_lo, table.row_hi, 0, _lib.KEEP_BAND_NEAR,
                                               want_transpose=True, band_shift=sh)
                subs += strip_partition(nkey[: near.nnz], nval, near.nnz, n_bins, sh)
                del near, nkey, nval
                blocks = self.band == "blocks"
                far, fkey, fval = filter_copy(copy, table.row_lo, table.row_hi, 0, _lib.KEEP_BAND_FAR,
                                              want_transpose=blocks, band_shift=sh)
                if far.nnz and blocks:            # far field in (row block, column tile) order: csrc/far.cuh
                    nfar = far.nnz
                    del far
                    fc = far_blocks(fkey, fval, nfar, n_bins, self.far_col_shift, compact=far_compact)
                    far_compact = fc.rec_bytes == 4   # the other orientation holds the same counts
                    self.far_copies.append(fc)
                elif far.nnz:                     # far field as one row-ordered copy gathered through L2
                    subs.append(far.tight())
                del fkey, fval
            self.subs = subs
            self.csr = self.csc = None
        tkey = tval = ckeys = cval = None

    # ------------------------------------------------------------------ streamed build
    @classmethod
    def from_host(cls, bin1_offset, bin2_id, count, chrom_offset, *, cis_only=False, trans_only=False,
                  ignore_diags=2, group=None, device="cuda", n_chunks=8, peer_reduce=None,
                  row_range=None):
        """Build the problem straight from host arrays, **pipelined with the host->device copy**.

        Whole-copy layout (bias vector served by L1/L2): row chunks of equal pixel counts.  Column-strip
        layout (10 kb class): one chunk per strip-aligned range of 16 384 ROWS; the bin2-major copy of
        such a chunk gathers only the chunk's own rows, i.e. it IS one strip of the bin2-major side, and
        its bin1-major pixels are cut into strip pieces that are concatenated per strip once the last
        chunk is in (`merge_row_pieces`).  The resulting sub-copies -- and therefore every bit of the
        sweep -- are identical to those of the two-step build (PixelTable + IceProblem), but only the
        last chunk's build and the concatenation (a few ms) are left when the last byte has crossed PCIe.

        The table is cut into row chunks; chunk c+1 crosses PCIe (copy stream) while chunk c is
        packed, filtered, radix-sorted into its own bin2-major copy and marginalised on the main
        stream.  Every chunk contributes a bin1-major and a bin2-major sub-copy (the column sums
        of the chunks add up exactly like column strips or multi-GPU shards), so nothing has to
        wait for the whole table.  Returns None when the layout needs far-field blocks (1 kb class):
        the caller then uses the two-step path (PixelTable + IceProblem).
        """
        if is_float_column(count):
            return None                               # floating-point counts: two-step build
        dev = _use_device(device)
        off_t = bin1_offset if isinstance(bin1_offset, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(bin1_offset, dtype=np.int64))
        off_np = off_t.numpy()
        n_bins = len(off_np) - 1
        b2_t = bin2_id if isinstance(bin2_id, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(bin2_id))
        c_t = counts_as_int32(_as_tensor(count))
        if b2_t.dtype not in (_I32, torch.int64):
            b2_t = b2_t.to(torch.int64)
        nnz = int(b2_t.numel())
        if n_bins >= 2**31 - 1:
            raise ValueError("n_bins must fit int32")
        lay = choose_layout(n_bins, nnz, None)
        if lay[1] or nnz == 0:
            return None
        strip_shift = lay[0]
        chrom_offset = np.asarray(chrom_offset, dtype=np.int64)
        info = _TableInfo(n_bins, chrom_offset, dev, row_range or (0, n_bins))
        if strip_shift is None:
            # row-aligned chunks of about nnz / n_chunks pixels
            targets = (np.arange(1, n_chunks) * (nnz / n_chunks)).astype(np.int64)
            rows = np.unique(np.r_[0, np.searchsorted(off_np, targets, side="left"), n_bins])
            rows = rows[np.r_[True, np.diff(off_np[rows]) > 0]] if len(rows) > 2 else rows
            rows[-1] = n_bins
        else:
            # column strips: one chunk per strip of ROWS, so that the bin2-major copy of a chunk is
            # exactly one strip of the bin2-major side (its gathered bins = the chunk's rows)
            S = 1 << strip_shift
            rows = np.unique(np.r_[np.arange(0, n_bins, S), n_bins]).astype(np.int64)
        main = torch.cuda.current_stream()
        copy_stream = _copy_stream(dev)
        indptr_d = off_t.to(dev, non_blocking=True)
        bin2_d = torch.empty(nnz, dtype=b2_t.dtype, device=dev)
        count_d = torch.empty(nnz, dtype=_I32, device=dev)
        copy_stream.wait_stream(main)
        events = []
        with torch.cuda.stream(copy_stream):
            for r0, r1 in zip(rows[:-1], rows[1:]):
                p0, p1 = int(off_np[r0]), int(off_np[r1])
                bin2_d[p0:p1].copy_(b2_t[p0:p1], non_blocking=True)
                count_d[p0:p1].copy_(c_t[p0:p1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                events.append(ev)
        bin2_d.record_stream(copy_stream)
        count_d.record_stream(copy_stream)

        self = cls.__new__(cls)
        self.table, self.group, self.use_smem = info, group, True
        self.float_counts = False
        self.mode = "cis" if cis_only else ("trans" if trans_only else "gw")
        self.peer_reduce = True if peer_reduce is None else bool(peer_reduce)
        self.strip_shift, self.band, self.csr, self.csc = strip_shift, False, None, None
        self.far_copies, self.far_col_shift, self.far_warps = [], 13, 16
        self.earlier = []
        ndiag = int(ignore_diags) if ignore_diags else 0
        keep = _lib.KEEP_CIS if cis_only else _lib.KEEP_ALL
        marg = torch.zeros((2, n_bins), dtype=torch.int64, device=dev)
        subs, nnz_eff = [], 0
        n_strips = ((max(n_bins, 1) - 1) >> strip_shift) + 1 if strip_shift is not None else 0
        row_pieces = [[] for _ in range(n_strips)]     # bin1-major side: pieces of strip s, one per row chunk
        col_strips = []                                # bin2-major side: chunk k IS strip k
        pack = "cb200_pack_pixels_i32" if b2_t.dtype == _I32 else "cb200_pack_pixels"
        for ev, r0, r1 in zip(events, rows[:-1], rows[1:]):
            r0, r1 = int(r0), int(r1)
            p0, p1 = int(off_np[r0]), int(off_np[r1])
            if p1 == p0:
                continue
            main.wait_event(ev)
            px = _new_px(p1 - p0, dev)
            _lib.call(pack, _ptr(bin2_d[p0:p1]), _ptr(count_d[p0:p1]), p1 - p0, _ptr(px), _stream())
            raw = TableCopy(px, indptr_d[r0:r1 + 1] - p0, p1 - p0, r1 - r0, row_base=r0)
            csr, tkey, tval = filter_copy(raw, info.row_lo, info.row_hi, ndiag, keep, want_transpose=True)
            if cis_only:
                e = earlier_chrom_pixels(raw, info.row_lo, info.row_hi)
                if e is not None:
                    self.earlier.append(e)
            del raw, px
            csc = transpose_from(tkey, tval, n_bins)
            if strip_shift is None or trans_only:
                del tkey, tval
            a_nnz, a_sum = marginals_i64(csr)
            marg[0, r0:r1] += a_nnz
            marg[1, r0:r1] += a_sum
            b_nnz, b_sum = marginals_i64(csc)
            marg[0] += b_nnz
            marg[1] += b_sum
            if trans_only and strip_shift is None:
                csr, _, _ = filter_copy(csr, info.row_lo, info.row_hi, 0, _lib.KEEP_TRANS)
                csc, _, _ = filter_copy(csc, info.row_lo, info.row_hi, 0, _lib.KEEP_TRANS)
            elif trans_only:
                csr, tkey, tval = filter_copy(csr, info.row_lo, info.row_hi, 0, _lib.KEEP_TRANS, want_transpose=True)
                csc = transpose_from(tkey, tval, n_bins) if csr.nnz else csc
            nnz_eff += csr.nnz
            if not csr.nnz:
                continue
            if strip_shift is None:
                subs += [csr.tight(), csc.tight()]
                continue
            # bin1-major side: this chunk's rows, cut into column strips (pieces of the final strips)
            for piece in strip_partition(tkey[: csr.nnz], tval, csr.nnz, n_bins, strip_shift):
                row_pieces[piece.strip].append(piece)
            del tkey, tval, csr
            csc = csc.tight()
            csc.strip = r0 >> strip_shift
            col_strips.append(csc)
        del bin2_d, count_d
        if strip_shift is not None:
            subs = [merge_row_pieces(p, s_) for s_, p in enumerate(row_pieces) if p] + col_strips
        _allreduce(marg, group)
        marg = to_host(marg)
        self.marg_nnz = marg[0].astype(np.float64)
        self.marg_sum = marg[1].astype(np.float64)
        self.subs = subs
        self._nnz_eff = nnz_eff
        return self

    @property
    def nnz_eff(self):
        """Stored pixels surviving the static filters (one copy)."""
        return self._nnz_eff

    def bytes_per_pass(self):
        """Algorithmic bytes of one marginal pass: 16 B/effective pixel + 40 B/bin (32 B/pixel with a
        floating-point count column: 8 B record + 8 B value in both copies)."""
        return (32 if self.float_counts else 16) * self.nnz_eff + 40 * self.table.n_bins

    def stream_bytes(self):
        """Bytes of pixel records one pass actually streams from HBM: 8 B per record of the row-ordered
        sub-copies, 8 or 4 B (compact) per record of the 2-D blocked copies, padding included."""
        far = [f for f in self.far_copies if f is not None]
        return (16 if self.float_counts else 8) * sum(c.nnz for c in self.subs) + \
            sum(f.n_chunks * 64 * f.rec_bytes for f in far)


L1_RESIDENT_BINS = 49_152      # bias vectors up to 384 KB are served well enough by L1 + L2
DEFAULT_STRIP_BINS = 16_384    # 128 KB of float64 bias per strip: stays in the SM's L1
MIN_PIXELS_PER_ROW_STRIP = 16  # below this the per-(row, strip) bookkeeping costs more than it saves
BLOCKS_MIN_BINS = 1 << 20      # sparse tables with at least this many bins use the 2-D blocked layout (csrc/far.cuh)
FAR_COL_SHIFT = 13             # 8192-column tiles of w (64 KB each, two in shared memory next to the 4096 row sums)
DEVICE_MEDIAN_MIN_BINS = 1 << 16   # from here on the medians of the bin filters are selected on the device
TWO_STAGE_BYTES = 32 << 20     # peer bytes per rank and pass from which the all-reduce runs in two stages
FAR_UNITS_PER_SM = 4           # work units of the blocked kernel per SM (2..16 measured equal)


def choose_layout(n_bins, nnz_eff, strip_bins=None, band=None):
    """(log2 strip width in bins | None, band: False | True | "blocks" | "allblocks") -- how the
    sweep gathers the bias.

    The sweep gathers bias[other] per pixel.  When the bias vector is much larger than
    L1 every gather costs a 32-byte L2 sector per 8-byte pixel and L2 bandwidth, not
    HBM, bounds the pass.  Splitting a copy into strips of the gathered index lets a
    persistent CTA stage the gathered slice in shared memory, at the price of one
    partial sum per (row, strip).  That pays while (row, strip) segments stay long
    ("strips": every strip).  For very sparse far fields (1 kb matrices) only the band
    of strips around the diagonal is cut into strips and the far remainder stays one
    row-ordered copy gathered through L2 ("band"), or -- the automatic choice from 2**20 bins
    on -- every pixel goes through the 2-D blocked kernel ("allblocks", csrc/far.cuh).
    """
    if strip_bins == 0:
        return None, False
    auto = strip_bins is None
    if auto:
        if n_bins <= L1_RESIDENT_BINS:
            return None, False
        strip_bins = DEFAULT_STRIP_BINS
    shift = int(strip_bins).bit_length() - 1
    if (1 << shift) != int(strip_bins):
        raise ValueError("strip_bins must be a power of two")
    n_strips = ((max(n_bins, 1) - 1) >> shift) + 1
    sparse = nnz_eff < MIN_PIXELS_PER_ROW_STRIP * n_bins * n_strips or n_strips > 1024
    if band is None:
        if auto and sparse:
            # measured on the three kinds of 1 kb shards (tools/tune_sweep.py, profiles/r01d_tune_1kb_*):
            # every pixel through the 2-D blocked kernel (4096 rows x 8192 columns) 1.56 / 1.42 / 1.42 ms
            # against 2.29 / 1.98 / 1.36 ms for whole copies gathered through L2; band strips gain
            # ~7 % over whole copies but cost two more filter passes -> never the default
            return None, ("allblocks" if n_bins >= BLOCKS_MIN_BINS else False)
        band = sparse
    return shift, (band if band in ("blocks", "allblocks") else bool(band))


def sub_descriptors(subs, n_bins, strip_shift, smem, dev, float_counts=False, val=None):
    """Device array of cb200_sub for the sub-copies ``subs`` plus their output buffers.  Returns
    (subs_dev, direct, pieces, claim, span, total_warps, total_tiles).  ``val`` overrides the value array
    of every sub-copy (prefilter marginals of floating-point columns: the raw column, records holding
    positions in the stored table)."""
    total_tiles = sum(n_tiles(c.nnz) for c in subs)
    span = int(min(8, max(1, -(-total_tiles // (148 * 32 * 8)))))
    if smem:
        span = min(span, 4)                       # persistent strip kernel: finer claims, see csrc/ice.cu
    arr = (_lib.Sub * max(len(subs), 1))()
    direct = torch.zeros(sum(c.n_rows for c in subs) + 1, dtype=_F64, device=dev)
    pieces = torch.zeros(2 * total_tiles + 2, dtype=_F64, device=dev)
    warp0, tile0, row0 = 0, 0, 0
    for k, c in enumerate(subs):
        arr[k].px, arr[k].indptr, arr[k].tile_row = c.px.data_ptr(), c.indptr.data_ptr(), c.tile_row.data_ptr()
        arr[k].nnz = c.nnz
        arr[k].row_base, arr[k].row_count = c.row_base, c.n_rows
        arr[k].direct = direct.data_ptr() + 8 * row0
        row0 += c.n_rows
        arr[k].pieces = pieces.data_ptr() + 8 * 2 * tile0
        arr[k].first_warp = warp0
        arr[k].first_tile = tile0
        if float_counts:
            arr[k].val = (val if val is not None else c.val).data_ptr()
        strip = getattr(c, "strip", None)
        if strip is None or strip_shift is None:
            arr[k].gather_base, arr[k].gather_len = 0, n_bins
        else:
            base = strip << strip_shift
            arr[k].gather_base, arr[k].gather_len = base, min(1 << strip_shift, n_bins - base)
        warp0 += -(-n_tiles(c.nnz) // span)
        tile0 += n_tiles(c.nnz)
    subs_dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
    claim = torch.zeros(len(subs) + 1, dtype=torch.int64, device=dev)
    return subs_dev, direct, pieces, claim, span, warp0, total_tiles


def float_marginals(copies, fval, n_bins):
    """Prefilter marginals of a floating-point count column (_balance.py:375-393): float64 [2, n_bins] =
    (number of nonzero values, sum of values) over the rows of every copy in ``copies`` -- the records
    hold positions into ``fval``.  One sweep (K3, whole-copy form, w = 1) + combine per quantity."""
    dev = fval.device
    out = torch.zeros((2, max(n_bins, 1)), dtype=_F64, device=dev)
    subs = [c.tight() for c in copies if c.nnz]
    if not subs or n_bins == 0:
        return out[:, :n_bins]
    ones = torch.ones(n_bins, dtype=_F64, device=dev)
    done = torch.zeros(1, dtype=_I32, device=dev)
    for k, val in enumerate((values_nonzero(fval), fval)):
        subs_dev, direct, pieces, claim, span, warps, tiles = sub_descriptors(
            subs, n_bins, None, False, dev, float_counts=True, val=val)
        st = _lib.Ice()
        st.subs, st.n_subs, st.span, st.total_warps, st.total_tiles = subs_dev.data_ptr(), len(subs), span, warps, tiles
        st.smem_bins, st.claim, st.n_bins = 0, claim.data_ptr(), n_bins
        st.w, st.s, st.all_done, st.float_counts = ones.data_ptr(), out[k].data_ptr(), done.data_ptr(), 1
        _lib.call("cb200_ice_sweep", C.byref(st), _stream())
        _lib.call("cb200_ice_combine", C.byref(st), _stream())
        torch.cuda.current_stream().synchronize()       # the descriptors are released at the end of the turn
    return out[:, :n_bins]


class IceState:
    """Device buffers + cb200_ice descriptor for the iteration loop."""

    def __init__(self, prob: IceProblem, bias0, cweights, seg_offset, tol, max_iters):
        dev = prob.table.device
        n = prob.table.n_bins
        seg_offset = np.asarray(seg_offset, dtype=np.int64)
        S = len(seg_offset) - 1
        # host block table: blocks of >= UPD_BINS bins, never straddling a segment.  Long vectors (1 kb:
        # 3.1e6 bins) get larger blocks: with 1024-bin blocks the 3016 tickets per launch, serialised on one
        # counter, made the two per-bin kernels 0.16 ms of a 1.75 ms pass at 8 GPUs.
        per_block = max(UPD_BINS, -(-n // UPD_BLOCKS_TARGET // 256) * 256)
        blk_seg, blk_lo, blk_hi, seg_blk = [], [], [], [0]
        for s in range(S):
            lo, hi = int(seg_offset[s]), int(seg_offset[s + 1])
            for b in range(lo, hi, per_block):
                blk_seg.append(s), blk_lo.append(b), blk_hi.append(min(b + per_block, hi))
            seg_blk.append(len(blk_seg))
        nb = len(blk_seg)

        def i32(a):
            return torch.tensor(np.asarray(a, dtype=np.int32), dtype=_I32, device=dev)

        def f64(k, fill=0.0):
            return torch.full((max(int(k), 1),), fill, dtype=_F64, device=dev)

        self.n, self.S, self.max_iters = n, S, int(max_iters)
        self.t = t = {}
        t["seg_offset"], t["blk_seg"], t["blk_lo"], t["blk_hi"] = i32(seg_offset), i32(blk_seg), i32(blk_lo), i32(blk_hi)
        t["seg_blk_offset"] = i32(seg_blk)
        t["bias"] = torch.from_numpy(np.ascontiguousarray(bias0, dtype=np.float64)).to(dev)
        if cweights is not None:
            t["cweights"] = torch.from_numpy(np.ascontiguousarray(cweights, dtype=np.float64)).to(dev)
            t["w"] = t["bias"] * t["cweights"]
        else:
            t["cweights"] = None
            t["w"] = t["bias"]
        # sub-copy descriptors (whole copies or column strips) for K3 / combine
        subs = prob.subs
        smem = prob.strip_shift is not None and prob.use_smem
        t["subs"], t["direct"], t["pieces"], t["claim"], span, warp0, total_tiles = sub_descriptors(
            subs, n, prob.strip_shift, smem, dev, float_counts=getattr(prob, "float_counts", False))
        t["s"], t["marg"] = f64(n), f64(n)
        # far-field blocks: units of about equal size, a few per SM (csrc/far.cuh)
        far = [f for f in getattr(prob, "far_copies", []) if f is not None]
        n_far_units = 0
        if far:
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            target = max(sum(f.nnz for f in far) // (sms * FAR_UNITS_PER_SM), 4096)
            units, tiles, rb_units = plan_far_units([(f.rb0, f.j0, f.tile_counts) for f in far], target)
            n_far_units = len(units)
            t["far_units"] = torch.from_numpy(np.ascontiguousarray(units)).to(dev)
            t["far_tiles"] = torch.from_numpy(np.ascontiguousarray(tiles)).to(dev)
            t["far_rb_unit"] = [torch.from_numpy(r).to(dev) for r in rb_units]
            farr = (_lib.FarCopy * len(far))()
            for k, f in enumerate(far):
                farr[k].px, farr[k].ptr, farr[k].nnz = f.px.data_ptr(), f.ptr.data_ptr(), f.nnz
                farr[k].rb0, farr[k].n_rb, farr[k].j0, farr[k].n_j = f.rb0, f.n_rb, f.j0, f.n_j
                farr[k].rb_unit = t["far_rb_unit"][k].data_ptr()
            t["far_copies"] = torch.frombuffer(bytearray(bytes(farr)), dtype=torch.uint8).to(dev)
            t["far_parts"] = torch.empty(max(n_far_units, 1) * far_rows(), dtype=_F64, device=dev)
            t["far_claim"] = torch.zeros(2, dtype=torch.int64, device=dev)
        self.n_far_units = n_far_units
        # multi-GPU: local s vectors live in peer-mapped symmetric memory (double-buffered) and
        # K4a sums them over NVLink; falls back to an NCCL all-reduce if that is unavailable
        self.peer = None
        self.two_stage = False
        if prob.group is not None and prob.peer_reduce:
            self.peer = _PeerVectors.create(n, prob.group, dev)
            if self.peer is not None:
                t["s"] = self.peer.local[: max(n, 1)]
                # every rank reading every rank's whole vector moves (world - 1) x 8 B per bin over NVLink;
                # beyond TWO_STAGE_BYTES the reduce-scatter + all-gather form (one more barrier) is cheaper
                self.two_stage = (self.peer.world - 1) * 8 * n > TWO_STAGE_BYTES
        t["blk_sum"], t["blk_cnt"], t["blk_dev"] = f64(nb), f64(nb), f64(nb)
        t["seg_mean"], t["seg_nnz"] = f64(S), f64(S)
        t["seg_scale"], t["seg_var"] = f64(S, 1.0), f64(S, float("nan"))
        for k in ("seg_iters", "seg_done", "seg_empty"):
            t[k] = torch.zeros(max(S, 1), dtype=_I32, device=dev)
        t["all_done"] = torch.zeros(1, dtype=_I32, device=dev)
        t["tickets"] = torch.zeros(2, dtype=_I32, device=dev)
        t["var_hist"] = f64(self.max_iters * S, float("nan"))
        st = _lib.Ice()
        st.n_subs, st.span, st.total_warps, st.total_tiles = len(subs), span, warp0, total_tiles
        st.smem_bins = (1 << prob.strip_shift) if (prob.strip_shift is not None and prob.use_smem) else 0
        st.n_bins, st.n_segs, st.n_blocks = n, S, nb
        for k in ("subs", "claim", "seg_offset", "blk_seg", "blk_lo", "blk_hi", "seg_blk_offset", "bias", "cweights", "w",
                  "s", "marg", "blk_sum", "blk_cnt",
                  "blk_dev", "seg_mean", "seg_nnz", "seg_scale", "seg_var", "seg_iters", "seg_done",
                  "seg_empty", "all_done", "tickets", "var_hist"):
            setattr(st, k, t[k].data_ptr() if t[k] is not None else None)
        st.tol, st.max_iters = float(tol), self.max_iters
        st.n_far_units, st.n_far_copies = n_far_units, len(far)
        if far:
            for k in ("far_copies", "far_units", "far_tiles", "far_parts", "far_claim"):
                setattr(st, k, t[k].data_ptr())
            st.far_col_shift, st.far_warps = far[0].col_shift, far[0].warps
            st.far_rec_bytes = far[0].rec_bytes
            assert all(f.rec_bytes == far[0].rec_bytes for f in far)
        st.float_counts = 1 if getattr(prob, "float_counts", False) else 0
        self.st = st
        self.prob = prob
        self.flag_host = torch.zeros(1, dtype=_I32).pin_memory() if torch.cuda.is_available() else None
        self.launches = 0
        self.sweep_events = None          # list of (start, end) CUDA events per pass when profiling
        self.comm_events = []             # same for the all-reduce (multi-GPU)
        self.stage_events = []            # multi-GPU: events at the stage boundaries of every profiled pass
        self._peer_tick = 0

    def reset(self, bias0):
        """Back to the initial state (same resident copies): lets a bench repeat the loop."""
        t = self.t
        t["bias"].copy_(torch.from_numpy(np.ascontiguousarray(bias0, dtype=np.float64)), non_blocking=True)
        if t["cweights"] is not None:
            torch.mul(t["bias"], t["cweights"], out=t["w"])
        for k in ("seg_iters", "seg_done", "seg_empty", "all_done"):
            t[k].zero_()
        t["seg_scale"].fill_(1.0)
        t["seg_var"].fill_(float("nan"))
        t["var_hist"].fill_(float("nan"))

    def one_pass(self, it):
        """sweep (K3) -> combine -> [allreduce] -> update (K4).  Asynchronous."""
        st, stream = C.byref(self.st), _stream()
        if self.sweep_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.call("cb200_ice_sweep", st, stream)
            e1.record()
            self.sweep_events.append((e0, e1))
        else:
            _lib.call("cb200_ice_sweep", st, stream)
        if self.prob.group is None:                   # single GPU: combine fused into K4a
            _lib.call("cb200_ice_update", st, int(it), 1, stream)
            self.launches += 3 + (1 if self.n_far_units else 0)
        elif self.peer is not None:                   # multi-GPU, all-reduce fused into K4a over NVLink
            half = self.peer.tick & 1             # keeps alternating across runs and calls
            self.peer.tick += 1
            self.st.s = self.peer.local.data_ptr() + 8 * half * self.peer.stride
            timed = self.sweep_events is not None
            marks = []

            def mark():                            # stage boundaries of one pass (profiling runs only)
                if timed:
                    ev = torch.cuda.Event(enable_timing=True)
                    ev.record()
                    marks.append(ev)

            mark()
            _lib.call("cb200_ice_combine", C.byref(self.st), stream)
            mark()
            self.peer.handle.barrier(channel=half)
            mark()
            if self.two_stage:
                # reduce-scatter (own slice, summed in rank order) -> barrier -> all-gather inside K4a
                red_off = (2 + half) * self.peer.stride
                _lib.call("cb200_ice_reduce_peer", C.byref(self.st), _ptr(self.peer.ptrs), self.peer.world,
                          self.peer.rank, half * self.peer.stride,
                          C.c_void_p(self.peer.local.data_ptr() + 8 * red_off), stream)
                mark()
                self.peer.handle.barrier(channel=2 + half)
                mark()
                _lib.call("cb200_ice_update_gather", C.byref(self.st), int(it), _ptr(self.peer.ptrs),
                          self.peer.world, red_off, stream)
                self.launches += 1
            else:
                _lib.call("cb200_ice_update_peer", C.byref(self.st), int(it), _ptr(self.peer.ptrs),
                          self.peer.world, half * self.peer.stride, stream)
            mark()
            if timed:
                self.comm_events.append((marks[1], marks[2]))
                self.stage_events.append(marks)
            self.launches += 4 + (1 if self.n_far_units else 0)
        else:
            _lib.call("cb200_ice_combine", st, stream)
            if self.sweep_events is not None:
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record()
                _allreduce(self.t["s"], self.prob.group)
                c1.record()
                self.comm_events.append((c0, c1))
            else:
                _allreduce(self.t["s"], self.prob.group)
            _lib.call("cb200_ice_update", st, int(it), 0, stream)
            self.launches += 4 + (1 if self.n_far_units else 0)

    def run(self, check_every=8):
        """Iterate until every segment converged or max_iters.  Returns passes launched."""
        it = 0
        while it < self.max_iters:
            k = min(check_every, self.max_iters - it)
            for j in range(k):
                self.one_pass(it + j)
            it += k
            self.flag_host.copy_(self.t["all_done"], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            if int(self.flag_host[0]):
                break
        return it

    def results(self):
        t = self.t
        out = {k: t[k][: self.S].cpu().numpy() for k in
               ("seg_scale", "seg_var", "seg_iters", "seg_done", "seg_empty")}
        out["bias"] = to_host(t["bias"][: self.n])
        out["var_hist"] = t["var_hist"][: self.max_iters * self.S].cpu().numpy().reshape(self.max_iters, self.S)
        return out


def prefilter_bias(prob: IceProblem, chrom_offset, *, mad_max, min_nnz, min_count, blacklist, x0):
    """Host-side bin filters, statement for statement (_balance.py:366-413)."""
    n_bins = prob.table.n_bins
    if x0 is not None:
        bias = x0
        bias[np.isnan(bias)] = 0
    else:
        bias = np.ones(n_bins, dtype=float)

    if min_nnz > 0:
        bias[prob.marg_nnz < min_nnz] = 0

    marg = prob.marg_sum.copy()
    if min_count:
        bias[marg < min_count] = 0

    if mad_max > 0:
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            dev = getattr(prob.table, "device", None)
            if n_bins >= DEVICE_MEDIAN_MIN_BINS and dev is not None and torch.device(dev).type == "cuda":
                # the three medians are order statistics: selected exactly on the device (bit-identical to
                # np.median); every floating-point operation (/, log, abs, exp) stays in numpy as in the
                # reference.  At 3.1e6 bins the numpy medians were 160 of the 400 ms of a 1 kb shard's build.
                def up(a):
                    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

                chrom_med = segment_medians(up(marg), chrom_offset, positive_only=True)
                for (lo, hi), m in zip(zip(chrom_offset[:-1], chrom_offset[1:]), chrom_med):
                    marg[lo:hi] /= m
                logNzMarg = np.log(marg[marg > 0])
                whole = np.array([0, len(logNzMarg)])
                med_logNzMarg = segment_medians(up(logNzMarg), whole)[0] if len(logNzMarg) else np.nan
                dev_logNzMarg = (segment_medians(up(np.abs(logNzMarg - med_logNzMarg)), whole)[0]
                                 if len(logNzMarg) else np.nan)
            else:
                for lo, hi in zip(chrom_offset[:-1], chrom_offset[1:]):
                    c_marg = marg[lo:hi]
                    marg[lo:hi] /= np.median(c_marg[c_marg > 0])
                logNzMarg = np.log(marg[marg > 0])
                med_logNzMarg = np.median(logNzMarg)
                dev_logNzMarg = mad(logNzMarg)
            cutoff = np.exp(med_logNzMarg - mad_max * dev_logNzMarg)
            bias[marg < cutoff] = 0

    if blacklist is not None:
        bias[blacklist] = 0
    return bias


MAX_CHROMS_CIS_TOUCH = 8192


def cis_touch_flags(prob, bias, chrom_offset, group=None):
    """cis_only: which chromosomes are linked by stored pixels (r, c) with c in a chromosome BEFORE r's.

    Returns None (no such pixel: every symmetric-upper table) or two bool matrices [n_chroms, n_chroms]
    ``(any, masked)``: ``any[a, b]`` = a row of chromosome a holds a pixel in a column of the earlier
    chromosome b; ``masked[a, b]`` = one of those columns is a masked bin (``bias == 0`` after the bin
    filters).  See ``poisoned_chromosomes`` and cb200_cis_touch (include/cooler_b200.h)."""
    table = prob.table
    dev = table.device
    n_chroms = len(chrom_offset) - 1
    earlier = getattr(prob, "earlier", [])
    have = torch.tensor([1 if earlier else 0], dtype=torch.int32, device=dev)
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(have, op=dist.ReduceOp.MAX, group=group)
    if not int(have.item()):
        return None
    if n_chroms > MAX_CHROMS_CIS_TOUCH:
        raise NotImplementedError(
            f"cis_only on a square-storage table with more than {MAX_CHROMS_CIS_TOUCH} chromosomes")
    nb = np.diff(np.asarray(chrom_offset, dtype=np.int64))
    bin_chrom = torch.from_numpy(np.repeat(np.arange(n_chroms, dtype=np.int32), nb)).to(dev)
    bias_d = torch.from_numpy(np.ascontiguousarray(bias, dtype=np.float64)).to(dev)
    flags = torch.zeros((2, n_chroms * n_chroms), dtype=torch.uint8, device=dev)
    for px, rows, n in earlier:
        _lib.call("cb200_cis_touch", _ptr(px), _ptr(rows), int(n), _ptr(bin_chrom), n_chroms, _ptr(bias_d),
                  _ptr(flags[0]), _ptr(flags[1]), _stream())
    if group is not None:
        import torch.distributed as dist
        wide = flags.to(torch.int32)
        dist.all_reduce(wide, op=dist.ReduceOp.MAX, group=group)
        flags = wide
    f = flags.cpu().numpy().astype(bool).reshape(2, n_chroms, n_chroms)
    return f[0], f[1]


def poisoned_chromosomes(touch, seg_empty):
    """The chromosomes on which the reference's sequential ``_balance_cisonly`` (_balance.py:127-196)
    ends in NaN on a square-storage table.  It scans the ROWS of a chromosome (:150-151); a stored pixel
    (r, c) with c in an earlier chromosome is zeroed by ``_zero_trans`` but still weighted
    ``bias[r] * bias[c] * 0`` (:57-60), which is NaN when ``bias[c]`` already is -- c a masked bin of a
    finished chromosome (:186), or that chromosome all-NaN (empty, :166-170, or itself poisoned).  Then
    r's marginal, the mean of the nonzero marginals and every bias of r's chromosome are NaN, the
    variance never drops below ``tol`` and all ``max_iters`` passes are made (:173-182)."""
    any_, masked = touch
    n_chroms = len(seg_empty)
    all_nan = np.asarray(seg_empty, dtype=bool).copy()
    poisoned = np.zeros(n_chroms, dtype=bool)
    for c in range(1, n_chroms):
        if (masked[c, :c] | (any_[c, :c] & all_nan[:c])).any():
            poisoned[c] = all_nan[c] = True
    return poisoned


def balance_table(table: PixelTable, *, cis_only=False, trans_only=False, ignore_diags=2, mad_max=5,
                  min_nnz=10, min_count=0, blacklist=None, rescale_marginals=True, x0=None,
                  tol=1e-5, max_iters=200, group=None, problem=None, chromnames=None,
                  return_info=False, strip_bins=None, band=None, far_col_shift=None, far_compact=None):
    """Balance a device-resident table.  Same semantics as ``balance_cooler``.

    ``problem`` lets callers reuse the filtered copies (bench: inputs resident
    in HBM).  ``group``: torch.distributed process group when the table is a
    row shard (marginals are all-reduced across ranks).
    """
    _use_device(table.device)
    chrom_offset = table.chrom_offset
    n_bins = table.n_bins
    n_chroms = len(chrom_offset) - 1
    prob = problem or IceProblem(table, cis_only=cis_only, trans_only=trans_only,
                                 ignore_diags=ignore_diags, group=group, strip_bins=strip_bins, band=band,
                                 far_col_shift=far_col_shift, far_compact=far_compact)
    bias = prefilter_bias(prob, chrom_offset, mad_max=mad_max, min_nnz=min_nnz, min_count=min_count,
                          blacklist=blacklist, x0=x0)

    cweights = None
    if cis_only:
        seg_offset = chrom_offset
    else:
        seg_offset = np.array([0, n_bins])
        if trans_only:                                            # _balance.py:214-220
            cweights = 1.0 / np.concatenate(
                [[(1 - (hi - lo) / n_bins)] * (hi - lo)
                 for lo, hi in zip(chrom_offset[:-1], chrom_offset[1:])]) if n_bins else np.zeros(0)
            raw = getattr(table, "raw", None)
            if raw is not None and raw.nnz == 0 and not np.all(np.isfinite(cweights)):
                # one chromosome only (weights 1 / 0) and not a single stored pixel: the reference's marginal is
                # exactly zero (_balance.py:98-102: "empty"), whereas w_i * (sum over no pixels) would be inf * 0
                cweights = np.ones(n_bins)

    touch = cis_touch_flags(prob, bias, chrom_offset, group) if cis_only else None
    state = IceState(prob, bias, cweights, seg_offset, tol, max_iters)
    passes = state.run()
    res = state.results()
    bias[:] = res["bias"]
    poisoned = poisoned_chromosomes(touch, res["seg_empty"]) if touch is not None else None

    hit_limit = False
    if cis_only:
        scales = np.ones(n_chroms)
        variances = np.full_like(scales, np.nan)
        for cid, lo, hi in zip(range(n_chroms), chrom_offset[:-1], chrom_offset[1:]):
            if chromnames is not None:
                logger.info(chromnames[cid])
            if poisoned is not None and poisoned[cid]:            # square storage: poisoned_chromosomes()
                for _ in range(max_iters):
                    logger.info(f"variance is {np.nan}")
                res["seg_iters"][cid] = max_iters
                scale = var = np.nan
                bias[lo:hi] = np.nan
                hit_limit = True
                name = chromnames[cid] if chromnames is not None else cid
                warnings.warn(f"Iteration limit reached without convergence on {name}.",
                              ConvergenceWarning, stacklevel=2)
                scales[cid], variances[cid] = scale, var
                continue
            for v in res["var_hist"][: res["seg_iters"][cid], cid]:
                logger.info(f"variance is {v}")
            if res["seg_empty"][cid]:                             # _balance.py:166-170
                scale, var = np.nan, 0.0
                bias[lo:hi] = np.nan
            else:
                scale, var = res["seg_scale"][cid], res["seg_var"][cid]
                if not res["seg_done"][cid]:
                    hit_limit = True
                    name = chromnames[cid] if chromnames is not None else cid
                    warnings.warn(f"Iteration limit reached without convergence on {name}.",
                                  ConvergenceWarning, stacklevel=2)
            b = bias[lo:hi]
            b[b == 0] = np.nan
            scales[cid] = scale
            variances[cid] = var
            if rescale_marginals:
                bias[lo:hi] /= np.sqrt(scale)
        scale, var = scales, variances
    else:
        for v in res["var_hist"][: res["seg_iters"][0], 0]:
            logger.info(f"variance is {v}")
        if res["seg_empty"][0]:                                   # _balance.py:98-102
            scale, var = np.nan, 0.0
            bias[:] = np.nan
        else:
            scale, var = res["seg_scale"][0], res["seg_var"][0]
            if not res["seg_done"][0]:
                hit_limit = True
                warnings.warn("Iteration limit reached without convergence.",
                              ConvergenceWarning, stacklevel=2)
        bias[bias == 0] = np.nan
        if rescale_marginals:
            bias /= np.sqrt(scale)

    stats = {
        "tol": tol, "min_nnz": min_nnz, "min_count": min_count, "mad_max": mad_max,
        "cis_only": cis_only, "ignore_diags": ignore_diags, "scale": scale,
        "converged": var < tol, "var": var, "divisive_weights": False,
    }
    if return_info:
        info = {"passes": res["seg_iters"].tolist(), "launched_passes": passes,
                "nnz_eff": prob.nnz_eff, "bytes_per_pass": prob.bytes_per_pass(),
                "hit_limit": hit_limit, "gpu_launches": state.launches,
                "n_subcopies": len(prob.subs), "strip_shift": prob.strip_shift, "band": prob.band,
                "far_units": state.n_far_units,
                "far_pixels": sum(getattr(f, "n_pixels", f.nnz) for f in getattr(prob, "far_copies", []) if f is not None),
                "far_records": sum(f.nnz for f in getattr(prob, "far_copies", []) if f is not None),
                "far_rec_bytes": [f.rec_bytes for f in getattr(prob, "far_copies", []) if f is not None]}
        return bias, stats, info
    return bias, stats


def balance_arrays(bin1_offset, bin2_id, count, chrom_offset, *, cis_only=False, trans_only=False,
                   ignore_diags=2, device="cuda", group=None, row_range=None, return_info=False,
                   chromnames=None, strip_bins=None, band=None, **kwargs):
    """Balance a pixel table given as HOST arrays (numpy or torch, ideally pinned):
    ``bin1_offset`` int64[n_bins+1] (the cooler's indexes/bin1_offset), ``bin2_id`` int64 or int32,
    ``count`` any integer type within int32 range, or float32 / float64 (kept as float64: the reference
    computes in whatever dtype the column has, _balance.py:25-26).  The build is pipelined with the host->device copy (IceProblem.from_host);
    tables that need column strips go through PixelTable + IceProblem.  Other keyword
    arguments and the return value are those of ``balance_table``."""
    prob = None
    on_device = isinstance(bin2_id, torch.Tensor) and bin2_id.is_cuda       # read_cooler_device streamed it already
    if strip_bins is None and band is None and not on_device:
        prob = IceProblem.from_host(bin1_offset, bin2_id, count, chrom_offset, cis_only=cis_only,
                                    trans_only=trans_only, ignore_diags=ignore_diags, group=group,
                                    device=device, row_range=row_range)
    if prob is not None:
        table = prob.table
    else:
        table = PixelTable.from_arrays(bin1_offset, bin2_id, count, chrom_offset, device=device,
                                       row_range=row_range, float_values=True)
    return balance_table(table, cis_only=cis_only, trans_only=trans_only, ignore_diags=ignore_diags,
                         group=group, problem=prob, return_info=return_info, chromnames=chromnames,
                         strip_bins=strip_bins, band=band, **kwargs)


def balance_cooler(clr, *, cis_only=False, trans_only=False, ignore_diags=2, mad_max=5, min_nnz=10,
                   min_count=0, blacklist=None, rescale_marginals=True, x0=None, tol=1e-5,
                   max_iters=200, chunksize=10_000_000, map=map, use_lock=False, store=False,
                   store_name="weight", device="cuda", group=None):
    """Iterative correction of a cooler on the GPU.

    Same signature, return value ``(bias, stats)`` and side effects as
    ``cooler.balance_cooler`` (_balance.py:263-477).  ``chunksize``, ``map`` and
    ``use_lock`` are accepted for compatibility and ignored: the pixel table is
    read once and stays resident in HBM.  ``clr`` is a ``cooler.Cooler`` or any
    object with the same read protocol (e.g. ``cooler_b200.store.MemCooler``).

    ``group``: a ``torch.distributed`` process group (NCCL, one rank per GPU).  Where the
    reference hands pixel spans to a process pool (``map=Pool.imap_unordered``,
    cli/balance.py:237-259), every rank here reads the stored pixels of ITS contiguous bin1 row
    range ``pixels[bin1_offset[r_g]:bin1_offset[r_g+1]]`` (about nnz / world pixels,
    table.shard_rows_by_pixels), keeps that shard resident on its GPU, and the per-pass partial
    marginal vectors are summed over NVLink.  Every rank returns the same ``(bias, stats)``;
    rank 0 alone stores.  Call it on all ranks, e.g.
    ``torchrun --nproc-per-node 8 -m cooler_b200.cli balance FILE.cool``.
    """
    rows = None
    rank = 0
    if group is not None:
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        full_offset = np.asarray(clr._load_dset("indexes/bin1_offset"))
        rows = shard_rows_by_pixels(full_offset, rank, world)
    bin1_offset, bin2, count, chrom_offset = read_cooler_device(clr, rows=rows, device=device)
    try:
        chromnames = list(clr.chroms()["name"][:])
    except Exception:  # pragma: no cover - duck-typed stores without a chroms selector
        chromnames = None
    bias, stats = balance_arrays(
        bin1_offset, bin2, count, chrom_offset, device=device, row_range=rows, group=group,
        cis_only=cis_only, trans_only=trans_only, ignore_diags=ignore_diags, mad_max=mad_max,
        min_nnz=min_nnz, min_count=min_count, blacklist=blacklist,
        rescale_marginals=rescale_marginals, x0=x0, tol=tol, max_iters=max_iters,
        chromnames=chromnames if rank == 0 else None)

    if store and rank == 0:                                       # _balance.py:469-475
        with clr.open("r+") as grp:
            if store_name in grp["bins"]:
                del grp["bins"][store_name]
            h5opts = {"compression": "gzip", "compression_opts": 6}
            grp["bins"].create_dataset(store_name, data=bias, **h5opts)
            grp["bins"][store_name].attrs.update(stats)
    if store and group is not None:
        import torch.distributed as dist
        dist.barrier(group)                                       # the column exists when any rank returns
    return bias, stats


iterative_correction = balance_cooler  # alias (_balance.py:480)
