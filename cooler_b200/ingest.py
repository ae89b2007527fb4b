"""Pairs -> pixels on the GPU: the binning / aggregation core of ``cooler cload pairs``.

Mirrors ``sanitize_records(bins, schema="pairs", decode_chroms=False, ...)`` followed by
``aggregate_records(sort=True, count=True)`` (reference src/cooler/create/_ingest.py:68-214,
452-504; call site cli/cload.py:634-650) for integer chromosome ids and fixed-size bins: records
on unknown chromosomes (id < 0) are dropped, coordinates become 0-based, bounds are checked
(``BadInputError``), lower-triangle records are reflected / dropped / rejected, every mate gets
``bin = chrom_binoffset[chrom] + pos // binsize`` and the records are counted per (bin1, bin2)
by a stable radix sort on bin2, then on bin1, and a reduce-by-key -- the kernels of the
coarsening path (csrc/coarsen.cu, csrc/prims.cu).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .reduce import _aggregate_runs, _reduce_runs, _settle_dtypes
from .table import _new_px, _ptr, _stream, _use_device, exclusive_scan_i32, key_bits_for, n_tiles, sort_pairs, \
    swap_key_other

__all__ = ["bin_pairs", "BadInputError", "sanitize_pixels", "sanitize_records", "sanitize_bg2", "UnorderedPixels"]

_TRIL = {None: 0, "reflect": 1, "drop": 2, "raise": 3}


class BadInputError(ValueError):
    """Same role as cooler.create.BadInputError (create/_ingest.py:38)."""


def bin_pairs(chrom1_ids, pos1, chrom2_ids, pos2, chromsizes, binsize, *, is_one_based=True,
              tril_action="reflect", validate=True, device="cuda"):
    """Count read pairs per pixel.

    ``chrom*_ids``: integer chromosome ids in the order of ``chromsizes`` (< 0 = not requested,
    dropped); ``pos*``: anchor positions (1-based by default like the ``pairs`` schema).
    Returns ``{"bin1_id": int64, "bin2_id": int64, "count": int32}`` sorted by (bin1_id, bin2_id)
    -- what ``aggregate_records(sort=True)(sanitize_records(...)(chunk))`` yields.
    """
    if tril_action not in _TRIL:
        raise ValueError(f"Unknown tril_action value: '{tril_action}'")
    if binsize is None:
        raise NotImplementedError("variable-size bins are not supported by the CUDA ingest path")
    dev = _use_device(device)
    chromsizes = np.asarray(chromsizes, dtype=np.int64)
    nb = -(-chromsizes // int(binsize))
    binoff = np.r_[0, np.cumsum(nb)].astype(np.int64)
    n_bins = int(binoff[-1])
    if n_bins >= 2**31 - 2:
        raise ValueError("n_bins must fit int32")

    def dev_arr(a, dtype, what):
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        if t.numel() and validate and (t.dtype.is_floating_point or t.dtype == torch.bool):
            if what == "chrom":
                raise BadInputError("`chrom` column is non-integer. If string, convert to an enumeration first")
            raise BadInputError("Found a non-integer anchor column")
        return t.to(dtype).to(dev, non_blocking=True)

    c1, c2 = dev_arr(chrom1_ids, torch.int32, "chrom"), dev_arr(chrom2_ids, torch.int32, "chrom")
    p1, p2 = dev_arr(pos1, torch.int64, "pos"), dev_arr(pos2, torch.int64, "pos")
    n = int(c1.numel())
    if not (c2.numel() == p1.numel() == p2.numel() == n):
        raise ValueError("input columns differ in length")
    empty = {"bin1_id": np.zeros(0, np.int64), "bin2_id": np.zeros(0, np.int64), "count": np.zeros(0, np.int32)}
    if n == 0:
        return empty
    if validate and (int(c1.max()) >= len(chromsizes) or int(c2.max()) >= len(chromsizes)):
        raise BadInputError("chromosome id out of range")
    key = torch.empty(n, dtype=torch.int32, device=dev)
    val = torch.empty((n + 2, 2), dtype=torch.int32, device=dev)
    flags = torch.zeros(4, dtype=torch.int32, device=dev)
    binoff_d = torch.from_numpy(binoff.astype(np.int32)).to(dev)
    sizes_d = torch.from_numpy(chromsizes).to(dev)
    _lib.call("cb200_bin_pairs", _ptr(c1), _ptr(p1), _ptr(c2), _ptr(p2), n, _ptr(binoff_d), _ptr(sizes_d),
              int(binsize), int(bool(is_one_based)), _TRIL[tril_action], n_bins, _ptr(key), _ptr(val), _ptr(flags),
              _stream())
    f = flags.cpu().numpy()
    if validate and f[0]:
        raise BadInputError("Found an anchor position with negative value. Make sure your coordinates are "
                            "1-based or use the --zero-based option when loading.")
    if validate and f[1] & 1:
        raise BadInputError("Found an anchor position exceeding chromosome length")
    if f[1] & 2:                                          # create/_ingest.py:407-410 (boundscheck)
        raise BadInputError("Found a bin ID that exceeds the declared number of bins. "
                            "Check whether your bin table is correct.")
    if f[2]:
        raise BadInputError("Found lower triangle pairs")
    bits = key_bits_for(n_bins + 1)
    ks, vs = sort_pairs(key, val, bits)                   # by bin2
    del key, val
    kr, vr = swap_key_other(ks, vs, n)                    # key = bin1, val = {bin2, 1}
    del ks, vs
    kr, vr = sort_pairs(kr, vr, bits)                     # by bin1 (stable): (bin1, bin2) order
    n_valid = n - int(f[3])                               # dropped records sort behind every bin: cut them off
    if n_valid <= 0:                                      # (one thread per run would walk them serially)
        return empty
    k2, v2, n_out, overflow = _reduce_runs(kr, vr, n_valid, 0, dev)
    if n_out <= 0:
        return empty
    b1 = torch.empty(n_out, dtype=torch.int64, device=dev)
    b2 = torch.empty(n_out, dtype=torch.int64, device=dev)
    cnt = torch.empty(n_out, dtype=torch.int32, device=dev)
    _lib.call("cb200_unpack_pixels", _ptr(k2), _ptr(v2), n_out, _ptr(b1), _ptr(b2), _ptr(cnt), _stream())
    return {"bin1_id": b1.cpu().numpy(), "bin2_id": b2.cpu().numpy(), "count": cnt.cpu().numpy()}


# ------------------------------------------------------------------------------------------------
# Already-binned or bedGraph-2D text records in any order: the sanitizers of ``cooler load``
# (host, vectorised numpy -- the text parser dominates them) and the device side of
# ``create_from_unordered`` (create/_create.py:716-812): where the reference sorts every chunk into a
# temporary cooler file and k-way merges the files (CoolerMerger, every value column summed), the
# chunks are parked on the device and ONE pair of stable radix sorts + one reduce-by-key per value
# column does the merge (the kernels of the coarsening / merging path).
# ------------------------------------------------------------------------------------------------
def sanitize_pixels(chunk, *, is_one_based=False, tril_action="reflect"):
    """``_sanitize_pixels`` (create/_ingest.py:307-349) for a dict / frame of arrays with ``bin1_id``,
    ``bin2_id`` and value columns: one-based ids become zero-based; pixels below the diagonal are
    mirrored ('reflect'), discarded ('drop'), refused ('raise') or kept (None).  Order is kept (the
    device sorts).  Returns a dict of numpy arrays."""
    if tril_action not in _TRIL:
        raise ValueError(f"Unknown tril_action value: '{tril_action}'")
    out = {k: np.asarray(v) for k, v in dict(chunk).items()}
    b1, b2 = out["bin1_id"].astype(np.int64), out["bin2_id"].astype(np.int64)
    if is_one_based:
        b1, b2 = b1 - 1, b2 - 1
    if tril_action is not None:
        low = b1 > b2
        if low.any():
            if tril_action == "raise":
                raise BadInputError("Found bin1_id greater than bin2_id")
            if tril_action == "reflect":
                b1, b2 = np.where(low, b2, b1), np.where(low, b1, b2)
            else:
                keep = ~low
                b1, b2 = b1[keep], b2[keep]
                out = {k: v[keep] for k, v in out.items()}
    out["bin1_id"], out["bin2_id"] = b1, b2
    return out


def sanitize_records(chunk, chromnames, chromsizes, bins_chrom, bins_start, bins_end, *, anchor_field="start",
                     is_one_based=False, tril_action="reflect", validate=True):
    """``_sanitize_records`` (create/_ingest.py:68-214) with string chromosome columns (``decode_chroms``):
    columns ``chrom1, <anchor>1, chrom2, <anchor>2`` (+ value columns) -> ``bin1_id, bin2_id`` (+ value
    columns); ``anchor_field`` is ``start`` for the ``bg2`` preset (:45-55) and ``pos`` for ``pairs`` (:56-66).
    Records on chromosomes that are not in the bin table are dropped, anchors are checked against the
    chromosome lengths, records below the diagonal (by chromosome order, then anchor) are mirrored /
    dropped / refused, and every anchor gets the bin that contains it -- ``offset[chrom] + anchor //
    binsize`` for fixed-size bins, else the last bin of the chromosome starting at or before it.  Order is
    kept (the device sorts)."""
    if tril_action not in _TRIL:
        raise ValueError(f"Unknown tril_action value: '{tril_action}'")
    import pandas as pd
    f1, f2 = anchor_field + "1", anchor_field + "2"
    out = {k: np.asarray(v) for k, v in dict(chunk).items()}
    idx = pd.Index([str(c) for c in chromnames])
    c1 = idx.get_indexer(pd.Index(out["chrom1"]).astype(str)).astype(np.int64)
    c2 = idx.get_indexer(pd.Index(out["chrom2"]).astype(str)).astype(np.int64)
    a1, a2 = out[f1], out[f2]
    if validate and not (np.issubdtype(a1.dtype, np.integer) and np.issubdtype(a2.dtype, np.integer)):
        raise BadInputError("Found a non-integer anchor column")
    known = (c1 >= 0) & (c2 >= 0)
    if not known.all():
        c1, c2, a1, a2 = c1[known], c2[known], a1[known], a2[known]
        out = {k: v[known] for k, v in out.items()}
    a1, a2 = a1.astype(np.int64), a2.astype(np.int64)
    if is_one_based:
        a1, a2 = a1 - 1, a2 - 1
    sizes = np.asarray(chromsizes, dtype=np.int64)
    if validate and len(a1):
        if ((a1 < 0) | (a2 < 0)).any():
            raise BadInputError("Found an anchor position with negative value. Make sure your coordinates are "
                                "1-based or use the --zero-based option when loading.")
        if ((a1 > sizes[c1]) | (a2 > sizes[c2])).any():
            raise BadInputError("Found an anchor position exceeding chromosome length")
    if tril_action is not None and len(a1):
        low = (c1 > c2) | ((c1 == c2) & (a1 > a2))
        if low.any():
            if tril_action == "raise":
                raise BadInputError("Found lower triangle pairs")
            if tril_action == "reflect":
                c1, c2 = np.where(low, c2, c1), np.where(low, c1, c2)
                a1, a2 = np.where(low, a2, a1), np.where(low, a1, a2)
            else:
                keep = ~low
                c1, c2, a1, a2 = c1[keep], c2[keep], a1[keep], a2[keep]
                out = {k: v[keep] for k, v in out.items()}
    bins_chrom = np.asarray(bins_chrom, dtype=np.int64)
    bins_start = np.asarray(bins_start, dtype=np.int64)
    nb = np.bincount(bins_chrom, minlength=len(sizes))
    binoff = np.r_[0, np.cumsum(nb)]                                   # GenomeSegmentation.chrom_binoffset
    from .create import _binsize
    binsize = _binsize(bins_chrom, bins_start, np.asarray(bins_end, dtype=np.int64))
    if binsize is not None:
        b1, b2 = binoff[c1] + a1 // binsize, binoff[c2] + a2 // binsize
    else:
        abspos = np.r_[0, np.cumsum(sizes)]                            # chrom_abspos / start_abspos
        start_abs = abspos[bins_chrom] + bins_start
        b1 = np.minimum(np.searchsorted(start_abs, abspos[c1] + a1, side="right") - 1, binoff[c1 + 1] - 1)
        b2 = np.minimum(np.searchsorted(start_abs, abspos[c2] + a2, side="right") - 1, binoff[c2 + 1] - 1)
    res = {"bin1_id": b1.astype(np.int64), "bin2_id": b2.astype(np.int64)}
    sided = {"chrom1", "chrom2", "start1", "start2", "end1", "end2", "pos1", "pos2"}
    res.update({k: v for k, v in out.items() if k not in sided})
    return res


def sanitize_bg2(chunk, chromnames, chromsizes, bins_chrom, bins_start, bins_end, **kwargs):
    """``sanitize_records`` with the ``bg2`` preset: anchors are the ``start`` columns."""
    return sanitize_records(chunk, chromnames, chromsizes, bins_chrom, bins_start, bins_end,
                            anchor_field="start", **kwargs)


class UnorderedPixels:
    """Device side of ``create_from_unordered``: ``add`` parks a sanitized chunk of (bin1, bin2) ids on
    the device -- its duplicate check (``create(..., dupcheck=True)`` of every temporary cooler,
    create/_ingest.py:421-427) is a sort of the chunk + a run count --; ``merge`` sorts all parked
    pixels by (bin1, bin2), stable, and sums every value column per pixel in input order."""

    def __init__(self, n_bins, device="cuda"):
        self.dev = _use_device(device)
        self.n_bins = int(n_bins)
        if self.n_bins >= 2**31 - 2:
            raise ValueError("n_bins must fit int32")
        self.bits = key_bits_for(max(self.n_bins, 1))
        self.keys, self.vals, self.n = [], [], 0

    def _sorted(self, key, val, n):
        ks, vs = sort_pairs(key, val, self.bits)                       # by bin2
        kr, vr = swap_key_other(ks, vs, n)                             # key = bin1, val = {bin2, position}
        del ks, vs
        return sort_pairs(kr, vr, self.bits)                           # by bin1 (stable): (bin1, bin2) order

    def add(self, bin1, bin2, dupcheck=True):
        m = len(bin1)
        if m == 0:
            return
        if self.n + m >= 2**31 - 1:
            raise NotImplementedError("the unordered ingest path indexes pixels with int32")
        b1h, b2h = np.asarray(bin1), np.asarray(bin2)
        if min(int(b1h.min()), int(b2h.min())) < 0 or max(int(b1h.max()), int(b2h.max())) >= self.n_bins:
            # the radix sorts look at the bits of a valid bin id only: an id outside [0, n_bins) must not get here
            raise BadInputError("Found a bin ID that exceeds the declared number of bins. "
                                "Check whether your bin table is correct.")
        b1 = torch.from_numpy(np.ascontiguousarray(bin1, dtype=np.int32)).to(self.dev)
        key = torch.from_numpy(np.ascontiguousarray(bin2, dtype=np.int32)).to(self.dev)
        pos = torch.arange(self.n, self.n + m, dtype=torch.int32, device=self.dev)
        val = _new_px(m, self.dev)
        _lib.call("cb200_pack_pixels_i32", _ptr(b1), _ptr(pos), m, _ptr(val), _stream())   # {bin1, position}
        if dupcheck:
            kr, vr = self._sorted(key, val, m)
            tile_count = torch.zeros(max(n_tiles(m), 1), dtype=torch.int32, device=self.dev)
            _lib.call("cb200_runs_count", _ptr(kr), _ptr(vr), m, _ptr(tile_count), _stream())
            if int(exclusive_scan_i32(tile_count[: n_tiles(m)])[-1].item()) != m:
                raise BadInputError("Found duplicate pixels")
        self.keys.append(key)
        self.vals.append(val[:m])
        self.n += m

    def merge(self, columns, agg=None):
        """``columns``: {name: host array over ALL added pixels in the order they were added}; ``agg``:
        {name: "sum" | "min" | "max" | "first" | "last" | "mean"}, default sum.  Returns (bin1 int64, bin2
        int64, {name: aggregated column}) sorted by (bin1, bin2), one entry per pixel -- CoolerMerger's
        ``groupby([bin1_id, bin2_id], sort=True).sum()`` over the chunks, or the ``aggregate_records`` of one
        chunk of records (create/_ingest.py:452-504)."""
        agg = dict(agg or {})
        n = self.n
        if n == 0:
            e = np.zeros(0, dtype=np.int64)
            return e, e, {c: np.asarray(v)[:0] for c, v in columns.items()}
        key = torch.cat(self.keys)
        val = torch.cat(self.vals + [torch.zeros((2, 2), dtype=torch.int32, device=self.dev)])
        self.keys, self.vals = [], []
        kr, vr = self._sorted(key, val, n)
        del key, val
        cols = {c: np.asarray(v) for c, v in columns.items()}
        b1, b2, out = _aggregate_runs(kr, vr, n, cols, agg, self.dev)
        return b1, b2, _settle_dtypes(out, {c: v.dtype for c, v in cols.items()}, agg)
