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
from .reduce import _reduce_runs
from .table import _ptr, _stream, _use_device, key_bits_for, sort_pairs, swap_key_other

__all__ = ["bin_pairs", "BadInputError"]

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
