"""Balanced matrix fetch on the GPU -- the consumer side of the balancing weights.

Mirrors the query surface of ``cooler.Cooler.matrix`` (reference src/cooler/api.py:326-409,
665-800; core/_selectors.py:130-190; core/_rangequery.py): ``matrix(clr, balance=True,
sparse=False).fetch("chr1:10,000,000-20,000,000", "chr2")`` or ``[i0:i1, j0:j1]``.  The pixel
table is uploaded once (``PixelTable``) and every query is a 2-D range query on the resident
bin1-major copy, the lower-triangle fill for symmetric-upper storage and ``w_i * w_j * x`` --
all in ``libcooler_b200.so`` (csrc/fetch.cu).  Dense windows come back as numpy arrays
(int32 counts, or float64 when balanced); ``sparse=True`` gives a ``scipy.sparse.coo_matrix``;
``as_pixels=True`` a data frame of the explicitly stored pixels (+ ``balanced`` column).
"""
from __future__ import annotations

import ctypes as C
import re

import numpy as np
import torch

from . import _lib
from .table import PixelTable, _ptr, _stream, _use_device, exclusive_scan_i32

__all__ = ["matrix", "MatrixSelector", "parse_region", "region_to_extent"]


# ------------------------------------------------------------------ regions (util.py:59-189)
def parse_humanized(s):
    _, value, unit = re.compile("([0-9,.]+)").split(s.replace(",", ""))
    if not len(unit):
        return int(value)
    value = float(value)
    unit = unit.upper().strip()
    mult = {"K": 1e3, "KB": 1e3, "M": 1e6, "MB": 1e6, "G": 1e9, "GB": 1e9}
    if unit not in mult:
        raise ValueError(f"Unknown unit '{unit}'")
    return int(value * mult[unit])


_REGION_TOKEN = re.compile(r"\s*(?:(-)|([0-9,]+(?:\.[0-9]*)?(?:[a-z]+)?)|(.+))", re.IGNORECASE)


def parse_region_string(s):
    """'chr5:10,100,000-30,000,000' -> (chrom, start | None, end | None), with the reference's token rules
    (util.parse_region_string, util.py:78-140): after the colon come a coordinate, a hyphen and optionally a
    second coordinate; whatever follows them is ignored."""
    parts = s.split(":")
    chrom = parts[0].strip()
    if not len(chrom):
        raise ValueError("Chromosome name cannot be empty")
    if len(parts) < 2:
        return chrom, None, None
    tokens = [("hyphen" if m.group(1) is not None else "coord" if m.group(2) is not None else "other",
               m.group(m.lastindex)) for m in _REGION_TOKEN.finditer(parts[1])]

    def take(k, kind):
        if k >= len(tokens):
            raise ValueError(f"Expected {kind.upper()} token missing")
        if tokens[k][0] != kind:
            raise ValueError(f'Unexpected token "{tokens[k][1]}"')
        return tokens[k][1]

    start = parse_humanized(take(0, "coord"))
    take(1, "hyphen")
    if len(tokens) < 3:
        return chrom, start, None
    end = parse_humanized(take(2, "coord"))
    if end < start:
        raise ValueError("End coordinate less than start")
    return chrom, start, end


def parse_region(reg, chromsizes=None):
    """(chrom, start, end) of a UCSC-style string or triple (util.parse_region, util.py:143-189)."""
    if isinstance(reg, str):
        chrom, start, end = parse_region_string(reg)
    else:
        chrom, start, end = reg
        start = int(start) if start is not None else start
        end = int(end) if end is not None else end
    try:
        clen = chromsizes[chrom] if chromsizes is not None else None
    except KeyError as e:
        raise ValueError(f"Unknown sequence label: {chrom}") from e
    start = 0 if start is None else start
    if end is None:
        if clen is None:
            raise ValueError("Cannot determine end coordinate.")
        end = clen
    if end < start:
        raise ValueError("End cannot be less than start")
    if start < 0 or (clen is not None and end > clen):
        raise ValueError(f"Genomic region out of bounds: [{start}, {end})")
    return chrom, int(start), int(end)


def region_to_extent(clr, region):
    """Bin range [lo, hi) of a genomic region (core/_rangequery.py:11-31)."""
    chromsizes = clr.chromsizes
    chrom, start, end = parse_region(region, chromsizes)
    cid = list(chromsizes.index).index(chrom)
    chrom_offset = np.asarray(clr._load_dset("indexes/chrom_offset"))
    binsize = clr.binsize
    if binsize is not None:
        return (int(chrom_offset[cid]) + int(np.floor(start / binsize)),
                int(chrom_offset[cid]) + int(np.ceil(end / binsize)))
    lo, hi = int(chrom_offset[cid]), int(chrom_offset[cid + 1])
    starts = np.asarray(clr._load_dset("bins/start"))[lo:hi]
    return (lo + int(np.searchsorted(starts, start, "right") - 1),
            lo + int(np.searchsorted(starts, end, "left")))


# ------------------------------------------------------------------ the selector
class MatrixSelector:
    """2-D range selector over a device-resident pixel table (``RangeSelector2D`` semantics)."""

    def __init__(self, clr, table, weights, *, sparse=False, as_pixels=False, divisive_weights=False,
                 fill_lower=True, weight_name="weight", field="count", values=None):
        self.clr, self.table = clr, table
        self.sparse, self.as_pixels = sparse, as_pixels
        # another value column than the table's int32 counts: float64 on the device, one value per stored
        # pixel in file order; results are cast back to the column's dtype (exact: every cell holds one pixel)
        self.field = field
        self.values, self.value_dtype = None, None
        if values is not None:
            v = np.asarray(values)
            if len(v) != table.raw.nnz:
                raise ValueError("values must have one entry per stored pixel")
            if v.dtype.kind in "iu" and len(v) and max(abs(int(v.max())), abs(int(v.min()))) > 2**53:
                raise NotImplementedError("integer value columns beyond 2**53 are not supported by the CUDA path")
            self.value_dtype = v.dtype
            self.values = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(table.device)
        self.divisive = bool(divisive_weights)
        self.weight_name = weight_name
        n = table.n_bins
        self.shape = (n, n)
        symmetric = getattr(clr, "storage_mode", "symmetric-upper") == "symmetric-upper" if clr is not None else True
        self.fill_lower = bool(fill_lower and symmetric and not as_pixels)
        self.weights = None
        if weights is not None:
            w = np.ascontiguousarray(weights, dtype=np.float64)
            if len(w) != n:
                raise ValueError("weights must have one entry per bin")
            self.weights = torch.from_numpy(w).to(table.device)

    # -- queries -------------------------------------------------------------------------
    def fetch(self, region, region2=None):
        if self.clr is None:
            raise ValueError("region queries need a cooler (bin table); use [i0:i1, j0:j1]")
        i0, i1 = region_to_extent(self.clr, region)
        j0, j1 = (i0, i1) if region2 is None else region_to_extent(self.clr, region2)
        return self._slice(i0, i1, j0, j1)

    def __getitem__(self, key):
        n = self.shape[0]
        if not isinstance(key, tuple):
            key = (key, slice(None))
        if len(key) != 2:
            raise IndexError("too many indices for a 2-D selector")

        def rng(k):
            if isinstance(k, (int, np.integer)):
                k = int(k)
                if k < 0:
                    k += n
                if not 0 <= k < n:                       # core/_selectors.py:46-51
                    raise IndexError("index is out of bounds")
                return k, k + 1
            lo, hi, step = k.indices(n)
            if step != 1:
                raise IndexError("slice step is not supported")
            return lo, max(hi, lo)
        (i0, i1), (j0, j1) = rng(key[0]), rng(key[1])
        return self._slice(i0, i1, j0, j1)

    def _slice(self, i0, i1, j0, j1):
        dev = _use_device(self.table.device)
        raw = self.table.raw
        cs = raw.c_struct()
        n = self.table.n_bins
        h, w = i1 - i0, j1 - j0
        fill = 1 if self.fill_lower else 0
        fv = self.values is not None
        vdt = torch.float64 if fv else torch.int32
        sfx = "_f64" if fv else ""
        vargs = (_ptr(self.values),) if fv else ()

        def native(a):                                   # back to the column's own dtype
            return a.astype(self.value_dtype, copy=False) if fv else a
        if not (self.sparse or self.as_pixels):
            dense = torch.empty(max(h * w, 1), dtype=vdt, device=dev)
            _lib.call("cb200_fetch_dense" + sfx, C.byref(cs), *vargs, n, i0, i1, j0, j1, fill, _ptr(dense), _stream())
            if self.weights is None:
                return native(dense[: h * w].reshape(h, w).cpu().numpy())
            out = torch.empty(max(h * w, 1), dtype=torch.float64, device=dev)
            _lib.call("cb200_fetch_balance_dense" + sfx, _ptr(dense), h, w, _ptr(self.weights[i0:i1]),
                      _ptr(self.weights[j0:j1]), int(self.divisive), _ptr(out), _stream())
            return out[: h * w].reshape(h, w).cpu().numpy()
        q = h + (w if fill else 0)
        counts = torch.zeros(max(q, 1), dtype=torch.int32, device=dev)
        _lib.call("cb200_fetch_count", C.byref(cs), n, i0, i1, j0, j1, fill, _ptr(counts), _stream())
        off = exclusive_scan_i32(counts[:q])
        total = int(off[-1].item())
        row = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        col = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        val = torch.empty(max(total, 1), dtype=vdt, device=dev)
        _lib.call("cb200_fetch_coo" + sfx, C.byref(cs), *vargs, n, i0, i1, j0, j1, fill, _ptr(off), _ptr(row),
                  _ptr(col), _ptr(val), _stream())
        bal = None
        if self.weights is not None:
            bal = torch.empty(max(total, 1), dtype=torch.float64, device=dev)
            _lib.call("cb200_fetch_balance_coo" + sfx, _ptr(row), _ptr(col), _ptr(val), total,
                      _ptr(self.weights[i0:i1]), _ptr(self.weights[j0:j1]), int(self.divisive), _ptr(bal), _stream())
            bal = bal[:total].cpu().numpy()
        row, col, val = (t[:total].cpu().numpy() for t in (row, col, val))
        val = native(val)
        if self.as_pixels:
            import pandas as pd
            df = pd.DataFrame({"bin1_id": row.astype(np.int64) + i0, "bin2_id": col.astype(np.int64) + j0,
                               self.field: val})
            if bal is not None:
                df["balanced"] = bal
            return df
        from scipy.sparse import coo_matrix
        data = bal if bal is not None else val
        return coo_matrix((data, (row, col)), shape=(h, w))


# The 4DN data portal and hic2cool store these weight vectors in divisive form (api.py:28)
_4DN_DIVISIVE_WEIGHTS = {"KR", "VC", "VC_SQRT"}


def matrix(clr, *, field=None, balance=True, sparse=False, as_pixels=False, join=False,
           ignore_index=True, divisive_weights=None, chunksize=10_000_000, fill_lower=True,
           weights=None, table=None, device="cuda"):
    """``clr.matrix(...)`` on the GPU: returns a selector supporting ``.fetch(region[, region2])``
    and ``[i0:i1, j0:j1]``.

    ``balance``: True -> the ``weight`` column of the bin table, a string -> that column, False ->
    raw counts; ``weights`` passes the weight vector directly (e.g. the bias just returned by
    ``balance_cooler``).  ``table`` reuses an uploaded ``PixelTable``.  ``field``: any numeric column of the
    pixel table (default ``count``); columns other than an int32-class ``count`` are fetched through the
    float64 kernels and returned in their own dtype.  ``as_pixels`` with ``join=True`` and
    ``ignore_index=False`` are not supported by the CUDA path.
    """
    field = "count" if field is None else field
    if isinstance(balance, str) and balance in _4DN_DIVISIVE_WEIGHTS and divisive_weights is None:
        divisive_weights = True                                        # api.py:367-368
    if join and as_pixels:
        raise NotImplementedError("as_pixels with join=True is not supported by the CUDA path")
    name = balance if isinstance(balance, str) else "weight"
    if weights is None and balance:
        with clr.open("r") as grp:
            if name not in grp["bins"]:
                raise ValueError(f"No column 'bins/{name}'found. Use ``cooler.balance_cooler`` to "
                                 "calculate balancing weights or set balance=False.")   # api.py:733-738
            weights = np.asarray(grp["bins"][name][:], dtype=np.float64)
    if not balance:
        weights = None
    dtypes = clr.pixels().dtypes
    if field not in dtypes:
        raise KeyError(field)
    dt = np.dtype(dtypes[field])
    # int32-class counts live in the table's records; any other column (another field, floating-point or
    # wide counts) is fetched as float64 values next to a table that only carries the structure
    values = None
    if not (field == "count" and dt.kind in "iu" and dt.itemsize <= 4):
        with clr.open("r") as grp:
            values = np.asarray(grp["pixels"][field][:])
        if values.dtype.kind not in "iuf":
            raise TypeError(f"pixel column '{field}' is not numeric")
    if table is None:
        table = PixelTable.from_cooler(clr, device=device, structure_only=values is not None)
    return MatrixSelector(clr, table, weights, sparse=sparse, as_pixels=as_pixels,
                          divisive_weights=divisive_weights, fill_lower=fill_lower, weight_name=name,
                          field=field, values=values)
