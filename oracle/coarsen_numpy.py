"""TEST INFRASTRUCTURE ONLY (oracle) -- CPU restatement of the reference's
coarsening path, open2c/cooler v0.10.4 ``src/cooler/_reduce.py:501-642``
(``CoolerCoarsener``), in plain numpy.  Only tests/, ``smoke()`` and
``bench.py``'s CPU legs may import it; the product never does.

Parity status: PINNED against (a) the reference's own golden files
(tests/data/toy.symm.upper.{2,4}.bg2, toy.asymm.{2,4,8,16,32}.bg2 -- the text
twins of the .cool goldens compared in tests/test_reduce.py:83-175) and
(b) golden outputs of the unmodified reference run through oracle/refshim.py on
GM12878 2 Mb (factors 2, 5, 10), odd.1 (factor 4, tests/test_reduce.py:127-137)
and a variable-bin table; see tests/golden/make_golden.py.

The reference re-derives bin ids from genomic coordinates (join with the bin
table, then ``chrom_binoffset[chrom] + floor(start / new_binsize)`` for fixed
bins, _reduce.py:611-614, or ``searchsorted(start_abspos, abs_start,
'right') - 1`` for variable bins, _reduce.py:601-609).  This restatement does
exactly that, coordinate-based, so it is an independent check of the
closed-form id map the CUDA path uses.
"""
from __future__ import annotations

import numpy as np


def coarsen_bins(bins_chrom, bins_start, bins_end, chromsizes, factor):
    """New bin table (_reduce.py:563-580): every ``factor``-th start per
    chromosome; end = end of the last pooled bin, or the chromosome length."""
    bins_chrom = np.asarray(bins_chrom)
    n_chroms = len(chromsizes)
    out_c, out_s, out_e = [], [], []
    for c in range(n_chroms):
        sel = np.flatnonzero(bins_chrom == c)
        if len(sel) == 0:
            continue
        start = np.asarray(bins_start)[sel][::factor]
        end = np.asarray(bins_end)[sel][factor - 1::factor]
        if len(end) < len(start):
            end = np.r_[end, chromsizes[c]]
        out_c.append(np.full(len(start), c, dtype=np.int32))
        out_s.append(start)
        out_e.append(end)
    return (np.concatenate(out_c), np.concatenate(out_s).astype(np.int64),
            np.concatenate(out_e).astype(np.int64))


def _binsize_of(bins_chrom, bins_start, bins_end):
    """util.get_binsize (util.py:393-411): uniform width ignoring each
    chromosome's last bin, else None."""
    sizes = set()
    bins_chrom = np.asarray(bins_chrom)
    width = np.asarray(bins_end) - np.asarray(bins_start)
    for c in np.unique(bins_chrom):
        w = width[bins_chrom == c][:-1]
        sizes.update(np.unique(w).tolist())
        if len(sizes) > 1:
            return None
    return next(iter(sizes)) if len(sizes) == 1 else None


def coarsen_pixels(px, bins_chrom, bins_start, bins_end, chromsizes, factor,
                   count_dtype=None):
    """Restatement of ``CoolerCoarsener._aggregate`` over the whole table
    (chunking never changes content: edges respect coarse rows, _reduce.py:550-561).

    Returns (bin1_id int64, bin2_id int64, count, new_bins tuple).
    """
    bins_chrom = np.asarray(bins_chrom)
    bins_start = np.asarray(bins_start, dtype=np.int64)
    chromsizes = np.asarray(chromsizes, dtype=np.int64)
    new_c, new_s, new_e = coarsen_bins(bins_chrom, bins_start, bins_end, chromsizes, factor)
    n_chroms = len(chromsizes)
    nb_per = np.bincount(new_c, minlength=n_chroms)
    chrom_binoffset = np.r_[0, np.cumsum(nb_per)]            # util.py:803
    chrom_abspos = np.r_[0, np.cumsum(chromsizes)]           # util.py:804
    new_binsize = _binsize_of(new_c, new_s, new_e)

    b1 = np.asarray(px["bin1_id"], dtype=np.int64)
    b2 = np.asarray(px["bin2_id"], dtype=np.int64)
    c1, c2 = bins_chrom[b1], bins_chrom[b2]
    s1, s2 = bins_start[b1], bins_start[b2]
    if new_binsize is None:                                  # 601-609
        start_abspos = chrom_abspos[new_c] + new_s
        n1 = np.searchsorted(start_abspos, chrom_abspos[c1] + s1, side="right") - 1
        n2 = np.searchsorted(start_abspos, chrom_abspos[c2] + s2, side="right") - 1
    else:                                                    # 611-614
        n1 = chrom_binoffset[c1] + np.floor(s1 / new_binsize).astype(int)
        n2 = chrom_binoffset[c2] + np.floor(s2 / new_binsize).astype(int)

    # groupby([bin1_id, bin2_id], sort=True).sum()  (616-620)
    cnt = np.asarray(px["count"])
    order = np.lexsort((n2, n1))
    n1, n2, cnt = n1[order], n2[order], cnt[order]
    if len(n1):
        head = np.r_[True, (n1[1:] != n1[:-1]) | (n2[1:] != n2[:-1])]
        starts = np.flatnonzero(head)
        acc = np.add.reduceat(cnt.astype(np.int64 if cnt.dtype.kind in "iu" else np.float64), starts)
        n1, n2 = n1[starts], n2[starts]
    else:
        acc = cnt.astype(np.int64)
    if count_dtype is None:
        count_dtype = cnt.dtype
    return n1.astype(np.int64), n2.astype(np.int64), acc.astype(count_dtype), (new_c, new_s, new_e)


def merge_pixels(tables, count_dtype=None):
    """Restatement of ``CoolerMerger.__iter__`` over whole tables (_reduce.py:170-208):
    concatenate the pixel tables, ``groupby([bin1_id, bin2_id], sort=True).sum()``.
    Epoch boundaries (merge_breakpoints) only chunk the output; content is independent.
    tables: list of dicts with bin1_id, bin2_id, count."""
    b1 = np.concatenate([np.asarray(t["bin1_id"], dtype=np.int64) for t in tables])
    b2 = np.concatenate([np.asarray(t["bin2_id"], dtype=np.int64) for t in tables])
    cnt = np.concatenate([np.asarray(t["count"]) for t in tables])
    if count_dtype is None:
        count_dtype = np.result_type(*[np.asarray(t["count"]).dtype for t in tables])   # merge_coolers: 286-291
    order = np.lexsort((b2, b1))
    b1, b2, cnt = b1[order], b2[order], cnt[order]
    if len(b1) == 0:
        return b1, b2, cnt.astype(count_dtype)
    head = np.r_[True, (b1[1:] != b1[:-1]) | (b2[1:] != b2[:-1])]
    starts = np.flatnonzero(head)
    acc = np.add.reduceat(cnt.astype(np.int64 if cnt.dtype.kind in "iu" else np.float64), starts)
    return b1[starts], b2[starts], acc.astype(count_dtype)


def merge_columns(tables, columns, agg=None):
    """Restatement of ``CoolerMerger.__iter__`` with general value columns / aggregators
    (_reduce.py:144-149 defaults every column to 'sum'; :188-203 ``pd.concat`` of the inputs in
    cooler order, then ``groupby([bin1_id, bin2_id], sort=True).aggregate(agg)``): a stable sort
    keeps the concatenation order inside a group (what 'first' / 'last' and the floating-point
    summation order see); numeric dtypes promote as ``pd.concat`` promotes them; integer sums /
    min / max / first / last keep the promoted dtype, 'mean' is float64 (float32 for a float32 column).
    ``tables``: dicts with bin1_id, bin2_id and the value columns.  Returns a dict of arrays.
    PINNED by tests/golden/merge_agg_golden.npz (unmodified reference, make_merge_golden.py)."""
    ops = {c: "sum" for c in columns}
    ops.update(agg or {})
    b1 = np.concatenate([np.asarray(t["bin1_id"], dtype=np.int64) for t in tables])
    b2 = np.concatenate([np.asarray(t["bin2_id"], dtype=np.int64) for t in tables])
    order = np.lexsort((b2, b1))                             # stable
    b1, b2 = b1[order], b2[order]
    head = np.r_[True, (b1[1:] != b1[:-1]) | (b2[1:] != b2[:-1])] if len(b1) else np.zeros(0, bool)
    starts = np.flatnonzero(head)
    ends = np.r_[starts[1:], len(b1)]
    out = {"bin1_id": b1[starts], "bin2_id": b2[starts]}
    for c, op in ops.items():
        v = np.concatenate([np.asarray(t[c]) for t in tables])[order]
        wide = v.astype(np.int64 if v.dtype.kind in "iu" else np.float64)
        if not len(starts):
            out[c] = v[:0] if op != "mean" else wide[:0].astype(np.float64)
        elif op == "sum":
            out[c] = np.add.reduceat(wide, starts).astype(v.dtype)
        elif op == "mean":                                   # float64, or float32 for a float32 column (pandas)
            out[c] = np.add.reduceat(wide.astype(np.float64), starts) / (ends - starts)
            if v.dtype == np.float32:
                out[c] = out[c].astype(np.float32)
        elif op == "min":
            out[c] = np.minimum.reduceat(v, starts)
        elif op == "max":
            out[c] = np.maximum.reduceat(v, starts)
        elif op == "first":
            out[c] = v[starts]
        elif op == "last":
            out[c] = v[ends - 1]
        else:
            raise ValueError(op)
    return out
