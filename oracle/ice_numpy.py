"""TEST INFRASTRUCTURE ONLY (oracle) -- CPU restatement of the reference's ICE
matrix balancing, written from the algorithm in open2c/cooler v0.10.4
``src/cooler/_balance.py``.  Only tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module;
the product package ``cooler_b200`` never does.

Parity status: PINNED.  The reference's own balance tests are property tests
(tests/test_balance.py:13-132); this restatement is additionally checked
(tests/test_oracle.py) against golden vectors produced by running the
UNMODIFIED reference in the build container through ``oracle/refshim.py``
(generator: tests/golden/make_golden.py): bias bit-identical, masks, iteration
counts, scale and var identical, in all three modes (symmetric-upper and square
storage, integer and floating-point counts), on SURVEY.md 8c's own known answers,
and -- wherever a copy of the reference exists -- against the LIVE reference on
seeded random tables and option sets (tests/_fuzz_oracle_vs_reference.py, run in
a subprocess by the CPU suite: bit-identical).

The chunked map-reduce of the reference (parallel.py:53-320) is restated as a
loop over pixel spans with per-span ``np.bincount`` partials folded in span
order -- the fold order of the reference's default serial ``map``.
"""
from __future__ import annotations

import numpy as np


class ConvergenceWarning(UserWarning):
    pass


_SHARED = {}   # arrays inherited by fork()ed pool workers, see share()


def _spans(nnz, chunksize):
    # _balance.py:351-357
    if chunksize is None:
        return [(0, nnz)]
    edges = np.arange(0, nnz + chunksize, chunksize)
    return list(zip(edges[:-1], edges[1:]))


def _partition(start, stop, step):
    # util.py:18-33
    return [(i, min(i + step, stop)) for i in range(start, stop, step)]


def _span_marginal(px, chrom_ids, n_bins, lo, hi, *, binarize, zero_trans, n_diags,
                   zero_cis, vec):
    """One chunk through the reference's pipe: _init, filters, outer product,
    _marginalize (_balance.py:25-69)."""
    b1 = px["bin1_id"][lo:hi]
    b2 = px["bin2_id"][lo:hi]
    data = np.copy(px["count"][lo:hi])                       # _init (25-26)
    if binarize:                                             # _binarize (29-31)
        data[data != 0] = 1
    if zero_trans:                                           # _zero_trans (41-46)
        data[chrom_ids[b1] != chrom_ids[b2]] = 0
    if n_diags:                                              # _zero_diags (34-38)
        data[np.abs(b1 - b2) < n_diags] = 0
    if zero_cis:                                             # _zero_cis (49-54)
        data[chrom_ids[b1] == chrom_ids[b2]] = 0
    if vec is not None:                                      # _timesouterproduct (57-60)
        data = vec[b1] * vec[b2] * data
    # _marginalize (63-69)
    return np.bincount(b1, weights=data, minlength=n_bins) + np.bincount(
        b2, weights=data, minlength=n_bins)


def span_marginal_task(args):
    """Picklable wrapper so a process pool can map over spans (bench reference arm)."""
    px, chrom_ids, n_bins, lo, hi, kw = args
    return _span_marginal(px, chrom_ids, n_bins, lo, hi, **kw)


def _marginal(px, chrom_ids, n_bins, spans, map_, **kw):
    acc = np.zeros(n_bins)
    if _SHARED.get("px") is px:
        # the arrays were share()d with fork()ed pool workers: only the span travels, as in the
        # reference, whose workers re-open the file (parallel.py:289-304)
        for part in map_(shared_span_task, [(lo, hi, kw) for lo, hi in spans]):
            acc = acc + part
        return acc
    tasks = [(px, chrom_ids, n_bins, lo, hi, kw) for lo, hi in spans]
    for part in map_(span_marginal_task, tasks):            # reduce(add, zeros) parallel.py:253-273
        acc = acc + part
    return acc


def _iterate(bias_view, get_marg, tol, max_iters, on_var):
    """Inner loop shared by the three modes (_balance.py:87-124 / 154-194 / 222-260).
    ``bias_view`` is updated in place.  Returns (scale, var, n_passes, hit_limit)."""
    var = np.nan
    passes = 0
    hit_limit = True
    nzmarg = None
    for _ in range(max_iters):
        marg = get_marg()
        nzmarg = marg[marg != 0]
        if not len(nzmarg):
            bias_view[:] = np.nan
            var = 0.0
            hit_limit = False
            break
        marg = marg / nzmarg.mean()
        marg[marg == 0] = 1
        bias_view /= marg
        var = nzmarg.var()
        passes += 1              # == one "variance is" log line (_balance.py:109)
        if on_var is not None:
            on_var(var)
        if var < tol:
            hit_limit = False
            break
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            scale = nzmarg.mean() if nzmarg is not None else np.nan
    return scale, var, passes, hit_limit


def initial_bias(n_bins, marg_nnz, marg, chrom_offset, *, mad_max=5, min_nnz=10, min_count=0,
                 blacklist=None, x0=None):
    """Bias vector after the bin filters (_balance.py:366-413) given the two prefilter
    marginals: ``marg_nnz`` (binarised counts; only read when min_nnz > 0) and ``marg`` (raw
    counts; modified in place by the MAD-max normalisation exactly as in the reference)."""
    import warnings

    if x0 is not None:                                       # 368-372
        bias = x0
        bias[np.isnan(bias)] = 0
    else:
        bias = np.ones(n_bins, dtype=float)

    if min_nnz > 0:                                          # 375-384
        bias[marg_nnz < min_nnz] = 0

    if min_count:                                            # 396-397
        bias[marg < min_count] = 0

    if mad_max > 0:                                          # 400-409
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for lo, hi in zip(chrom_offset[:-1], chrom_offset[1:]):
                c_marg = marg[lo:hi]
                marg[lo:hi] /= np.median(c_marg[c_marg > 0])
            logNzMarg = np.log(marg[marg > 0])
            med = np.median(logNzMarg)
            dev = np.median(np.abs(logNzMarg - np.median(logNzMarg)))   # util.mad (util.py:513-514)
            cutoff = np.exp(med - mad_max * dev)
            bias[marg < cutoff] = 0

    if blacklist is not None:                                # 412-413
        bias[blacklist] = 0
    return bias


def ice_step(bias, marg):
    """One update of the genome-wide loop (_balance.py:97-108) given the marginal ``marg`` of the
    current bias: returns (var, mean) and divides ``bias`` in place; (0.0, nan) and bias = NaN when
    every marginal is zero."""
    nzmarg = marg[marg != 0]
    if not len(nzmarg):
        bias[:] = np.nan
        return 0.0, np.nan
    mean = nzmarg.mean()
    marg = marg / mean
    marg[marg == 0] = 1
    bias /= marg
    return nzmarg.var(), mean


def balance_arrays(px, chrom_offset, bin1_offset, *, cis_only=False, trans_only=False,
                   ignore_diags=2, mad_max=5, min_nnz=10, min_count=0, blacklist=None,
                   rescale_marginals=True, x0=None, tol=1e-5, max_iters=200,
                   chunksize=10_000_000, map=map, return_passes=False):
    """Restatement of ``balance_cooler`` (_balance.py:263-477) on raw arrays.

    px: dict with bin1_id, bin2_id (int64) and count; chrom_offset/bin1_offset
    are the cooler indexes.  Returns (bias, stats) or (bias, stats, passes, warned).
    """
    import warnings

    chrom_offset = np.asarray(chrom_offset)
    n_bins = int(chrom_offset[-1])
    nnz = len(px["bin1_id"])
    n_chroms = len(chrom_offset) - 1
    chrom_ids = np.repeat(np.arange(n_chroms, dtype=np.int32), np.diff(chrom_offset))
    spans = _spans(nnz, chunksize)
    n_diags = ignore_diags if ignore_diags else 0

    base = dict(zero_trans=cis_only, n_diags=n_diags)
    marg_nnz = None
    if min_nnz > 0:                                          # 375-384
        marg_nnz = _marginal(px, chrom_ids, n_bins, spans, map, binarize=True,
                             zero_cis=False, vec=None, **base)
    marg = _marginal(px, chrom_ids, n_bins, spans, map, binarize=False,   # 386-393
                     zero_cis=False, vec=None, **base)
    bias = initial_bias(n_bins, marg_nnz, marg, chrom_offset, mad_max=mad_max, min_nnz=min_nnz,
                        min_count=min_count, blacklist=blacklist, x0=x0)

    passes = []
    warned = False
    if cis_only:                                             # _balance_cisonly (127-196)
        scales = np.ones(n_chroms)
        variances = np.full_like(scales, np.nan)
        for cid, lo, hi in zip(range(n_chroms), chrom_offset[:-1], chrom_offset[1:]):
            cspans = _partition(int(bin1_offset[lo]), int(bin1_offset[hi]), chunksize or max(nnz, 1))

            def get_marg(cspans=cspans, lo=lo, hi=hi):
                return _marginal(px, chrom_ids, n_bins, cspans, map, binarize=False,
                                 zero_cis=False, vec=bias, **base)[lo:hi]

            scale, var, n, lim = _iterate(bias[lo:hi], get_marg, tol, max_iters, None)
            passes.append(n)
            warned |= lim
            b = bias[lo:hi]
            b[b == 0] = np.nan
            scales[cid] = scale
            variances[cid] = var
            if rescale_marginals:
                bias[lo:hi] /= np.sqrt(scale)
        scale, var = scales, variances
    else:
        if trans_only:                                       # 214-220
            cweights = 1.0 / np.concatenate(
                [[(1 - (hi - lo) / n_bins)] * (hi - lo)
                 for lo, hi in zip(chrom_offset[:-1], chrom_offset[1:])])

            def get_marg():
                return _marginal(px, chrom_ids, n_bins, spans, map, binarize=False,
                                 zero_cis=True, vec=bias * cweights, **base)
        else:                                                # genome-wide (72-124)
            def get_marg():
                return _marginal(px, chrom_ids, n_bins, spans, map, binarize=False,
                                 zero_cis=False, vec=bias, **base)

        scale, var, n, warned = _iterate(bias, get_marg, tol, max_iters, None)
        passes.append(n)
        bias[bias == 0] = np.nan
        if rescale_marginals:
            bias /= np.sqrt(scale)

    if warned:
        warnings.warn("Iteration limit reached without convergence.",
                      ConvergenceWarning, stacklevel=2)

    stats = {"tol": tol, "min_nnz": min_nnz, "min_count": min_count, "mad_max": mad_max,
             "cis_only": cis_only, "ignore_diags": ignore_diags, "scale": scale,
             "converged": var < tol, "var": var, "divisive_weights": False}
    if return_passes:
        return bias, stats, passes, warned
    return bias, stats


# ---------------------------------------------------------------------------
# Process-pool form of one marginal pass (the reference's `cooler balance -p N`
# path: cli/balance.py:237-243 -> Pool.imap_unordered over pixel spans).  The
# pixel arrays are inherited by fork()ed workers through this module global --
# the stand-in for every worker re-opening the file (parallel.py:289-304) --
# while the per-task bias vector and the per-span n_bins result vector travel
# by pickle exactly as in the reference.
# ---------------------------------------------------------------------------
def share(px, chrom_ids, n_bins):
    _SHARED.clear()
    _SHARED.update(px=px, chrom_ids=chrom_ids, n_bins=n_bins)


def shared_span_task(args):
    lo, hi, kw = args
    return _span_marginal(_SHARED["px"], _SHARED["chrom_ids"], _SHARED["n_bins"], lo, hi, **kw)


def pool_marginal(map_, spans, **kw):
    """reduce(add, zeros) over spans mapped with ``map_`` (e.g. Pool.imap_unordered)."""
    acc = np.zeros(_SHARED["n_bins"])
    for part in map_(shared_span_task, [(lo, hi, kw) for lo, hi in spans]):
        acc = acc + part
    return acc
