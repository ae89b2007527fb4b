"""Run in a SUBPROCESS by tests/test_host_logic.py: the HOST glue of the balancing path -- balance.balance_table:
bin filters, mode dispatch (segments, trans weights), per-chromosome logging, the empty-segment and
iteration-limit branches, NaN conversion, rescaling, the stats dict, warnings, warm-start aliasing and the
square-storage cascade -- against the LIVE unmodified reference's balance_cooler on seeded random tables.
The device is replaced by a numpy restatement of what the kernels compute (csrc/ice.cu ice_marg_kernel /
ice_update_kernel: all segments iterated concurrently, a segment stops when its variance drops below tol or its
marginals are all zero), so everything ABOVE the C ABI runs exactly as it does on a GPU."""
import logging
import os
import sys
import warnings
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

from tests._fuzz_oracle_vs_reference import random_kwargs, random_table  # noqa: E402


class FakeProblem:
    """What IceProblem offers to balance_table, computed with numpy."""

    def __init__(self, t, chrom_offset, cis_only, trans_only, ignore_diags):
        n = len(t["bc"])
        self.table = SimpleNamespace(n_bins=n, chrom_offset=chrom_offset, device=None,
                                     raw=SimpleNamespace(nnz=len(t["b1"])))
        self.group, self.earlier, self.subs, self.strip_shift, self.band = None, [], [], None, False
        b1, b2, x = t["b1"], t["b2"], t["count"].astype(np.float64)
        chrom = np.repeat(np.arange(len(chrom_offset) - 1), np.diff(chrom_offset))
        keep = np.ones(len(b1), bool)
        nd = int(ignore_diags) if ignore_diags else 0
        if nd > 0:
            keep &= np.abs(b1 - b2) >= nd
        cis = chrom[b1] == chrom[b2]
        if cis_only:
            keep &= cis
        self.marg_nnz = (np.bincount(b1[keep], x[keep] != 0, n) + np.bincount(b2[keep], x[keep] != 0, n)).astype(float)
        self.marg_sum = (np.bincount(b1[keep], x[keep], n) + np.bincount(b2[keep], x[keep], n)).astype(float)
        if trans_only:
            keep = keep & ~cis                       # _zero_cis only inside the iteration (_balance.py:227)
        self.b1, self.b2, self.x = b1[keep], b2[keep], x[keep]
        self.n = n

    def sweep(self, w):
        return (np.bincount(self.b1, w[self.b2] * self.x, self.n) + np.bincount(self.b2, w[self.b1] * self.x, self.n))


class FakeState:
    """numpy restatement of the device loop (cb200_ice_sweep + cb200_ice_update)."""

    def __init__(self, prob, bias0, cweights, seg_offset, tol, max_iters):
        self.prob, self.tol, self.max_iters = prob, float(tol), int(max_iters)
        self.bias = np.array(bias0, dtype=np.float64)
        self.cw = None if cweights is None else np.asarray(cweights, dtype=np.float64)
        self.seg = np.asarray(seg_offset, dtype=np.int64)
        S = len(self.seg) - 1
        self.S = S
        self.scale, self.var = np.ones(S), np.full(S, np.nan)
        self.iters, self.done, self.empty = np.zeros(S, np.int32), np.zeros(S, np.int32), np.zeros(S, np.int32)
        self.var_hist = np.full((self.max_iters, S), np.nan)
        self.launches, self.n_far_units = 0, 0

    def run(self, check_every=8):
        it = 0
        with np.errstate(all="ignore"):
            while it < self.max_iters and not self.done.all():
                w = self.bias if self.cw is None else self.bias * self.cw
                marg = w * self.prob.sweep(w)
                for s in range(self.S):
                    if self.done[s]:
                        continue
                    lo, hi = self.seg[s], self.seg[s + 1]
                    m = marg[lo:hi]
                    nz = m[m != 0]
                    if len(nz) == 0:
                        self.scale[s], self.var[s], self.empty[s], self.done[s] = np.nan, 0.0, 1, 1
                        continue
                    mean = nz.sum() / len(nz)
                    q = m / mean
                    q[q == 0] = 1.0
                    self.bias[lo:hi] /= q
                    var = ((nz - mean) ** 2).sum() / len(nz)
                    self.scale[s], self.var[s] = mean, var
                    self.iters[s] += 1
                    self.var_hist[it, s] = var
                    if var < self.tol:
                        self.done[s] = 1
                it += 1
        return it

    def results(self):
        return {"seg_scale": self.scale, "seg_var": self.var, "seg_iters": self.iters, "seg_done": self.done,
                "seg_empty": self.empty, "bias": self.bias, "var_hist": self.var_hist}


class Count(logging.Handler):
    def __init__(self):
        super().__init__(level=logging.INFO)
        self.n = 0

    def emit(self, record):
        self.n += record.getMessage().startswith("variance is")


def close(a, b, rtol=1e-9, atol=0.0):
    a, b = np.atleast_1d(np.asarray(a, dtype=float)), np.atleast_1d(np.asarray(b, dtype=float))
    return a.shape == b.shape and bool(np.all(np.isnan(a) == np.isnan(b))) and \
        bool(np.allclose(a[~np.isnan(a)], b[~np.isnan(b)], rtol=rtol, atol=atol))


def main(n_cases, seed):
    assert refshim.available()
    import cooler_b200.balance as B
    B._use_device = lambda device: device
    B.IceState = FakeState
    current = {}

    def numpy_touch(prob, bias, chrom_offset, group=None):
        t = current["t"]
        chrom = np.repeat(np.arange(len(chrom_offset) - 1), np.diff(chrom_offset))
        c1, c2 = chrom[t["b1"]], chrom[t["b2"]]
        early = c2 < c1
        if not early.any():
            return None
        nc = len(chrom_offset) - 1
        any_, masked = np.zeros((nc, nc), bool), np.zeros((nc, nc), bool)
        any_[c1[early], c2[early]] = True
        m = early & ((bias[t["b2"]] == 0) | np.isnan(bias[t["b2"]]))
        masked[c1[m], c2[m]] = True
        return any_, masked

    B.cis_touch_flags = numpy_touch
    rng = np.random.default_rng(seed)
    n_runs = n_skip = 0
    lg = logging.getLogger("cooler")
    for case in range(n_cases):
        t = random_table(rng)
        current["t"] = t
        n = len(t["bc"])
        name = f"fuzzglue{case}.cool"
        refshim.register(name, t["names"], t["sizes"], t["bc"], t["bs"], t["be"], t["b1"], t["b2"], t["count"],
                         t["binsize"], symmetric_upper=not t["square"])
        chrom_offset = np.searchsorted(t["bc"], np.arange(len(t["names"]) + 1)).astype(np.int64)
        for _ in range(2):
            kw = random_kwargs(rng, n)
            kw_ref = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
            kw_our = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                np.seterr(all="ignore")
                want, wstats, wpasses, wwarn = refshim.ref_balance(name, **kw_ref)
            prob = FakeProblem(t, chrom_offset, kw.get("cis_only", False), kw.get("trans_only", False),
                               kw.get("ignore_diags", 2))
            h = Count()
            lg.addHandler(h)
            old = lg.level
            lg.setLevel(logging.INFO)
            try:
                with warnings.catch_warnings(record=True) as rec:
                    warnings.simplefilter("always")
                    got, gstats = B.balance_table(prob.table, problem=prob, chromnames=t["names"], **kw_our)
            finally:
                lg.removeHandler(h)
                lg.setLevel(old)
            gwarn = any(issubclass(x.category, B.ConvergenceWarning) for x in rec)
            tag = (case, {k: v for k, v in kw.items() if k != "x0"}, t["square"], str(t["count"].dtype))
            # a variance within rounding of tol can stop one side a pass earlier: such draws are not comparable
            if h.n != sum(wpasses):
                v = np.atleast_1d(np.asarray(wstats["var"], dtype=float))
                near = np.any(np.abs(v / kw.get("tol", 1e-5) - 1) < 1e-6)
                assert near, ("passes", tag, h.n, wpasses)
                n_skip += 1
                continue
            assert gwarn == wwarn, ("warned", tag)
            assert sorted(gstats) == sorted(wstats), (sorted(gstats), sorted(wstats))
            for k in ("tol", "min_nnz", "min_count", "mad_max", "cis_only", "ignore_diags", "divisive_weights"):
                assert gstats[k] == wstats[k] or (gstats[k] is wstats[k]), (k, tag)
            assert close(got, want), ("bias", tag, np.nanmax(np.abs(np.asarray(got) / np.asarray(want) - 1)))
            assert close(gstats["scale"], wstats["scale"]) and close(gstats["var"], wstats["var"], 1e-6, 1e-20), ("stats", tag,
                                                                                                           gstats, wstats)
            assert np.array_equal(np.atleast_1d(gstats["converged"]), np.atleast_1d(wstats["converged"])), tag
            if "x0" in kw:
                assert got is kw_our["x0"], "x0 must be mutated in place and returned"
            n_runs += 1
    print(f"ok: {n_runs} balance_table runs agree with the live reference ({n_skip} draws at the tolerance edge skipped)")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
