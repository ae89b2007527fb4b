"""Stage timings of the end-to-end path (pinned host arrays -> bias on host)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cooler_b200  # noqa: E402
from cooler_b200 import _lib, synth, table as T  # noqa: E402
from cooler_b200.balance import IceProblem, IceState, prefilter_bias  # noqa: E402

binsize = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
target = int(float(sys.argv[2])) if len(sys.argv) > 2 else 200_000_000
plan = synth.SynthPlan(binsize, target, seed=0)
data = synth.generate(plan)
print("n_bins", plan.n_bins, "nnz", data["nnz"], flush=True)


class Timer:
    def __init__(self):
        self.t = {}

    def __call__(self, name):
        torch.cuda.synchronize()
        now = time.perf_counter()
        if hasattr(self, "_last"):
            self.t[self._name] = self.t.get(self._name, 0.0) + now - self._last
        self._last, self._name = now, name


for rep in range(3):
    tm = Timer()
    tm("h2d")
    dev = torch.device("cuda")
    indptr = data["bin1_offset"].to(dev, non_blocking=True)
    bin2 = data["bin2_id"].to(dev, non_blocking=True)
    cnt = data["count"].to(dev, non_blocking=True)
    tm("pack+tile_rows")
    px = T._new_px(data["nnz"], dev)
    _lib.call("cb200_pack_pixels", T._ptr(bin2), T._ptr(cnt), data["nnz"], T._ptr(px), T._stream())
    raw = T.TableCopy(px, indptr, data["nnz"], plan.n_bins)
    tab = cooler_b200.PixelTable(raw, data["chrom_offset"])
    tm("filter")
    csr, tkey, tval = T.filter_copy(tab.raw, tab.row_lo, tab.row_hi, 2, 0, want_transpose=True)
    tm("sort(transpose)")
    csc, ckeys = T.transpose_from(tkey, tval, plan.n_bins, return_keys=True)
    tm("marginals")
    a = T.marginals_i64(csr)
    b = T.marginals_i64(csc)
    m = torch.stack([a[0] + b[0], a[1] + b[1]]).cpu()
    tm("problem(total,incl strips)")
    del csr, csc, tkey, tval, ckeys, a, b
    prob = IceProblem(tab)
    tm("prefilter(host)")
    bias0 = prefilter_bias(prob, tab.chrom_offset, mad_max=5, min_nnz=10, min_count=0, blacklist=None, x0=None)
    tm("state")
    state = IceState(prob, bias0, None, np.array([0, tab.n_bins]), 1e-5, 200)
    tm("loop")
    state.run()
    tm("results")
    res = state.results()
    tm("end")
    if rep:
        print({k: round(v * 1e3, 2) for k, v in tm.t.items()}, "passes", int(res["seg_iters"][0]), flush=True)
    del prob, state, tab, raw, px
    torch.cuda.empty_cache()
