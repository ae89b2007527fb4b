"""Sweep time of the floating-point-count variant of K3 next to the integer one, same synthetic table
(values = count x lognormal noise):  python tools/bench_float.py [binsize] [nnz]   (default: C2, 100 kb, 2e8)"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cooler_b200  # noqa: E402
from cooler_b200 import _lib, synth  # noqa: E402
from cooler_b200.balance import IceProblem, IceState, prefilter_bias  # noqa: E402
from cooler_b200.table import _stream  # noqa: E402

binsize = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
target = int(float(sys.argv[2])) if len(sys.argv) > 2 else 200_000_000
plan = synth.SynthPlan(binsize, target, seed=0)
data = synth.generate(plan)
peak = 6536.7
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        peak = float(json.load(f)["hbm_gbs"])
rng = np.random.default_rng(0)
fval = data["count"].numpy().astype(np.float64) * rng.lognormal(0.0, 0.3, data["nnz"])
out = {"binsize": binsize, "n_bins": plan.n_bins, "nnz": data["nnz"], "peak_GBps": peak}
for kind, count in (("int32", data["count"]), ("float64", fval)):
    tab = cooler_b200.PixelTable.from_arrays(data["bin1_offset"], data["bin2_id"], count, data["chrom_offset"],
                                             float_values=True)
    prob = IceProblem(tab)
    bias0 = prefilter_bias(prob, tab.chrom_offset, mad_max=5, min_nnz=10, min_count=0, blacklist=None, x0=None)
    state = IceState(prob, bias0, None, np.array([0, tab.n_bins]), 1e-5, 200)
    st = C.byref(state.st)
    for _ in range(3):
        _lib.call("cb200_ice_sweep", st, _stream())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        _lib.call("cb200_ice_sweep", st, _stream())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out[kind] = {"sweep_ms": ms, "layout": "strips" if prob.strip_shift is not None else "whole copies",
                 "algorithmic_bytes": prob.bytes_per_pass(), "GBps": prob.bytes_per_pass() / ms / 1e6,
                 "frac_of_measured_hbm": prob.bytes_per_pass() / ms / 1e6 / peak}
    print(kind, out[kind], flush=True)
    del state, prob, tab
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "bench_float.json"), "w") as f:
    json.dump(out, f, indent=1)
