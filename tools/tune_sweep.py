"""Time cb200_ice_sweep variants on a resident synthetic table (CUDA events).
    python tools/tune_sweep.py [binsize] [nnz_target]
"""
import ctypes as C
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
world = int(sys.argv[3]) if len(sys.argv) > 3 else 1
rank = int(sys.argv[4]) if len(sys.argv) > 4 else 0
variants = [(2, 0, 0), (2, 16384, 1), (2, 16384, 2), (2, 8192, 2), (2, 32768, 2)]   # (unroll, strip_bins, use_smem)
if os.environ.get('TUNE_VARIANTS'):
    variants = [tuple(int(x) for x in v.split(':')) for v in os.environ['TUNE_VARIANTS'].split(',')]
plan = synth.SynthPlan(binsize, target, seed=0)
data = synth.generate(plan, rows=plan.shard_rows(rank, world))
print("n_bins", plan.n_bins, "nnz(shard)", data["nnz"], "rows", data["row_range"], flush=True)
tab = cooler_b200.PixelTable.from_arrays(data["bin1_offset"], data["bin2_id"], data["count"], data["chrom_offset"])
del data
peak = 6536.7
ref = None
lib = _lib.load()
last_strip = None
for unroll, strip_bins, use_smem in variants:
    if (strip_bins, use_smem) != last_strip:
        prob = state = None
        torch.cuda.empty_cache()
        prob = IceProblem(tab, strip_bins=strip_bins, use_smem=bool(use_smem), band=(use_smem == 2) if strip_bins else None)
        bias0 = prefilter_bias(prob, tab.chrom_offset, mad_max=5, min_nnz=10, min_count=0, blacklist=None, x0=None)
        state = IceState(prob, bias0, None, np.array([0, tab.n_bins]), 1e-5, 200)
        last_strip = (strip_bins, use_smem)
    assert lib.cb200_set_tuning(unroll % 10, -1 if unroll < 10 else (unroll // 10) - 1) == 0   # 1x: selective off, 2x: on
    state.reset(bias0)
    st = C.byref(state.st)
    for _ in range(3):
        _lib.call("cb200_ice_sweep", st, _stream())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        _lib.call("cb200_ice_sweep", st, _stream())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    e0.record()
    for _ in range(n):
        _lib.call("cb200_ice_combine", st, _stream())
    e1.record()
    torch.cuda.synchronize()
    ms_c = e0.elapsed_time(e1) / n
    s = state.t["s"].cpu().numpy().copy()
    if ref is None:
        ref = s
    err = float(np.max(np.abs(s - ref) / np.maximum(np.abs(ref), 1e-300)))
    gbs = prob.bytes_per_pass() / (ms * 1e-3) / 1e9
    gbs2 = prob.bytes_per_pass() / ((ms + ms_c) * 1e-3) / 1e9
    print(f"unroll={unroll} smem={use_smem} strip_bins={strip_bins:6d} subs={len(prob.subs):3d} span={state.st.span}  sweep {ms*1e3:8.1f} us "
          f"{gbs:7.1f} GB/s frac={gbs/peak:.3f} | +combine {ms_c*1e3:6.1f} us frac={gbs2/peak:.3f}  max_rel_diff_vs_first={err:.1e}", flush=True)
