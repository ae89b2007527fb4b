"""Time cb200_ice_sweep variants on a resident synthetic table (CUDA events).
    python tools/tune_sweep.py [binsize] [nnz_target]
"""
import ctypes as C
import os
import sys
import time

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
# variant = unroll:strip_bins:mode[:far_col_shift:far_warps:far_unroll]; mode 0 = gather through L1/L2,
# 1 = shared-memory strips, 2 = band strips + far remainder through L2, 3 = band strips + far-field blocks
for var in variants:
    unroll, strip_bins, use_smem = var[:3]
    far_cs, far_w, far_u = (list(var[3:]) + [12, 16, 0])[:3] if len(var) > 3 else (12, 16, 0)
    if (strip_bins, use_smem, far_cs, far_w) != last_strip:
        prob = state = None
        torch.cuda.empty_cache()
        t0 = time.time()
        band = {2: True, 3: "blocks", 4: "allblocks"}.get(use_smem, False) if strip_bins else None
        prob = IceProblem(tab, strip_bins=strip_bins, use_smem=bool(use_smem), band=band,
                          far_col_shift=far_cs,
                          far_compact={"0": False, "1": True}.get(os.environ.get("TUNE_COMPACT", ""), None))
        bias0 = prefilter_bias(prob, tab.chrom_offset, mad_max=5, min_nnz=10, min_count=0, blacklist=None, x0=None)
        state = IceState(prob, bias0, None, np.array([0, tab.n_bins]), 1e-5, 200)
        torch.cuda.synchronize()
        print(f"  build {time.time() - t0:.2f} s  far_px={sum(f.nnz for f in prob.far_copies)} "
              f"far_units={state.n_far_units} near_px={sum(c.nnz for c in prob.subs)} "
              f"padded={sum(f.n_chunks * 64 for f in prob.far_copies)} rec_bytes={[f.rec_bytes for f in prob.far_copies]} "
              f"records={sum(f.nnz for f in prob.far_copies)}", flush=True)
        last_strip = (strip_bins, use_smem, far_cs, far_w)
    assert lib.cb200_set_tuning(0, -1 if unroll < 10 else (unroll // 10) - 1) == 0   # 1x: selective off, 2x: on
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
    split = ""
    if state.n_far_units:                      # time the near (strips) and far (blocks) launches separately
        nf, nw = state.st.n_far_units, state.st.total_warps
        parts = []
        for far_on in (0, 1):
            state.st.n_far_units = nf if far_on else 0
            state.st.total_warps = 0 if far_on else nw
            e0.record()
            for _ in range(n):
                _lib.call("cb200_ice_sweep", C.byref(state.st), _stream())
            e1.record()
            torch.cuda.synchronize()
            parts.append(e0.elapsed_time(e1) / n)
        state.st.n_far_units, state.st.total_warps = nf, nw
        _lib.call("cb200_ice_sweep", C.byref(state.st), _stream())
        far_px = sum(f.nnz for f in prob.far_copies)
        near_px = sum(c.nnz for c in prob.subs)
        split = (f" | near {parts[0]*1e3:.0f} us ({8*near_px/parts[0]/1e6/peak:.2f}) "
                 f"far {parts[1]*1e3:.0f} us ({8*far_px/parts[1]/1e6/peak:.2f})")
    e0.record()
    for _ in range(n):
        _lib.call("cb200_ice_combine", st, _stream())
    e1.record()
    torch.cuda.synchronize()
    ms_c = e0.elapsed_time(e1) / n
    upd = []
    for fused in (0, 1):                           # the two per-bin kernels (K4): s given / s from the sweep outputs
        e0.record()
        for k in range(n):
            _lib.call("cb200_ice_update", st, k, fused, _stream())
        e1.record()
        torch.cuda.synchronize()
        upd.append(e0.elapsed_time(e1) / n)
    state.reset(bias0)
    _lib.call("cb200_ice_sweep", st, _stream())
    _lib.call("cb200_ice_combine", st, _stream())
    print(f"    per-bin kernels (marg + update, {state.st.n_blocks} blocks): s given {upd[0]*1e3:.1f} us, fused combine {upd[1]*1e3:.1f} us", flush=True)
    s = state.t["s"].cpu().numpy().copy()
    if ref is None:
        ref = s
    err = float(np.max(np.abs(s - ref) / np.maximum(np.abs(ref), 1e-300)))
    gbs = prob.bytes_per_pass() / (ms * 1e-3) / 1e9
    gbs2 = prob.bytes_per_pass() / ((ms + ms_c) * 1e-3) / 1e9
    print(f"unroll={unroll} smem={use_smem} far={far_cs}/{far_w}/{far_u} strip_bins={strip_bins:6d} subs={len(prob.subs):3d} span={state.st.span}  sweep {ms*1e3:8.1f} us "
          f"{gbs:7.1f} GB/s frac={gbs/peak:.3f} | +combine {ms_c*1e3:6.1f} us frac={gbs2/peak:.3f}  max_rel_diff_vs_first={err:.1e}{split}", flush=True)
