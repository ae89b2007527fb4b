"""BASELINE configs[4] (C5): zoomify a synthetic hg38 1 kb matrix to the 4DN resolution list
(2 kb ... 10 Mb) with a default ICE balance per level, everything chained on the device
(level k+1 is coarsened from the resident level it derives from, src/cooler/_reduce.py:752-877;
per-level balance as cli/zoomify.py:20-46 invokes it).

    python tools/bench_zoomify.py [nnz_target] [out.json]

Reported per level: coarsening time (CUDA events) against 12 B x (nnz_in + nnz_out) of algorithmic
traffic, balance time (wall: build + prefilter + iteration loop), passes, convergence; sum(count)
is checked to be preserved by every level.  The reference's CoolerCoarsener (pandas groupby,
unmodified, via oracle/refshim.py) is timed beside it on a bounded sample of the base table.
"""
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cooler_b200  # noqa: E402
from cooler_b200 import synth  # noqa: E402
from cooler_b200.reduce import ZOOMS_4DN, coarsen_table, get_multiplier_sequence  # noqa: E402

target = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "bench_zoomify.json")
peak = 6536.7
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        peak = float(json.load(f)["hbm_gbs"])
plan = synth.SynthPlan(1000, target, seed=0)
t0 = time.perf_counter()
data = synth.generate(plan)
gen_s = time.perf_counter() - t0
base = cooler_b200.PixelTable.from_arrays(data["bin1_offset"], data["bin2_id"], data["count"], data["chrom_offset"])
resn, pred, mult = get_multiplier_sequence(ZOOMS_4DN)
total_count = int(data["count"].sum(dtype=torch.int64))
out = {"config": "BASELINE configs[4] C5: zoomify synthetic hg38 1 kb -> 2 kb ... 10 Mb + default ICE per level, one B200",
       "n_bins_base": plan.n_bins, "nnz_base": data["nnz"], "sum_count": total_count, "levels": [],
       "generate_s": gen_s, "peak_GBps": peak}
tables = {0: base}


def balance(tab):
    torch.cuda.synchronize()
    t = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bias, stats, info = cooler_b200.balance_table(tab, return_info=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t, stats, info


t_all = time.perf_counter()
coarsen_ms_total, balance_s_total = 0.0, 0.0
for i in range(len(resn)):
    if pred[i] == -1:
        tab, ms, nnz_in = base, 0.0, base.nnz
    else:
        src = tables[int(pred[i])]
        coarsen_table(src, int(mult[i]))           # warm-up (allocator)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        tab, _ = coarsen_table(src, int(mult[i]))
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        ms, nnz_in = e0.elapsed_time(e1), src.nnz
        s = int(tab.raw.px[: tab.nnz, 1].sum(dtype=torch.int64))
        assert s == total_count, (int(resn[i]), s, total_count)
    tables[i] = tab
    bal_s, stats, info = balance(tab)
    algo = 12 * (nnz_in + tab.nnz)
    lvl = {"resolution": int(resn[i]), "from": int(resn[pred[i]]) if pred[i] != -1 else None,
           "factor": int(mult[i]) if pred[i] != -1 else None, "n_bins": tab.n_bins, "nnz_in": nnz_in, "nnz_out": tab.nnz,
           "coarsen_ms": ms, "coarsen_frac_of_hbm": (algo / (ms * 1e-3) / 1e9 / peak) if ms else None,
           "coarsen_wall_ms": (wall * 1e3) if pred[i] != -1 else None,
           "balance_s": bal_s, "passes": int(max(info["passes"])), "converged": bool(np.all(stats["converged"])),
           "layout": str(info["band"]) if info["band"] else ("strips" if info["strip_shift"] else "whole")}
    coarsen_ms_total += ms
    balance_s_total += bal_s
    out["levels"].append(lvl)
    print(lvl, flush=True)
    # free levels nobody derives from any more
    for k in list(tables):
        if k != i and not any(int(pred[j]) == k for j in range(i + 1, len(resn))):
            del tables[k]
    torch.cuda.empty_cache()
torch.cuda.synchronize()
out["total_s"] = time.perf_counter() - t_all
out["coarsen_ms_total"] = coarsen_ms_total
out["balance_s_total"] = balance_s_total
out["coarsen_input_pixels_per_s"] = sum(l["nnz_in"] for l in out["levels"] if l["factor"]) / (coarsen_ms_total * 1e-3)

# the reference's own coarsener on a bounded sample: the first ~2e7 pixels of the base table, x2
try:
    from oracle import refshim
    if refshim.available():
        n = plan.n_bins
        off = data["bin1_offset"].numpy()
        m = int(off[np.searchsorted(off, 20_000_000)])
        b1 = np.repeat(np.arange(n, dtype=np.int64), np.diff(off))[:m]
        nb = np.diff(plan.chrom_offset)
        start = np.concatenate([np.arange(k, dtype=np.int64) * 1000 for k in nb])
        refshim.register("c5.cool", synth.HG38_NAMES, plan.chromsizes,
                         np.repeat(np.arange(len(nb), dtype=np.int32), nb), start,
                         np.minimum(start + 1000, np.repeat(plan.chromsizes, nb)), b1,
                         data["bin2_id"].numpy()[:m].astype(np.int64), data["count"].numpy()[:m], 1000)
        t0 = time.perf_counter()
        refshim.ref_coarsen("c5.cool", 2)
        dt = time.perf_counter() - t0
        out["cpu_reference_coarsen"] = {"kind": "reference", "code": "unmodified cooler._reduce.CoolerCoarsener (pandas groupby), 1 process",
                                        "pixels": m, "seconds": dt, "input_pixels_per_s": m / dt, "cores": 1}
        print(out["cpu_reference_coarsen"], flush=True)
except Exception as e:  # noqa: BLE001 - the CPU figure is optional
    out["cpu_reference_coarsen"] = {"error": repr(e)}
os.makedirs(os.path.dirname(out_path), exist_ok=True)
with open(out_path, "w") as fjson:
    json.dump(out, fjson, indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "levels"}))
