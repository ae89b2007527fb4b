"""Coarsening throughput on a resident synthetic table + zoomify chain, with the CPU oracle
(numpy restatement of CoolerCoarsener._aggregate) timed on a bounded sample beside it.

    python tools/bench_coarsen.py [binsize] [nnz_target] [factor ...]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cooler_b200  # noqa: E402
from cooler_b200 import synth  # noqa: E402
from cooler_b200.reduce import coarsen_table, table_to_host  # noqa: E402

binsize = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
target = int(float(sys.argv[2])) if len(sys.argv) > 2 else 200_000_000
factors = [int(x) for x in sys.argv[3:]] or [2, 5, 2, 5, 2]
plan = synth.SynthPlan(binsize, target, seed=0)
data = synth.generate(plan)
tab = cooler_b200.PixelTable.from_arrays(data["bin1_offset"], data["bin2_id"], data["count"], data["chrom_offset"])
peak = 6536.7
out = {"binsize": binsize, "n_bins": plan.n_bins, "nnz": data["nnz"], "levels": []}
cur = tab
res = binsize
for f in factors:
    for _ in range(4):                       # 3 warm-ups (allocator) + timed
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        new, ids = coarsen_table(cur, f)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    algo = 12 * (cur.nnz + new.nnz)
    lvl = {"from": res, "factor": f, "nnz_in": cur.nnz, "nnz_out": new.nnz, "ms": ms,
           "input_pixels_per_s": cur.nnz / (ms * 1e-3), "algorithmic_GBps": algo / (ms * 1e-3) / 1e9,
           "frac_of_measured_hbm": algo / (ms * 1e-3) / 1e9 / peak}
    out["levels"].append(lvl)
    print(lvl, flush=True)
    cur, res = new, res * f

# CPU oracle on a bounded sample: first 2e7 pixels of the base table, factor = factors[0]
from oracle import coarsen_numpy  # noqa: E402
n = plan.n_bins
m = min(20_000_000, data["nnz"])
off = data["bin1_offset"].numpy()
b1 = np.repeat(np.arange(n, dtype=np.int64), np.diff(off))[:m]
px = dict(bin1_id=b1, bin2_id=data["bin2_id"].numpy()[:m], count=data["count"].numpy()[:m])
co = plan.chrom_offset
bc = np.repeat(np.arange(len(co) - 1, dtype=np.int32), np.diff(co))
bs = np.concatenate([np.arange(k, dtype=np.int64) * binsize for k in np.diff(co)])
be = np.minimum(bs + binsize, np.repeat(plan.chromsizes, np.diff(co)))
t0 = time.perf_counter()
w = coarsen_numpy.coarsen_pixels(px, bc, bs, be, plan.chromsizes, factors[0])
dt = time.perf_counter() - t0
out["cpu_oracle"] = {"pixels": m, "seconds": dt, "input_pixels_per_s": m / dt, "cores": 1, "kind": "port"}
print(out["cpu_oracle"], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "bench_coarsen.json"), "w") as fjson:
    json.dump(out, fjson, indent=1)
