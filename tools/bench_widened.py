"""Throughput of the widened rows on a resident synthetic table, CUDA events, with the
CPU oracle timed on a bounded sample beside each:
  * balanced matrix fetch (dense windows of a 100 kb hg38 matrix),
  * pairs binning (cload core) on synthetic read pairs.

    python tools/bench_widened.py [nnz_target] [n_pairs]
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
from cooler_b200 import ingest, synth  # noqa: E402
from cooler_b200.api import MatrixSelector  # noqa: E402

target = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000
n_pairs = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000_000
peak = 6536.7
out = {}

# ---- matrix fetch -----------------------------------------------------------------------------
plan = synth.SynthPlan(100_000, target, seed=0)
data = synth.generate(plan)
tab = cooler_b200.PixelTable.from_arrays(data["bin1_offset"], data["bin2_id"], data["count"], data["chrom_offset"])
n = plan.n_bins
w = np.random.default_rng(0).lognormal(0, 0.3, n)
sel = MatrixSelector(None, tab, w)
co = plan.chrom_offset
windows = {"chr1 x chr1": (0, int(co[1]), 0, int(co[1])), "chr1 x chr2": (0, int(co[1]), int(co[1]), int(co[2])),
           "chr2 x chr1 (below the diagonal)": (int(co[1]), int(co[2]), 0, int(co[1])),
           "8192 x 8192 at the diagonal": (10000, 18192, 10000, 18192)}
off = data["bin1_offset"].numpy()
b2 = data["bin2_id"].numpy()
res = []
for name, (i0, i1, j0, j1) in windows.items():
    for _ in range(3):
        sel[i0:i1, j0:j1]
    torch.cuda.synchronize()
    # device-side time: the two kernels (fetch + balance), result left on the device
    import ctypes as C
    from cooler_b200 import _lib
    from cooler_b200.table import _ptr, _stream
    cs = tab.raw.c_struct()
    h, wd = i1 - i0, j1 - j0
    dense = torch.empty(h * wd, dtype=torch.int32, device="cuda")
    outd = torch.empty(h * wd, dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        _lib.call("cb200_fetch_dense", C.byref(cs), n, i0, i1, j0, j1, 1, _ptr(dense), _stream())
        _lib.call("cb200_fetch_balance_dense", _ptr(dense), h, wd, _ptr(sel.weights[i0:i1]), _ptr(sel.weights[j0:j1]),
                  0, _ptr(outd), _stream())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    npx = int((dense != 0).sum().item())
    algo = 8 * npx + (4 + 4 + 8) * h * wd                # pixels read + window written (int32), read, written (f64)
    t0 = time.perf_counter()
    arr = sel[i0:i1, j0:j1]
    t_api = time.perf_counter() - t0
    res.append({"window": name, "shape": [h, wd], "pixels": npx, "kernel_ms": ms, "algorithmic_GBps": algo / ms / 1e6,
                "frac_of_measured_hbm": algo / ms / 1e6 / peak, "api_call_s_incl_d2h": t_api})
    print(res[-1], flush=True)
# CPU oracle on one window
from oracle import matrix_numpy  # noqa: E402
i0, i1, j0, j1 = windows["chr1 x chr1"]
p1 = int(off[i1])
b1 = np.repeat(np.arange(i1, dtype=np.int64), np.diff(off[: i1 + 1]))
t0 = time.perf_counter()
ref = matrix_numpy.fetch_dense(b1, b2[:p1], data["count"].numpy()[:p1], i0, i1, j0, j1, weights=w)
dt = time.perf_counter() - t0
assert np.array_equal(ref, sel[i0:i1, j0:j1], equal_nan=True)
out["matrix_fetch"] = {"n_bins": n, "nnz": data["nnz"], "windows": res,
                       "cpu_oracle": {"window": "chr1 x chr1", "seconds": dt, "cores": 1, "kind": "port",
                                      "note": "numpy restatement over the rows of the window only"}}
del tab, sel, data
torch.cuda.empty_cache()

# ---- pairs binning -----------------------------------------------------------------------------
rng = np.random.default_rng(1)
sizes = np.asarray(synth.HG38, dtype=np.int64)
prob = sizes / sizes.sum()
c1 = rng.choice(len(sizes), n_pairs, p=prob).astype(np.int32)
cis = rng.random(n_pairs) < 0.7
c2 = np.where(cis, c1, rng.choice(len(sizes), n_pairs, p=prob)).astype(np.int32)
p1 = (rng.random(n_pairs) * sizes[c1]).astype(np.int64) + 1
d = (np.exp(rng.random(n_pairs) * np.log(1e8))).astype(np.int64)
p2 = np.where(cis, np.minimum(p1 + d, sizes[c2]), (rng.random(n_pairs) * sizes[c2]).astype(np.int64) + 1)
tc = [torch.from_numpy(a).pin_memory() for a in (c1, p1, c2, p2)]
res = []
for binsize in (1_000_000, 10_000, 1_000):
    for _ in range(2):
        px = ingest.bin_pairs(tc[0], tc[1], tc[2], tc[3], sizes, binsize)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    px = ingest.bin_pairs(tc[0], tc[1], tc[2], tc[3], sizes, binsize)
    dt = time.perf_counter() - t0
    res.append({"binsize": binsize, "pairs": n_pairs, "pixels_out": int(len(px["count"])), "seconds_e2e_from_pinned_host": dt,
                "pairs_per_s": n_pairs / dt, "h2d_bytes": 24 * n_pairs})
    print(res[-1], flush=True)
from oracle import ingest_numpy  # noqa: E402
m = min(10_000_000, n_pairs)
t0 = time.perf_counter()
want = ingest_numpy.bin_pairs(c1[:m], p1[:m], c2[:m], p2[:m], sizes, 10_000)
dt = time.perf_counter() - t0
got = ingest.bin_pairs(c1[:m], p1[:m], c2[:m], p2[:m], sizes, 10_000)
assert all(np.array_equal(got[k], want[k]) for k in want)
out["pairs_binning"] = {"runs": res, "cpu_oracle": {"pairs": m, "binsize": 10_000, "seconds": dt, "pairs_per_s": m / dt,
                                                     "cores": 1, "kind": "port"}}
print(out["pairs_binning"]["cpu_oracle"], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "bench_widened.json"), "w") as f:
    json.dump(out, f, indent=1)
