"""Throughput of the one-time build primitives on random data: stable radix sort of (key, 8-byte value)
pairs per pass, filter compaction, exclusive scan:  python tools/bench_build_prims.py [n]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cooler_b200 import table as T  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000
dev = T._use_device("cuda")
out = {"n": n}
for bits in (8, 16, 22):
    keys = torch.randint(0, 1 << bits, (n,), dtype=torch.int32, device=dev)
    vals = torch.empty((n + 2, 2), dtype=torch.int32, device=dev)
    vals[:n, 0] = torch.arange(n, dtype=torch.int32, device=dev)
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ks, vs = T.sort_pairs(keys, vals, bits)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    passes = (bits + 7) // 8
    ok = bool((ks[1:] >= ks[:-1]).all())
    out[f"sort_{bits}bit"] = {"ms": ms, "passes": passes, "GBps_per_pass_rw": 24 * n * passes / ms / 1e6, "sorted": ok}
    print(bits, out[f"sort_{bits}bit"], flush=True)
    del keys, vals, ks, vs
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "bench_build_prims.json"), "w") as f:
    json.dump(out, f, indent=1)
