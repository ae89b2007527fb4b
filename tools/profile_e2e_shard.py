"""Where the end-to-end time of one rank goes (pinned host arrays -> bias on the host) for a row
shard of a big table: python tools/profile_e2e_shard.py <binsize> <nnz_total> <world> <rank>.
Every stage function is wrapped with a device synchronisation, so the sum is a little above the
pipelined wall time that is printed beside it."""
import os
import sys
import time
from collections import defaultdict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cooler_b200  # noqa: E402
from cooler_b200 import balance as B, synth, table as T  # noqa: E402

binsize = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
target = int(float(sys.argv[2])) if len(sys.argv) > 2 else 3_000_000_000
world = int(sys.argv[3]) if len(sys.argv) > 3 else 8
rank = int(sys.argv[4]) if len(sys.argv) > 4 else 0
plan = synth.SynthPlan(binsize, target, seed=0)
data = synth.generate(plan, rows=plan.shard_rows(rank, world))
data["bin2_id"] = data["bin2_id"].to(torch.int32).pin_memory()
print("n_bins", plan.n_bins, "nnz(shard)", data["nnz"], flush=True)
acc = defaultdict(float)


def wrap(mod, name, label=None):
    fn = getattr(mod, name)

    def inner(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize()
        acc[label or name] += time.perf_counter() - t0
        return out
    setattr(mod, name, inner)


NOWRAP = bool(os.environ.get("NOWRAP"))      # NOWRAP=1: no per-stage synchronisation, the true pipelined wall time
for mod, name in () if NOWRAP else ((B, "filter_copy"), (B, "transpose_from"), (B, "marginals_i64"), (B, "far_blocks"),
                  (B, "swap_key_other"), (B, "strip_partition"), (B, "prefilter_bias"), (B, "plan_far_units")):
    wrap(mod, name)
if not NOWRAP:
    wrap(B.PixelTable, "from_arrays", "PixelTable.from_arrays (H2D + pack)")
    wrap(B.IceState, "run", "IceState.run (loop)")
    wrap(B.IceState, "results", "IceState.results (D2H)")
_init = B.IceState.__init__


def timed_init(self, *a, **k):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _init(self, *a, **k)
    torch.cuda.synchronize()
    acc["IceState.__init__ (incl. plan_far_units)"] += time.perf_counter() - t0


if not NOWRAP:
    B.IceState.__init__ = timed_init
for rep in range(3):
    acc.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bias, stats, info = cooler_b200.balance_arrays(data["bin1_offset"], data["bin2_id"], data["count"],
                                                   data["chrom_offset"], row_range=data["row_range"], return_info=True)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if rep:
        print({k: round(v * 1e3, 1) for k, v in acc.items()}, "| wall ms", round(wall * 1e3, 1),
              "passes", max(info["passes"]), "layout", info["band"], flush=True)
