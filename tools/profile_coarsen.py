"""Stage timings of K5 (cooler_b200.reduce.coarsen_table: merge rounds, fused merge + reduce, scan,
compaction) on a resident synthetic table, for several chunk widths of the merge kernel, each checked
bit for bit against the unfused path (merge rounds -> runs_count -> runs_write):

    python tools/profile_coarsen.py <binsize> <nnz> <factor> [factor ...]
"""
import ctypes as C
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cooler_b200  # noqa: E402
from cooler_b200 import _lib, synth  # noqa: E402
from cooler_b200 import reduce as R  # noqa: E402
from cooler_b200 import table as T  # noqa: E402

binsize, target = int(sys.argv[1]), int(float(sys.argv[2]))
factors = [int(x) for x in sys.argv[3:]]
VTS = [int(x) for x in os.environ.get("VTS", "7").split(",")]
plan = synth.SynthPlan(binsize, target, seed=0)
data = synth.generate(plan)
tab = cooler_b200.PixelTable.from_arrays(data["bin1_offset"], data["bin2_id"], data["count"], data["chrom_offset"])
dev = tab.device
lib = _lib.load()
PEAK = 6536.7


class Tm:
    def __init__(self):
        self.t = {}

    def __call__(self, name):
        torch.cuda.synchronize()
        now = time.perf_counter()
        if hasattr(self, "_l"):
            self.t[self._n] = self.t.get(self._n, 0) + now - self._l
        self._l, self._n = now, name


def unfused(table, factor):
    keys, vals, n_new, new_off = R._merge_coarse_rows(table, factor)
    k2, v2, n_out, _ = R._reduce_runs(keys, vals, table.raw.nnz, 0, dev)
    return k2[:n_out].clone(), v2[:n_out].clone()


out = []
for factor in factors:
    want_k, want_v = unfused(tab, factor)
    for vt in VTS:
        if hasattr(lib, "cb200_dev_set_merge_vt"):
            lib.cb200_dev_set_merge_vt(vt)
        for rep in range(3):
            tm = Tm()
            tm("maps")
            csr = tab.raw
            n = csr.nnz
            maps = R._coarse_maps(tab, factor)
            bin_map_d, group_first_d, n_new, new_off, rounds, _, remap = maps
            tm("merge_rounds")
            src = R._merge_rounds(csr, maps, rounds - 1)
            tm("alloc")
            tmp = torch.empty((n + 2, 2), dtype=torch.int32, device=dev)
            row_count = torch.zeros(n_new, dtype=torch.int32, device=dev)
            ov = torch.zeros(1, dtype=torch.int32, device=dev)
            claim = torch.zeros(1, dtype=torch.int64, device=dev)
            tm("merge_reduce")
            _lib.call("cb200_coarsen_merge_reduce", T._ptr(src), T._ptr(tmp), T._ptr(csr.indptr), T._ptr(group_first_d),
                      n_new, rounds - 1, C.byref(remap) if rounds == 1 else None, T._ptr(row_count),
                      T._ptr(ov), T._ptr(claim), T._stream())
            tm("scan")
            indptr = T.exclusive_scan_i32(row_count)
            n_out = int(indptr[-1].item())
            tm("compact")
            k2 = torch.empty(n_out, dtype=torch.int32, device=dev)
            v2 = torch.empty((n_out + 2, 2), dtype=torch.int32, device=dev)
            _lib.call("cb200_coarsen_compact", T._ptr(tmp), T._ptr(csr.indptr), T._ptr(group_first_d), T._ptr(indptr),
                      n_new, T._ptr(v2), T._ptr(k2), T._stream())
            tm("end")
        ok = bool(n_out == want_k.numel() and torch.equal(k2, want_k) and torch.equal(v2[:n_out], want_v))
        R.coarsen_table(tab, factor)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        new, ids = R.coarsen_table(tab, factor)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        algo = 12 * (n + n_out)
        rec = {"factor": factor, "vt": vt, "nnz_in": n, "nnz_out": n_out, "rounds": rounds, "equal_to_unfused": ok,
               "stages_ms": {k: round(v * 1e3, 3) for k, v in tm.t.items()}, "coarsen_table_ms": ms,
               "algorithmic_GBps": algo / ms / 1e6, "frac_of_measured_hbm": algo / ms / 1e6 / PEAK}
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del tmp, k2, v2, src
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "profile_coarsen.json"), "w") as f:
    json.dump(out, f, indent=1)
