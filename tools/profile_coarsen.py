"""Stage timings of the merge form of K5 (cooler_b200.reduce.coarsen_table) on a resident
synthetic table:  python tools/profile_coarsen.py <binsize> <nnz> <factor>"""
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
from cooler_b200 import table as T  # noqa: E402
from cooler_b200.reduce import coarsen_bin_map, coarsen_table  # noqa: E402

binsize, target, factor = int(sys.argv[1]), int(float(sys.argv[2])), int(sys.argv[3])
plan = synth.SynthPlan(binsize, target, seed=0)
data = synth.generate(plan)
tab = cooler_b200.PixelTable.from_arrays(data["bin1_offset"], data["bin2_id"], data["count"], data["chrom_offset"])
dev = tab.device


class Tm:
    def __init__(self):
        self.t = {}

    def __call__(self, name):
        torch.cuda.synchronize()
        now = time.perf_counter()
        if hasattr(self, "_l"):
            self.t[self._n] = self.t.get(self._n, 0) + now - self._l
        self._l, self._n = now, name


for rep in range(3):
    tm = Tm()
    tm("setup")
    csr = tab.raw
    bin_map, new_off = coarsen_bin_map(tab.chrom_offset, factor)
    n_new = int(new_off[-1])
    n = csr.nnz
    gf = np.searchsorted(bin_map, np.arange(n_new + 1), side="left").astype(np.int32)
    mg = int(np.diff(gf).max())
    rounds = max(1, int(mg - 1).bit_length())
    bm = torch.from_numpy(bin_map).to(dev)
    gfd = torch.from_numpy(gf).to(dev)
    claim = torch.zeros(1, dtype=torch.int64, device=dev)
    keys = torch.empty(n, dtype=torch.int32, device=dev)
    bufs = [torch.empty((n + 2, 2), dtype=torch.int32, device=dev) for _ in range(min(rounds, 2))]
    src = csr.px
    for r in range(rounds):
        tm(f"merge_round_{r}")
        dst = bufs[r % 2]
        lists = -(-mg // (1 << r))
        _lib.call("cb200_coarsen_merge_round", T._ptr(src), T._ptr(dst), T._ptr(csr.indptr), T._ptr(gfd), n_new, r,
                  -(-lists // 2), T._ptr(bm) if r == 0 else C.c_void_p(0),
                  T._ptr(keys) if r == rounds - 1 else C.c_void_p(0), T._ptr(claim), T._stream())
        src = dst
    tm("runs_count+scan")
    nt = T.n_tiles(n)
    tc = torch.zeros(nt, dtype=torch.int32, device=dev)
    _lib.call("cb200_runs_count", T._ptr(keys), T._ptr(src), n, T._ptr(tc), T._stream())
    toff = T.exclusive_scan_i32(tc)
    n_out = int(toff[-1].item())
    tm("runs_write")
    k2 = torch.empty(n_out, dtype=torch.int32, device=dev)
    v2 = torch.empty((n_out + 2, 2), dtype=torch.int32, device=dev)
    ov = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.call("cb200_runs_write", T._ptr(keys), T._ptr(src), n, T._ptr(toff), 0, T._ptr(k2), T._ptr(v2),
              C.c_void_p(0), T._ptr(ov), T._stream())
    tm("indptr")
    ind = T.indptr_from_sorted(k2, n_out, n_new)
    tm("end")
    if rep:
        print(n, n_out, "rounds", rounds, {k: round(v * 1e3, 2) for k, v in tm.t.items()}, flush=True)
    del keys, bufs, src, k2, v2
coarsen_table(tab, factor)
torch.cuda.synchronize()
t0 = time.perf_counter()
new, ids = coarsen_table(tab, factor)
torch.cuda.synchronize()
print("coarsen_table wall ms", (time.perf_counter() - t0) * 1e3)
