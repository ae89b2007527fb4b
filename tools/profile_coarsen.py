import ctypes as C, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import cooler_b200
from cooler_b200 import synth, _lib
from cooler_b200 import table as T
from cooler_b200.reduce import coarsen_bin_map
binsize, target, factor = int(sys.argv[1]), int(float(sys.argv[2])), int(sys.argv[3])
plan = synth.SynthPlan(binsize, target, seed=0)
data = synth.generate(plan)
tab = cooler_b200.PixelTable.from_arrays(data["bin1_offset"], data["bin2_id"], data["count"], data["chrom_offset"])
dev = tab.device
class Tm:
    def __init__(s): s.t = {}
    def __call__(s, name):
        torch.cuda.synchronize(); now = time.perf_counter()
        if hasattr(s, "_l"): s.t[s._n] = s.t.get(s._n, 0) + now - s._l
        s._l, s._n = now, name
for rep in range(3):
    tm = Tm(); tm("setup")
    csr = tab.raw
    bin_map, new_off = coarsen_bin_map(tab.chrom_offset, factor)
    n_new = int(new_off[-1]); bits = T.key_bits_for(n_new)
    bm = torch.from_numpy(bin_map).to(dev); n = csr.nnz
    key = torch.empty(n, dtype=torch.int32, device=dev); val = torch.empty((n + 2, 2), dtype=torch.int32, device=dev)
    cs = csr.c_struct()
    tm("remap")
    _lib.call("cb200_coarsen_remap", C.byref(cs), csr.n_rows, T._ptr(bm), T._ptr(key), T._ptr(val), T._stream())
    tm("sort1")
    ks, vs = T.sort_pairs(key, val, bits)
    tm("runs_count+scan")
    nt = T.n_tiles(n); tc = torch.zeros(nt, dtype=torch.int32, device=dev)
    _lib.call("cb200_runs_count", T._ptr(ks), T._ptr(vs), n, T._ptr(tc), T._stream())
    toff = T.exclusive_scan_i32(tc); n_out = int(toff[-1].item())
    tm("runs_write")
    k2 = torch.empty(n_out, dtype=torch.int32, device=dev); v2 = torch.empty((n_out + 2, 2), dtype=torch.int32, device=dev)
    ov = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.call("cb200_runs_write", T._ptr(ks), T._ptr(vs), n, T._ptr(toff), 1, T._ptr(k2), T._ptr(v2), C.c_void_p(0), T._ptr(ov), T._stream())
    tm("sort2")
    k3, v3 = T.sort_pairs(k2, v2, bits)
    tm("indptr")
    ind = T.indptr_from_sorted(k3, n_out, n_new)
    tm("end")
    if rep: print(bits, n, n_out, {k: round(v * 1e3, 2) for k, v in tm.t.items()}, flush=True)
    del key, val, ks, vs, k2, v2, k3, v3
import cProfile, pstats
from cooler_b200.reduce import coarsen_table
coarsen_table(tab, factor); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter(); new, ids = coarsen_table(tab, factor); torch.cuda.synchronize(); print("coarsen_table wall ms", (time.perf_counter() - t0) * 1e3)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
