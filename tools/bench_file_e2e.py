"""File-to-file timings of the drop-in calls on a real ``.cool`` (no h5py in the image:
``h5lite`` reads, ``h5write`` writes): ``balance_cooler(clr, store=True)``,
``coarsen_cooler(in.cool, out.cool, 2)`` and ``zoomify_cooler(in.cool, out.mcool, [...])`` + per-level
balance, on a synthetic hg38 matrix written to disk first.  Results are checked against the
array-level API on the same pixels (weights equal; coarsened pixels bit-exact).

    python tools/bench_file_e2e.py [binsize] [nnz_target] [out.json]
"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cooler_b200  # noqa: E402
from cooler_b200 import create, store, synth  # noqa: E402
from cooler_b200.reduce import coarsen_cooler, coarsen_table, table_to_host, zoomify_cooler  # noqa: E402

binsize = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
target = int(float(sys.argv[2])) if len(sys.argv) > 2 else 50_000_000
out_json = sys.argv[3] if len(sys.argv) > 3 else None

plan = synth.SynthPlan(binsize, target, seed=0)
data = synth.generate(plan)
n = plan.n_bins
off = data["bin1_offset"].numpy()
b2, cnt = data["bin2_id"].numpy(), data["count"].numpy()
nnz = data["nnz"]
co = plan.chrom_offset
nb = np.diff(co)
bins = pd.DataFrame({
    "chrom": pd.Categorical.from_codes(np.repeat(np.arange(len(nb)), nb), categories=synth.HG38_NAMES, ordered=True),
    "start": np.concatenate([np.arange(k, dtype=np.int64) * binsize for k in nb]),
})
bins["end"] = np.minimum(bins["start"] + binsize, np.repeat(plan.chromsizes, nb))
res = {"binsize": binsize, "n_bins": n, "nnz": nnz, "host_cores": os.cpu_count()}
tmp = tempfile.mkdtemp(prefix="cb200_")
path = os.path.join(tmp, "synth.cool")


def chunks(step=10_000_000):
    b1 = np.repeat(np.arange(n, dtype=np.int64), np.diff(off))
    for lo in range(0, nnz, step):
        yield {"bin1_id": b1[lo:lo + step], "bin2_id": b2[lo:lo + step], "count": cnt[lo:lo + step]}


t0 = time.perf_counter()
create.create_cooler(path, bins, chunks(), dupcheck=False)
res["write_cool_s"] = time.perf_counter() - t0
res["cool_bytes"] = os.path.getsize(path)

# ---- balance: file -> bins/weight in the file --------------------------------------------------
clr = store.open_cooler(path)
t0 = time.perf_counter()
arrs = cooler_b200.table.read_cooler_arrays(clr)
res["read_decode_s"] = time.perf_counter() - t0
del arrs
for rep in range(2):                                    # first call allocates the pinned ring
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    arrs = cooler_b200.table.read_cooler_device(store.open_cooler(path))
    torch.cuda.synchronize()
    res["read_to_device_s"] = time.perf_counter() - t0
assert arrs[1].is_cuda and bool((arrs[1].cpu() == data["bin2_id"].to(arrs[1].dtype)).all())
del arrs
for rep in range(2):                                    # first call pays CUDA context + allocator warm-up
    clr = store.open_cooler(path)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bias, stats = cooler_b200.balance_cooler(clr, store=True, store_name=f"weight{rep}")
    torch.cuda.synchronize()
    res["balance_file_s"] = time.perf_counter() - t0
tab = cooler_b200.PixelTable.from_arrays(data["bin1_offset"], data["bin2_id"], data["count"], co)
want, _, info = cooler_b200.balance_table(tab, return_info=True)
with store.open_cooler(path).open("r") as g:
    got = g["bins/weight1"][:]
    st = dict(g["bins/weight1"].attrs)
assert np.array_equal(got, bias, equal_nan=True) and np.array_equal(np.isnan(got), np.isnan(want))
np.testing.assert_allclose(got, want, rtol=1e-12, equal_nan=True)
passes = info["passes"]
passes = int(max(passes)) if isinstance(passes, (list, tuple, np.ndarray)) else int(passes)
res["balance_passes"] = passes
res["balance_converged"] = bool(st["converged"])
res["balance_pixels_per_s_file_to_file"] = nnz * passes / res["balance_file_s"]

# ---- coarsen: file -> file ------------------------------------------------------------------------
out = os.path.join(tmp, "half.cool")
torch.cuda.synchronize()
t0 = time.perf_counter()
coarsen_cooler(path, out, 2, chunksize=10_000_000)
torch.cuda.synchronize()
res["coarsen_file_s"] = time.perf_counter() - t0
new, ids = coarsen_table(tab, 2)
w1, w2, wc = table_to_host(new, ids)
half = store.open_cooler(out)
assert np.array_equal(half._load_dset("pixels/bin1_id"), w1) and np.array_equal(half._load_dset("pixels/bin2_id"), w2)
assert np.array_equal(half._load_dset("pixels/count"), wc)
res["coarsen_nnz_out"] = len(w1)
res["coarsen_input_pixels_per_s_file_to_file"] = nnz / res["coarsen_file_s"]

# ---- zoomify + per-level balance: file -> .mcool ------------------------------------------------------
levels = [binsize * k for k in (2, 5, 10, 50)]
mc = os.path.join(tmp, "synth.mcool")
t0 = time.perf_counter()
zoomify_cooler(path, mc, levels, chunksize=10_000_000)
res["zoomify_file_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
for r in levels:
    cooler_b200.balance_cooler(store.open_cooler(f"{mc}::resolutions/{r}"), store=True)
res["zoomify_balance_levels_s"] = time.perf_counter() - t0
res["zoomify_levels"] = levels
res["mcool_bytes"] = os.path.getsize(mc)
assert sorted(store.list_coolers(mc)) == sorted(f"/resolutions/{r}" for r in [binsize, *levels])
lvl = store.open_cooler(f"{mc}::resolutions/{2 * binsize}")
assert np.array_equal(lvl._load_dset("pixels/count"), wc)
with lvl.open("r") as g:
    assert "weight" in g["bins"]
print(json.dumps(res))
if out_json:
    with open(out_json, "w") as f:
        json.dump(res, f, indent=1)
for p in (path, out, mc):
    os.remove(p)
os.rmdir(tmp)
