#!/usr/bin/env python
"""Benchmark of the ICE balancing hot path (BASELINE.json metric: ICE pixels/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workloads are BASELINE.json's configs (synthetic hg38 power-law Hi-C, cooler_b200/synth.py, seed 0):

  N = 1, 2, 4   configs[2]: hg38 10 kb genome-wide ICE, 308 839 bins, ~1e9 stored pixels, the whole
                table on one GPU or sharded by contiguous bin1 row ranges (STRONG scaling);
  N = 8         configs[3]: hg38 1 kb genome-wide ICE, 3 088 298 bins, ~3e9 stored pixels, sharded
                over the 8 GPUs, per-pass all-reduce of the n_bins float64 vector over NVLink.

A "step" is one complete ICE iteration loop (reference defaults: mad_max=5, min_nnz=10,
ignore_diags=2, tol=1e-5) to convergence over the HBM-resident table; value = stored pixels x
marginal passes / time.  `e2e` is the same metric through the host-buffer API:
pinned host arrays -> H2D -> build (filter, radix sorts, prefilter marginals) -> loop -> bias back
on the host, all inside the timed region.  `parity` compares this very run with the CPU oracle
(oracle/ice_numpy.py) outside the timed region.  `--impl reference` (and the `cpu_baseline` leg at
N = 1) time the UNMODIFIED reference `cooler.balance_cooler` (installed copy in oracle/_ref,
multiprocess.Pool(P).imap_unordered over an in-memory store) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ICE_KW = dict(ignore_diags=2, mad_max=5, min_nnz=10, min_count=0, tol=1e-5, max_iters=200)
CONFIGS = {   # n_gpus -> (BASELINE config, binsize, stored-pixel target of the WHOLE table)
    1: ("configs[2] C3", 10_000, 1_000_000_000), 2: ("configs[2] C3", 10_000, 1_000_000_000),
    4: ("configs[2] C3", 10_000, 1_000_000_000), 8: ("configs[3] C4", 1_000, 3_000_000_000),
}
REF_PIXEL_PASS_BUDGET = 4.0e10   # reference arm: pixels x passes it may process (a few minutes of host time)


def workload(n_gpus, binsize, nnz):
    name, b, t = CONFIGS.get(n_gpus, CONFIGS[1])
    if binsize or nnz:
        name = "custom"
    return name, int(binsize or b), int(nnz or t)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])), mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons), "window": "warm-up + timed steps, 20 ms period"}


# ------------------------------------------------------------------------------- CPU side
def host_px(data):
    """numpy arrays in the reference's on-disk dtypes (bin1_id materialised from the offsets)."""
    off = data["bin1_offset"].numpy()
    b1 = np.repeat(np.arange(data["n_bins"], dtype=np.int64), np.diff(off))
    return {"bin1_id": b1, "bin2_id": data["bin2_id"].numpy().astype(np.int64, copy=False),
            "count": data["count"].numpy()}


def reference_passes(plan, data, n_passes, warmup, nproc=None):
    """Time ``n_passes`` marginal passes of the UNMODIFIED reference (cooler._balance.balance_cooler,
    installed copy oracle/_ref or /root/reference) over ``data`` registered as an in-memory store
    (oracle/refshim.py), with ``multiprocess.Pool(P).imap_unordered`` as the CLI does
    (cli/balance.py:237-243).  tol = 0 keeps it iterating; every pass logs "variance is ..."
    (_balance.py:109) and the log timestamps delimit the passes.  Falls back to the numpy port
    (oracle/ice_numpy.py, kind "port") when the reference package is not available."""
    import logging
    import warnings

    from oracle import refshim
    P = nproc or os.cpu_count() or 1
    px = host_px(data)
    nnz = len(px["bin2_id"])
    n = data["n_bins"]
    co = np.asarray(data["chrom_offset"])
    chunk = int(max(1_000_000, nnz // (4 * P)))
    note = (f"{n_passes} timed marginal passes (after {warmup} warm-up passes and the 2 prefilter passes) over "
            f"{nnz} stored pixels, chunksize {chunk}, Pool({P}).imap_unordered; in-memory store (no HDF5 "
            "decode) => upper bound on the file-backed reference; tol=0 so that it keeps iterating")
    if not refshim.available():
        return port_passes(data, px, n_passes, warmup, P, chunk, note)
    import multiprocess as mp
    from cooler_b200 import synth
    nb = np.diff(co)
    chrom = np.repeat(np.arange(len(nb), dtype=np.int32), nb)
    start = np.concatenate([np.arange(k, dtype=np.int64) * plan.binsize for k in nb])
    end = np.minimum(start + plan.binsize, np.repeat(plan.chromsizes, nb))
    name = f"bench_{plan.binsize}_{nnz}.cool"
    refshim.register(name, synth.HG38_NAMES, plan.chromsizes, chrom, start, end, px["bin1_id"], px["bin2_id"],
                     px["count"], plan.binsize)
    cooler = refshim.import_reference()
    stamps = []

    class Stamp(logging.Handler):
        def emit(self, record):
            if record.getMessage().startswith("variance is"):
                stamps.append(time.perf_counter())

    lg = logging.getLogger("cooler")
    h, old, old_handlers, old_prop = Stamp(level=logging.INFO), lg.level, lg.handlers[:], lg.propagate
    lg.handlers[:] = [h]                           # only the time stamps: nothing but the JSON line on stdout
    lg.propagate = False
    lg.setLevel(logging.INFO)
    pool = mp.Pool(P)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            kw = {k: v for k, v in ICE_KW.items() if k not in ("tol", "max_iters")}
            cooler._balance.balance_cooler(cooler.Cooler(name), tol=0.0, max_iters=warmup + n_passes,
                                           chunksize=chunk, map=pool.imap_unordered, **kw)
    finally:
        pool.terminate()
        lg.handlers[:] = old_handlers
        lg.propagate = old_prop
        lg.setLevel(old)
        refshim._REGISTRY.pop(name, None)
    assert len(stamps) == warmup + n_passes, (len(stamps), warmup, n_passes)
    t = stamps[-1] - stamps[warmup - 1]
    return {"value": nnz * n_passes / t, "unit": "pixels/s", "cores": P, "kind": "reference",
            "sample": note + "; code: unmodified cooler._balance.balance_cooler v0.10.4 via oracle/refshim.py",
            "seconds": t, "nnz": nnz}


def port_passes(data, px, n_passes, warmup, P, chunk, note):
    import multiprocessing as mp

    from oracle import ice_numpy
    nnz, n = len(px["bin2_id"]), data["n_bins"]
    co = np.asarray(data["chrom_offset"])
    chrom_ids = np.repeat(np.arange(len(co) - 1, dtype=np.int32), np.diff(co))
    ice_numpy.share(px, chrom_ids, n)
    spans = [(lo, min(lo + chunk, nnz)) for lo in range(0, nnz, chunk)]
    bias = np.ones(n)
    kw = dict(binarize=False, zero_trans=False, n_diags=ICE_KW["ignore_diags"], zero_cis=False)
    times = []
    with mp.get_context("fork").Pool(P) as pool:
        for _ in range(warmup + n_passes):
            t0 = time.perf_counter()
            ice_numpy.ice_step(bias, ice_numpy.pool_marginal(pool.imap_unordered, spans, vec=bias, **kw))
            times.append(time.perf_counter() - t0)
    t = float(np.sum(times[warmup:]))
    return {"value": nnz * n_passes / t, "unit": "pixels/s", "cores": P, "kind": "port",
            "sample": note + "; code: oracle/ice_numpy.py (reference package not installed in oracle/_ref)",
            "seconds": t, "nnz": nnz}


def bounded_rows(plan, nnz_total, passes_total, mem_bytes_per_px=60):
    """Rows [0, r) of the table the CPU arm runs on: the whole table when the pixel-pass budget and
    the host memory allow, else the leading rows holding about the affordable number of pixels."""
    import psutil
    afford = REF_PIXEL_PASS_BUDGET / max(passes_total, 1)
    afford = min(afford, 0.5 * psutil.virtual_memory().available / mem_bytes_per_px)
    if afford >= nnz_total:
        return (0, plan.n_bins), "the whole table"
    cw = plan.row_weight_cum
    r = int(np.searchsorted(cw, cw[-1] * afford / nnz_total))
    r = max(1, min(r, plan.n_bins))
    return (0, r), f"rows [0, {r}) of {plan.n_bins} (~{afford:.3g} of {nnz_total:.3g} pixels: pixel-pass / host-memory budget)"


def oracle_parity(data, prob, state_factory, bias0_gpu, group, world, rank, full, device):
    """Parity of THIS run against the CPU oracle (oracle/ice_numpy.py), outside the timed region.

    Every rank evaluates the oracle's span pipeline (_init, filters, outer product, _marginalize) on
    its own host shard; partial marginals are summed across ranks (integers exactly).  Checked:
    the two prefilter marginals (exact), the filter mask (exact), then ``n`` passes of the oracle's
    update loop against the same number of passes of the kernels (bias and variance to 1e-9) --
    all passes to convergence (mask, pass count, scale, var, bias) when ``full``."""
    import torch
    from oracle import ice_numpy
    px = host_px(data)
    n = data["n_bins"]
    co = np.asarray(data["chrom_offset"])
    chrom_ids = np.repeat(np.arange(len(co) - 1, dtype=np.int32), np.diff(co))
    nnz = len(px["bin2_id"])
    pool = None
    if world == 1 and nnz > 20_000_000:
        import multiprocessing as mp
        ice_numpy.share(px, chrom_ids, n)
        pool = mp.get_context("fork").Pool(os.cpu_count() or 1)
    chunk = int(max(1_000_000, nnz // (4 * (os.cpu_count() or 1)))) if pool else 10_000_000
    spans = [(lo, min(lo + chunk, nnz)) for lo in range(0, nnz, chunk)]

    def marginal(**kw):
        if pool is not None:
            part = ice_numpy.pool_marginal(pool.imap_unordered, spans, **kw)
        else:
            part = ice_numpy._marginal(px, chrom_ids, n, spans, map, **kw)
        if world > 1:
            import torch.distributed as dist
            t = torch.from_numpy(part).to(device)
            dist.all_reduce(t, group=group)
            part = t.cpu().numpy()
        return part

    t0 = time.perf_counter()
    base = dict(zero_trans=False, n_diags=ICE_KW["ignore_diags"], zero_cis=False)
    try:
        m_nnz = marginal(binarize=True, vec=None, **base)
        m_sum = marginal(binarize=False, vec=None, **base)
        out = {"oracle": "oracle/ice_numpy.py", "prefilter_marginals_exact":
               bool(np.array_equal(m_nnz, prob.marg_nnz) and np.array_equal(m_sum, prob.marg_sum))}
        bias = ice_numpy.initial_bias(n, m_nnz, m_sum.copy(), co, mad_max=ICE_KW["mad_max"],
                                      min_nnz=ICE_KW["min_nnz"], min_count=ICE_KW["min_count"])
        out["mask_exact"] = bool(np.array_equal(bias == 0, np.asarray(bias0_gpu) == 0))
        out["masked_bins"] = int((bias == 0).sum())
        n_passes = ICE_KW["max_iters"] if full else 3
        variances, mean = [], np.nan
        for _ in range(n_passes):
            var, mean = ice_numpy.ice_step(bias, marginal(binarize=False, vec=bias, **base))
            variances.append(float(var))
            if var < ICE_KW["tol"]:
                break
        state = state_factory(len(variances))
        state.run()
        res = state.results()
        g_bias, g_var = res["bias"], res["var_hist"][: len(variances), 0]
        keep = bias != 0
        out["passes_compared"] = len(variances)
        out["max_rel_bias_diff"] = float(np.max(np.abs(g_bias[keep] / bias[keep] - 1))) if keep.any() else 0.0
        out["max_rel_var_diff"] = float(np.max(np.abs(g_var / np.asarray(variances) - 1)))
        out["zero_bias_bins_equal"] = bool(np.array_equal(g_bias == 0, bias == 0))
        if full:
            out["converged_pass_count_equal"] = bool(int(res["seg_iters"][0]) == len(variances)
                                                     and bool(res["seg_done"][0]) == (variances[-1] < ICE_KW["tol"]))
            out["rel_scale_diff"] = float(abs(res["seg_scale"][0] / mean - 1))
        out["scope"] = ("all passes to convergence" if full else "first 3 passes") + \
                       (" on the whole table" if world == 1 else f"; every rank on its own shard, partials summed over {world} ranks")
        out["ok"] = bool(out["prefilter_marginals_exact"] and out["mask_exact"] and out["zero_bias_bins_equal"]
                         and out["max_rel_bias_diff"] <= 1e-9 and out["max_rel_var_diff"] <= 1e-9
                         and out.get("converged_pass_count_equal", True))
    finally:
        if pool is not None:
            pool.terminate()
    out["seconds"] = time.perf_counter() - t0
    return out


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if part:
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_host_to_gpu_node(torch, index):
    """Run this rank's host threads on the NUMA node its GPU hangs off, so that the pinned host buffers
    (first touched by cudaHostAlloc in this process) are local to the GPU's PCIe root -- what `numactl
    --cpunodebind` does for a one-process-per-GPU job.  Returns a dict for the JSON line, or None when the
    topology is unknown (single-node hosts report numa_node -1).  Undone by ``os.sched_setaffinity(0, all)``
    before the CPU-baseline leg, which uses every core."""
    try:
        p = torch.cuda.get_device_properties(index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _cpulist(f.read()) & os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"gpu_pci": bus, "numa_node": node, "cpus": len(cpus)}
    except Exception:  # pragma: no cover - depends on the host
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--binsize", type=int, default=0, help="override the config's bin size")
    ap.add_argument("--nnz", type=float, default=0, help="override the config's stored-pixel target (whole table)")
    ap.add_argument("--mode", default="gw", choices=["gw", "cis", "trans"])
    ap.add_argument("--cpu-passes", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--parity", default="auto", choices=["auto", "full", "3"],
                    help="oracle passes compared: all to convergence, or the first 3 (auto: full on one GPU)")
    ap.add_argument("--e2e-steps", type=int, default=3,
                    help="timed end-to-end calls after one warm-up call (at most --steps); their mean is reported")
    ap.add_argument("--host-bin2", default="int32", choices=["int64", "int32"],
                    help="dtype of the pinned host bin2_id column of the e2e leg: int32 = narrowed while "
                         "reading the file, as read_cooler_arrays does (default); int64 = the on-disk dtype")
    ap.add_argument("--calibrate", type=int, default=3,
                    help="multi-GPU: rounds of cost calibration of the row shards (one sweep per rank, then the "
                         "row boundaries move so that the sweep times agree); 0 = shards of equal pixel counts")
    ap.add_argument("--two-stage", action="store_true",
                    help="multi-GPU: force the reduce-scatter + all-gather form of the fused all-reduce (default: "
                         "only for vectors beyond balance.TWO_STAGE_BYTES, i.e. the 1 kb config at 8 GPUs)")
    ap.add_argument("--layout", default=None, help="force a sweep layout (IceProblem band=...), e.g. rows, allblocks")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="do not move this rank's host threads to the NUMA node of its GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)          # timing rule: at least 3 warm-up steps

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)
    cfg_name, binsize, nnz_target = workload(n_gpus, args.binsize, args.nnz)

    import torch

    from cooler_b200 import synth

    plan = synth.SynthPlan(binsize, nnz_target, seed=0)
    # identical in both arms (the driver compares it): only what follows from the command line
    cfg = {"workload": f"BASELINE {cfg_name}: synthetic hg38 {binsize} bp genome-wide ICE, {plan.n_bins} bins, "
                       f"~{nnz_target:.3g} stored pixels (whole table), mode={args.mode}, defaults mad_max=5 "
                       "min_nnz=10 ignore_diags=2 tol=1e-5",
           "binsize": binsize, "n_bins": plan.n_bins, "nnz_target": nnz_target, "mode": args.mode,
           "sharding": f"{n_gpus} contiguous bin1 row ranges" if n_gpus > 1 else "single GPU",
           "l2": "inputs larger than L2 (no flush needed)"}
    scaling = "weak" if cfg_name.endswith("C4") else "strong"

    # ------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        K = max(args.steps, 1)
        rows, what = bounded_rows(plan, nnz_target, args.warmup + K + 2)
        dev = "cuda:0" if torch.cuda.is_available() else "cpu"
        data = synth.generate(plan, rows=rows, device=dev, pin=False)
        res = reference_passes(plan, data, K, args.warmup)
        res["sample"] = f"{what}; " + res["sample"]
        val = res["value"]
        line = {"impl": "reference", "metric": "ICE pixels/s", "value": val, "unit": "pixels/s",
                "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * res["seconds"] / K, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg, "cpu_baseline": res,
                "e2e": {"value": val, "unit": "pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    host_numa = None if args.no_numa_bind else bind_host_to_gpu_node(torch, local_rank)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        group = dist.group.WORLD

    import cooler_b200
    from cooler_b200 import balance as _balance
    from cooler_b200.balance import IceProblem, IceState, prefilter_bias
    if args.two_stage:
        _balance.TWO_STAGE_BYTES = 0

    dev = f"cuda:{local_rank}"
    kw = dict(cis_only=args.mode == "cis", trans_only=args.mode == "trans")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def build(rows):
        """This rank's shard resident in HBM: host arrays, uploaded table, filtered copies, loop state."""
        data = synth.generate(plan, rows=rows, device=dev)
        torch.cuda.empty_cache()
        table = cooler_b200.PixelTable.from_arrays(data["bin1_offset"], data["bin2_id"], data["count"],
                                                   data["chrom_offset"], device=dev, row_range=data["row_range"])
        prob = IceProblem(table, ignore_diags=ICE_KW["ignore_diags"], group=group, band=args.layout, **kw)
        bias0 = prefilter_bias(prob, table.chrom_offset, mad_max=ICE_KW["mad_max"], min_nnz=ICE_KW["min_nnz"],
                               min_count=ICE_KW["min_count"], blacklist=None, x0=None)
        n = table.n_bins
        cweights = None
        if args.mode == "cis":
            seg_offset = table.chrom_offset
        else:
            seg_offset = np.array([0, n])
            if args.mode == "trans":
                co = table.chrom_offset
                cweights = 1.0 / np.repeat(1 - np.diff(co) / n, np.diff(co))
        state = IceState(prob, bias0, cweights, seg_offset, ICE_KW["tol"], ICE_KW["max_iters"])
        return data, table, prob, bias0, seg_offset, state

    # --- resident leg: table and filtered copies already in HBM ---------------
    rows = plan.shard_rows(rank, world)
    calibration = None
    shares = np.full(world, 1.0 / world)
    best = (float("inf"), rows)
    for rnd in range(args.calibrate + 1 if world > 1 and args.calibrate > 0 else 0):
        # cost-balanced shards: pixels are not equally expensive (a row near the start of the genome spreads
        # its pixels over more column strips / tiles than one near the end), so every rank times its sweep and
        # the pixel shares are scaled by (mean time / own time) ** 0.8; the split with the smallest slowest
        # sweep seen is kept (the last round only measures)
        data, table, prob, bias0, seg_offset, state = build(rows)
        import ctypes as C
        from cooler_b200 import _lib
        from cooler_b200.table import _stream
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for k in range(2):                       # warm-up launch set, then the timed one
            barrier()
            ev[0].record()
            for _ in range(4):
                _lib.call("cb200_ice_sweep", C.byref(state.st), _stream())
            ev[1].record()
            torch.cuda.synchronize()
        t_all = torch.zeros(world, dtype=torch.float64, device="cuda")
        t_all[rank] = ev[0].elapsed_time(ev[1]) / 4
        dist.all_reduce(t_all)
        t_all = t_all.cpu().numpy()
        calibration = (calibration or []) + [{"round": rnd + 1, "pixel_share": [round(float(x), 4) for x in shares],
                                             "sweep_ms": [round(float(x), 4) for x in t_all]}]
        if t_all.max() < best[0]:
            best = (float(t_all.max()), rows)
        del data, table, prob, state
        torch.cuda.empty_cache()
        if t_all.max() / t_all.min() < 1.03 or rnd == args.calibrate:
            break
        shares = shares * (t_all.mean() / t_all) ** 0.8
        shares /= shares.sum()
        rows = synth.shard_rows_by_share(plan, shares)[rank]
    rows = best[1]
    data, table, prob, bias0, seg_offset, state = build(rows)
    n = table.n_bins
    nnz_local = data["nnz"]
    nnz_total = torch.tensor([nnz_local], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(nnz_total)
    nnz_total = int(nnz_total.item())
    nnz_eff = torch.tensor([prob.nnz_eff], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(nnz_eff)
    nnz_eff = int(nnz_eff.item())

    # the timed region lasts tens of milliseconds, less than nvidia-smi needs to start: the sampler
    # runs from before the warm-up steps (same kernels, same load) to the end of the timed steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    for _ in range(args.warmup):
        state.reset(bias0)
        state.run()
    passes_per_step = int(state.results()["seg_iters"].max())
    barrier()
    state.sweep_events = []
    state.comm_events = []
    state.stage_events = []
    state.launches = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launched = 0
    ev0.record()
    for _ in range(args.steps):
        state.reset(bias0)
        launched += state.run()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    t_ms = float(t_ms.item())
    res = state.results()
    passes_per_step = int(res["seg_iters"].max())
    converged = bool(res["seg_done"].all())
    # sweep kernel: average duration of the passes that did work (speculative no-op launches excluded)
    per_step = launched // max(args.steps, 1)
    sweep_ms = [e0.elapsed_time(e1) for k, (e0, e1) in enumerate(state.sweep_events)
                if (k % per_step) < passes_per_step]
    sweep_ms_avg = float(np.mean(sweep_ms)) if sweep_ms else float("nan")
    comm_ms = [c0.elapsed_time(c1) for k, (c0, c1) in enumerate(state.comm_events)
               if (k % per_step) < passes_per_step]
    comm_ms_avg = float(np.mean(comm_ms)) if comm_ms else 0.0
    # multi-GPU: mean duration of the stages after the sweep (this rank; the slowest rank's line is reported)
    stage_names = (["combine", "barrier", "reduce_scatter", "barrier2", "gather+update"]
                   if getattr(state, "two_stage", False) else ["combine", "barrier", "allreduce+update"])
    stage_ms = [0.0] * len(stage_names)
    used = [m for k, m in enumerate(state.stage_events) if (k % per_step) < passes_per_step]
    for m in used:
        for i in range(len(stage_names)):
            stage_ms[i] += m[i].elapsed_time(m[i + 1]) / len(used)
    gpu_launches = state.launches
    state.sweep_events = None
    value = nnz_total * passes_per_step * args.steps / (t_ms * 1e-3)

    peak, peak_src = load_peaks()
    bytes_pass_local = prob.bytes_per_pass()
    achieved = bytes_pass_local / (sweep_ms_avg * 1e-3) / 1e9
    # the roofline line is the SLOWEST rank's: the loop runs at its pace
    per_rank = torch.zeros(world, 3 + 5, dtype=torch.float64, device="cuda")
    per_rank[rank, :3] = torch.tensor([sweep_ms_avg, achieved / peak, comm_ms_avg], dtype=torch.float64)
    per_rank[rank, 3:3 + len(stage_ms)] = torch.tensor(stage_ms, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(per_rank)
    per_rank = per_rank.cpu().numpy()
    slow = int(np.argmin(per_rank[:, 1]))
    # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel on
    # this workload (profiles/*traffic.json), reported only when the pixel count matches
    traffic = None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp) and world == 1 and traffic is None:
            with open(tp) as f:
                for ent in json.load(f):
                    if ent["binsize"] == binsize and ent["mode"] == args.mode and ent["nnz_effective"] == prob.nnz_eff:
                        traffic = ent["dram_bytes_per_launch"]
    far = bool(getattr(prob, "far_copies", None))
    if far and not prob.subs:
        layout = "2-D blocks, lane rows" if prob.band == "rows" else "2-D blocks, lane pieces"
        kernel_name = "ice_sweep_rows_kernel" if prob.band == "rows" else "ice_sweep_far_kernel"
    elif far:
        layout, kernel_name = "band strips + far blocks", "ice_sweep_strips_kernel + ice_sweep_far_kernel"
    elif prob.strip_shift is not None:
        layout, kernel_name = f"column strips of {1 << prob.strip_shift} bins", "ice_sweep_strips_kernel"
    else:
        layout, kernel_name = "whole copies", "ice_sweep_kernel"
    bytes_all = torch.tensor([float(bytes_pass_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        gathered = [torch.zeros_like(bytes_all) for _ in range(world)]
        dist.all_gather(gathered, bytes_all)
        bytes_slow = float(gathered[slow].item())
    else:
        bytes_slow = float(bytes_pass_local)
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": float(per_rank[slow, 1] * peak), "peak": peak,
                "unit": "GB/s", "frac": float(per_rank[slow, 1]), "traffic": traffic, "peak_source": peak_src,
                "rank": slow, "algorithmic_bytes_per_launch": int(bytes_slow),
                "kernel_ms_avg": float(per_rank[slow, 0]),
                "record_stream_bytes_per_launch_rank0": int(prob.stream_bytes()),
                "formula": "16 B x effective pixels + 40 B x n_bins (per rank), SURVEY.md 8d; the slowest rank's launch"
                           + ("; the 2-D blocked layout stores compact 4-byte records, so the bytes actually streamed "
                              "(record_stream_bytes_per_launch_rank0) are about half the algorithmic ones" if far and
                              all(f.rec_bytes == 4 for f in prob.far_copies if f is not None) else ""),
                "frac_per_rank": [round(float(x), 4) for x in per_rank[:, 1]],
                "frac_of_nominal_8TBs": float(per_rank[slow, 1] * peak / 8000.0)}

    # --- parity gate (outside the timed region): this run against the CPU oracle ---------------
    parity = None
    if not args.no_parity and args.mode == "gw":
        full = args.parity == "full" or (args.parity == "auto" and world == 1)

        def state_factory(k):
            return IceState(prob, bias0.copy(), None, seg_offset, ICE_KW["tol"], k)

        os.sched_setaffinity(0, all_cpus)          # the oracle uses every core
        parity = oracle_parity(data, prob, state_factory, bias0, group, world, rank, full, dev)
        if host_numa is not None:
            bind_host_to_gpu_node(torch, local_rank)
        ok = torch.tensor([1 if parity["ok"] else 0], device="cuda")
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        parity["ok"] = bool(ok.item())

    # --- e2e leg: host buffers in, bias on the host out ---------------------------
    del state, prob, table
    torch.cuda.empty_cache()
    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    if args.host_bin2 == "int32":
        data["bin2_id"] = data["bin2_id"].to(torch.int32).pin_memory()
    h2d = (data["bin2_id"].numel() * data["bin2_id"].element_size() + data["count"].numel() * 4
           + data["bin1_offset"].numel() * 8)
    d2h = n * 8 * 3   # int64 nnz/sum marginals + float64 bias
    t_e2e = []
    info = None
    for k in range(1 + e2e_steps):       # first run = warm-up
        barrier()
        t0 = time.perf_counter()
        bias, stats, info = cooler_b200.balance_arrays(
            data["bin1_offset"], data["bin2_id"], data["count"], data["chrom_offset"],
            device=dev, row_range=data["row_range"], group=group, return_info=True, band=args.layout,
            **kw, **{k_: v for k_, v in ICE_KW.items()})
        barrier()
        if k > 0:
            t_e2e.append(time.perf_counter() - t0)
    t_e = torch.tensor([float(np.mean(t_e2e))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_passes = int(max(info["passes"]))
    e2e = {"value": nnz_total * e2e_passes / float(t_e.item()), "unit": "pixels/s",
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "seconds_per_call": float(t_e.item()), "passes": e2e_passes,
           "host_bin2_dtype": args.host_bin2,
           "host_numa": host_numa,     # this rank's threads and pinned buffers on the GPU's NUMA node (None: unknown / off)
           "n_subcopies": info["n_subcopies"],
           "api": "cooler_b200.balance_arrays(pinned host arrays) -> numpy bias (per rank: its row shard)"}

    cpu = None
    os.sched_setaffinity(0, all_cpus)              # the CPU leg uses every core
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rows, what = bounded_rows(plan, nnz_total, args.cpu_passes + 3 + 2)
        sample = data if rows == (0, plan.n_bins) else synth.generate(plan, rows=rows, device=dev, pin=False)
        cpu = reference_passes(plan, sample, args.cpu_passes, 3)
        cpu["sample"] = f"{what}; " + cpu["sample"]

    if rank == 0:
        detail = {"shard_calibration": calibration if world > 1 else None, "sweep_layout": layout, "nnz_stored": nnz_total, "nnz_effective": nnz_eff,
                  "passes_per_step": passes_per_step, "converged": converged,
                  "masked_bins": int(np.isnan(bias).sum())}
        line = {"metric": "ICE pixels/s", "value": value, "unit": "pixels/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_ms / max(args.steps, 1),
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": cfg, "detail": detail, "clocks": clocks, "e2e": e2e,
                "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
                "time_to_converge_s": {"resident_loop": t_ms * 1e-3 / max(args.steps, 1),
                                       "e2e_from_pinned_host": float(t_e.item())},
                "per_pass_ms": {"total": t_ms / max(args.steps, 1) / max(passes_per_step, 1),
                                "sweep_slowest_rank": float(per_rank[:, 0].max()),
                                "sweep_fastest_rank": float(per_rank[:, 0].min()),
                                "sweep_per_rank": [round(float(x), 4) for x in per_rank[:, 0]],
                                "allreduce_incl_wait_slowest_rank": float(per_rank[:, 2].max()),
                                "stages_after_sweep_on_slowest_rank": (
                                    {nm: round(float(per_rank[int(np.argmax(per_rank[:, 0])), 3 + i]), 4)
                                     for i, nm in enumerate(stage_names)} if world > 1 else None)}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
