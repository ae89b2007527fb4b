#!/usr/bin/env python
"""Benchmark of the ICE balancing hot path (BASELINE.json metric: ICE pixels/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (weak scaling, ~2e8 stored pixels per GPU, synthetic hg38 power-law Hi-C,
cooler_b200/synth.py):  N=1 hg38 100 kb genome-wide ICE (BASELINE configs[1]);
N=2 50 kb / 4e8 px; N=4 25 kb / 8e8 px; N=8 12.5 kb / 1.6e9 px, sharded by
contiguous bin1 row ranges with one NCCL all-reduce of the n_bins marginal vector
per pass.

A "step" is one complete ICE iteration loop (reference defaults: mad_max=5,
min_nnz=10, ignore_diags=2, tol=1e-5) to convergence over the HBM-resident table;
value = stored pixels x marginal passes / time.  `e2e` is the same
metric through the host-buffer API: pinned host arrays -> H2D -> build (filter, radix
sort, prefilter marginals) -> loop -> bias back on the host, all inside the timed region.
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

PX_PER_GPU = 200_000_000
BASE_BINSIZE = 100_000
ICE_KW = dict(ignore_diags=2, mad_max=5, min_nnz=10, min_count=0, tol=1e-5, max_iters=200)


def workload(n_gpus, px_per_gpu, binsize):
    binsize = binsize or BASE_BINSIZE // n_gpus
    return binsize, px_per_gpu * n_gpus


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


def host_arrays_for_cpu(data):
    """numpy views in the reference's dtypes (bin1_id materialised from the offsets)."""
    off = data["bin1_offset"].numpy()
    n = data["n_bins"]
    b1 = np.repeat(np.arange(n, dtype=np.int64), np.diff(off))
    return {"bin1_id": b1, "bin2_id": data["bin2_id"].numpy(), "count": data["count"].numpy()}


def cpu_ice_passes(data, n_passes, warmup=0, nproc=None):
    """Time ``n_passes`` ICE marginal passes (PASS C of _balance.py:88-95: filters,
    outer product with the bias, bincount marginals, fold) of the oracle over the
    host arrays with a fork()ed process pool, the reference's `-p N` path."""
    import multiprocessing as mp

    from oracle import ice_numpy
    px = host_arrays_for_cpu(data)
    nnz = len(px["bin2_id"])
    n = data["n_bins"]
    P = nproc or os.cpu_count() or 1
    chrom_ids = np.repeat(np.arange(len(data["chrom_offset"]) - 1, dtype=np.int32), np.diff(data["chrom_offset"]))
    ice_numpy.share(px, chrom_ids, n)
    chunk = int(max(1_000_000, nnz // (4 * P)))
    spans = [(lo, min(lo + chunk, nnz)) for lo in range(0, nnz, chunk)]
    bias = np.ones(n)
    kw = dict(binarize=False, zero_trans=False, n_diags=ICE_KW["ignore_diags"], zero_cis=False)
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(P) as pool:
        for k in range(warmup + n_passes):
            t0 = time.perf_counter()
            marg = ice_numpy.pool_marginal(pool.imap_unordered, spans, vec=bias, **kw)
            nz = marg[marg != 0]
            m = marg / nz.mean()
            m[m == 0] = 1
            bias = bias / m
            times.append(time.perf_counter() - t0)
    t = float(np.sum(times[warmup:]))
    return {"value": nnz * n_passes / t, "unit": "pixels/s", "cores": P, "kind": "port",
            "sample": f"{n_passes} ICE marginal passes over {nnz} stored pixels "
                      f"({len(spans)} spans of {chunk}), Pool({P}).imap_unordered, oracle/ice_numpy.py; "
                      "in-memory arrays (no HDF5 decode) => upper bound on the file-backed reference",
            "seconds": t, "nnz": nnz}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--px-per-gpu", type=int, default=PX_PER_GPU)
    ap.add_argument("--binsize", type=int, default=0)
    ap.add_argument("--mode", default="gw", choices=["gw", "cis", "trans"])
    ap.add_argument("--cpu-passes", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--host-bin2", default="int64", choices=["int64", "int32"],
                    help="dtype of the pinned host bin2_id column of the e2e leg: int64 = the on-disk "
                         "dtype (default); int32 = narrowed by the HDF5 read as PixelTable.from_cooler does")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)          # timing rule: at least 3 warm-up steps

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)
    binsize, nnz_target = workload(n_gpus, args.px_per_gpu, args.binsize)

    import torch

    from cooler_b200 import synth

    plan = synth.SynthPlan(binsize, nnz_target, seed=0)
    cfg = {"workload": f"synthetic hg38 {binsize} bp genome-wide ICE ({plan.n_bins} bins, ~{nnz_target:.3g} px target), "
                       f"mode={args.mode}, defaults mad_max=5 min_nnz=10 ignore_diags=2 tol=1e-5",
           "binsize": binsize, "n_bins": plan.n_bins, "mode": args.mode,
           "sharding": f"{world} contiguous bin1 row ranges" if world > 1 else "single GPU",
           "l2": "inputs larger than L2 (no flush needed)"}

    # ------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        blocks = plan.shard_rows(0, n_gpus)
        data = synth.generate(plan, rows=blocks, pin=False)
        res = cpu_ice_passes(data, max(args.steps, 1), warmup=args.warmup)
        val = res["value"]
        line = {"impl": "reference", "metric": "ICE pixels/s", "value": val, "unit": "pixels/s",
                "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * res["seconds"] / max(args.steps, 1), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg, "cpu_baseline": res,
                "e2e": {"value": val, "unit": "pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        group = dist.group.WORLD

    import cooler_b200
    from cooler_b200.balance import IceProblem, IceState, prefilter_bias

    blocks = plan.shard_rows(rank, world)
    data = synth.generate(plan, rows=blocks, device=f"cuda:{local_rank}")
    torch.cuda.empty_cache()
    nnz_local = data["nnz"]
    nnz_total = torch.tensor([nnz_local], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(nnz_total)
    nnz_total = int(nnz_total.item())
    kw = dict(cis_only=args.mode == "cis", trans_only=args.mode == "trans")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # --- resident leg: table and filtered copies already in HBM ---------------
    table = cooler_b200.PixelTable.from_arrays(data["bin1_offset"], data["bin2_id"], data["count"],
                                               data["chrom_offset"], device=f"cuda:{local_rank}",
                                               row_range=data["row_range"])
    prob = IceProblem(table, ignore_diags=ICE_KW["ignore_diags"], group=group, **kw)
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
    # slowest / fastest rank of the dominant kernel (shard balance) and the all-reduce share
    rk = torch.tensor([sweep_ms_avg, -sweep_ms_avg, comm_ms_avg], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(rk, op=dist.ReduceOp.MAX)
    sweep_ms_max, sweep_ms_min, comm_ms_max = float(rk[0]), -float(rk[1]), float(rk[2])
    gpu_launches = state.launches
    state.sweep_events = None
    value = nnz_total * passes_per_step * args.steps / (t_ms * 1e-3)

    peak, peak_src = load_peaks()
    bytes_pass_local = prob.bytes_per_pass()
    achieved = bytes_pass_local / (sweep_ms_avg * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel on
    # this workload (profiles/r01_traffic.json), reported only when the pixel count matches
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp) and world == 1:
        with open(tp) as f:
            for ent in json.load(f):
                if ent["binsize"] == binsize and ent["mode"] == args.mode and ent["nnz_effective"] == prob.nnz_eff:
                    traffic = ent["dram_bytes_per_launch"]
    layout = ("all blocks" if getattr(prob, "far_copies", None) and not prob.subs else
              "band strips + far blocks" if getattr(prob, "far_copies", None) else
              f"column strips of {1 << prob.strip_shift} bins" if prob.strip_shift is not None else "whole copies")
    if getattr(prob, "far_copies", None) and not prob.subs:
        kernel_name = "ice_sweep_far_kernel"          # every pixel through the 2-D blocked kernel (csrc/far.cuh)
    elif getattr(prob, "far_copies", None):
        kernel_name = "ice_sweep_strips_kernel + ice_sweep_far_kernel"
    else:
        kernel_name = "ice_sweep_strips_kernel" if prob.strip_shift is not None else "ice_sweep_kernel"
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_pass_local, "kernel_ms_avg": sweep_ms_avg,
                "formula": "16 B x effective pixels + 40 B x n_bins (per rank), SURVEY.md 8d",
                "frac_of_nominal_8TBs": achieved / 8000.0}

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
            device=f"cuda:{local_rank}", row_range=data["row_range"], group=group, return_info=True,
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
           "n_subcopies": info["n_subcopies"],
           "api": "cooler_b200.balance_arrays(pinned host arrays) -> numpy bias (build pipelined with the H2D copy)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_ice_passes(data, args.cpu_passes)

    if rank == 0:
        cfg.update({"sweep_layout": layout, "nnz_stored": nnz_total, "nnz_effective": nnz_eff, "passes_per_step": passes_per_step,
                    "converged": converged, "masked_bins": int(np.isnan(bias).sum())})
        line = {"metric": "ICE pixels/s", "value": value, "unit": "pixels/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_ms / max(args.steps, 1),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e,
                "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu,
                "time_to_converge_s": {"resident_loop": t_ms * 1e-3 / max(args.steps, 1),
                                       "e2e_from_pinned_host": float(t_e.item())},
                "per_pass_ms": {"total": t_ms / max(args.steps, 1) / max(passes_per_step, 1),
                                "sweep_rank0": sweep_ms_avg, "sweep_slowest_rank": sweep_ms_max,
                                "sweep_fastest_rank": sweep_ms_min,
                                "allreduce_incl_wait_slowest_rank": comm_ms_max}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
