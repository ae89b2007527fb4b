"""Multi-GPU entry through the reference API, run under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29519 tools/check_dist_cooler.py [workdir]

1. ``balance_cooler(clr, group=WORLD)`` on the reference's own bundled GM12878 2 Mb ``.cool``
   (every rank reads its bin1 row range of the file) against the golden vectors made by the
   UNMODIFIED reference (tests/golden/make_golden.py): masks exact, bias within 1e-9, all modes.
2. A ~1e7-pixel synthetic ``.cool`` written by rank 0 (``create_cooler``), balanced across the ranks
   with ``store=True``, against the CPU oracle (oracle/ice_numpy.py) run to convergence on rank 0:
   mask, pass count (from the stored stats: converged), bias, scale, var; the stored column is read back.
3. ``coarsen_cooler(..., group=WORLD)`` on the bundled GM12878 file (every rank coarsens its own
   row range, rank 0 gathers) against the reference's coarsening goldens x2 / x5 / x10, and
   ``zoomify_cooler(..., group=WORLD)`` of the written file to an ``.mcool`` whose levels must be
   bit-identical to a single-rank run of the same chain.
4. ``python -m cooler_b200.cli balance`` semantics under torchrun are covered by
   tests/test_dist_gpu.py, which launches the CLI itself.
"""
import os
import sys
import tempfile
import warnings

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import cooler_b200
    from cooler_b200 import store, synth
    from tests import _golden
    group = dist.group.WORLD
    ok = True
    workdir = sys.argv[1] if len(sys.argv) > 1 else tempfile.gettempdir()

    # ---- 1. the bundled GM12878 file vs the reference goldens -------------------------------
    BAL, META = _golden.balance_golden()
    path = os.path.join(ROOT, "tests", "golden", "cool", "hg19.GM12878-MboI.matrix.2000kb.cool")
    for key in ("gm/defaults", "gm/cis_only", "gm/trans_only", "gm/diag0"):
        kw = dict(META[key]["kwargs"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            bias, stats = cooler_b200.balance_cooler(store.open_cooler(path), group=group, **kw)
        want = BAL[key]
        same = np.array_equal(np.isnan(bias), np.isnan(want))
        err = float(np.nanmax(np.abs(bias / want - 1))) if np.isfinite(want).any() else 0.0
        good = same and err < 1e-9
        ok &= bool(good)
        if rank == 0:
            print(f"[check_dist_cooler] world={world} {key}: mask_equal={same} max_rel_err={err:.2e} "
                  f"{'OK' if good else 'FAIL'}", flush=True)

    # ---- 2. a written ~1e7-pixel .cool, store=True, vs the oracle -----------------------------
    synth.ROWS_PER_BLOCK_TARGET = 4_000_000
    plan = synth.SynthPlan(250_000, 10_000_000, seed=3)
    fn = os.path.join(workdir, "check_dist_1e7.cool")
    if rank == 0:
        import pandas as pd
        full = synth.generate(plan, device=f"cuda:{local}", pin=False)
        off = full["bin1_offset"].numpy()
        b1 = np.repeat(np.arange(plan.n_bins, dtype=np.int64), np.diff(off))
        nb = np.diff(plan.chrom_offset)
        start = np.concatenate([np.arange(k, dtype=np.int64) * plan.binsize for k in nb])
        bins = pd.DataFrame({"chrom": np.repeat(np.array(synth.HG38_NAMES), nb), "start": start,
                             "end": np.minimum(start + plan.binsize, np.repeat(plan.chromsizes, nb))})
        px = {"bin1_id": b1, "bin2_id": full["bin2_id"].numpy(), "count": full["count"].numpy()}
        from cooler_b200.create import create_cooler
        create_cooler(fn, bins, px)
    dist.barrier()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bias, stats = cooler_b200.balance_cooler(store.open_cooler(fn), group=group, store=True)
    if rank == 0:
        from oracle import ice_numpy
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want, wstats, wpasses, _ = ice_numpy.balance_arrays(px, plan.chrom_offset, off, return_passes=True)
        same = np.array_equal(np.isnan(bias), np.isnan(want))
        err = float(np.nanmax(np.abs(bias / want - 1)))
        with store.open_cooler(fn).open("r") as grp:
            stored = np.asarray(grp["bins"]["weight"][:])
            attrs = dict(grp["bins"]["weight"].attrs)
        good = (same and err < 1e-9 and bool(stats["converged"]) == bool(wstats["converged"])
                and abs(stats["scale"] / wstats["scale"] - 1) < 1e-9 and abs(stats["var"] / wstats["var"] - 1) < 1e-9
                and np.array_equal(stored, bias, equal_nan=True) and bool(attrs["converged"]))
        ok &= bool(good)
        print(f"[check_dist_cooler] world={world} written .cool nnz={len(b1)} oracle passes={wpasses} "
              f"mask_equal={same} max_rel_err={err:.2e} stored_column_equal={np.array_equal(stored, bias, equal_nan=True)} "
              f"{'OK' if good else 'FAIL'}", flush=True)

    # ---- 3. sharded coarsening / zoomify through the reference API ----------------------------
    from cooler_b200.reduce import coarsen_cooler, zoomify_cooler
    CO = _golden.coarsen_golden()
    for factor in (2, 5, 10):
        out = coarsen_cooler(store.open_cooler(path), None, factor, chunksize=50_000, group=group)
        if rank == 0:
            with out.open("r") as grp:
                got = [np.asarray(grp["pixels"][c][:]) for c in ("bin1_id", "bin2_id", "count")]
            good = all(np.array_equal(g, CO[f"gm/x{factor}/{c}"]) for g, c in zip(got, ("bin1_id", "bin2_id", "count")))
            ok &= bool(good)
            print(f"[check_dist_cooler] world={world} sharded coarsen_cooler x{factor} vs reference golden: "
                  f"nnz_out={len(got[0])} {'OK (bit-exact)' if good else 'FAIL'}", flush=True)
    # general aggregators / extra value columns on row shards (CoolerCoarsener(group=...)), against the goldens
    # of the unmodified reference (tests/golden/make_agg_golden.py)
    from cooler_b200 import MemCooler
    from cooler_b200.reduce import CoolerCoarsener
    from tests import _agg_cases
    gm = _golden.gm12878()
    fcount, big = _agg_cases.derived(gm)
    G = dict(np.load(os.path.join(_golden.GOLD, "coarsen_agg_golden.npz")))
    for key in ("gm_mean_x2", "gm_last_x7", "gm_cols_x2", "gmf_mean_x10"):
        tab, factor, columns, agg = _agg_cases.CASES[key]
        count = gm["count"] if tab == "gm" else (gm["count"] * 0.5).astype(np.float32)
        clr = MemCooler.from_arrays([str(x) for x in gm["chromnames"]], gm["chromsizes"], gm["bins_chrom"],
                                    gm["bins_start"], gm["bins_end"], gm["bin1_id"], gm["bin2_id"], count,
                                    binsize=2_000_000, extra_pixel_columns={"fcount": fcount, "big": big})
        chunks = list(CoolerCoarsener(clr, factor, 5000, columns=columns, agg=agg, group=group))
        if rank == 0:
            got = {k: np.concatenate([c[k] for c in chunks]) for k in chunks[0]}
            good = sorted(got) == sorted(k.split("/", 1)[1] for k in G if k.startswith(key + "/"))
            for k, v in got.items():
                want = G[f"{key}/{k}"]
                good = good and v.dtype == want.dtype and (
                    np.array_equal(v, want) if want.dtype.kind in "iu" else
                    np.allclose(v, want, rtol=1e-6 if want.dtype == np.float32 else 1e-12, atol=0))
            ok &= bool(good)
            print(f"[check_dist_cooler] world={world} sharded general aggregation {key} vs reference golden: "
                  f"{'OK' if good else 'FAIL'}", flush=True)
        else:
            assert not chunks or all(len(c["bin1_id"]) == 0 for c in chunks)
    resolutions = [500_000, 1_000_000, 2_500_000, 5_000_000]
    mc = os.path.join(workdir, "check_dist_zoom.mcool")
    zoomify_cooler(fn, mc, resolutions, chunksize=2_000_000, group=group)
    if rank == 0:
        single = zoomify_cooler(fn, None, resolutions, chunksize=2_000_000)
        good = True
        for res in resolutions:
            with store.open_cooler(f"{mc}::resolutions/{res}").open("r") as grp:
                got = [np.asarray(grp["pixels"][c][:]) for c in ("bin1_id", "bin2_id", "count")]
            with single[res].open("r") as grp:
                want = [np.asarray(grp["pixels"][c][:]) for c in ("bin1_id", "bin2_id", "count")]
            good &= all(np.array_equal(a, b) for a, b in zip(got, want)) and len(got[0]) > 0
        ok &= bool(good)
        print(f"[check_dist_cooler] world={world} sharded zoomify_cooler {resolutions} vs single rank: "
              f"{'OK (bit-exact)' if good else 'FAIL'}", flush=True)
        os.unlink(mc)
    dist.barrier()
    if rank == 0:
        os.unlink(fn)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
