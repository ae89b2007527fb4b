"""Multi-GPU parity check, run under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29517 tools/check_dist.py

Every rank balances its bin1 row-range shard of a synthetic table with one
all-reduce of the marginal vector per pass (fused into K4a over peer memory; also in its
two-stage reduce-scatter / all-gather form); rank 0 runs the CPU ORACLE (oracle/ice_numpy.py)
on the whole table.  Masks and pass counts must be identical, bias / scale / var within 1e-9.
The all-blocks layout on shards is checked the same way.
"""
import os
import sys

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
    from cooler_b200 import synth
    synth.ROWS_PER_BLOCK_TARGET = 400_000
    ok = True
    for binsize, target in ((1_000_000, 2_000_000), (250_000, 12_000_000)):
        plan = synth.SynthPlan(binsize, target, seed=1)
        for mode in ("gw", "cis", "trans"):
            kw = dict(cis_only=mode == "cis", trans_only=mode == "trans")
            shard = synth.generate(plan, rows=plan.shard_rows(rank, world), device=f"cuda:{local}")
            tab = cooler_b200.PixelTable.from_arrays(shard["bin1_offset"], shard["bin2_id"], shard["count"],
                                                     shard["chrom_offset"], device=f"cuda:{local}",
                                                     row_range=shard["row_range"])
            bias, stats, info = cooler_b200.balance_table(tab, group=dist.group.WORLD, return_info=True, **kw)
            if rank == 0:
                from oracle import ice_numpy
                full = synth.generate(plan, device=f"cuda:{local}", pin=False)
                off = full["bin1_offset"].numpy()
                px = dict(bin1_id=np.repeat(np.arange(plan.n_bins, dtype=np.int64), np.diff(off)),
                          bin2_id=full["bin2_id"].numpy(), count=full["count"].numpy())
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    want, wstats, wpasses, _ = ice_numpy.balance_arrays(px, plan.chrom_offset, off,
                                                                        return_passes=True, **kw)

                def judge(b, st, inf):
                    same_mask = np.array_equal(np.isnan(b), np.isnan(want))
                    err = float(np.nanmax(np.abs(b / want - 1))) if np.isfinite(want).any() else 0.0
                    sc = np.nanmax(np.abs(np.atleast_1d(st["scale"]) / np.atleast_1d(wstats["scale"]) - 1))
                    vr = np.nanmax(np.abs(np.atleast_1d(st["var"]) / np.where(np.atleast_1d(wstats["var"]) == 0, 1,
                                                                            np.atleast_1d(wstats["var"])) - 1)
                                   * (np.atleast_1d(wstats["var"]) != 0))
                    good = (same_mask and inf["passes"] == wpasses and err < 1e-9 and not sc > 1e-9 and not vr > 1e-9)
                    return good, same_mask, err

                good, same_mask, err = judge(bias, stats, info)
                ok &= bool(good)
                print(f"[check_dist] world={world} binsize={binsize} mode={mode} nnz={full['nnz']} "
                      f"passes={max(info['passes'])} vs ORACLE: mask_equal={same_mask} max_rel_err={err:.2e} "
                      f"{'OK' if good else 'FAIL'}", flush=True)
            dist.barrier()
            if mode == "gw":
                # the layout huge sparse tables get automatically (2-D blocked kernel), forced here on the shards
                b2, s2, i2 = cooler_b200.balance_table(tab, group=dist.group.WORLD, return_info=True,
                                                      band="allblocks", far_col_shift=9, **kw)
                if rank == 0:
                    good, same_mask, err = judge(b2, s2, i2)
                    good = good and i2["band"] == "allblocks"
                    ok &= bool(good)
                    print(f"[check_dist] world={world} binsize={binsize} all-blocks layout on shards vs ORACLE: "
                          f"max_rel_err={err:.2e} {'OK' if good else 'FAIL'}", flush=True)
                dist.barrier()
                # the two-stage (reduce-scatter + all-gather) form of the fused all-reduce, forced
                import cooler_b200.balance as _bal
                old_thr, _bal.TWO_STAGE_BYTES = _bal.TWO_STAGE_BYTES, -1
                try:
                    b3, s3, i3 = cooler_b200.balance_table(tab, group=dist.group.WORLD, return_info=True, **kw)
                finally:
                    _bal.TWO_STAGE_BYTES = old_thr
                if rank == 0:
                    good, same_mask, err = judge(b3, s3, i3)
                    ok &= bool(good)
                    print(f"[check_dist] world={world} binsize={binsize} two-stage all-reduce vs ORACLE: "
                          f"max_rel_err={err:.2e} {'OK' if good else 'FAIL'}", flush=True)
                dist.barrier()
                # floating-point count column on the shards (values = count x a position-keyed factor, so every
                # rank derives its own slice without communication); float prefilter marginals all-reduced
                def fvals(cnt, b2id):
                    return cnt.numpy().astype(np.float64) * (0.5 + (b2id.numpy() % 7) / 7.0)
                tabf = cooler_b200.PixelTable.from_arrays(shard["bin1_offset"], shard["bin2_id"],
                                                          fvals(shard["count"], shard["bin2_id"]), shard["chrom_offset"],
                                                          device=f"cuda:{local}", row_range=shard["row_range"],
                                                          float_values=True)
                b4, s4, i4 = cooler_b200.balance_table(tabf, group=dist.group.WORLD, return_info=True, **kw)
                if rank == 0:
                    pxf = dict(px, count=fvals(full["count"], full["bin2_id"]))
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        want, wstats, wpasses, _ = ice_numpy.balance_arrays(pxf, plan.chrom_offset, off,
                                                                            return_passes=True, **kw)
                    good, same_mask, err = judge(b4, s4, i4)
                    ok &= bool(good)
                    print(f"[check_dist] world={world} binsize={binsize} float64 counts on shards vs ORACLE: "
                          f"mask_equal={same_mask} max_rel_err={err:.2e} {'OK' if good else 'FAIL'}", flush=True)
                dist.barrier()
    # coarsening shards with no exchange: rows cut at coarse-bin boundaries, every rank coarsens its
    # shard, rank-ordered concatenation == single-GPU result (bit-exact)
    from cooler_b200.reduce import align_rows_to_factor, coarsen_table, table_to_host
    plan = synth.SynthPlan(250_000, 6_000_000, seed=2)
    for factor in (2, 5):
        lo, hi = plan.shard_rows(rank, world)
        lo = align_rows_to_factor(plan.chrom_offset, factor, lo)
        hi = align_rows_to_factor(plan.chrom_offset, factor, hi) if rank + 1 < world else plan.n_bins
        shard = synth.generate(plan, rows=(lo, hi), device=f"cuda:{local}")
        tab = cooler_b200.PixelTable.from_arrays(shard["bin1_offset"], shard["bin2_id"], shard["count"],
                                                 shard["chrom_offset"], device=f"cuda:{local}", row_range=(lo, hi))
        new, ids = coarsen_table(tab, factor)
        g = [torch.from_numpy(x).cuda() for x in table_to_host(new, ids)]
        sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([g[0].numel()], dtype=torch.int64, device="cuda"))
        sizes = [int(x.item()) for x in sizes]
        parts = []
        for col in g:
            bufs = [torch.zeros(max(sizes), dtype=col.dtype, device="cuda") for _ in range(world)]
            pad = torch.zeros(max(sizes), dtype=col.dtype, device="cuda")
            pad[: col.numel()] = col
            dist.all_gather(bufs, pad)
            parts.append(torch.cat([b[:k] for b, k in zip(bufs, sizes)]).cpu().numpy())
        if rank == 0:
            full = synth.generate(plan, device=f"cuda:{local}")
            tab1 = cooler_b200.PixelTable.from_arrays(full["bin1_offset"], full["bin2_id"], full["count"],
                                                      full["chrom_offset"], device=f"cuda:{local}")
            w = table_to_host(*coarsen_table(tab1, factor))
            good = all(np.array_equal(a, b) for a, b in zip(parts, w))
            ok &= bool(good)
            print(f"[check_dist] world={world} sharded coarsen x{factor}: nnz_out={len(w[0])} "
                  f"{'OK (bit-exact)' if good else 'FAIL'}", flush=True)
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
