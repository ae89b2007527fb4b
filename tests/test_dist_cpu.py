"""world_size-2 gloo test of the multi-GPU host logic (no CUDA): row-range shards of
the synthetic table, per-shard partial marginals all-reduced across ranks == the
single-process marginal; every rank then derives identical bias updates."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    torch.set_num_threads(1)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cooler_b200 import synth
    from cooler_b200.balance import _allreduce
    from oracle import ice_numpy
    synth.ROWS_PER_BLOCK_TARGET = 200_000
    plan = synth.SynthPlan(1_000_000, 900_000, seed=5)
    shard = synth.generate(plan, rows=plan.shard_rows(rank, world), device="cpu", pin=False)
    n = shard["n_bins"]
    off = shard["bin1_offset"].numpy()
    b1 = np.repeat(np.arange(n, dtype=np.int64), np.diff(off))
    px = dict(bin1_id=b1, bin2_id=shard["bin2_id"].numpy(), count=shard["count"].numpy())
    chrom_ids = np.repeat(np.arange(25, dtype=np.int32), np.diff(plan.chrom_offset))
    rng = np.random.default_rng(0)
    w = rng.lognormal(0, 0.2, n)                      # same vector on every rank
    part = ice_numpy._span_marginal(px, chrom_ids, n, 0, len(b1), binarize=False, zero_trans=False,
                                    n_diags=2, zero_cis=False, vec=w)
    t = torch.from_numpy(part.copy())
    _allreduce(t, dist.group.WORLD)                   # the exchange step of one ICE pass
    nnz = torch.tensor([len(b1)])
    dist.all_reduce(nnz)
    if rank == 0:
        np.save(out, t.numpy())
        np.save(out + ".nnz.npy", nnz.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_marginal_allreduce_equals_global(tmp_path):
    world = 2
    out = str(tmp_path / "marg.npy")
    port = _free_port()
    mp.start_processes(_worker, args=(world, port, out), nprocs=world, join=True, start_method="spawn")
    got = np.load(out)
    sys.path.insert(0, ROOT)
    from cooler_b200 import synth
    from oracle import ice_numpy
    synth.ROWS_PER_BLOCK_TARGET, old = 200_000, synth.ROWS_PER_BLOCK_TARGET
    try:
        plan = synth.SynthPlan(1_000_000, 900_000, seed=5)
        full = synth.generate(plan, device="cpu", pin=False)
    finally:
        synth.ROWS_PER_BLOCK_TARGET = old
    n = full["n_bins"]
    off = full["bin1_offset"].numpy()
    b1 = np.repeat(np.arange(n, dtype=np.int64), np.diff(off))
    px = dict(bin1_id=b1, bin2_id=full["bin2_id"].numpy(), count=full["count"].numpy())
    chrom_ids = np.repeat(np.arange(25, dtype=np.int32), np.diff(plan.chrom_offset))
    w = np.random.default_rng(0).lognormal(0, 0.2, n)
    want = ice_numpy._span_marginal(px, chrom_ids, n, 0, len(b1), binarize=False, zero_trans=False,
                                    n_diags=2, zero_cis=False, vec=w)
    assert int(np.load(out + ".nnz.npy")[0]) == len(b1)
    np.testing.assert_allclose(got, want, rtol=1e-12)
