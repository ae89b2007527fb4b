"""gloo tests (world size 2 and 3, no CUDA) of the multi-GPU host logic: row-range shards of the synthetic
table and of a real .cool file -- per-shard partial marginals all-reduced across ranks == the single-process
marginal --, and exchange-free sharded coarsening (shards cut at coarse-bin boundaries, pieces gathered in
rank order == the reference's coarsened table).  The numpy oracle stands in for the device kernels."""
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


def _coarsen_worker(rank, world, port, out, factor):
    sys.path.insert(0, ROOT)
    torch.set_num_threads(1)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cooler_b200 import MemCooler
    from cooler_b200.reduce import gather_pixels_to_rank0, shard_rows_for_coarsening
    from oracle import coarsen_numpy
    from tests import _golden
    t = _golden.gm12878()
    clr = MemCooler.from_arrays([str(x) for x in t["chromnames"]], t["chromsizes"], t["bins_chrom"], t["bins_start"],
                                t["bins_end"], t["bin1_id"], t["bin2_id"], t["count"], binsize=2_000_000)
    lo, hi = shard_rows_for_coarsening(clr, dist.group.WORLD, factor)
    sel = (t["bin1_id"] >= lo) & (t["bin1_id"] < hi)          # this rank's rows, cut at coarse-bin boundaries
    px = {k: t[k][sel] for k in ("bin1_id", "bin2_id", "count")}
    b1, b2, c, _ = coarsen_numpy.coarsen_pixels(px, t["bins_chrom"], t["bins_start"], t["bins_end"], t["chromsizes"],
                                                factor)
    g1, g2, gc = gather_pixels_to_rank0([b1, b2, c], dist.group.WORLD)
    if rank == 0:
        np.save(out, np.stack([g1, g2, gc.astype(np.int64)], axis=1))
        np.save(out + ".rows.npy", np.array([lo, hi]))
    else:
        assert len(g1) == 0
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,factor", [(2, 5), (3, 2)])
def test_sharded_coarsening_needs_no_exchange(tmp_path, world, factor):
    """coarsen_cooler(group=...): every rank coarsens the rows `shard_rows_for_coarsening` gives it (the numpy
    oracle stands in for the device coarsener) and `gather_pixels_to_rank0` concatenates the pieces in rank
    order -- the result equals the unmodified reference's coarsened table: no coarse row is split, nothing
    is exchanged."""
    out = str(tmp_path / "coarse.npy")
    mp.start_processes(_coarsen_worker, args=(world, _free_port(), out, factor), nprocs=world, join=True,
                       start_method="spawn")
    got = np.load(out)
    sys.path.insert(0, ROOT)
    from tests import _golden
    CO = _golden.coarsen_golden()
    want = np.stack([CO[f"gm/x{factor}/{k}"].astype(np.int64) for k in ("bin1_id", "bin2_id", "count")], axis=1)
    np.testing.assert_array_equal(got, want)
    lo, hi = np.load(out + ".rows.npy")
    assert lo == 0 and 0 < hi < 1561


def _file_shard_worker(rank, world, port, out, path):
    sys.path.insert(0, ROOT)
    torch.set_num_threads(1)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cooler_b200 import store
    from cooler_b200.balance import _allreduce
    from cooler_b200.table import read_cooler_arrays, shard_rows_by_pixels
    from oracle import ice_numpy
    clr = store.open_cooler(path)
    full_offset = np.asarray(clr._load_dset("indexes/bin1_offset"))
    rows = shard_rows_by_pixels(full_offset, rank, world)            # what balance_cooler(group=...) reads
    off, bin2, count, chrom_offset = read_cooler_arrays(clr, rows=rows)
    n = len(off) - 1
    assert int(off[rows[0]]) == 0 and int(off[-1]) == len(bin2) == int(full_offset[rows[1]] - full_offset[rows[0]])
    b1 = np.repeat(np.arange(n, dtype=np.int64), np.diff(off))
    assert len(b1) == 0 or (b1.min() >= rows[0] and b1.max() < rows[1])
    px = dict(bin1_id=b1, bin2_id=np.asarray(bin2, dtype=np.int64), count=np.asarray(count))
    chrom_ids = np.repeat(np.arange(len(chrom_offset) - 1, dtype=np.int32), np.diff(chrom_offset))
    w = np.random.default_rng(1).lognormal(0, 0.2, n)
    part = ice_numpy._span_marginal(px, chrom_ids, n, 0, len(b1), binarize=False, zero_trans=False, n_diags=2,
                                    zero_cis=False, vec=w)
    t = torch.from_numpy(part.copy())
    _allreduce(t, dist.group.WORLD)
    sizes = torch.tensor([len(b1)])
    dist.all_reduce(sizes)
    if rank == 0:
        np.save(out, t.numpy())
        np.save(out + ".nnz.npy", sizes.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_file_row_shards_cover_the_table_once(tmp_path):
    """balance_cooler(clr, group=...) on a real .cool file: rank g reads pixels[bin1_offset[r_g] :
    bin1_offset[r_g+1]] through the built-in HDF5 reader; the shards' partial marginals summed over
    the ranks equal the marginal of the whole file (the oracle stands in for the device sweep)."""
    path = os.path.join(ROOT, "tests", "golden", "cool", "hg19.GM12878-MboI.matrix.2000kb.cool")
    out = str(tmp_path / "marg.npy")
    world = 3
    mp.start_processes(_file_shard_worker, args=(world, _free_port(), out, path), nprocs=world, join=True,
                       start_method="spawn")
    sys.path.insert(0, ROOT)
    from oracle import ice_numpy
    from tests import _golden
    t = _golden.gm12878()
    chrom_offset, _ = _golden.offsets(t)
    n = len(t["bins_chrom"])
    chrom_ids = np.repeat(np.arange(len(chrom_offset) - 1, dtype=np.int32), np.diff(chrom_offset))
    w = np.random.default_rng(1).lognormal(0, 0.2, n)
    px = {k: t[k] for k in ("bin1_id", "bin2_id", "count")}
    want = ice_numpy._span_marginal(px, chrom_ids, n, 0, len(t["count"]), binarize=False, zero_trans=False,
                                    n_diags=2, zero_cis=False, vec=w)
    assert int(np.load(out + ".nnz.npy")[0]) == len(t["count"])
    np.testing.assert_allclose(np.load(out), want, rtol=1e-12)
