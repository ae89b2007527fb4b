"""Runs tools/check_dist.py under torchrun when >= 2 GPUs are visible."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_balance_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "check_dist.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(res.stdout[-3000:], res.stderr[-3000:])
    assert res.returncode == 0


def _torchrun(args, port, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.gpu
def test_balance_cooler_group_reads_row_shards_of_real_files(tmp_path):
    """balance_cooler(clr, group=...) on the bundled GM12878 .cool (reference goldens) and on a
    written 1e7-pixel .cool (oracle), two ranks, every rank reading its own bin1 row range; then
    coarsen_cooler / zoomify_cooler(group=...) against the reference's coarsening goldens and a
    single-rank run."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    res = _torchrun([os.path.join(ROOT, "tools", "check_dist_cooler.py"), str(tmp_path)], 29519)
    print(res.stdout[-3000:], res.stderr[-3000:])
    assert res.returncode == 0 and "FAIL" not in res.stdout and res.stdout.count(" OK") >= 13   # + 3 coarsen, 4 general aggregations, 1 zoomify


@pytest.mark.gpu
def test_balance_cli_under_torchrun(tmp_path):
    """`torchrun ... -m cooler_b200.cli balance FILE`: rank 0 checks and writes, exit codes as the
    reference's CLI (cli/balance.py:180-206)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import shutil

    import numpy as np

    from cooler_b200 import store
    from tests import _golden
    src = os.path.join(ROOT, "tests", "golden", "cool", "hg19.GM12878-MboI.matrix.2000kb.cool")
    fn = str(tmp_path / "gm.cool")
    shutil.copy(src, fn)
    res = _torchrun(["-m", "cooler_b200.cli", "balance", fn, "--check"], 29521)
    assert res.returncode == 1 and "No 'weight' column found" in res.stdout, (res.stdout, res.stderr[-2000:])
    res = _torchrun(["-m", "cooler_b200.cli", "balance", fn], 29523)
    assert res.returncode == 0, res.stderr[-3000:]
    BAL, _ = _golden.balance_golden()
    with store.open_cooler(fn).open("r") as grp:
        np.testing.assert_allclose(np.asarray(grp["bins"]["weight"][:]), BAL["gm/defaults"], rtol=1e-9, equal_nan=True)
    res = _torchrun(["-m", "cooler_b200.cli", "balance", fn], 29525)          # exists, no --force
    assert res.returncode == 1 and "already exists" in res.stderr
    res = _torchrun(["-m", "cooler_b200.cli", "balance", fn, "--force", "--cis-only"], 29527)
    assert res.returncode == 0, res.stderr[-3000:]
    with store.open_cooler(fn).open("r") as grp:
        np.testing.assert_allclose(np.asarray(grp["bins"]["weight"][:]), BAL["gm/cis_only"], rtol=1e-9, equal_nan=True)


@pytest.mark.gpu
def test_zoomify_and_coarsen_cli_under_torchrun(tmp_path):
    """`torchrun ... -m cooler_b200.cli zoomify FILE -r ... --balance` and `... coarsen FILE -k 5`:
    the files rank 0 writes hold the same pixels and weights as a single-process run."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import shutil

    import numpy as np

    from cooler_b200 import store
    src = os.path.join(ROOT, "tests", "golden", "cool", "hg19.GM12878-MboI.matrix.2000kb.cool")
    fn = str(tmp_path / "gm.cool")
    shutil.copy(src, fn)
    outs = {}
    for tag, run in (("multi", lambda a, port: _torchrun(["-m", "cooler_b200.cli", *a], port)),
                     ("single", lambda a, port: subprocess.run([sys.executable, "-m", "cooler_b200.cli", *a],
                                                               capture_output=True, text=True, cwd=ROOT))):
        mc, co = str(tmp_path / f"{tag}.mcool"), str(tmp_path / f"{tag}.x5.cool")
        res = run(["zoomify", fn, "-r", "4000000,8000000,40000000", "--balance", "-o", mc], 29531)
        assert res.returncode == 0, res.stderr[-3000:]
        res = run(["coarsen", fn, "-k", "5", "-o", co], 29533)
        assert res.returncode == 0, res.stderr[-3000:]
        outs[tag] = (mc, co)
    for res_ in (4000000, 8000000, 40000000):
        a = store.open_cooler(f"{outs['multi'][0]}::resolutions/{res_}")
        b = store.open_cooler(f"{outs['single'][0]}::resolutions/{res_}")
        for col in ("bin1_id", "bin2_id", "count"):
            np.testing.assert_array_equal(np.asarray(a._load_dset("pixels/" + col)), np.asarray(b._load_dset("pixels/" + col)))
        with a.open("r") as ga, b.open("r") as gb:
            np.testing.assert_allclose(np.asarray(ga["bins"]["weight"][:]), np.asarray(gb["bins"]["weight"][:]),
                                       rtol=1e-9, equal_nan=True)
    a, b = store.open_cooler(outs["multi"][1]), store.open_cooler(outs["single"][1])
    for col in ("bin1_id", "bin2_id", "count"):
        np.testing.assert_array_equal(np.asarray(a._load_dset("pixels/" + col)), np.asarray(b._load_dset("pixels/" + col)))
