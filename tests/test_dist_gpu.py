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
