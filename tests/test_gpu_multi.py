"""Sharded (2-rank, NCCL) evaluation against the single-GPU one; needs two devices (``-m gpu``)."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_match_one_rank():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    command = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", "29611", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    result = subprocess.run(command, capture_output=True, text=True, timeout=900)
    assert result.returncode == 0, result.stdout[-2000:] + result.stderr[-4000:]
    assert "multi-GPU check ok" in result.stdout
