"""Sharded (NCCL + NVLink peer memory) evaluation and MD against the single-GPU ones; needs several devices (``-m gpu``)."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_ranks_match_one_rank(nranks):
    """tools/multi_gpu_check.py under torchrun: forces / energies / virials of every system family, ten MD steps, and
    device-resident MD across neighbour-list rebuilds on a 274 625-atom box, N ranks against one."""
    import torch

    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    command = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks), "--master-addr", "127.0.0.1",
               "--master-port", str(29611 + nranks), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    result = subprocess.run(command, capture_output=True, text=True, timeout=1500)
    assert result.returncode == 0, result.stdout[-2000:] + result.stderr[-4000:]
    assert "multi-GPU check ok" in result.stdout


@pytest.mark.parametrize("ndevices", [2, 8])
def test_one_process_many_devices(ndevices):
    """tools/multi_device_check.py: a context made by lumol_cuda_create_multi (one host process, one library thread per
    device) against a single-device context: forces / energies / virials, MD across rebuilds, error propagation."""
    import torch

    if torch.cuda.device_count() < ndevices:
        pytest.skip(f"needs {ndevices} GPUs")
    command = [sys.executable, os.path.join(ROOT, "tools", "multi_device_check.py"), str(ndevices)]
    result = subprocess.run(command, capture_output=True, text=True, timeout=1500)
    assert result.returncode == 0, result.stdout[-2000:] + result.stderr[-4000:]
    assert "multi-device check ok" in result.stdout
