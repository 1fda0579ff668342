"""Sharded (multi-GPU) parity: N ranks over NCCL must reproduce the unsharded run BIT FOR BIT
(the chunking of the partial sums is a function of N only).  Needs >= 2 GPUs; the 1-GPU
round-end run skips it.  `gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import traceit_b200 as tr
    import ctypes as C
    n = C.c_int(0)
    tr.lib().tr_device_count(C.byref(n))
    return n.value


@pytest.mark.parametrize("n", [20000, 20001])
def test_two_ranks_equal_unsharded(n):
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29544", os.path.join(ROOT, "tools", "sharded_check.py"), str(n), "4"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_OK" in r.stdout


def test_two_ranks_with_replicated_collisions_equal_unsharded():
    """gravity + collisions on 2 ranks == unsharded, bit for bit."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29545", os.path.join(ROOT, "tools", "sharded_check.py"), "12001", "5", "collide"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_OK" in r.stdout


@pytest.mark.parametrize("mode", ["grow", "growc"])
def test_two_ranks_world_grown_mid_run_equals_unsharded(mode):
    """Adding bodies to an already-stepped sharded world moves the block boundaries: bodies change
    owner.  The new owner must continue from the old owner's velocity / pending force / collider
    centre (all-gathered under the old partition before the switch), i.e. == the unsharded run."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29546", os.path.join(ROOT, "tools", "sharded_check.py"), "9001", "6", mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_OK" in r.stdout
