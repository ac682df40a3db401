"""Row-partitioned solver on >= 2 GPUs of one box: runs tests/dist_worker.py under torchrun (one
process per GPU) and in its single-process form (folp_create_multi). Skipped on single-GPU boxes."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_row_partitioned_parity(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
           os.path.join(HERE, "dist_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    sys.stdout.write(out.stdout[-6000:])
    sys.stderr.write(out.stderr[-6000:])
    assert out.returncode == 0
    assert "ALL OK" in out.stdout


@pytest.mark.parametrize("world", [2, 4])
def test_single_process_multi_gpu_parity(world):
    """folp_create_multi: one process, one call, `world` devices."""
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, os.path.join(HERE, "dist_worker.py"), "--single-process", str(world)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    sys.stdout.write(out.stdout[-6000:])
    sys.stderr.write(out.stderr[-6000:])
    assert out.returncode == 0
    assert "ALL OK" in out.stdout


def test_peer_timeout_is_configurable_and_surfaces_as_error():
    """tests/dist_timeout_worker.py: a stalled rank is tolerated up to FOLP_P2P_TIMEOUT_MS, then both
    ranks get an error instead of a hung device."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(HERE, "dist_timeout_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(out.stdout[-4000:])
    sys.stderr.write(out.stderr[-4000:])
    assert out.returncode == 0
    assert out.stdout.count("TIMEOUT OK") == 2
