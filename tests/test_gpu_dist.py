"""Row-partitioned solver over NCCL on >= 2 GPUs of one box: runs tests/dist_worker.py under
torchrun (one process per GPU). Skipped on single-GPU boxes."""
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
