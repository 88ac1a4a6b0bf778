"""NCCL transport of the slab path: needs >= 2 GPUs on the box (skipped on a 1-GPU box, where tests/test_dist_gpu.py
covers the same phases through the LOCAL transport).  Launches torchrun with one rank per GPU."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _gpus() -> int:
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_slab_ranks_match_single_device(gpu, world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29510 + world), str(ROOT / "tests" / "nccl_slab_worker.py"), "6"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "NCCL_SLAB_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
