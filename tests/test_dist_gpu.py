"""Direction-split path over NCCL on >= 2 GPUs of one box (skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_direction_split_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    worker = os.path.join(os.path.dirname(__file__), "dist_gpu_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29577", worker]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=170)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "MISMATCH" not in r.stdout
