"""Multi-GPU test of the fine-tune path's only collective (needs >= 2 GPUs: run with
`gpurun --gpus 2 -- python -m pytest tests/test_ddp_gpu.py -m gpu`; skipped on a 1-GPU box)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_averaged_grads_equal_one_gpu_full_batch():
    """1-GPU full-batch gradients == N-GPU all-reduced gradients (NCCL)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29517",
           os.path.join(ROOT, "tests", "ddp_grad_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
