"""Time the long-sequence attention kernel at the segmentation size (16 images x 12 heads x 1025)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "dynamic-tuning_b200"), ROOT]
from dyt_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, N, H = int(os.environ.get("AB_B", "16")), 1025, 12
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, N, 3 * 64 * H, generator=g).half().to(dev)
bias = torch.randn(H, N, N, generator=g).to(dev) if os.environ.get("AB_BIAS", "0") == "1" else None
for _ in range(3):
    ops.attn_bias(qkv, H, bias)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    ops.attn_bias(qkv, H, bias)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 10 * 1e3
fl = 4.0 * B * H * N * N * 64
print(f"attn_bias B={B} N={N} H={H} bias={bias is not None}: {us:.1f} us  {fl / us * 1e-6:.1f} TFLOP/s")
