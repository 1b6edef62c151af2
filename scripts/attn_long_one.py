import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "dynamic-tuning_b200"))
from dyt_b200 import ops
B, N, H = 16, int(os.environ.get("LA_N", "1025")), 12
wb = os.environ.get("LA_BIAS", "0") == "1"
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
qkv = (torch.randn(B, N, 3 * 64 * H, generator=g) * 1.2).half().to(dev)
bias = (torch.randn(H, N, N, generator=g) * 1.5).to(dev) if wb else None
for _ in range(3):
    ops.attn_bias(qkv, H, bias)
torch.cuda.synchronize()
