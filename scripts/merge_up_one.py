"""One launch sequence of the fused merge kernel at the headline size (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
import torch
from dyt_b200 import ops
dev = torch.device("cuda:0")
B, N, C, K = 256, 197, 768, 64
g = torch.Generator().manual_seed(0)
x1 = torch.randn(B, N, C, generator=g).to(dev)
down = torch.relu(torch.randn(B, N, K, generator=g)).half().to(dev)
up_w = (torch.randn(C, K, generator=g) * 0.05).half().to(dev)
up_b = (torch.randn(C, generator=g) * 0.1).half().to(dev)
mask = torch.rand(B * N, generator=g) > 0.5
idx = mask.nonzero().flatten()
mlp = torch.randn(idx.numel(), C, generator=g).half().to(dev)
pos = torch.full((B * N,), -1, dtype=torch.int32)
pos[idx] = torch.arange(idx.numel(), dtype=torch.int32)
pos = pos.to(dev)
lw, lb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
for _ in range(4):
    ops.merge_up(down, up_w, up_b, 0.1, x1, mlp, pos, next_ln=(lw, lb))
torch.cuda.synchronize()
