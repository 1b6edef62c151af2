import os, sys
sys.path.insert(0, "/root/repo/dynamic-tuning_b200")
import torch
from dyt_b200 import ops
dev = torch.device("cuda:0")
qkv = (torch.randn(256, 197, 3 * 12 * 64, device=dev) * 1.5).half()
for _ in range(3):
    ops.attn_varlen(qkv, 12)
torch.cuda.synchronize()
