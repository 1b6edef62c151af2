import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dynamic-tuning_b200"))
import torch
from dyt_b200 import ops
dev = torch.device("cuda:0")
qkv = (torch.randn(64, 197, 3 * 768, device=dev) * 1.5).half()
o = ops.attn_varlen(qkv, 12); d_o = torch.randn(64, 197, 768, device=dev).half()
for _ in range(3):
    ops.attn_varlen_bwd(qkv, o, d_o, 12)
torch.cuda.synchronize()
