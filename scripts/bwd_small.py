import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
import torch
from dyt_b200 import ops
dev = torch.device("cuda:0")
B, N, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
qkv = (torch.randn(B, N, 3 * H * 64, device=dev) * 1.5).half()
o = ops.attn_varlen(qkv, H)
d_o = torch.randn(B, N, H * 64, device=dev).half()
g = ops.attn_varlen_bwd(qkv, o, d_o, H)
torch.cuda.synchronize()
print("ok", float(g.float().abs().mean()))
