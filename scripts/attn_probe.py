"""GPU probe: time the attention kernel alone at the bench shape; DYT_ATTN_TRACE=1 prints CTA 0's
clock64 timeline (debug aid, not a bench)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
import torch
from dyt_b200 import ops
dev = torch.device("cuda:0")
B, H, N = int(os.environ.get("ATTN_B", "256")), int(os.environ.get("ATTN_H", "12")), int(os.environ.get("ATTN_N", "197"))
qkv = torch.randn(B, N, 3 * H * 64, device=dev, dtype=torch.float16)
def run(n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): ops.attn_varlen(qkv, H)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
if os.environ.get("DYT_ATTN_TRACE"):
    ops.attn_varlen(qkv, H); torch.cuda.synchronize()
else:
    run(3)
    print("attn us", run(20))
