import os, sys, torch
sys.path.insert(0, "dynamic-tuning_b200")
from dyt_b200 import ops
dev = torch.device("cuda:0")
def timed(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
for B, N, H in ((256, 197, 12), (256, 150, 12), (256, 256, 12)):
    qkv = (torch.randn(B, N, 3 * 64 * H) * 1.2).half().to(dev)
    a = ops.attn_bias(qkv, H, None); v = ops.attn_varlen(qkv, H)
    print(B, N, "maxdiff", (a.float() - v.float()).abs().max().item(), "long kernel %.1f us" % timed(lambda: ops.attn_bias(qkv, H, None)), "varlen kernel %.1f us" % timed(lambda: ops.attn_varlen(qkv, H)))
