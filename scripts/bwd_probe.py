"""Time the backward kernels alone at the fine-tune shape (64 images x 197 tokens, ViT-B)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "dynamic-tuning_b200"), ROOT]
from dyt_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, N, H, C = int(os.environ.get("BWD_B", "64")), 197, 12, 768
T = B * N
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, N, 3 * C, generator=g).half().to(dev)
d_o = torch.randn(B, N, C, generator=g).half().to(dev)
o = ops.attn_varlen(qkv, H)


def timeit(name, fn, flops=0.0, bytes_=0.0, iters=10):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    extra = ""
    if flops:
        extra += f"  {flops / us * 1e-6:8.1f} TFLOP/s"
    if bytes_:
        extra += f"  {bytes_ / us * 1e-3:8.1f} GB/s"
    print(f"{name:34s} {us:9.1f} us{extra}")


attn_flops = B * H * (2 * N * N * 64) * 5   # S, dP, dV, dQ, dK
timeit("attn_bwd", lambda: ops.attn_varlen_bwd(qkv, o, d_o, H), flops=attn_flops)
for bott in (16, 64):
    g16 = torch.randn(T, C, generator=g).half().to(dev)
    hd = torch.randn(T, bott, generator=g).half().to(dev)
    timeit(f"wgrad up   [{C}x{bott}]", lambda: ops.wgrad_f16(g16, hd), flops=2.0 * T * C * bott,
           bytes_=T * (C + bott) * 2)
    timeit(f"wgrad down [{bott}x{C}]", lambda: ops.wgrad_f16(hd, g16), flops=2.0 * T * C * bott,
           bytes_=T * (C + bott) * 2)
x = torch.randn(T, C, generator=g).to(dev)
w = torch.ones(C, device=dev)
timeit("layernorm_bwd (+resid)", lambda: ops.layernorm_bwd(g16, x, w, 1e-6, resid=x),
       bytes_=T * C * (2 + 4 + 4 + 4))
