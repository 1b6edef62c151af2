"""GPU probe (development aid): the fused adapter-up + scatter-merge kernel against the two launches
it replaces, at the headline size (256 x 197 tokens, C = 768, K = 64, half of the rows kept)."""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=30):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n * 1e3


def unfused():
    adapt, _ = ops.linear_f16(down.reshape(B * N, K), up_w, up_b, scale=0.1)
    return ops.scatter_merge(x1, adapt.reshape(B, N, C), mlp, pos, next_ln=(lw, lb))


def fused():
    return ops.merge_up(down, up_w, up_b, 0.1, x1, mlp, pos, next_ln=(lw, lb))


down_w = (torch.randn(K, C, generator=g) * 0.03).half().to(dev)
down_b = (torch.randn(K, generator=g) * 0.1).half().to(dev)
x1h = x1.half()
from dyt_b200 import _lib
def down_then_fused():
    dn, _ = ops.linear_f16(x1h.reshape(B * N, C), down_w, down_b, epilogue=_lib.EPI_BIAS_RELU)
    return ops.merge_up(dn.reshape(B, N, K), up_w, up_b, 0.1, x1, mlp, pos, next_ln=(lw, lb))
print("down GEMM + fused merge_up us:", round(timed(down_then_fused), 1))
print("whole adapter branch fused (adapter_merge) us:", round(timed(lambda: ops.adapter_merge(down_w, down_b, up_w, up_b, 0.1, x1, mlp, pos, next_ln=(lw, lb))), 1))
print("unfused (up GEMM + scatter_merge) us:", round(timed(unfused), 1))
print("fused merge_up us:", round(timed(fused), 1))
print("fused, no LayerNorm us:", round(timed(lambda: ops.merge_up(down, up_w, up_b, 0.1, x1, mlp, pos)), 1))
bytes_alg = B * N * C * (4 + 4 + 2) + idx.numel() * C * 2 + B * N * K * 2
print("algorithmic MB:", bytes_alg / 1e6)
