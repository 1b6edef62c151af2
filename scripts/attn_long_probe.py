"""Times dyt_attn_bias_fwd (tcgen05 flash attention with additive bias) at the segmentation
backbone's shape (16 images x 12 heads x 1025 tokens) with and without bias, checks it against a
torch fp32 reference, and times torch SDPA beside it."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "dynamic-tuning_b200"))
from dyt_b200 import ops

dev = torch.device("cuda:0")


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


for B, N, H in ((16, 1025, 12), (16, 1024, 12), (64, 577, 12), (2, 300, 2)):
    g = torch.Generator().manual_seed(B + N)
    C = 64 * H
    qkv = (torch.randn(B, N, 3 * C, generator=g) * 1.2).half().to(dev)
    bias = (torch.randn(H, N, N, generator=g) * 1.5).to(dev)
    for bb in (None, bias, ops.pad_attn_bias(bias)):
        got = ops.attn_bias(qkv, H, bb)
        q, k, v = qkv.float().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
        s = ((q * 0.125).half().float() @ k.transpose(-1, -2)).half().float()
        if bb is not None:
            s = s + bb
        ref = (torch.softmax(s, -1).half().float() @ v).transpose(1, 2).reshape(B, N, C)
        err = (got.float() - ref).abs().max().item()
        t = timed(lambda: ops.attn_bias(qkv, H, bb))
        flops = 4.0 * B * H * N * N * 64
        print(f"B={B} N={N} H={H} bias={'no' if bb is None else 'pitch %d' % bb.stride(1)}: {t:8.1f} us  {flops / t / 1e6:7.1f} TFLOP/s  max|err|={err:.2e}",
              flush=True)
    qh, kh, vh = qkv.reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
    t = timed(lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh))
    print(f"   torch SDPA (no bias): {t:8.1f} us", flush=True)
    t = timed(lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh, attn_mask=bias.half().unsqueeze(0)))
    print(f"   torch SDPA (fp16 bias mask): {t:8.1f} us", flush=True)
