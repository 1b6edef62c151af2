import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "dynamic-tuning_b200"))
from dyt_b200 import ops
dev = torch.device("cuda:0")
CASES = ((16, 1024, 12, True), (2, 577, 2, True), (16, 1025, 12, False), (16, 1025, 12, True), (3, 5, 2, True), (40, 300, 4, True))
for B, N, H, wb in CASES:
    g = torch.Generator().manual_seed(B + N)
    C = 64 * H
    qkv = (torch.randn(B, N, 3 * C, generator=g) * 1.2).half().to(dev)
    bias = (torch.randn(H, N, N, generator=g) * 1.5).to(dev) if wb else None
    got = ops.attn_bias(qkv, H, bias)
    q, k, v = qkv.float().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
    s = ((q * 0.125).half().float() @ k.transpose(-1, -2)).half().float()
    if bias is not None:
        s = s + bias
    ref = (torch.softmax(s, -1).half().float() @ v)          # B H N 64
    e = (got.float().reshape(B, N, H, 64).permute(0, 2, 1, 3) - ref).abs()
    print(f"B={B} N={N} H={H} bias={wb}: max {e.max().item():.3e}")
    bad = e > 5e-3
    print("  bad elements", int(bad.sum()), "of", bad.numel())
    if bad.any():
        print("  per (image, head) bad counts:", bad.sum(dim=(2, 3)).flatten().tolist()[:32])
        rows = bad.any(dim=3)                          # B H N
        qt = torch.arange(N, device=dev) // 128
        print("  bad rows per q tile:", [int(rows[..., qt == t].sum()) for t in range((N + 127) // 128)])
        lane_q = (torch.arange(N, device=dev) % 128) // 32
        print("  bad rows per lane quarter:", [int(rows[..., lane_q == t].sum()) for t in range(4)])
        print("  bad per column part:", [int(bad[..., 16 * t:16 * t + 16].sum()) for t in range(4)])
        b_, h_, n_ = [int(x[0]) for x in torch.nonzero(rows, as_tuple=True)]
        print("  first bad row", b_, h_, n_, "got", got.reshape(B, N, H, 64)[b_, n_, h_, :8].tolist(), "ref", ref[b_, h_, n_, :8].tolist())
