"""GPU probe: four-stream attention kernel (attn_split.cu) vs the two-stream one, 256 x 12 x 197."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
import torch
from dyt_b200 import ops, _lib
dev = torch.device("cuda:0")
lib = _lib.lib()
B, H, N = int(os.environ.get("B", 256)), 12, int(os.environ.get("N", 197))
qkv = torch.randn(B, N, 3 * H * 64, device=dev, dtype=torch.float16)
def t(n=30):
    for _ in range(3): ops.attn_varlen(qkv, H)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): ops.attn_varlen(qkv, H)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
lib.dyt_configure(_lib.OPT_ATTN_SPLIT, 0); o0 = ops.attn_varlen(qkv, H).clone(); t0 = t()
lib.dyt_configure(_lib.OPT_ATTN_SPLIT, 1); o1 = ops.attn_varlen(qkv, H).clone(); t1 = t()
q, k, v = qkv.float().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B, N, H * 64)
print(f"two-stream {t0:.1f} us  err {float((o0.float() - ref).abs().max()):.2e} | four-stream {t1:.1f} us  err {float((o1.float() - ref).abs().max()):.2e} | max diff {float((o0.float() - o1.float()).abs().max()):.2e}")
