"""GPU probe: attention backward kernel time at the fine-tune shape (64 x 12 heads x 197)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
import torch
from dyt_b200 import ops
dev = torch.device("cuda:0")
for B in (64, 256):
    qkv = (torch.randn(B, 197, 3 * 768, device=dev) * 1.5).half()
    o = ops.attn_varlen(qkv, 12)
    d_o = torch.randn(B, 197, 768, device=dev).half()
    for _ in range(3):
        ops.attn_varlen_bwd(qkv, o, d_o, 12)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20):
        ops.attn_varlen_bwd(qkv, o, d_o, 12)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 20 * 1e3
    fl = 5 * 2 * 197 * 197 * 64 * 12 * B
    print(f"attn bwd B={B}: {us:.1f} us  {fl / us / 1e6:.1f} TFLOP/s", flush=True)

import ctypes as C
from dyt_b200 import _lib
lib = _lib.lib()
if hasattr(lib, "dyt_debug_bwd_trace"):
    lib.dyt_debug_bwd_trace.restype = C.c_int; lib.dyt_debug_bwd_trace.argtypes = [C.c_void_p]
    buf = torch.zeros(4 * 16 * 8, dtype=torch.int64, device=dev)
    qkv = (torch.randn(64, 197, 3 * 768, device=dev) * 1.5).half()
    o = ops.attn_varlen(qkv, 12); d_o = torch.randn(64, 197, 768, device=dev).half()
    ops.attn_varlen_bwd(qkv, o, d_o, 12); torch.cuda.synchronize()
    lib.dyt_debug_bwd_trace(buf.data_ptr())
    ops.attn_varlen_bwd(qkv, o, d_o, 12); torch.cuda.synchronize()
    lib.dyt_debug_bwd_trace(None)
    t = buf.cpu().view(4, 16, 8); t0 = int(t[t > 0].min())
    for s_, name in ((0, "mma "), (1, "math"), (2, "smx "), (3, "ds  ")):
        for j in range(4):
            print(f"trace {name} tile={j}: " + " ".join(f"{(int(v) - t0) if v > 0 else -1:7d}" for v in t[s_, j][:6]))
