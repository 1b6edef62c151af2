"""GPU probe (development aid, not a bench): new attention kernel vs the round-1 kernel
(dyt_attn_varlen_fwd_v1, only in -DDYT_AB_BUILD libraries) on the same box: max |diff| and µs."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from dyt_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.lib()
has_v1 = hasattr(lib, "dyt_attn_varlen_fwd_v1")
if has_v1:
    lib.dyt_attn_varlen_fwd_v1.restype = C.c_int
    lib.dyt_attn_varlen_fwd_v1.argtypes = _lib.SIGNATURES["dyt_attn_varlen_fwd"][1]


def v1(qkv, H):
    B, N, C3 = qkv.shape
    out = torch.empty(B * N, C3 // 3, dtype=torch.float16, device=dev)
    st = lib.dyt_attn_varlen_fwd_v1(qkv.data_ptr(), C3, None, B, N, N, B * N, H, 64, out.data_ptr(),
                                    C3 // 3, torch.cuda.current_stream().cuda_stream)
    assert st == 0, lib.dyt_last_error()
    return out.view(B, N, -1)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for (B, H, N) in [(256, 12, 197), (128, 16, 197)]:
    qkv = (torch.randn(B, N, 3 * H * 64, device=dev) * 1.5).half()
    new = ops.attn_varlen(qkv, H)
    torch.cuda.synchronize()
    line = f"B={B} H={H} N={N}: new {timeit(lambda: ops.attn_varlen(qkv, H)):7.1f} us"
    if has_v1:
        old = v1(qkv, H)
        torch.cuda.synchronize()
        d = (new.float() - old.float()).abs().max().item()
        line += f"  v1 {timeit(lambda: v1(qkv, H)):7.1f} us  max|new-v1| {d:.2e}"
    # torch reference on a slice
    q, k, v = qkv[:4].float().view(4, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(4, N, -1)
    line += f"  max|new-sdpa32| {(new[:4].float() - ref).abs().max().item():.2e}"
    print(line, flush=True)

# ---- timeline of CTA 0 (development builds) ----
if hasattr(lib, "dyt_debug_attn_trace") and os.environ.get("ATTN_TRACE", "1") == "1":
    lib.dyt_debug_attn_trace.restype = C.c_int
    lib.dyt_debug_attn_trace.argtypes = [C.c_void_p]
    ROLES, JOBS, EV = 8, 32, 8
    buf = torch.zeros(ROLES * JOBS * EV, dtype=torch.int64, device=dev)
    qkv = (torch.randn(256, 197, 3 * 12 * 64, device=dev) * 1.5).half()
    ops.attn_varlen(qkv, 12); torch.cuda.synchronize()
    assert lib.dyt_debug_attn_trace(buf.data_ptr()) == 0
    ops.attn_varlen(qkv, 12); torch.cuda.synchronize()
    lib.dyt_debug_attn_trace(None)
    t = buf.cpu().view(ROLES, JOBS, EV)
    t0 = int(t[t > 0].min())
    names = ["smx0", "smx1", "mma0", "mma1", "out ", "tma "]
    for r in range(5):
        for j in range(12):
            row = " ".join(f"{(int(v) - t0) if v > 0 else -1:7d}" for v in t[r, j])
            print(f"trace {names[r]} job={j:2d}: {row}")
