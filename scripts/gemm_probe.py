"""GPU probe: time the tcgen05 GEMM at the bench shapes (qkv / proj / fc1 / fc2) -- debug aid."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
import torch
from dyt_b200 import ops, _lib
dev = torch.device("cuda:0")
T, C, HID = 256 * 197, 768, 3072
K_KEPT = 25300
which = sys.argv[1] if len(sys.argv) > 1 else "all"
def t(fn, n=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
h = torch.float16
x = torch.randn(T, C, device=dev, dtype=h)
hid = torch.randn(T, HID, device=dev, dtype=h)
res = torch.randn(T, C, device=dev, dtype=torch.float32)
w_qkv = torch.randn(3 * C, C, device=dev, dtype=h) * 0.02; b_qkv = torch.zeros(3 * C, device=dev, dtype=h)
w_c = torch.randn(C, C, device=dev, dtype=h) * 0.02; b_c = torch.zeros(C, device=dev, dtype=h)
w_fc1 = torch.randn(HID, C, device=dev, dtype=h) * 0.02; b_fc1 = torch.zeros(HID, device=dev, dtype=h)
w_fc2 = torch.randn(C, HID, device=dev, dtype=h) * 0.02
o_qkv = torch.empty(T, 3 * C, device=dev, dtype=h); o_c = torch.empty(T, C, device=dev, dtype=h)
o_f = torch.empty(T, C, device=dev, dtype=torch.float32); o_h = torch.empty(T, HID, device=dev, dtype=h)
m_dev = torch.tensor([K_KEPT], dtype=torch.int32, device=dev)
cases = {
 "qkv": (lambda: ops.linear_f16(x, w_qkv, b_qkv, out=o_qkv), 2.0 * T * 3 * C * C),
 "proj": (lambda: ops.linear_f16(x, w_c, b_c, epilogue=_lib.EPI_BIAS_RESID, resid=res, out=o_f, want_f16_copy=False), 2.0 * T * C * C),
 "fc1": (lambda: ops.linear_f16(x, w_fc1, b_fc1, epilogue=_lib.EPI_BIAS_GELU, m_dev=m_dev, out=o_h), 2.0 * K_KEPT * HID * C),
 "fc2": (lambda: ops.linear_f16(hid, w_fc2, b_c, m_dev=m_dev, out=o_c), 2.0 * K_KEPT * HID * C),
}
for name, (fn, fl) in cases.items():
    if which not in ("all", name): continue
    us = t(fn)
    print(f"{name}: {us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s")
