"""GPU probe: effect of the last-round sub-tiling (DYT_OPT_GEMM_TAIL_SPLIT) on fc1 / fc2 at kept-row
counts around the bench's (leftover tiles 0 .. few), same box, interleaved; also checks bit-equality."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
import torch
from dyt_b200 import ops, _lib
lib = _lib.lib()
dev = torch.device("cuda:0")
T, C, HID = 256 * 197, 768, 3072
h = torch.float16
def t(fn, n=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
x = torch.randn(T, C, device=dev, dtype=h)
hid = torch.randn(T, HID, device=dev, dtype=h) * 0.3
w_fc1 = torch.randn(HID, C, device=dev, dtype=h) * 0.02; b_fc1 = torch.randn(HID, device=dev, dtype=h) * 0.1
w_fc2 = torch.randn(C, HID, device=dev, dtype=h) * 0.02; b_c = torch.randn(C, device=dev, dtype=h) * 0.1
w_qkv = torch.randn(3 * C, C, device=dev, dtype=h) * 0.02; b_qkv = torch.randn(3 * C, device=dev, dtype=h) * 0.1
for kept in (25305, 24800, 25100, 25700, 26500):
    m_dev = torch.tensor([kept], dtype=torch.int32, device=dev)
    outs = {}
    line = f"kept {kept}: "
    for name, fn_mk in (("fc1", lambda o: (lambda: ops.linear_f16(x, w_fc1, b_fc1, epilogue=_lib.EPI_BIAS_GELU, m_dev=m_dev, out=o))),
                        ("fc2", lambda o: (lambda: ops.linear_f16(hid, w_fc2, b_c, m_dev=m_dev, out=o)))):
        res = []
        for opt in (0, 1):
            lib.dyt_configure(_lib.OPT_GEMM_TAIL_SPLIT, opt)
            o = torch.zeros(T, HID if name == "fc1" else C, device=dev, dtype=h)
            fn = fn_mk(o)
            us = t(fn)
            res.append((us, o[:kept].clone()))
        same = torch.equal(res[0][1], res[1][1])
        line += f"{name} off {res[0][0]:6.1f} on {res[1][0]:6.1f} us equal {same}   "
    print(line, flush=True)
lib.dyt_configure(_lib.OPT_GEMM_TAIL_SPLIT, 1)
o = torch.zeros(T, 3 * C, device=dev, dtype=h)
print("qkv", t(lambda: ops.linear_f16(x, w_qkv, b_qkv, out=o)))
