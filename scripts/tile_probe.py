import os, sys, torch
sys.path.insert(0, "dynamic-tuning_b200")
from dyt_b200 import ops, _lib
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
T, C, Hd = 50432, 768, 3072
hid = torch.randn(T, Hd, generator=g).half().to(dev)
w2 = (torch.randn(C, Hd, generator=g) * 0.02).half().to(dev)
b2 = torch.zeros(C).half().to(dev)
xn = torch.randn(T, C, generator=g).half().to(dev)
wp = (torch.randn(C, C, generator=g) * 0.02).half().to(dev)
x32 = torch.randn(T, C, generator=g).to(dev)
outf = torch.empty(T, C, device=dev)
outc = torch.empty(T, C, dtype=torch.half, device=dev)
def t(fn, it=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3
for kept in (25344, 25088, 26000, 24000):
    m_dev = torch.tensor([kept], dtype=torch.int32, device=dev)
    print("fc2 kept", kept, "%.1f us" % t(lambda: ops.linear_f16(hid, w2, b2, m_dev=m_dev, out=outc)))
print("proj resid (no dot) %.1f us" % t(lambda: ops.linear_f16(xn, wp, b2, epilogue=_lib.EPI_BIAS_RESID, resid=x32, out=outf)))
