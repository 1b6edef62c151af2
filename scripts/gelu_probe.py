import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
import torch
from dyt_b200 import ops, _lib
dev = torch.device("cuda:0")
bits = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(torch.float16)
x = bits[torch.isfinite(bits)]
pad = (-x.numel()) % 8
x = torch.cat([x, torch.zeros(pad, dtype=torch.float16)]).reshape(-1, 8)
w = torch.eye(8, dtype=torch.float16)
out, _ = ops.linear_f16(x.to(dev), w.to(dev), torch.zeros(8, dtype=torch.float16, device=dev), epilogue=_lib.EPI_BIAS_GELU)
ident, _ = ops.linear_f16(x.to(dev), w.to(dev), torch.zeros(8, dtype=torch.float16, device=dev))
print("identity exact:", torch.equal(ident.cpu(), x))
ref = torch.nn.functional.gelu(x.float()).half()
got = out.cpu()
d = (got.view(torch.int16).int() - ref.view(torch.int16).int()).abs()
d = torch.where((got == 0) & (ref == 0), torch.zeros_like(d), d)
idx = torch.nonzero(d > 1)
print("n>1:", idx.shape[0], "n>0:", int((d > 0).sum()))
for i in idx[:20]:
    r, c = int(i[0]), int(i[1])
    print(float(x[r, c]), float(got[r, c]), float(ref[r, c]), int(d[r, c]))
