"""GPU probe: the kernels one CUDA-graph replay of the headline forward launches (torch profiler)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from dyt_b200 import synthetic, GraphedForward
dev = torch.device("cuda:0")
from dyt_b200 import _lib
if os.environ.get("ATTN_SPLIT"):
    _lib.lib().dyt_configure(_lib.OPT_ATTN_SPLIT, 1)
model = synthetic.build_vit_b16(dev, num_classes=100, seed=0)
cal = torch.randn(64, 3, 224, 224, generator=torch.Generator().manual_seed(0)).to(dev)
synthetic.calibrate_keep_rate(model, cal, 0.5)
x = torch.randn(256, 3, 224, 224, generator=torch.Generator().manual_seed(1)).to(dev)
gm = GraphedForward(model)
buf = gm.input_buffer(x.shape, x.dtype, dev); buf.copy_(x)
for _ in range(3): gm.replay(x.shape, x.dtype, dev)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gm.replay(x.shape, x.dtype, dev)
    torch.cuda.synchronize()
c = collections.Counter(); t = collections.Counter()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = e.name.split("(")[0][:70]; c[n] += 1; t[n] += e.device_time
tot = sum(t.values())
for n, k in sorted(c.items(), key=lambda kv: -t[kv[0]]):
    print(f"{k:4d} {t[n]:9.1f} us {100 * t[n] / tot:5.1f}%  {n}")
print("kernels", sum(c.values()), "sum us", round(tot, 1))
