"""GPU probe: two graphs of different batch sizes captured on the same stream (shared workspace)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
import torch
from dyt_b200 import synthetic, GraphedForward
dev = torch.device("cuda:0")
model = synthetic.build_vit_b16(dev, num_classes=100, seed=0)
cal = torch.randn(64, 3, 224, 224, generator=torch.Generator().manual_seed(0)).to(dev)
synthetic.calibrate_keep_rate(model, cal, 0.5)
x256 = torch.randn(256, 3, 224, 224, generator=torch.Generator().manual_seed(1)).to(dev)
x128 = x256[:128].contiguous()
which = sys.argv[1]
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    ref128 = model(x128).clone()
torch.cuda.synchronize(); print("eager 128 ok", flush=True)
gm = GraphedForward(model)
if which == "big_first":
    o = gm(x256); torch.cuda.synchronize(); print("graph 256 ok", flush=True)
    o = gm(x128); torch.cuda.synchronize(); print("graph 128 ok", bool(torch.equal(o, ref128)), flush=True)
    o = gm(x256); torch.cuda.synchronize(); print("graph 256 again ok", flush=True)
    o = gm(x128); torch.cuda.synchronize(); print("graph 128 again ok", bool(torch.equal(o, ref128)), flush=True)
else:
    o = gm(x128); torch.cuda.synchronize(); print("graph 128 ok", bool(torch.equal(o, ref128)), flush=True)
    o = gm(x256); torch.cuda.synchronize(); print("graph 256 ok", flush=True)
    o = gm(x128); torch.cuda.synchronize(); print("graph 128 again ok", bool(torch.equal(o, ref128)), flush=True)
