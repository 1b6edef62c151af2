"""GPU probe: eager forward vs CUDA-graph replay of the same forward (how much is launch gaps?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
import torch
from dyt_b200 import synthetic
dev = torch.device("cuda:0")
model = synthetic.build_vit_b16(dev, seed=0)
cal = torch.randn(32, 3, 224, 224, generator=torch.Generator().manual_seed(0)).to(dev)
synthetic.calibrate_keep_rate(model, cal, 0.5)
img = torch.randn(256, 3, 224, 224, generator=torch.Generator().manual_seed(1)).to(dev)
def fwd():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        return model(img)
def t(fn, n=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print("eager ms", t(fwd))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): fwd()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        out = fwd()
    print("graph ms", t(g.replay))
except Exception as e:
    print("graph capture failed:", repr(e)[:300])
