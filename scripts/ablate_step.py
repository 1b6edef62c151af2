"""GPU probe (development aid): what the headline step would cost without one of its kernels
(results are wrong by construction; only the time is read).  Needs the temporary option 99."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
import torch
from dyt_b200 import GraphedForward, _lib, synthetic

dev = torch.device("cuda:0")
lib = _lib.lib()
model = synthetic.build_vit_b16(dev, num_classes=100, seed=0)
cal = torch.randn(64, 3, 224, 224, generator=torch.Generator().manual_seed(0)).to(dev)
synthetic.calibrate_keep_rate(model, cal, 0.5)
images = torch.randn(256, 3, 224, 224, generator=torch.Generator().manual_seed(0)).to(dev)
ARMS = [("full", 0), ("no_up", 1), ("no_up_no_adapt_read", 9), ("no_attn", 2), ("no_down", 4),
        ("no_merge", 16), ("no_adapter_at_all", 13)]
arms = {}
for name, mask in ARMS:
    assert lib.dyt_configure(99, mask) == 0
    g = GraphedForward(model)
    buf = g.input_buffer(images.shape, images.dtype, dev)
    buf.copy_(images)
    arms[name] = (g, buf)
lib.dyt_configure(99, 0)

def run(g, buf, n=20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        g(buf)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for name, (g, buf) in arms.items():
    run(g, buf, 10)
for rnd in range(4):
    print("round", rnd, "  ".join(f"{name} {run(g, buf):.3f}" for name, (g, buf) in arms.items()), flush=True)
