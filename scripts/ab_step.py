"""GPU probe (development aid, not the bench): same-box A/B of library options on the headline step
(ViT-B/16 DyT bs256, CUDA-graph replay): interleaved rounds so clock drift hits both arms."""
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

arms = {}
OPT = getattr(_lib, os.environ.get("AB_OPT", "OPT_GEMM_TAIL_SPLIT"))
# AB_VALUES="0,7,3": one arm per option value (default: off / on)
VALUES = [int(v) for v in os.environ.get("AB_VALUES", "0,1").split(",")]
for name, pdl in [((f"opt_{v}" if VALUES != [0, 1] else ("opt_off", "opt_on")[v]), v) for v in VALUES]:
    assert lib.dyt_configure(OPT, pdl) == 0
    g = GraphedForward(model)
    buf = g.input_buffer(images.shape, images.dtype, dev)
    buf.copy_(images)
    arms[name] = (g, buf)
lib.dyt_configure(OPT, {_lib.OPT_FUSE_ADAPTER_DOWN: 0, _lib.OPT_TILE_ORDER: 7, _lib.OPT_SIDE_PLAN: 0}.get(OPT, 1))   # back to the default

ref = None
for name, (g, buf) in arms.items():
    out = g(buf).clone()
    if ref is None:
        ref = out
    print(name, "logits equal to first arm:", bool(torch.equal(out, ref)))


def run(g, buf, n=int(os.environ.get("AB_N", 20))):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        g(buf)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for name, (g, buf) in arms.items():
    run(g, buf, 10)
for rnd in range(int(os.environ.get("AB_ROUNDS", 4))):
    print("round", rnd, "  ".join(f"{name} {run(g, buf):.3f} ms" for name, (g, buf) in arms.items()), flush=True)
