"""One ViT-B/16 DyT forward (bs256, r~0.5) inside a cudaProfilerStart/Stop window, for ncu:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python profiles/prof_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:'gemm_tn|attn_fwd|dispatch_kernel|scatter_merge|layernorm' -c 10 \
      -o gpurun_out/prof python profiles/prof_step.py --layers 1
Numbers printed under ncu are never bench values."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
import torch  # noqa: E402
from dyt_b200 import engine, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--layers", type=int, default=12)
ap.add_argument("--tile-order", type=int, default=None, help="dyt_configure(DYT_OPT_TILE_ORDER, v)")
args = ap.parse_args()
dev = torch.device("cuda:0")
if args.tile_order is not None:
    from dyt_b200 import _lib
    assert _lib.lib().dyt_configure(_lib.OPT_TILE_ORDER, args.tile_order) == 0
model = synthetic.build_vit_b16(dev, seed=0)
cal = torch.randn(32, 3, 224, 224, generator=torch.Generator().manual_seed(0)).to(dev)
print("calibrated keep rate", synthetic.calibrate_keep_rate(model, cal, 0.5))
img = torch.randn(args.batch, 3, 224, 224, generator=torch.Generator().manual_seed(1)).to(dev)
blocks = list(model.blocks)[:args.layers]
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    for _ in range(3):
        model(img)
    x = model._embed(img).float()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    if args.layers == 12:
        model(img)
    else:
        engine.run_blocks(x, blocks)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
