"""Time one fine-tuning step (BASELINE configs[2] shape: ViT-B/16, 64 images per GPU, ffn_num 16,
adapter scale 1) and print the kernel table of one step (torch profiler, CUDA activities)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "dynamic-tuning_b200"), ROOT]

from dyt_b200 import synthetic  # noqa: E402
from dyt_b200.ddp import GradArena, trainable_parameters  # noqa: E402
from dyt_b200.finetune import FinetuneStep  # noqa: E402

B = int(os.environ.get("TRAIN_B", "64"))
if "TRAIN_SPARSE_MIN" in os.environ:
    from dyt_b200 import train as _train
    _train.SPARSE_STUDENT_MIN_TOKENS = int(os.environ["TRAIN_SPARSE_MIN"])
dev = torch.device("cuda:0")
model = synthetic.build_vit_b16(dev, flavour="train", ffn_num=16, scalar="1.0")
g = torch.Generator().manual_seed(0)
img = torch.randn(B, 3, 224, 224, generator=g).to(dev)
tgt = torch.randint(0, 100, (B,), generator=g).to(dev)
synthetic.calibrate_keep_rate(model, img, 0.5)
params = trainable_parameters(model)
model.train()
arena = GradArena(params)
opt = torch.optim.AdamW(params, lr=1e-3)
step = FinetuneStep(model, opt, arena, cuda_graph=os.environ.get("TRAIN_GRAPH", "0") == "1")
for _ in range(6):
    loss = step(img, tgt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    loss = step(img, tgt)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"finetune step bs{B}: {ms:.2f} ms  {B / ms * 1e3:.1f} img/s  loss {loss.item():.4f} "
      f"arena {arena.nbytes / 1e6:.2f} MB / {len(arena.params)} tensors")
if os.environ.get("TRAIN_NCU", "0") == "1":
    # one eager step inside a cudaProfilerStart/Stop window:
    #   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    #       --log-file gpurun_out/finetune_launches.csv python scripts/train_probe.py   (TRAIN_NCU=1)
    eager = FinetuneStep(model, opt, arena)
    eager(img, tgt)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    eager(img, tgt)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    sys.exit(0)
if os.environ.get("TRAIN_PROF", "1") == "1":
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step(img, tgt)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=32, max_name_column_width=60))
