"""GPU probe: the kernels of one fine-tune step (CUDA-graph replay + all-reduce + optimizer), torch profiler."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from dyt_b200 import synthetic
from dyt_b200.ddp import GradArena, trainable_parameters
from dyt_b200.finetune import FinetuneStep
dev = torch.device("cuda:0")
model = synthetic.build_vit_b16(dev, flavour="train", ffn_num=16, scalar="1.0", seed=0)
g = torch.Generator().manual_seed(0)
images = torch.randn(64, 3, 224, 224, generator=g).to(dev)
targets = torch.randint(0, 100, (64,), generator=g).to(dev)
synthetic.calibrate_keep_rate(model, images, 0.5)
params = trainable_parameters(model)
model.train()
arena = GradArena(params)
step = FinetuneStep(model, torch.optim.AdamW(params, lr=1e-3, weight_decay=0.05), arena, cuda_graph=True)
for _ in range(8): step(images, targets)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(images, targets)
    torch.cuda.synchronize()
c = collections.Counter(); t = collections.Counter()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = e.name.split("(")[0][:90]; c[n] += 1; t[n] += e.device_time
tot = sum(t.values())
for n, k in sorted(c.items(), key=lambda kv: -t[kv[0]])[:45]:
    print(f"{k:4d} {t[n]:9.1f} us {100 * t[n] / tot:5.1f}%  {n}")
print("kernels", sum(c.values()), "sum us", round(tot, 1))
