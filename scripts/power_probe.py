"""GPU probe (development aid): board power and SM clock (nvidia-smi, 20 ms samples) while looping
(a) the whole forward step, (b) the qkv GEMM alone, (c) the scatter-merge alone, (d) attention
alone.  Answers: is the step limited by the 1000 W cap throughout (then time ~ energy)?"""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
import torch
from dyt_b200 import GraphedForward, _lib, ops, synthetic

dev = torch.device("cuda:0")
model = synthetic.build_vit_b16(dev, num_classes=100, seed=0)
cal = torch.randn(64, 3, 224, 224, generator=torch.Generator().manual_seed(0)).to(dev)
synthetic.calibrate_keep_rate(model, cal, 0.5)
images = torch.randn(256, 3, 224, 224, generator=torch.Generator().manual_seed(0)).to(dev)
g = GraphedForward(model)
buf = g.input_buffer(images.shape, images.dtype, dev); buf.copy_(images)
T = 256 * 197
blk = model.blocks[0]
xn = torch.randn(T, 768, device=dev, dtype=torch.float16)
w = blk.attn.qkv.weight.detach().half().contiguous(); b = blk.attn.qkv.bias.detach().half().contiguous()
out_qkv = torch.empty(T, 2304, device=dev, dtype=torch.float16)
x32 = torch.randn(256, 197, 768, device=dev)
ad = torch.randn(256, 197, 768, device=dev, dtype=torch.float16)
d = ops.dispatch(x32, blk.mlp_token_select.mlp_head.weight.detach(), blk.mlp_token_select.mlp_head.bias.detach(),
                 ln_w=blk.norm2.weight.detach().float(), ln_b=blk.norm2.bias.detach().float())
qkv3 = torch.randn(256, 197, 2304, device=dev, dtype=torch.float16)
ln = (blk.norm2.weight.detach().float(), blk.norm2.bias.detach().float())

def sample(fn, seconds=2.0):
    fd, path = tempfile.mkstemp(suffix=".csv"); os.close(fd)
    proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                             "-lms", "20", "-i", "0"], stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    time.sleep(0.3)
    t0 = time.time(); n = 0
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    while time.time() - t0 < seconds:
        for _ in range(10):
            fn()
        n += 10
        torch.cuda.synchronize()
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / n
    proc.terminate(); proc.wait()
    rows = [l.split(",") for l in open(path) if "," in l]
    rows = rows[len(rows) // 3:]           # steady part
    clk = sorted(float(r[0]) for r in rows); pw = sorted(float(r[1]) for r in rows)
    os.unlink(path)
    return ms, clk[len(clk) // 2], pw[len(pw) // 2], pw[-1]

for name, fn in (("full step (graph)", lambda: g(buf)),
                 ("qkv GEMM only", lambda: ops.linear_f16(xn, w, b, out=out_qkv)),
                 ("scatter-merge only", lambda: ops.scatter_merge(x32, ad, ad.reshape(T, 768), d["token_pos"], next_ln=ln)),
                 ("attention only", lambda: ops.attn_varlen(qkv3, 12)),
                 ("full step (graph) again", lambda: g(buf))):
    ms, clk, pw, pmax = sample(fn)
    print(f"{name:26s}: {ms*1e3:9.1f} us/iter  sm clock median {clk:6.0f} MHz  power median {pw:6.0f} W max {pmax:6.0f} W", flush=True)
