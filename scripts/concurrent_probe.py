"""GPU probe (development aid): does running two half-batch forwards concurrently (two CUDA graphs on
two streams, own workspaces) beat one full-batch forward?  Tests whether overlapping the HBM-bound
phases of one half with the tensor-bound phases of the other buys anything under the power cap."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
import torch
from dyt_b200 import synthetic

dev = torch.device("cuda:0")
from dyt_b200 import _lib
if os.environ.get("NOFUSE"):
    _lib.lib().dyt_configure(_lib.OPT_FUSE_ADAPTER_UP, 0)
model = synthetic.build_vit_b16(dev, num_classes=100, seed=0)
cal = torch.randn(64, 3, 224, 224, generator=torch.Generator().manual_seed(0)).to(dev)
synthetic.calibrate_keep_rate(model, cal, 0.5)


def fwd(x):
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        return model(x)


_keep = []   # static inputs must outlive their graphs (torch.cuda.graph empties the cache on entry)


def capture(bs, stream):
    x = torch.randn(bs, 3, 224, 224, generator=torch.Generator().manual_seed(1)).to(dev)
    _keep.append(x)
    warm = torch.cuda.Stream()
    warm.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(warm):
        for _ in range(2):
            fwd(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        out = fwd(x)
    return g, x, out


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
g256, _, o256 = capture(256, s1); torch.cuda.synchronize(); print("captured 256/s1", flush=True)
# SM_LIMIT=74: the graphs that run two at a time are captured with every persistent grid sized for half
# of the SMs, so that the two streams really share the machine
if os.environ.get("SM_LIMIT"):
    assert _lib.lib().dyt_configure(_lib.OPT_SM_LIMIT, int(os.environ["SM_LIMIT"])) == 0
ga, _, oa = capture(128, s1); torch.cuda.synchronize(); print("captured 128/s1", flush=True)
gb, _, ob = capture(128, s2); torch.cuda.synchronize(); print("captured 128/s2", flush=True)
gc, _, oc = capture(256, s2); torch.cuda.synchronize(); print("captured 256/s2", flush=True)
gd, _, od = capture(256, s1); torch.cuda.synchronize(); print("captured 256/s1 (limited)", flush=True)
_lib.lib().dyt_configure(_lib.OPT_SM_LIMIT, 0)


def timed(run, n=20):
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        run()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def one():
    g256.replay()


def two_halves():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        ga.replay()
    with torch.cuda.stream(s2):
        gb.replay()
    cur.wait_stream(s1); cur.wait_stream(s2)


def two_full():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        gd.replay()
    with torch.cuda.stream(s2):
        gc.replay()
    cur.wait_stream(s1); cur.wait_stream(s2)


def halves_serial():
    ga.replay(); gb.replay()


for rnd in range(3):
    print(f"round {rnd}: one x256 {timed(one):.3f} ms | two x128 concurrent {timed(two_halves):.3f} ms | "
          f"two x128 serial {timed(halves_serial):.3f} ms | two x256 concurrent {timed(two_full) / 2:.3f} ms per 256",
          flush=True)
print("outputs equal (limited 256 vs full 256):", bool(torch.equal(od, o256)))
