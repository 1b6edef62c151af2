#!/usr/bin/env python
"""Benchmark of the DyT token-dispatched ViT forward (BASELINE.json metric: images/s, ViT-B/16 DyT,
r = 0.5, batch 256 per GPU, 224x224 synthetic images; reference protocol speed.py:247-275).

  python bench.py --gpus N --steps K --warmup W            # our sm_100a path (one rank per GPU)
  python bench.py --impl reference --gpus N ...            # reference algorithm on the host CPU

A step = one forward of the whole model over one batch.  `value` is device-timed (CUDA events,
inputs resident in HBM); `e2e` goes through the public drop-in API (model(images) under autocast)
with pinned host images copied H2D and the logits read back D2H inside the timed region.
Multi-GPU: images shard over ranks with no data-path collective (weak scaling, 256 images / GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))

import torch  # noqa: E402

METRIC = "images/sec ViT-B/16 DyT r=0.5 bs256"
BATCH = 256
RATE = 0.5
N_TOK, C_DIM, DEPTH, HEADS, HIDDEN, BOTTLENECK, NUM_CLASSES = 197, 768, 12, 12, 3072, 64, 100


def flops_per_image(kept_tokens: float) -> float:
    """Algorithmic FLOPs (2 x MAC) of one ViT-B/16 DyT forward, SURVEY.md section 8d."""
    n, c = N_TOK, C_DIM
    per_block = (2 * n * c * 3 * c + 4 * n * n * c + 2 * n * c * c + 2 * (n - 1) * c +
                 4 * n * c * BOTTLENECK + 4 * kept_tokens * c * HIDDEN)
    return DEPTH * per_block + 2 * 196 * c * 768 + 2 * c * NUM_CLASSES


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback")


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "50", "-i", str(self.idx)], stdout=open(self.path, "w"),
                stderr=subprocess.DEVNULL)
            t0 = time.time()      # wait for the first sample so the loop is live before the load
            while time.time() - t0 < 5.0 and os.path.getsize(self.path) == 0:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, pw, mx, reasons = [], [], None, set()
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1]))
                mx = float(f[2])
                try:
                    pw.append(float(f[3]))
                except ValueError:
                    pass
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        sm.sort()
        pw.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx,
                    reasons=sorted(reasons), samples=len(sm),
                    # board power under load: the samples cover idle set-up phases too, so report the
                    # upper quartile and the maximum (cap: 1000 W)
                    power_w=(pw[(3 * len(pw)) // 4] if pw else None),
                    power_w_max=(pw[-1] if pw else None))


def dist_setup(n_gpus: int):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL prints its version banner to stdout at the
        # VERSION and WARN levels, so run those at NONE (INFO / TRACE are left to whoever asked)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG", None)
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend="gloo")
    return world, rank, local


def bind_to_gpu_numa_node(local: int):
    """Multi-rank runs: keep this rank (and the pinned host buffers it is about to allocate) on the
    NUMA node its GPU hangs off, so that the H2D copies of the end-to-end loop do not cross the
    socket interconnect.  Best effort: returns the node, or None when the topology is not exposed."""
    try:
        bdf = torch.cuda.get_device_properties(local).pci_bus_id  # torch >= 2.5
    except Exception:
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
            bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
            if isinstance(bdf, bytes):
                bdf = bdf.decode()
        except Exception:
            return None
    try:
        bdf = str(bdf).lower()
        if len(bdf.split(":")[0]) == 8:      # 00000000:1B:00.0 -> 0000:1b:00.0
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def shard_images(total: int, world: int, rank: int):
    """Contiguous image shard of rank `rank`: images are independent (no cross-image state), the
    reference shards the same way (strided Subset, speed.py:188).  Returns (start, count)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def max_over_ranks(seconds: float, world: int, device) -> float:
    if world == 1:
        return seconds
    import torch.distributed as dist
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world: int):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU oracle (restatement of the reference forward, pinned to the
# reference's own outputs by tests/test_oracle_golden.py) on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_reference_run(state_dict, steps: int, warmup: int, sample_batch: int, seed: int = 0):
    """The reference's own CPU forward of the path on a bounded sample: models/model_speed_test.py
    UNMODIFIED (oracle/_ref/, made by oracle/build_ref.py from the read-only checkout, imported
    through oracle/ref_shim.py's timm / easydict stand-ins) -> kind "reference"; the oracle port
    only if those files are not there -> kind "port".  fp32, eval, no_grad, all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dyt_oracle as O
    import ref_shim
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    img = torch.randn(sample_batch, 3, 224, 224, generator=torch.Generator().manual_seed(seed))
    sd = {k: v.detach().float().cpu() for k, v in state_dict.items()}
    kind = "port"
    fwd = lambda: O.vit_forward(img, sd, DEPTH, HEADS, 0.1, policy="fp32", sparse=True)["logits"]
    if ref_shim.reference_available():
        ref = ref_shim.import_reference("models.model_speed_test")
        tuning, select = ref_shim.reference_configs(ffn_num=BOTTLENECK, scalar="0.1", d_model=C_DIM,
                                                    ratio=RATE)
        model = ref.vit_base_patch16_224_in21k(num_classes=NUM_CLASSES, drop_path_rate=0.0,
                                               tuning_config=tuning, select_config=select)
        model.load_state_dict(sd, strict=True)
        model.eval()
        fwd = lambda: model(img)
        kind = "reference"
    with torch.no_grad():
        for _ in range(warmup):
            fwd()
        t0 = time.perf_counter()
        for _ in range(steps):
            fwd()
        dt = time.perf_counter() - t0
        keep = O.vit_forward(img[:4], sd, DEPTH, HEADS, 0.1, policy="fp32",
                             sparse=True)["token_select"].float().mean().item()
    what = ("reference models/model_speed_test.py (unmodified, oracle/_ref)" if kind == "reference"
            else "oracle port of the reference forward")
    return dict(value=sample_batch * steps / dt, ms_per_step=dt / steps * 1e3, cores=cores,
                keep_rate=keep, kind=kind,
                sample=f"{steps} forward(s) of {sample_batch} seed-{seed} images, {what}, fp32, "
                       f"torch {torch.__version__} CPU, {cores} threads")


def cpu_finetune_run(state_dict, sample_batch: int = 8, seed: int = 0):
    """CPU baseline of the fine-tune step: the oracle's train-mode restatement (student + teacher
    forward, loss, autograd backward through the frozen backbone) on a bounded sample, fp32, all
    host threads; one untimed + one timed step."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dyt_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(sample_batch, 3, 224, 224, generator=g)
    tgt = torch.randint(0, NUM_CLASSES, (sample_batch,), generator=g)
    sd = {}
    for k, v in state_dict.items():
        v = v.detach().float().cpu().clone()
        if ("adaptmlp" in k) or ("mlp_token_select" in k) or k.startswith("head."):
            v.requires_grad_(True)
        sd[k] = v
    n_tok = sd["pos_embed"].shape[1]
    bott = sd["blocks.0.adaptmlp.down_proj.weight"].shape[0]

    def step():
        noises = [(-torch.empty(sample_batch, n_tok - 1, 1).exponential_(generator=g).log(),
                   -torch.empty(sample_batch, n_tok - 1, 1).exponential_(generator=g).log())
                  for _ in range(2 * DEPTH)]
        drops = [(torch.rand(sample_batch, n_tok, bott, generator=g) >= 0.1).float() / 0.9
                 for _ in range(2 * DEPTH)]
        s = O.vit_train_forward(img, sd, DEPTH, HEADS, 1.0, noises[:DEPTH], drops[:DEPTH], False)
        t = O.vit_train_forward(img, sd, DEPTH, HEADS, 1.0, noises[DEPTH:], drops[DEPTH:], True)
        loss = O.finetune_loss(s["logits"], s["token_select"], t["logits"], tgt)
        loss.backward()
        for v in sd.values():
            v.grad = None

    step()
    t0 = time.perf_counter()
    step()
    dt = time.perf_counter() - t0
    return dict(value=sample_batch / dt, unit="images/s", cores=cores, kind="port",
                sample=f"1 fine-tune step (student + teacher forward, backward) of {sample_batch} "
                       f"images, fp32 autograd through the oracle, torch {torch.__version__} CPU, "
                       f"{cores} threads")


def synthetic_cpu_state_dict(seed: int = 0):
    """Same construction as dyt_b200.synthetic.build_vit_b16 but on the CPU, with the selector bias
    calibrated by the oracle (used when no GPU is present, i.e. the pure reference arm)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dyt_oracle as O
    sd = O.synthetic_state_dict(seed=seed)
    img = torch.randn(8, 3, 224, 224, generator=torch.Generator().manual_seed(seed))
    return O.calibrate_selector_bias(sd, img, DEPTH, HEADS, 0.1, RATE)


def run_reference(args, world, rank):
    if rank != 0:
        return
    sample = 32
    sd = synthetic_cpu_state_dict(0)
    r = cpu_reference_run(sd, max(1, args.steps), max(1, min(args.warmup, 2)), sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ViT-B/16 DyT inference 224x224 r~0.5 (speed.py path), CPU sample "
                               f"of {sample} images per step", "keep_rate": r["keep_rate"]},
        "cpu_baseline": {"value": r["value"], "unit": "images/s", "cores": r["cores"],
                         "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def time_kernel(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def kernel_table(model, x_blocks, kept_total, peaks, device):
    """Each hot-path kernel timed alone (CUDA events on the launching stream) at the bench shapes:
    algorithmic bytes / FLOPs per launch over the measured duration."""
    from dyt_b200 import _lib, ops
    blk = model.blocks[0]
    T = x_blocks.shape[0] * N_TOK
    h16 = torch.float16
    f = lambda t: t.detach().to(h16).contiguous()
    xn = torch.randn(T, C_DIM, device=device, dtype=h16)
    qkv_w, qkv_b = f(blk.attn.qkv.weight), f(blk.attn.qkv.bias)
    proj_w, proj_b = f(blk.attn.proj.weight), f(blk.attn.proj.bias)
    fc1_w, fc1_b = f(blk.mlp.fc1.weight), f(blk.mlp.fc1.bias)
    fc2_w, fc2_b = f(blk.mlp.fc2.weight), f(blk.mlp.fc2.bias)
    dn_w, dn_b = f(blk.adaptmlp.down_proj.weight), f(blk.adaptmlp.down_proj.bias)
    up_w, up_b = f(blk.adaptmlp.up_proj.weight), f(blk.adaptmlp.up_proj.bias)
    x32 = x_blocks.reshape(T, C_DIM).contiguous()
    qkv = torch.randn(x_blocks.shape[0], N_TOK, 3 * C_DIM, device=device, dtype=h16)
    K = int(kept_total)
    m_dev = torch.tensor([K], dtype=torch.int32, device=device)
    hid = torch.randn(T, HIDDEN, device=device, dtype=h16)
    dn = torch.randn(T, BOTTLENECK, device=device, dtype=h16)
    out_qkv = torch.empty(T, 3 * C_DIM, device=device, dtype=h16)
    out_c = torch.empty(T, C_DIM, device=device, dtype=h16)
    out_f = torch.empty(T, C_DIM, device=device, dtype=torch.float32)
    out_h = torch.empty(T, HIDDEN, device=device, dtype=h16)
    out_d = torch.empty(T, BOTTLENECK, device=device, dtype=h16)
    ln_w, ln_b = blk.norm2.weight.detach().float(), blk.norm2.bias.detach().float()
    sel_w, sel_b = blk.mlp_token_select.mlp_head.weight.detach(), blk.mlp_token_select.mlp_head.bias.detach()
    d = ops.dispatch(x_blocks, sel_w, sel_b, ln_w=ln_w, ln_b=ln_b)
    rows = []

    def add(name, fn, flops, bytes_, bound):
        t = time_kernel(fn)
        ach = (flops / t / 1e12) if bound == "tensor" else (bytes_ / t / 1e9)
        peak = peaks["tf_burst"] if bound == "tensor" else peaks["hbm"]
        rows.append(dict(kernel=name, us=t * 1e6, bound=bound, achieved=ach,
                         unit="TFLOP/s" if bound == "tensor" else "GB/s", frac=ach / peak,
                         flops=flops, bytes=bytes_))

    add("gemm qkv [T,768]x[2304,768]", lambda: ops.linear_f16(xn, qkv_w, qkv_b, out=out_qkv),
        2.0 * T * 3 * C_DIM * C_DIM, T * C_DIM * 2 + T * 3 * C_DIM * 2, "tensor")
    add("attention 12 heads x 197", lambda: ops.attn_varlen(qkv, HEADS),
        4.0 * x_blocks.shape[0] * HEADS * N_TOK * N_TOK * 64, T * 4 * C_DIM * 2, "tensor")
    add("gemm proj + residual", lambda: ops.linear_f16(xn, proj_w, proj_b, epilogue=_lib.EPI_BIAS_RESID,
                                                       resid=x32, out=out_f, want_f16_copy=False),
        2.0 * T * C_DIM * C_DIM, T * C_DIM * (2 + 4 + 4), "tensor")
    add("dispatcher (score+gate+compact+LN2 pack)",
        lambda: ops.dispatch(x_blocks, sel_w, sel_b, ln_w=ln_w, ln_b=ln_b),
        2.0 * T * C_DIM, T * C_DIM * 4 + K * C_DIM * (4 + 2), "hbm")
    add("gemm fc1 + GELU (kept rows)", lambda: ops.linear_f16(xn, fc1_w, fc1_b, epilogue=_lib.EPI_BIAS_GELU,
                                                              m_dev=m_dev, out=out_h),
        2.0 * K * HIDDEN * C_DIM, K * (C_DIM + HIDDEN) * 2, "tensor")
    add("gemm fc2 (kept rows)", lambda: ops.linear_f16(hid, fc2_w, fc2_b, m_dev=m_dev, out=out_c),
        2.0 * K * HIDDEN * C_DIM, K * (C_DIM + HIDDEN) * 2, "tensor")
    add("gemm adapter down + ReLU", lambda: ops.linear_f16(xn, dn_w, dn_b, epilogue=_lib.EPI_BIAS_RELU, out=out_d),
        2.0 * T * BOTTLENECK * C_DIM, T * (C_DIM + BOTTLENECK) * 2, "hbm")
    # adapter up-projection + scatter-merge + next LN1 in one kernel (the default block path)
    dn_relu = torch.relu(dn)
    add("adapter up + scatter-merge + next LN1 (fused)",
        lambda: ops.merge_up(dn_relu.reshape(x_blocks.shape[0], N_TOK, BOTTLENECK), up_w, up_b, 0.1,
                             x_blocks, out_c, d["token_pos"], next_ln=(ln_w, ln_b)),
        2.0 * T * BOTTLENECK * C_DIM, T * C_DIM * (4 + 4 + 2) + K * C_DIM * 2 + T * BOTTLENECK * 2, "hbm")
    return rows


def run_ours(args, world, rank, local):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sm_100a kernels are the only implementation "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    # several ranks on one host: each stays on its GPU's NUMA node (pinned buffers, launch thread)
    numa_node = bind_to_gpu_numa_node(local) if (world > 1 and not os.environ.get("DYT_NO_NUMA_BIND")) else None
    from dyt_b200 import engine, lib, synthetic
    lib()
    peaks = load_peaks()
    model = synthetic.build_vit_b16(device, num_classes=NUM_CLASSES, seed=0)
    # every rank generates the same seed-0 global batch and keeps its shard
    start, count = shard_images(BATCH * world, world, rank)
    assert count == BATCH
    gen = torch.Generator().manual_seed(rank)          # rank-specific images, same distribution
    host_images = torch.randn(BATCH, 3, 224, 224, generator=gen).pin_memory()
    images = host_images.to(device, non_blocking=True)
    cal = torch.randn(64, 3, 224, 224, generator=torch.Generator().manual_seed(0)).to(device)
    keep_cal = synthetic.calibrate_keep_rate(model, cal, RATE)

    # The call a user makes: dyt_b200.GraphedForward(model)(images) = model(images) under no_grad +
    # fp16 autocast, captured once into a CUDA graph and replayed (the ~115 launches of one forward
    # otherwise leave host-dependent gaps: 0.3-0.8 ms per step across boxes).
    from dyt_b200 import GraphedForward
    graphed = GraphedForward(model)

    def forward(imgs, slot=0):
        return graphed(imgs, slot=slot)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # samples cover warm-up + both timed regions (all under load)
    warm = max(3, args.warmup)
    static_images = graphed.input_buffer(images.shape, images.dtype, device, slot=0)
    static_images.copy_(images)  # device-resident batch, in the graph's own input buffer (no copy per step)
    for _ in range(warm):
        logits = forward(static_images)
    torch.cuda.synchronize()
    # realised keep rate on the bench batch
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        x0 = model._embed(images).float()
    _, masks, _, _ = engine.run_blocks(x0, list(model.blocks))
    keep_rate = masks[:, :, 1:].mean().item()
    kept_tokens = masks.sum(dim=2).mean().item()       # mean kept tokens per image per layer

    # ---- device-resident timing ----
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        logits = forward(static_images)
    ev1.record()
    torch.cuda.synchronize()
    barrier(world)
    sec = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3, world, device)
    value = BATCH * world * args.steps / sec
    # the reference's own timing style (speed.py:254-275): wall clock around the loop with a
    # torch.cuda.synchronize() on both sides; reported next to the CUDA-event number
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        logits = forward(static_images)
    torch.cuda.synchronize()
    wall_value = BATCH * args.steps / (time.perf_counter() - t0)

    # ---- end to end: pinned host images -> H2D -> model(images) -> logits D2H, every step ----
    # Double-buffered like a pinned-memory DataLoader (the reference's loader, speed.py:165-193):
    # the H2D copy of step i+1 runs on a copy stream while step i computes; every step's copy and
    # logits read-back are inside the timed region.
    host_logits = torch.empty(BATCH, NUM_CLASSES, dtype=torch.float16).pin_memory()
    # two static input slots of the graphed forward: the H2D copy lands directly in the graph's input
    dev_in = [graphed.input_buffer(images.shape, images.dtype, device, slot=1),
              graphed.input_buffer(images.shape, images.dtype, device, slot=2)]
    copy_stream = torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_loop(n):
        for b in range(2):
            consumed[b].record(main_stream)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[0])
            dev_in[0].copy_(host_images, non_blocking=True)
            copied[0].record(copy_stream)
        for i in range(n):
            cur, nxt = i % 2, (i + 1) % 2
            if i + 1 < n:
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[nxt])
                    dev_in[nxt].copy_(host_images, non_blocking=True)
                    copied[nxt].record(copy_stream)
            main_stream.wait_event(copied[cur])
            host_logits.copy_(forward(dev_in[cur], slot=cur + 1), non_blocking=True)
            consumed[cur].record(main_stream)

    e2e_loop(2)
    barrier(world)
    torch.cuda.synchronize()
    ev0.record()
    e2e_loop(args.steps)
    ev1.record()
    torch.cuda.synchronize()
    barrier(world)
    clocks = sampler.stop() if rank == 0 else None
    sec_e2e = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3, world, device)
    e2e_value = BATCH * world * args.steps / sec_e2e

    # ---- extras that involve every rank (fine-tune step with its gradient all-reduce), measured
    # after and outside the headline's timed regions ----
    del graphed, dev_in, static_images
    torch.cuda.empty_cache()
    extras = {}
    if not args.no_extras:
        try:
            extras = gather_extras_after(args, world, rank, local)
        except Exception as e:       # an extra must never take the headline down
            extras = {"error": repr(e)[:300]}
    if rank != 0:
        return
    fl_img = flops_per_image(kept_tokens)
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": sec / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": "ViT-B/16 DyT inference bs256 224x224 r~0.5 per B200 (speed.py path; "
                               "BASELINE configs[1]), images sharded over the GPUs",
                   "batch_per_gpu": BATCH, "keep_rate": round(keep_rate, 4),
                   "kept_tokens_per_image_layer": round(kept_tokens, 2),
                   "l2": "per-step working set ~1.1 GB of activations >> 126 MB L2 (no flush needed)",
                   "weights": "random init seed 0, selector bias calibrated on the GPU path",
                   "launch": "dyt_b200.GraphedForward: model(images) captured once, replayed as one "
                             "CUDA graph per step (both timed regions)",
                   "stem_head": "patch embed = own im2col + tcgen05 GEMM + assemble kernels; final LN "
                                "(cls rows, own row-gather LayerNorm kernel) + 768x100 head (own "
                                "tcgen05 GEMM, classes zero-padded to 104)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/s",
                "h2d_bytes_per_step": host_images.numel() * 4,
                "d2h_bytes_per_step": host_logits.numel() * 2,
                "host_numa_node_rank0": numa_node},
        # 9 per block + first LN1 + 3 stem kernels + final LN + head GEMM
        # per layer: qkv, attention, proj, adapter down, dispatcher, fc1, fc2, fused up + merge;
        # + first LN1, 3 stem kernels, final LN + head
        "gpu_launches": args.steps * (DEPTH * 8 + 1 + 3 + 2),
        "speed_py_style_images_per_s_per_gpu": wall_value,
        "model_flops_per_image": fl_img,
        "model_tflops": value / world * fl_img / 1e12,
        "frac_of_r_scaled_compute_roofline": value / world * fl_img / 1e12 / peaks["tf_sust"],
        "peaks": peaks,
    }
    if world == 1:
        rows = kernel_table(model, x0, kept_tokens * BATCH, peaks, device)
        top = max((r for r in rows if r["bound"] == "tensor"), key=lambda r: r["us"])
        line["roofline"] = {"kernel": top["kernel"], "bound": "tensor", "achieved": top["achieved"],
                            "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": top["frac"],
                            "traffic": ncu_traffic(top["kernel"]), "us_per_launch": top["us"],
                            "peak_source": peaks["source"] + " bf16 burst (kernel timed alone)",
                            # the whole step against the r-scaled compute roofline (the north-star
                            # figure): model FLOPs / step time / sustained tensor peak
                            "frac_model": line["frac_of_r_scaled_compute_roofline"],
                            "frac_model_peak": peaks["tf_sust"],
                            "frac_model_note": "whole forward: algorithmic FLOPs per image x images/s / "
                                               "measured sustained bf16 peak (kernel inside a long step)"}
        line["kernels"] = [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()}
                           for r in rows]
        if not args.no_cpu_baseline:
            sd = model.state_dict()
            r = cpu_reference_run(sd, steps=8, warmup=1, sample_batch=32)
            line["cpu_baseline"] = {"value": r["value"], "unit": "images/s", "cores": r["cores"],
                                    "kind": r["kind"], "sample": r["sample"]}
    if not args.no_extras and world == 1:
        try:
            extras["torch_eager_b200"] = torch_eager_b200(model, images)
        except Exception as e:
            extras["torch_eager_b200"] = {"error": repr(e)[:200]}
    line["extras"] = extras or None
    print(json.dumps(line), flush=True)


def gather_extras_after(args, world, rank, local):
    """Other BASELINE configs as one-liners inside the headline JSON: the fine-tune step at the same
    N ranks (BASELINE configs[2]; the only path with a collective: flat-arena gradient all-reduce,
    reported with its bytes and its own device time), and at N = 1 ViT-L/16 (configs[3] without the
    MoE-adapter) and the video model (configs[4])."""
    ex = {}
    ft = measure_finetune(8, 6, world, rank, local, cpu_baseline=False)
    if ft is not None:
        ex["finetune_b16"] = {k: ft[k] for k in ("value", "unit", "ms_per_step", "n_gpus", "model_tflops")}
        ex["finetune_b16"].update(allreduce_bytes_per_step=ft["config"]["allreduce_bytes_per_step"],
                                  allreduce_us=ft["config"]["allreduce_us"], loss=ft["config"]["loss"],
                                  images_per_gpu=64, workload=ft["config"]["workload"])
    if world > 1:
        st = measure_strong_scaling(world, rank, local)
        if st is not None:
            ex["strong_scaling_b256_global"] = st
    if world == 1:
        for wl in ("vit_l16", "vit_l16_moe", "video_b16"):
            r = measure_extra(wl, 8, 4, world, rank, local)
            if r is not None:
                ex[wl] = {"value": r["value"], "unit": r["unit"], "ms_per_step": r["ms_per_step"],
                          "workload": r["config"]["workload"]}
    return ex


def measure_strong_scaling(world, rank, local, steps=15, warmup=6):
    """The headline workload with the GLOBAL batch fixed at 256 images (256 / N per GPU): strong
    scaling.  At 32 images per GPU the 50 row tiles no longer fill the 74 CTA pairs of a GEMM, so
    this is the unfavourable reading of the metric; the headline line keeps 256 images per GPU."""
    device = torch.device("cuda", local)
    from dyt_b200 import GraphedForward, synthetic
    per = BATCH // world
    if per < 1:
        return None
    model = synthetic.build_vit_b16(device, num_classes=NUM_CLASSES, seed=0)
    cal = torch.randn(64, 3, 224, 224, generator=torch.Generator().manual_seed(0)).to(device)
    synthetic.calibrate_keep_rate(model, cal, RATE)
    images = torch.randn(per, 3, 224, 224, generator=torch.Generator().manual_seed(rank)).to(device)
    fwd = GraphedForward(model)
    buf = fwd.input_buffer(images.shape, images.dtype, device)
    buf.copy_(images)
    for _ in range(warmup):
        fwd.replay(images.shape, images.dtype, device)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(steps):
        fwd.replay(images.shape, images.dtype, device)
    ev1.record()
    torch.cuda.synchronize()
    barrier(world)
    sec = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3, world, device)
    del fwd, model
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {"value": per * world * steps / sec, "unit": "images/s", "ms_per_step": sec / steps * 1e3,
            "images_per_gpu": per, "n_gpus": world, "scaling": "strong",
            "workload": f"ViT-B/16 DyT inference, 256 images global = {per} per GPU, r~0.5"}


def measure_extra(workload, steps, warmup, world, rank, local):
    """Extra workloads (not the headline line): the other single-GPU BASELINE configs, same timing
    rules (CUDA events, >= 3 warm-up steps, inputs in HBM, working set >> L2).  Returns the JSON
    object (rank 0) or None."""
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    from dyt_b200 import synthetic
    args = argparse.Namespace(workload=workload, steps=steps, warmup=warmup)
    if args.workload in ("vit_l16", "vit_l16_moe"):
        if args.workload == "vit_l16":
            model = synthetic.build_vit_l16(device, seed=0)
            name = "ViT-L/16 DyT inference bs128 224x224 r~0.7 (plain adapter, as the reference's blocks)"
        else:
            model = synthetic.build_vit_l16_moe(device, seed=0, experts=4)
            name = ("ViT-L/16 DyT + MoE-adapter (4 experts; own restatement of the paper's MoE-adapter, not in "
                    "the reference: no reference parity) inference bs128 224x224 r~0.7")
        batch, rate, units = 128, 0.7, "images/s"
        images = torch.randn(batch, 3, 224, 224, generator=torch.Generator().manual_seed(rank)).to(device)
        cal = images[:32]
        per_step = batch
    elif args.workload == "seg_b16":
        # segmentation backbone: ViT-B/16 DyT at 512 x 512 = 1025 tokens per
        # image, reference default use_rel_pos_bias=False, feature maps after blocks 3 / 5 / 7 / 11 + FPN
        from dense_tasks.Segmentation.backbone.segmentation_vision_transformer_IN21K import VisionTransformer21K
        tuning, select = synthetic.reference_configs()
        select.update(token_ratio=2.0, token_minimal=0.1, token_minimal_weight=1.0)
        torch.manual_seed(0)
        model = VisionTransformer21K(img_size=512, patch_size=16, embed_dim=768, depth=12, num_heads=12,
                                     num_classes=0, tuning_config=tuning, select_config=select,
                                     use_rel_pos_bias=False)
        synthetic._randomise_dyt_parts(model, 0)
        model = model.eval().to(device)
        batch, rate, units = 16, 0.5, "images/s"
        name = "segmentation backbone ViT-B/16 DyT, 16 images 512x512 (1025 tokens), r~0.5, incl. FPN heads"
        images = torch.randn(batch, 3, 512, 512, generator=torch.Generator().manual_seed(rank)).to(device)
        cal = None
        per_step = batch
    else:
        model = synthetic.build_video_b16(device, seed=0)
        clips = 32 if world == 1 else 8   # BASELINE: 64 clips over 8 GPUs; one GPU alone takes 32
        batch, rate, units, name = clips, 0.5, "clips/s", f"video ViT-B DyT, {clips} clips x 8 x 224x224 per GPU, r~0.5"
        images = torch.randn(clips, 3, 8, 224, 224, generator=torch.Generator().manual_seed(rank)).to(device)
        cal = images[:4].permute(0, 2, 1, 3, 4).reshape(-1, 3, 224, 224)
        per_step = clips
    if cal is not None:
        keep = synthetic.calibrate_keep_rate(model, cal, rate)
    else:   # segmentation: calibrate the selector biases layer by layer on the bench images
        from dyt_b200 import engine
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            pe = model.patch_embed
            from dyt_b200 import ops as _ops
            x = _ops.patch_embed(images[:4], pe.proj.weight, pe.proj.bias, model.cls_token,
                                 model.pos_embed, pe.patch_size[0])
            kept = []
            for blk in model.blocks:
                blk.mlp_token_select.mlp_head.bias.zero_()
                _, _, lg, _ = engine.run_blocks(x, [blk], fuse_next_ln=False)
                blk.mlp_token_select.mlp_head.bias.fill_(-float(torch.quantile(lg.flatten().float(), 1.0 - rate)))
                x, mk, _, _ = engine.run_blocks(x, [blk], fuse_next_ln=False)
                kept.append(mk[:, :, 1:].mean().item())
        keep = sum(kept) / len(kept)

    from dyt_b200 import GraphedForward
    graphed = GraphedForward(model)

    static_images = graphed.input_buffer(images.shape, images.dtype, device)
    static_images.copy_(images)          # inputs resident in HBM, in the graph's own input buffer

    def forward():
        return graphed(static_images)

    warm = max(3, args.warmup)
    for _ in range(warm):
        forward()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        forward()
    ev1.record()
    torch.cuda.synchronize()
    barrier(world)
    sec = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3, world, device)
    del graphed, model
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {"metric": f"extra workload {args.workload}", "value": per_step * world * args.steps / sec,
            "unit": units, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": name, "keep_rate_calibration": round(keep, 4)}}


def run_extra(args, world, rank, local):
    line = measure_extra(args.workload, args.steps, args.warmup, world, rank, local)
    if line is not None:
        print(json.dumps(line), flush=True)


def finetune_flops_per_image(bottleneck: int = 16) -> float:
    """Algorithmic FLOPs (2 x MAC) of one fine-tune step per image: two dense forward passes
    (student, teacher) and their backward through the frozen backbone = data-gradient GEMMs of the
    same size as the forward ones, attention backward = 5 contractions against the forward's 2,
    adapter weight gradients; block 0 needs no gradient w.r.t. its input (stem frozen)."""
    n, c, hid = N_TOK, C_DIM, HIDDEN
    qkv, proj, mlp = 2.0 * n * c * 3 * c, 2.0 * n * c * c, 4.0 * n * c * hid
    attn = 4.0 * n * n * c
    adapter = 4.0 * n * c * bottleneck
    fwd = qkv + proj + mlp + attn + adapter + 2.0 * (n - 1) * c
    bwd = qkv + proj + mlp + 2.5 * attn + 2.0 * adapter
    bwd_first = bwd - (qkv + proj + 2.5 * attn)       # block 0: no input gradient
    per_pass = DEPTH * fwd + (DEPTH - 1) * bwd + bwd_first + 2.0 * (n - 1) * c * 3 * 256
    return 2.0 * per_pass


def measure_finetune(steps, warmup, world, rank, local, cpu_baseline=False):
    """BASELINE configs[2]: ViT-B/16 DyT fine-tune step on synthetic VTAB-shape data, 64 images per
    GPU (512 global at 8), ffn_num 16, adapter scale 1 (train_vtab.sh:8, main_vtab.py:351).  One step
    = student pass + teacher pass + loss + backward + ONE gradient all-reduce (NCCL, flat 74-tensor
    arena) + AdamW step, as engine_finetune.py:47-76."""
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    from dyt_b200 import synthetic
    from dyt_b200.ddp import GradArena, trainable_parameters
    from dyt_b200.finetune import FinetuneStep
    batch = 64
    model = synthetic.build_vit_b16(device, flavour="train", ffn_num=16, scalar="1.0", seed=0)
    gen = torch.Generator().manual_seed(rank)
    images = torch.randn(batch, 3, 224, 224, generator=gen).to(device)
    targets = torch.randint(0, NUM_CLASSES, (batch,), generator=gen).to(device)
    keep = synthetic.calibrate_keep_rate(model, images, RATE)
    params = trainable_parameters(model)
    model.train()
    arena = GradArena(params)
    step = FinetuneStep(model, torch.optim.AdamW(params, lr=1e-3, weight_decay=0.05), arena,
                        cuda_graph=True)
    args = argparse.Namespace(steps=steps, warmup=warmup)
    warm = max(6, args.warmup)          # the dynamic loss scale settles within the first steps
    for _ in range(warm):
        loss = step(images, targets)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        loss = step(images, targets)
    ev1.record()
    torch.cuda.synchronize()
    barrier(world)
    sec = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3, world, device)
    # the step's only collective on its own: the flat gradient arena all-reduce (NCCL over NVLink),
    # timed on the device, max over ranks
    ar_us = None
    if world > 1:
        for _ in range(5):
            arena.all_reduce_mean()
        barrier(world)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(20):
            arena.all_reduce_mean()
        ev1.record()
        torch.cuda.synchronize()
        ar_us = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3, world, device) / 20 * 1e6
    loss_val, scale_val, nbytes, ntens = float(loss), float(step.scaler.get_scale()), arena.nbytes, len(arena.params)
    sd_cpu = {k: v.detach().float().cpu() for k, v in model.state_dict().items()} if (cpu_baseline and rank == 0) else None
    del step, arena, model
    torch.cuda.empty_cache()
    if rank == 0:
        cpu = None
        if cpu_baseline and world == 1:
            cpu = cpu_finetune_run(sd_cpu)
        return ({
            "metric": "extra workload finetune_b16", "value": batch * world * args.steps / sec,
            "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "ViT-B/16 DyT fine-tune step (student + teacher forward, backward, "
                                   "grad all-reduce, AdamW), 64 images per GPU, ffn_num 16, scale 1",
                       "keep_rate_calibration": round(keep, 4), "loss": loss_val,
                       "loss_scale": scale_val,
                       "loss_recipe": "AdaLoss(token_loss_ratio 2, token_minimal 0, token_minimal_weight 0) "
                                      "+ teacher CE + KL (main_image.py:206-209, engine_finetune.py:47-65)",
                       "cuda_graph": "forward + backward of the step replayed as one CUDA graph",
                       "allreduce_bytes_per_step": nbytes if world > 1 else 0,
                       "allreduce_us": ar_us,
                       "trainable_tensors": ntens},
            "model_flops_per_image": finetune_flops_per_image(16),
            "model_tflops": batch * args.steps / sec * finetune_flops_per_image(16) / 1e12,
            "cpu_baseline": cpu})
    return None


def run_finetune(args, world, rank, local):
    line = measure_finetune(args.steps, args.warmup, world, rank, local,
                            cpu_baseline=not args.no_cpu_baseline)
    if line is not None:
        print(json.dumps(line), flush=True)


def torch_eager_b200(model, images, steps=5, warm=3):
    """The 'existing Blackwell kernels' bar: the reference's op sequence as plain PyTorch
    on this GPU -- the oracle restatement under REAL torch.autocast(fp16): cuBLAS, flash-SDPA, ATen
    nonzero / index gather-scatter with their host syncs -- timed with CUDA events.  A reported
    baseline (kind "port"), not part of the product path."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dyt_oracle as O
    import torch.nn.functional as F

    def attention_sdpa(x, p, prefix, num_heads, policy="fp32"):   # vision_transformer_IN21K.py:54-65
        B, N, C = x.shape
        qkv = F.linear(x, p[prefix + "qkv.weight"], p[prefix + "qkv.bias"])
        q, k, v = qkv.reshape(B, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4).unbind(0)
        o = F.scaled_dot_product_attention(q, k, v)
        return F.linear(o.transpose(1, 2).reshape(B, N, C), p[prefix + "proj.weight"], p[prefix + "proj.bias"])

    sd = {k: v.detach() for k, v in model.state_dict().items()}
    saved = O.attention
    O.attention = attention_sdpa
    try:
        def fwd():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                return O.vit_forward(images, sd, DEPTH, HEADS, 0.1, policy="fp32", sparse=True)
        for _ in range(warm):
            fwd()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fwd()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    finally:
        O.attention = saved
    return {"value": images.shape[0] / ms * 1e3, "unit": "images/s", "ms_per_step": ms, "kind": "port",
            "what": "oracle restatement of the reference forward under torch.autocast(fp16) on this GPU "
                    "(cuBLAS + flash SDPA + ATen nonzero/index), bs %d" % images.shape[0]}


def ncu_traffic(kernel_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel, from the
    committed ncu --set full summary of THIS build's kernel sources (profiles/r2_ncu_traffic.json;
    null when the kernel source changed since the capture)."""
    import hashlib
    path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if not os.path.exists(path):
        return None
    try:
        rec = json.load(open(path))
        h = hashlib.sha256()
        for f in rec["source_files"]:
            h.update(open(os.path.join(ROOT, f), "rb").read())
        if h.hexdigest() != rec["source_sha256"]:
            return None
        return rec["dram_bytes_per_launch"].get(kernel_name)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=15)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extras block (fine-tune step, ViT-L, video, torch-eager bar)")
    ap.add_argument("--workload", default="vit_b16", choices=["vit_b16", "vit_l16", "vit_l16_moe", "video_b16", "finetune_b16", "seg_b16"],
                    help="vit_b16 = the BASELINE metric (default); the others are extra lines")
    args = ap.parse_args()
    if args.impl == "reference":
        # CPU arm: rank 0 alone works, nobody needs a process group
        run_reference(args, int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")))
        return
    world, rank, local = dist_setup(args.gpus)
    try:
        if args.workload == "finetune_b16":
            run_finetune(args, world, rank, local)
        elif args.workload != "vit_b16":
            run_extra(args, world, rank, local)
        else:
            run_ours(args, world, rank, local)
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
