"""'Existing Blackwell kernels' bar: the reference algorithm as plain PyTorch
(the oracle restatement = the reference modules' own ops: nn.functional.linear / layer_norm / gelu,
scaled_dot_product_attention, nonzero + index gather / scatter) run ON THE B200 under
torch.autocast(fp16), i.e. cuBLAS / cuDNN-flash / ATen kernels, timed like bench.py (CUDA events,
6 warm-up + 15 timed steps, batch resident in HBM).  /root/reference itself is not on the GPU box;
the oracle is pinned to it by tests/golden.  Test infrastructure (it runs the oracle): not part of the product or of bench.py.

  python tests/baseline_torch_eager_b200.py            # inference bs256 + fine-tune step bs64
"""
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), ROOT]
import dyt_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
DEPTH, HEADS = 12, 12


def attention_sdpa(x, p, prefix, num_heads):
    """reference Attention.forward with fused_attn (vision_transformer_IN21K.py:54-65)"""
    B, N, C = x.shape
    qkv = F.linear(x, p[prefix + "qkv.weight"], p[prefix + "qkv.bias"])
    qkv = qkv.reshape(B, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    o = F.scaled_dot_product_attention(q, k, v)
    return F.linear(o.transpose(1, 2).reshape(B, N, C), p[prefix + "proj.weight"], p[prefix + "proj.bias"])


O.attention = lambda x, p, prefix, num_heads, policy="fp32": attention_sdpa(x, p, prefix, num_heads)


def timed(fn, warm=6, steps=15):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    g = torch.Generator().manual_seed(0)
    sd = O.synthetic_state_dict(seed=0)
    cal = torch.randn(8, 3, 224, 224, generator=g)
    sd = O.calibrate_selector_bias(sd, cal, DEPTH, HEADS, 0.1, 0.5)
    sd = {k: v.to(dev) for k, v in sd.items()}
    img = torch.randn(256, 3, 224, 224, generator=g).to(dev)

    def infer():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            return O.vit_forward(img, sd, DEPTH, HEADS, 0.1, policy="fp32", sparse=True)

    ms = timed(infer)
    keep = infer()["token_select"].float().mean().item()
    print(f"torch eager (autocast fp16, cuBLAS + SDPA) inference bs256: {ms:.2f} ms/step "
          f"{256 / ms * 1e3:.0f} img/s keep {keep:.3f}")

    # fine-tune step, 64 images: student + teacher + loss + backward through the oracle
    B = 64
    imgs = img[:B]
    tgt = torch.randint(0, 100, (B,), generator=g).to(dev)
    p = {}
    for k, v in sd.items():
        v = v.clone()
        if ("adaptmlp" in k) or ("mlp_token_select" in k) or k.startswith("head."):
            v.requires_grad_(True)
        p[k] = v

    def step():
        noises = [(-torch.empty(B, 196, 1, device=dev).exponential_().log(),
                   -torch.empty(B, 196, 1, device=dev).exponential_().log()) for _ in range(2 * DEPTH)]
        drops = [(torch.rand(B, 197, 64, device=dev) >= 0.1).float() / 0.9 for _ in range(2 * DEPTH)]
        with torch.autocast("cuda", dtype=torch.float16):
            s = O.vit_train_forward(imgs, p, DEPTH, HEADS, 0.1, noises[:DEPTH], drops[:DEPTH], False)
            t = O.vit_train_forward(imgs, p, DEPTH, HEADS, 0.1, noises[DEPTH:], drops[DEPTH:], True)
            loss = O.finetune_loss(s["logits"].float(), s["token_select"].float(), t["logits"].float(), tgt)
        (loss * 1024.0).backward()
        for v in p.values():
            v.grad = None

    ms = timed(step, warm=3, steps=8)
    print(f"torch eager (autocast fp16) fine-tune step bs64: {ms:.2f} ms/step {B / ms * 1e3:.0f} img/s")


if __name__ == "__main__":
    main()
