"""Pins the oracle's "amp16" arithmetic policy -- our emulation of what torch.autocast(fp16) does
to the reference's op sequence -- to REAL torch.autocast on the GPU, and holds the kernels to that.

The reference runs under `torch.cuda.amp.autocast()` (speed.py:254, engine_finetune.py:47).  Here
the oracle's restatement of the reference ops (policy "fp32" = plain F.linear / F.layer_norm /
F.gelu / sigmoid / nonzero gather-scatter, attention through F.scaled_dot_product_attention like
the reference's fused_attn path, vision_transformer_IN21K.py:54-65) is executed on CUDA tensors
inside a real `torch.autocast("cuda", dtype=torch.float16)` region: cuBLAS / SDPA / ATen decide every
rounding point.  Three-way comparison on identical inputs:
    real autocast (GPU, torch)  ~  oracle amp16 (CPU emulation)  ~  dyt_b200 kernels
Tolerance: max|err| <= 1e-3 * max|ref| on every row whose gate decision agrees; decisions may
differ only for tokens whose logit lies within one fp16 ulp of the gate threshold.
"""
import pytest
import torch
import torch.nn.functional as F

import dyt_oracle as O
from test_block_gpu import _check_masks, _rel, _rel_rows, _speed_model, vitb_sd  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _attention_sdpa(x, p, prefix, num_heads, policy="fp32"):
    """reference Attention.forward with fused_attn (models/vision_transformer_IN21K.py:54-65)"""
    B, N, C = x.shape
    qkv = F.linear(x, p[prefix + "qkv.weight"], p[prefix + "qkv.bias"])
    qkv = qkv.reshape(B, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    o = F.scaled_dot_product_attention(q, k, v)
    return F.linear(o.transpose(1, 2).reshape(B, N, C), p[prefix + "proj.weight"],
                    p[prefix + "proj.bias"])


@pytest.fixture()
def real_autocast(monkeypatch):
    monkeypatch.setattr(O, "attention", _attention_sdpa)

    def run(fn, *args, **kw):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            return fn(*args, **kw)
    return run


def test_block_three_way_real_autocast(dev, vitb_sd, real_autocast, monkeypatch):
    g, sd, img = vitb_sd
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    m = _speed_model(sd, dev)
    from dyt_b200 import engine
    x = torch.randn(3, 197, 768, generator=torch.Generator().manual_seed(31)) * 0.7
    for layer in (0, 5, 11):
        pre = f"blocks.{layer}."
        real = real_autocast(O.block_sparse, x.to(dev), sd_dev, pre, 12, g["scale"], "fp32")
        assert real["out"].dtype == torch.float32 and real["logits"].dtype == torch.float32
        monkeypatch.undo()                         # the CPU emulation uses its own attention_core
        emu = O.block_sparse(x, sd, pre, 12, g["scale"], "amp16")
        monkeypatch.setattr(O, "attention", _attention_sdpa)
        out, masks, logits, _ = engine.run_blocks(x.to(dev), [m.blocks[layer]], fuse_next_ln=False)
        real_mask, real_out = real["mask"].cpu(), real["out"].cpu()
        # (1) the emulation against real autocast: this is the pin of the amp16 policy
        _check_masks(emu["mask"], real_mask, real["logits"].cpu(), max_flips=2)
        same = emu["mask"][..., 0] == real_mask[..., 0]
        assert _rel_rows(emu["out"], real_out, same) <= 1e-3, layer
        assert _rel(emu["logits"], real["logits"]) <= 2e-3
        # (2) the kernels against real autocast
        _check_masks(masks[0].unsqueeze(-1).cpu(), real_mask, real["logits"].cpu(), max_flips=2)
        same = masks[0].cpu() == real_mask[..., 0]
        assert _rel_rows(out, real_out, same) <= 1e-3, layer
        assert _rel(logits[0].unsqueeze(-1), real["logits"]) <= 2e-3


def test_model_real_autocast_with_the_kernels_decisions(dev, vitb_sd, real_autocast):
    """Whole 12-layer model: the real-autocast run continues with the kernels' own gate decisions
    (so there are no data-dependent flips) and must give the same logits to fp16 accuracy."""
    g, sd, img = vitb_sd
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    m = _speed_model(sd, dev)
    from dyt_b200 import engine
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        x = m._embed(img.to(dev))
        x, masks, _, _ = engine.run_blocks(x, list(m.blocks))
        logits = m.forward_head(m.norm(x))
    forced = [masks[i].unsqueeze(-1) for i in range(12)]
    real = real_autocast(O.vit_forward, img.to(dev), sd_dev, 12, 12, g["scale"], policy="fp32",
                         sparse=True, forced_masks=forced)
    assert _rel(logits, real["logits"]) <= 5e-3
    # and the free-running real-autocast model keeps (almost) the same tokens
    free = real_autocast(O.vit_forward, img.to(dev), sd_dev, 12, 12, g["scale"], policy="fp32",
                         sparse=True)
    agree = (free["token_select"][..., 0].cpu() == masks.permute(1, 0, 2)[:, :, 1:].cpu()).float()
    assert bool((agree.mean(dim=(0, 2)) >= 0.99).all())
    assert agree[:, 0].mean() >= 0.995
