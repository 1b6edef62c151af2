"""Synthetic ViT-B/16 DyT model for benchmarks and smoke runs (no checkpoints are available
offline): reference-style random init, non-degenerate adapters / selectors, and selector biases
calibrated ON THE GPU PATH so the realised keep-rate is ~r."""
from __future__ import annotations

import torch

from . import engine


class AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def reference_configs(ffn_num: int = 64, scalar: str = "0.1", d_model: int = 768, ratio: float = 0.5):
    """tuning_config / select_config as main_image.py:186-210 and speed.py:196-224 build them."""
    tuning = AttrDict(ffn_adapt=True, ffn_option="parallel", ffn_adapter_layernorm_option="none",
                      ffn_adapter_init_option="lora", ffn_adapter_scalar=scalar, ffn_num=ffn_num,
                      d_model=d_model, vpt_on=False, vpt_num=0)
    select = AttrDict(open=True, keep_layers=0, token_target_ratio=ratio)
    return tuning, select


def build_vit_b16(device, num_classes: int = 100, seed: int = 0, flavour: str = "speed",
                  ffn_num: int = 64, scalar: str = "0.1"):
    if flavour == "speed":
        from models.model_speed_test import vit_base_patch16_224_in21k as ctor
    else:
        from models.vision_transformer_IN21K import vit_base_patch16_224_in21k as ctor
    tuning, select = reference_configs(ffn_num=ffn_num, scalar=scalar)
    torch.manual_seed(seed)
    model = ctor(num_classes=num_classes, drop_path_rate=0.0, tuning_config=tuning,
                 select_config=select)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for blk in model.blocks:
            blk.adaptmlp.up_proj.weight.copy_(
                0.02 * torch.randn(blk.adaptmlp.up_proj.weight.shape, generator=g))
            blk.mlp_token_select.mlp_head.weight.copy_(
                0.5 * torch.randn(blk.mlp_token_select.mlp_head.weight.shape, generator=g))
    return model.eval().to(device)


def _randomise_dyt_parts(model, seed: int):
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for blk in model.blocks:
            blk.adaptmlp.up_proj.weight.copy_(
                0.02 * torch.randn(blk.adaptmlp.up_proj.weight.shape, generator=g))
            blk.mlp_token_select.mlp_head.weight.copy_(
                0.5 * torch.randn(blk.mlp_token_select.mlp_head.weight.shape, generator=g))


def build_vit_l16(device, num_classes: int = 100, seed: int = 0, ffn_num: int = 64,
                  scalar: str = "0.1"):
    """BASELINE configs[3] backbone: ViT-L/16 (C = 1024, depth 24, 16 heads) with DyT blocks via the
    reference's generic ctor (vision_transformer_IN21K.py:199-231); the MoE-adapter of that config
    does not exist in the reference."""
    from models.model_speed_test import VisionTransformer
    tuning, select = reference_configs(ffn_num=ffn_num, scalar=scalar, d_model=1024, ratio=0.7)
    torch.manual_seed(seed)
    model = VisionTransformer(patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4.0,
                              qkv_bias=True, num_classes=num_classes, drop_path_rate=0.0,
                              tuning_config=tuning, select_config=select)
    _randomise_dyt_parts(model, seed)
    return model.eval().to(device)


def _randomise_moe_parts(model, seed: int):
    g = torch.Generator().manual_seed(seed + 3)
    with torch.no_grad():
        for blk in model.blocks:
            a = blk.adaptmlp
            a.router.weight.copy_(0.3 * torch.randn(a.router.weight.shape, generator=g))
            for lin in a.up_proj:
                lin.weight.copy_(0.02 * torch.randn(lin.weight.shape, generator=g))
                lin.bias.copy_(0.02 * torch.randn(lin.bias.shape, generator=g))
            blk.mlp_token_select.mlp_head.weight.copy_(
                0.5 * torch.randn(blk.mlp_token_select.mlp_head.weight.shape, generator=g))


def build_vit_l16_moe(device, num_classes: int = 100, seed: int = 0, ffn_num: int = 64,
                      scalar: str = "0.1", experts: int = 4):
    """BASELINE configs[3]: ViT-L/16 DyT with the MoE-adapter (`tuning_config.moe_experts`; the
    MoE-adapter is not in the reference: own restatement, no reference parity)."""
    from models.model_speed_test import VisionTransformer
    tuning, select = reference_configs(ffn_num=ffn_num, scalar=scalar, d_model=1024, ratio=0.7)
    tuning["moe_experts"] = experts
    torch.manual_seed(seed)
    model = VisionTransformer(patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4.0,
                              qkv_bias=True, num_classes=num_classes, drop_path_rate=0.0,
                              tuning_config=tuning, select_config=select)
    _randomise_moe_parts(model, seed)
    return model.eval().to(device)


def build_video_b16(device, num_classes: int = 174, seed: int = 0, ffn_num: int = 64,
                    scalar: str = "0.1"):
    """BASELINE configs[4] model: per-frame ViT-B/16 DyT + attentive pooling head."""
    from video_models.video_vision_transformer_IN21K import vit_base_patch16_224_in21k as ctor
    tuning, select = reference_configs(ffn_num=ffn_num, scalar=scalar)
    torch.manual_seed(seed)
    model = ctor(num_classes=num_classes, drop_path_rate=0.0, tuning_config=tuning,
                 select_config=select)
    _randomise_dyt_parts(model, seed)
    with torch.no_grad():
        model.query_token.normal_(std=0.5, generator=torch.Generator().manual_seed(seed + 2))
    return model.eval().to(device)


@torch.no_grad()
def calibrate_keep_rate(model, images: torch.Tensor, rate: float) -> float:
    """Layer by layer: bias_i = -quantile(logits_i, 1 - rate) on `images`, using the kernels."""
    with torch.autocast("cuda", dtype=torch.float16):
        x = model._embed(images).float()
    kept = []
    for blk in model.blocks:
        blk.mlp_token_select.mlp_head.bias.zero_()
        _, _, logits, _ = engine.run_blocks(x, [blk], fuse_next_ln=False)
        q = torch.quantile(logits.flatten().float()[:4_000_000], 1.0 - rate)
        blk.mlp_token_select.mlp_head.bias.fill_(-float(q))
        x, masks, _, _ = engine.run_blocks(x, [blk], fuse_next_ln=False)
        kept.append(masks[:, :, 1:].mean().item())
    return sum(kept) / len(kept)
