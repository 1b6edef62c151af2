"""Minimal stand-ins for the six timm==0.9.12 symbols the reference hot path uses
(SURVEY.md section 8c): Mlp, PatchEmbed, DropPath, PatchDropout (never enabled), trunc_normal_,
use_fused_attn.  Parameter / attribute names follow timm so real timm checkpoints load.
These modules only carry parameters and the stem/head glue; the block compute is in engine.py.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def trunc_normal_(tensor: torch.Tensor, mean: float = 0.0, std: float = 1.0, a: float = -2.0,
                  b: float = 2.0) -> torch.Tensor:
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def use_fused_attn() -> bool:
    # timm semantics; the dyt_b200 attention kernel is always the fused one
    return hasattr(F, "scaled_dot_product_attention") and int(os.environ.get("TIMM_FUSED_ATTN", "1")) > 0


class DropPath(nn.Module):
    """Stochastic depth; identity in eval and for p == 0 (every reference entry script uses 0)."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        m = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            m.div_(keep)
        return x * m


class Mlp(nn.Module):
    """fc1 -> act -> drop1 -> norm -> fc2 -> drop2 (attribute names of timm.layers.Mlp)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU,
                 norm_layer=None, bias=True, drop=0.0, use_conv=False):
        super().__init__()
        hidden_features = hidden_features or in_features
        out_features = out_features or in_features
        bias = to_2tuple(bias)
        drop = to_2tuple(drop)
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias[0])
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop[0])
        self.norm = norm_layer(hidden_features) if norm_layer is not None else nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias[1])
        self.drop2 = nn.Dropout(drop[1])

    def forward(self, x):
        # Only the FLOP probe (Block.forward_count_flops) and user code outside the dispatched
        # block call this; the block itself runs fc1/GELU/fc2 through the tcgen05 GEMM kernels.
        from . import ops, _lib
        if not x.is_cuda:
            raise _lib.DytError("dyt_b200 Mlp: CUDA only (no CPU fallback)")
        h, _ = ops.linear_f16(x.to(torch.float16), self.fc1.weight.to(torch.float16),
                              None if self.fc1.bias is None else self.fc1.bias.to(torch.float16),
                              epilogue=_lib.EPI_BIAS_GELU)
        y, _ = ops.linear_f16(h, self.fc2.weight.to(torch.float16),
                              None if self.fc2.bias is None else self.fc2.bias.to(torch.float16))
        return y


class PatchEmbed(nn.Module):
    """Conv2d(in_chans, embed_dim, k=patch, s=patch) named `proj`, flattened to [B, L, C]."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None,
                 flatten=True, bias=True, **_unused):
        super().__init__()
        self.img_size = to_2tuple(img_size)
        self.patch_size = to_2tuple(patch_size)
        self.grid_size = (self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=self.patch_size,
                              stride=self.patch_size, bias=bias)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        if x.shape[-2:] != tuple(self.img_size):
            raise ValueError(f"input size {tuple(x.shape[-2:])} != model size {self.img_size}")
        x = self.proj(x)
        if self.flatten:
            x = x.flatten(2).transpose(1, 2)
        return self.norm(x)


class PatchDropout(nn.Module):
    """Present for API parity; the reference never enables it (patch_drop_rate == 0)."""

    def __init__(self, prob: float = 0.5, num_prefix_tokens: int = 1, **_unused):
        super().__init__()
        if prob > 0:
            raise NotImplementedError("PatchDropout > 0 is outside the dyt_b200 hot path")
        self.prob, self.num_prefix_tokens = prob, num_prefix_tokens

    def forward(self, x):
        return x
