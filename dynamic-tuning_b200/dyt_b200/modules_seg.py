"""Segmentation backbone with DyT blocks, backed by the sm_100a kernels (inference forward).

Drop-in for the reference's dense_tasks/Segmentation/backbone/segmentation_vision_transformer_IN21K.py
(`VisionTransformer21K`, registered in mmseg's BACKBONES there): same constructor keywords,
state_dict keys (blocks.*.attn.relative_position_bias_table / relative_position_index when
use_rel_pos_bias, fpn1..fpn4) and forward contract
    forward(x [B, 3, H, W]) -> (tuple of 4 feature maps, dict(token_select, token_logits, loss)).
Differences from the image model that matter for the kernels: 512 x 512 inputs = 1025 tokens per
image and an optional additive relative-position bias in the attention (reference :181-203), both
served by dyt_attn_bias_fwd; every block reports its mask, and the maps after blocks `out_indices`
go through the FPN heads (ConvTranspose2d / MaxPool2d: plain torch modules, outside the hot path).
The dense masked block of the reference equals the sparse block in eval mode,
which is what runs.  Training this backbone is not built (raises).
"""
from __future__ import annotations

from functools import partial
from typing import Callable, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import engine, ops
from ._lib import DytError
from .layers import Mlp, PatchEmbed, trunc_normal_
from .modules import Adapter, TokenSelect, _act_dtype, _no_backward


def relative_position_index(window: Tuple[int, int]) -> torch.Tensor:
    """[Wh*Ww + 1, Wh*Ww + 1] index into the bias table: pairs of patches by their 2-D offset, plus
    three extra entries for cls->token, token->cls and cls->cls (reference :150-176)."""
    wh, ww = window
    ys, xs = torch.meshgrid(torch.arange(wh), torch.arange(ww), indexing="ij")
    pos = torch.stack([ys.reshape(-1), xs.reshape(-1)], dim=1)            # [L, 2]
    delta = pos[:, None, :] - pos[None, :, :]                             # [L, L, 2]
    idx = (delta[..., 0] + wh - 1) * (2 * ww - 1) + (delta[..., 1] + ww - 1)
    n_rel = (2 * wh - 1) * (2 * ww - 1) + 3
    out = torch.zeros((wh * ww + 1, wh * ww + 1), dtype=idx.dtype)
    out[1:, 1:] = idx
    out[0, :] = n_rel - 3
    out[:, 0] = n_rel - 2
    out[0, 0] = n_rel - 1
    return out


class SegAttention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_norm=False, attn_drop=0.0,
                 proj_drop=0.0, norm_layer=nn.LayerNorm, window_size=None):
        super().__init__()
        if qk_norm or attn_drop > 0 or proj_drop > 0:
            raise NotImplementedError("qk_norm / attention dropout are never enabled by the reference")
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.q_norm = nn.Identity()
        self.k_norm = nn.Identity()
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        if window_size:
            self.window_size = tuple(window_size)
            self.num_relative_distance = (2 * window_size[0] - 1) * (2 * window_size[1] - 1) + 3
            self.relative_position_bias_table = nn.Parameter(
                torch.zeros(self.num_relative_distance, num_heads))
            self.register_buffer("relative_position_index", relative_position_index(self.window_size))
        else:
            self.window_size = None
            self.relative_position_bias_table = None
            self.relative_position_index = None

    def bias(self) -> Optional[torch.Tensor]:
        """[heads, N, N] fp32 (reference :192-197), None without a table.  The values live in a
        buffer whose rows are padded to a multiple of 4 floats (ops.pad_attn_bias)."""
        if self.relative_position_bias_table is None:
            return None
        tab = self.relative_position_bias_table
        key = (tab.data_ptr(), tab._version, tab.device)
        cached = self.__dict__.get("_dyt_bias")
        if cached is None or cached[0] != key:     # gathered once per table version, not per forward
            n = self.relative_position_index.shape[0]
            t = tab.detach().float()
            b = ops.pad_attn_bias(t[self.relative_position_index.reshape(-1)].reshape(n, n, -1).permute(2, 0, 1))
            cached = (key, b)
            self.__dict__["_dyt_bias"] = cached
        return cached[1]

    def forward(self, x):
        _no_backward("Attention", x, self.qkv.weight, self.proj.weight)
        h16 = torch.float16
        qkv, _ = ops.linear_f16(x.to(h16), self.qkv.weight.to(h16),
                                None if self.qkv.bias is None else self.qkv.bias.to(h16))
        o = ops.attn_bias(qkv, self.num_heads, self.bias())
        y, _ = ops.linear_f16(o, self.proj.weight.to(h16), self.proj.bias.to(h16))
        return y.to(_act_dtype())


class SegBlock(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_norm=False, proj_drop=0.0,
                 attn_drop=0.0, init_values=None, drop_path=0.0, act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm, mlp_layer=Mlp, window_size=None, tuning_config=None,
                 select=False):
        super().__init__()
        if init_values or drop_path > 0:
            raise NotImplementedError("LayerScale / DropPath are not enabled by the reference configs")
        if act_layer is not nn.GELU:
            raise NotImplementedError("the fc1 epilogue implements exact-erf GELU only")
        self.tuning_config = tuning_config
        self.norm1 = norm_layer(dim)
        self.attn = SegAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_norm=qk_norm,
                                 attn_drop=attn_drop, proj_drop=proj_drop, norm_layer=norm_layer,
                                 window_size=window_size)
        self.ls1 = nn.Identity()
        self.drop_path1 = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = mlp_layer(in_features=dim, hidden_features=int(dim * mlp_ratio),
                             act_layer=act_layer, drop=proj_drop)
        self.ls2 = nn.Identity()
        self.drop_path2 = nn.Identity()
        self.adaptmlp = Adapter(self.tuning_config, dropout=0.1, bottleneck=tuning_config.ffn_num,
                                init_option=tuning_config.ffn_adapter_init_option,
                                adapter_scalar=tuning_config.ffn_adapter_scalar,
                                adapter_layernorm_option=tuning_config.ffn_adapter_layernorm_option)
        self.mlp_token_select = TokenSelect(dim, num_sub_layer=1) if select else None

    def forward(self, x):
        if self.mlp_token_select is None:
            raise NotImplementedError("dyt_b200 segmentation block: every reference config selects in "
                                      "all layers (keep_layers = 0)")
        _no_backward("Block", x, *self.parameters())
        if self.training:
            raise NotImplementedError("dyt_b200 segmentation backbone: inference only (model.eval())")
        out, masks, logits, _ = engine.run_blocks(x, [self], eps=float(self.norm1.eps),
                                                  fuse_next_ln=False, attn_biases=[self.attn.bias()])
        dt = _act_dtype()
        return out, dict(sub_token_select=masks[0].unsqueeze(-1).to(dt),
                         token_logits=logits[0].unsqueeze(-1).to(dt))


class TokenRateLoss(nn.Module):
    """AdaLoss of the segmentation file (reference :30-87): token_ratio * ((mean keep - target)^2 +
    w * sum(clamp(minimal - keep, 0)))."""

    def __init__(self, token_target_ratio=0.5, token_loss_ratio=2.0, token_minimal=0.1,
                 token_minimal_weight=1.0, **unused):
        super().__init__()
        self.token_target_ratio = token_target_ratio
        self.token_loss_ratio = token_loss_ratio
        self.token_minimal = token_minimal
        self.token_minimal_weight = token_minimal_weight

    def forward(self, outputs):
        sel = outputs["token_select"]
        loss = ((sel.mean() - self.token_target_ratio) ** 2).mean()
        if self.token_minimal_weight > 0:
            loss = loss + self.token_minimal_weight * (self.token_minimal - sel.mean(-1)).clamp(min=0.0).sum()
        return self.token_loss_ratio * loss


def resize_pos_embed(pos_embed, src_shape, dst_shape, mode="bicubic", num_extra_tokens=1):
    """Bicubic resize of the patch part of a [1, L, C] position embedding (reference :88-118)."""
    if tuple(src_shape) == tuple(dst_shape):
        return pos_embed
    extra, grid = pos_embed[:, :num_extra_tokens], pos_embed[:, num_extra_tokens:]
    c = grid.shape[-1]
    grid = grid.reshape(1, src_shape[0], src_shape[1], c).permute(0, 3, 1, 2)
    grid = F.interpolate(grid.float(), size=tuple(dst_shape), align_corners=False, mode=mode)
    grid = grid.flatten(2).transpose(1, 2).to(pos_embed.dtype)
    return torch.cat((extra, grid), dim=1)


class VisionTransformer21K(nn.Module):
    def __init__(self, img_size: Union[int, Tuple[int, int]] = 224,
                 patch_size: Union[int, Tuple[int, int]] = 16, in_chans: int = 3,
                 num_classes: int = 1000, global_pool: str = "token", embed_dim: int = 768,
                 depth: int = 12, num_heads: int = 12, mlp_ratio: float = 4.0, qkv_bias: bool = True,
                 qk_norm: bool = False, init_values: Optional[float] = None, class_token: bool = True,
                 no_embed_class: bool = False, pre_norm: bool = False, fc_norm: Optional[bool] = None,
                 drop_rate: float = 0.0, pos_drop_rate: float = 0.0, patch_drop_rate: float = 0.0,
                 proj_drop_rate: float = 0.0, attn_drop_rate: float = 0.0, drop_path_rate: float = 0.0,
                 weight_init: str = "", embed_layer: Callable = PatchEmbed,
                 norm_layer: Optional[Callable] = None, act_layer: Optional[Callable] = None,
                 block_fn: Callable = SegBlock, mlp_layer: Callable = Mlp, tuning_config=None,
                 select_config=None, out_indices=(3, 5, 7, 11), use_rel_pos_bias=False):
        super().__init__()
        if not class_token or no_embed_class or pre_norm or patch_drop_rate > 0 or pos_drop_rate > 0:
            raise NotImplementedError("dyt_b200 keeps the reference configuration: class token, "
                                      "embedded class position, no pre-norm, no patch / pos dropout")
        self.tuning_config = tuning_config
        self.select_config = select_config
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        act_layer = act_layer or nn.GELU
        self.num_classes = num_classes
        self.global_pool = global_pool
        self.num_features = self.embed_dim = embed_dim
        self.num_prefix_tokens = 1
        self.no_embed_class = no_embed_class
        self.grad_checkpointing = False
        self.patch_embed = embed_layer(img_size=img_size, patch_size=patch_size, in_chans=in_chans,
                                       embed_dim=embed_dim, bias=True)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, num_patches + 1, embed_dim) * 0.02)
        self.pos_drop = nn.Dropout(p=pos_drop_rate)
        self.patch_drop = nn.Identity()
        self.norm_pre = nn.Identity()
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.Sequential(*[
            block_fn(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                     qk_norm=qk_norm, init_values=init_values, proj_drop=proj_drop_rate,
                     attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer,
                     act_layer=act_layer, mlp_layer=mlp_layer,
                     window_size=self.patch_embed.grid_size if use_rel_pos_bias else None,
                     tuning_config=tuning_config,
                     select=select_config.open and i >= select_config.keep_layers)
            for i in range(depth)])
        self.fpn1 = nn.Sequential(nn.ConvTranspose2d(embed_dim, embed_dim, kernel_size=2, stride=2),
                                  nn.GELU(),
                                  nn.ConvTranspose2d(embed_dim, embed_dim, kernel_size=2, stride=2))
        self.fpn2 = nn.Sequential(nn.ConvTranspose2d(embed_dim, embed_dim, kernel_size=2, stride=2))
        self.fpn3 = nn.Identity()
        self.fpn4 = nn.MaxPool2d(kernel_size=2, stride=2)
        self.out_indices = list(out_indices)
        self._register_load_state_dict_pre_hook(self._prepare_pos_embed)
        nn.init.normal_(self.cls_token, std=1e-6)
        self.apply(self.init_weights)
        def g(key, default):       # EasyDict-like configs raise KeyError for missing attributes
            try:
                return getattr(select_config, key)
            except (AttributeError, KeyError):
                return default
        self.token_loss = TokenRateLoss(token_target_ratio=g("token_target_ratio", 0.5),
                                        token_loss_ratio=g("token_ratio", 2.0),
                                        token_minimal=g("token_minimal", 0.1),
                                        token_minimal_weight=g("token_minimal_weight", 1.0))

    def init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif hasattr(m, "_init_weights"):
            m._init_weights()

    def _prepare_pos_embed(self, state_dict, prefix, *args, **kwargs):
        name = prefix + "pos_embed"
        if name in state_dict and state_dict[name].shape != self.pos_embed.shape:
            src = int(round((state_dict[name].shape[1] - 1) ** 0.5))
            dst = int(round((self.pos_embed.shape[1] - 1) ** 0.5))
            state_dict[name] = resize_pos_embed(state_dict[name], (src, src), (dst, dst))

    @staticmethod
    def resize_pos_embed(*args, **kwargs):
        return resize_pos_embed(*args, **kwargs)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed", "cls_token", "dist_token"}

    def forward_features(self, x):
        if self.training:
            raise NotImplementedError("dyt_b200 segmentation backbone: inference only (model.eval() "
                                      "under torch.no_grad())")
        _no_backward("VisionTransformer21K", x, *self.blocks.parameters())
        if not x.is_cuda:
            raise DytError("dyt_b200 needs CUDA inputs (no CPU fallback)")
        for blk in self.blocks:
            if blk.mlp_token_select is None:
                raise NotImplementedError("dyt_b200 segmentation backbone: keep_layers must be 0")
        B, _, H, W = x.shape
        pe = self.patch_embed
        Hp, Wp = H // pe.patch_size[0], W // pe.patch_size[1]
        x = ops.patch_embed(x, pe.proj.weight, pe.proj.bias, self.cls_token, self.pos_embed,
                            pe.patch_size[0])
        blocks = list(self.blocks)
        biases = [blk.attn.bias() for blk in blocks]
        features, masks_all, logits_all = [], [], []
        start = 0
        stops = sorted(set(i for i in self.out_indices if 0 <= i < len(blocks)))
        for stop in stops + ([len(blocks) - 1] if (not stops or stops[-1] != len(blocks) - 1) else []):
            x, masks, logits, _ = engine.run_blocks(x, blocks[start:stop + 1],
                                                    eps=float(blocks[0].norm1.eps),
                                                    attn_biases=biases[start:stop + 1])
            masks_all.append(masks)
            logits_all.append(logits)
            if stop in self.out_indices:
                features.append(x[:, 1:, :].permute(0, 2, 1).reshape(B, -1, Hp, Wp).contiguous())
            start = stop + 1
        dt = _act_dtype()
        masks = torch.cat(masks_all, dim=0)                     # [L, B, N]
        logits = torch.cat(logits_all, dim=0)                   # [L, B, N-1]
        token_select = masks.permute(1, 0, 2)[:, :, 1:].unsqueeze(-1).to(dt)
        token_logits = logits.permute(1, 0, 2).unsqueeze(-1).to(dt)
        heads = [self.fpn1, self.fpn2, self.fpn3, self.fpn4]
        features = [heads[i](f.to(dt) if dt != torch.float32 else f) for i, f in enumerate(features)]
        loss = self.token_loss(dict(token_select=token_select))
        return tuple(features), dict(token_select=token_select, token_logits=token_logits, loss=loss)

    def forward(self, x):
        return self.forward_features(x)


def vit_base_patch16_224_in21k(**kwargs):
    return VisionTransformer21K(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                                qkv_bias=True, **kwargs)
