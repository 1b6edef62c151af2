"""timm-style nn.Module surface of the DyT hot path, backed by the sm_100a kernels.

Same class names, constructor keywords, forward signatures / return types and state_dict keys as
the reference, so `main_image.py`, `main_vtab.py` and `speed.py` import these
through the drop-in `models` package unchanged.  Reference counterparts:
  TokenSelect / Adapter      models/dynamic_adapter.py:58-140 (speed copies model_speed_test.py:41-114)
  Attention                  models/vision_transformer_IN21K.py:27-75
  Block (speed flavour)      models/model_speed_test.py:180-310      forward(x) -> x
  Block (train flavour)      models/vision_transformer_IN21K.py:88-185  forward(x, complete_model)
  VisionTransformer          models/vision_transformer_IN21K.py:192-385, model_speed_test.py:318-496
"""
from __future__ import annotations

import math
from functools import partial
from typing import Callable, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import _lib, engine, ops, train
from ._lib import DytError
from .layers import DropPath, Mlp, PatchDropout, PatchEmbed, trunc_normal_, use_fused_attn


def _act_dtype() -> torch.dtype:
    """dtype the reference would hand back for Linear outputs: fp16 under autocast, else fp32."""
    if torch.is_autocast_enabled():
        return torch.get_autocast_dtype("cuda") if hasattr(torch, "get_autocast_dtype") \
            else torch.get_autocast_gpu_dtype()
    return torch.float32


def _no_backward(what: str, *tensors) -> None:
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError(
            f"dyt_b200 {what}: only the forward pass is implemented (backward is the next scope "
            "row, SURVEY.md section 8f); call under torch.no_grad()")


def _wants_grad(module: nn.Module, x: torch.Tensor) -> bool:
    """True when the call must be differentiable: autograd on and the input or a parameter of the
    module requires grad (the reference's fine-tuning step, engine_finetune.py:47-76)."""
    return torch.is_grad_enabled() and (x.requires_grad or
                                        any(p.requires_grad for p in module.parameters()))


def _gumbel_sigmoid(logits, tau=1, hard=False, eps=1e-10, training=True, threshold=0.5):
    """API twin of models/dynamic_adapter.py:25-54 operating on caller-provided logits.  Kept for
    import compatibility; the fused dispatcher kernel evaluates the same gate on the device."""
    raise DytError("_gumbel_sigmoid is fused into the dispatcher kernel (TokenSelect.forward); "
                   "it is not a standalone op in dyt_b200")


def draw_gumbel_pair(shape, dtype, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """The two Gumbel draws of the train-mode gate, in the reference's order and dtype
    (models/dynamic_adapter.py:30-39): -log(Exp(1)) twice."""
    g1 = -torch.empty(shape, dtype=dtype, device=device).exponential_().log()
    g2 = -torch.empty(shape, dtype=dtype, device=device).exponential_().log()
    return g1, g2


class TokenSelect(nn.Module):
    def __init__(self, dim_in, num_sub_layer, tau=5, is_hard=True, threshold=0.5, bias=True):
        super().__init__()
        if num_sub_layer != 1:
            raise NotImplementedError("TokenSelect: the reference only ever uses num_sub_layer=1")
        self.mlp_head = nn.Linear(dim_in, num_sub_layer, bias=bias)
        self.is_hard = is_hard
        self.tau = tau
        self.threshold = threshold
        # train-mode Gumbel noise as in models/dynamic_adapter.py:25-54; the speed model's gate
        # (models/model_speed_test.py:27-38) never draws noise, whatever the module mode
        self.noise_in_training = True

    def set_tau(self, tau):
        self.tau = tau

    def forward(self, x, noise=None):
        """x [B, N, C] -> (token_select [B, N, 1] with cls = 1, logits [B, N-1, 1])."""
        _no_backward("TokenSelect", x, self.mlp_head.weight)
        if not self.is_hard:
            raise NotImplementedError("TokenSelect(is_hard=False) is never used by the reference")
        dt = _act_dtype()
        if self.training and self.noise_in_training and noise is None:
            noise = draw_gumbel_pair((x.shape[0], x.shape[1] - 1, 1),
                                     dt if dt != torch.float32 else torch.float32, x.device)
        bias = self.mlp_head.bias if self.mlp_head.bias is not None else x.new_zeros(1)
        r = ops.dispatch(x.float(), self.mlp_head.weight, bias,
                         logit_dtype=torch.float16 if dt != torch.float32 else torch.float32,
                         threshold=float(self.threshold), noise=noise, tau=float(self.tau),
                         pack=False)
        return r["mask"].to(dt), r["logits"].to(dt)


class Adapter(nn.Module):
    def __init__(self, config=None, d_model=None, bottleneck=None, dropout=0.0, init_option="bert",
                 adapter_scalar="1.0", adapter_layernorm_option="in"):
        super().__init__()
        self.n_embd = config.d_model if d_model is None else d_model
        self.down_size = config.attn_bn if bottleneck is None else bottleneck
        self.adapter_layernorm_option = adapter_layernorm_option
        self.adapter_layer_norm_before = None
        if adapter_layernorm_option in ("in", "out"):
            self.adapter_layer_norm_before = nn.LayerNorm(self.n_embd)
        if adapter_scalar == "learnable_scalar":
            self.scale = nn.Parameter(torch.ones(1))
        else:
            self.scale = float(adapter_scalar)
        self.down_proj = nn.Linear(self.n_embd, self.down_size)
        self.non_linear_func = nn.ReLU()
        self.up_proj = nn.Linear(self.down_size, self.n_embd)
        self.dropout = dropout

    def _init_weights(self):
        with torch.no_grad():
            nn.init.kaiming_uniform_(self.down_proj.weight, a=math.sqrt(5))
            nn.init.zeros_(self.up_proj.weight)
            nn.init.zeros_(self.down_proj.bias)
            nn.init.zeros_(self.up_proj.bias)

    def forward(self, x, add_residual=True, residual=None):
        _no_backward("Adapter", x, self.down_proj.weight, self.up_proj.weight)
        if self.adapter_layernorm_option in ("in", "out"):
            raise NotImplementedError("adapter LayerNorm options 'in'/'out' are dead code in every "
                                      "reference entry script (ffn_adapter_layernorm_option='none')")
        if self.training and self.dropout > 0:
            raise NotImplementedError("train-mode adapter dropout needs the backward path (next row)")
        dt = _act_dtype()
        h16 = torch.float16
        down, _ = ops.linear_f16(x.to(h16), self.down_proj.weight.to(h16),
                                 self.down_proj.bias.to(h16), epilogue=_lib.EPI_BIAS_RELU)
        scale = float(self.scale.item()) if torch.is_tensor(self.scale) else self.scale
        up, _ = ops.linear_f16(down, self.up_proj.weight.to(h16), self.up_proj.bias.to(h16),
                               epilogue=_lib.EPI_BIAS, scale=scale)
        up = up.to(dt)
        if add_residual:
            up = up + (x if residual is None else residual)
        return up


class MoEAdapter(nn.Module):
    """MoE-adapter of the DyT paper (arXiv 2403.11808; BASELINE configs[3]).  NOT part of the reference
    repository: the arithmetic is this repository's own restatement
    (oracle/dyt_oracle.py `moe_adapter`), there is no reference parity.
      alpha = softmax(router(mean over the tokens of x));  W_mix = sum_i alpha_i W^i (down, up, biases)
      out = scale * (relu(x W_down_mix^T + b_down_mix) W_up_mix^T + b_up_mix)
    Enabled by `tuning_config.moe_experts > 1`; inference only (the block forward computes it through
    dyt_moe_adapter_fwd)."""

    def __init__(self, config=None, d_model=None, bottleneck=None, num_experts=4, dropout=0.0,
                 adapter_scalar="1.0"):
        super().__init__()
        self.n_embd = config.d_model if d_model is None else d_model
        self.down_size = config.attn_bn if bottleneck is None else bottleneck
        self.num_experts = int(num_experts)
        if not 2 <= self.num_experts <= 8:
            raise NotImplementedError("MoEAdapter: 2..8 experts")
        if adapter_scalar == "learnable_scalar":
            raise NotImplementedError("MoEAdapter: learnable_scalar is not built")
        self.scale = float(adapter_scalar)
        self.router = nn.Linear(self.n_embd, self.num_experts)
        self.down_proj = nn.ModuleList(nn.Linear(self.n_embd, self.down_size) for _ in range(self.num_experts))
        self.non_linear_func = nn.ReLU()
        self.up_proj = nn.ModuleList(nn.Linear(self.down_size, self.n_embd) for _ in range(self.num_experts))
        self.dropout = dropout

    def forward(self, x, add_residual=True, residual=None):
        raise NotImplementedError("MoEAdapter runs inside the block forward (dyt_block_fwd); it has no "
                                  "stand-alone forward")


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_norm=False, attn_drop=0.0,
                 proj_drop=0.0, norm_layer=nn.LayerNorm):
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        if qk_norm or attn_drop > 0 or proj_drop > 0:
            raise NotImplementedError("qk_norm / attention dropout are never enabled by the reference")
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.fused_attn = use_fused_attn()
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.q_norm = nn.Identity()
        self.k_norm = nn.Identity()
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        _no_backward("Attention", x, self.qkv.weight, self.proj.weight)
        h16 = torch.float16
        B, N, _ = x.shape
        qkv, _ = ops.linear_f16(x.to(h16), self.qkv.weight.to(h16),
                                None if self.qkv.bias is None else self.qkv.bias.to(h16))
        o = ops.attn_varlen(qkv.reshape(B, N, -1), self.num_heads)
        y, _ = ops.linear_f16(o, self.proj.weight.to(h16), self.proj.bias.to(h16))
        return y.to(_act_dtype())


class LayerScale(nn.Module):
    def __init__(self, dim, init_values=1e-5, inplace=False):
        super().__init__()
        raise NotImplementedError("LayerScale (init_values) is never enabled by the reference")


class _BlockBase(nn.Module):
    """Parameters and sub-module names of the reference Block; compute goes through engine."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_norm=False, proj_drop=0.0,
                 attn_drop=0.0, init_values=None, drop_path=0.0, act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm, mlp_layer=Mlp, tuning_config=None, layer_id=None,
                 select=False):
        super().__init__()
        if init_values:
            raise NotImplementedError("LayerScale is never enabled by the reference entry scripts")
        if act_layer is not nn.GELU:
            raise NotImplementedError("the fc1 epilogue implements exact-erf GELU only")
        self.tuning_config = tuning_config
        self.layer_id = layer_id
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_norm=qk_norm,
                              attn_drop=attn_drop, proj_drop=proj_drop, norm_layer=norm_layer)
        self.ls1 = nn.Identity()
        self.drop_path1 = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = mlp_layer(in_features=dim, hidden_features=int(dim * mlp_ratio),
                             act_layer=act_layer, drop=proj_drop)
        self.ls2 = nn.Identity()
        self.drop_path2 = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        n_exp = int(tuning_config.get("moe_experts", 0) if hasattr(tuning_config, "get")
                    else getattr(tuning_config, "moe_experts", 0))
        if n_exp > 1:   # MoE-adapter (BASELINE configs[3]; not in the reference, see MoEAdapter)
            self.adaptmlp = MoEAdapter(self.tuning_config, bottleneck=tuning_config.ffn_num,
                                       num_experts=n_exp, dropout=0.1,
                                       adapter_scalar=tuning_config.ffn_adapter_scalar)
        else:
            self.adaptmlp = Adapter(self.tuning_config, dropout=0.1, bottleneck=tuning_config.ffn_num,
                                    init_option=tuning_config.ffn_adapter_init_option,
                                    adapter_scalar=tuning_config.ffn_adapter_scalar,
                                    adapter_layernorm_option=tuning_config.ffn_adapter_layernorm_option)

    def _eps(self) -> float:
        return float(self.norm1.eps)

    def _run(self, x, forced_mask=None, report_gate=False):
        out, masks, logits, _ = engine.run_blocks(
            x, [self], eps=self._eps(), forced_masks=None if forced_mask is None else [forced_mask],
            fuse_next_ln=False, report_gate=report_gate)
        return out, masks[0].unsqueeze(-1), logits[0].unsqueeze(-1)


class SpeedBlock(_BlockBase):
    """Sparse inference block: reference models/model_speed_test.py Block (forward(x) -> x)."""

    def __init__(self, *args, select=False, **kwargs):
        super().__init__(*args, select=select, **kwargs)
        if select:
            self.mlp_token_select = TokenSelect(self.attn.qkv.in_features, num_sub_layer=1)
            self.mlp_token_select.noise_in_training = False   # model_speed_test.py:27-38
        else:
            # reference quirk (model_speed_test.py:229-232): without a selector the block cannot run
            self.token_select = None

    def forward(self, x):
        if not hasattr(self, "mlp_token_select"):
            raise AttributeError("'Block' object has no attribute 'mlp_token_select' "
                                 "(select=False; every reference entry script uses keep_layers=0)")
        _no_backward("Block (speed flavour: inference only; fine-tune with "
                     "models.vision_transformer_IN21K)", x, *self.parameters())
        out, _, _ = self._run(x)
        return out

    # same entry points as the reference (B == 1 and B > 1 share the packed kernel path here)
    single_forward = forward
    batch_forward = forward


class TrainBlock(_BlockBase):
    """Dense masked block: reference models/vision_transformer_IN21K.py Block.
    forward(x, complete_model=False) -> (x, dict(sub_token_select, token_logits)).  In eval mode the
    masked dense form equals the sparse form, which is what runs here."""

    def __init__(self, *args, select=False, **kwargs):
        super().__init__(*args, select=select, **kwargs)
        self.mlp_token_select = TokenSelect(self.attn.qkv.in_features, num_sub_layer=1)  # `select` ignored, as in the reference
        self.count_flops = None
        self.token_select_num = None

    def forward(self, x, complete_model=False):
        if self.count_flops:
            return self.forward_count_flops(x)
        B, N, _ = x.shape
        if _wants_grad(self, x) or self.training:
            # fine-tuning: dense masked block with autograd (Gumbel gate + adapter dropout in
            # train() mode), forward and backward on the sm_100a kernels (dyt_b200.train).  A
            # train()-mode call without autograd (no_grad, everything frozen) takes the same path:
            # the reference draws the Gumbel noise and the adapter dropout whenever self.training
            out, sel, logit = train.block_train(self, x, complete_model)
            dt = _act_dtype()
            return out, dict(sub_token_select=sel.to(dt), token_logits=logit.to(dt))
        if complete_model:
            # teacher pass: MLP on every token; the selector's own decision is still reported
            out, sel, logit = self._run(x, forced_mask=torch.ones(B, N, device=x.device),
                                        report_gate=True)
        else:
            out, sel, logit = self._run(x)
        dt = _act_dtype()
        return out, dict(sub_token_select=sel.to(dt), token_logits=logit.to(dt))

    def forward_count_flops(self, x):
        """FLOP probe for block_flops_dict.get_block_flops (reference
        models/vision_transformer_IN21K.py:167-185): MLP on the first `token_select_num` tokens.
        Implemented with the imposed-mask path of the kernels."""
        assert self.token_select_num is not None
        B, N, _ = x.shape
        fm = torch.zeros(B, N, device=x.device)
        fm[:, :self.token_select_num] = 1.0
        out, _, _ = self._run(x, forced_mask=fm)
        return out


class VisionTransformer(nn.Module):
    """ViT with DyT blocks.  `flavour` = "speed" (forward(x) -> logits) or "train"
    (forward(x, complete_model=False) -> (logits, dict(token_select, token_logits)))."""

    flavour = "speed"
    block_cls = SpeedBlock

    def __init__(self, img_size: Union[int, Tuple[int, int]] = 224,
                 patch_size: Union[int, Tuple[int, int]] = 16, in_chans: int = 3,
                 num_classes: int = 1000, global_pool: str = "token", embed_dim: int = 768,
                 depth: int = 12, num_heads: int = 12, mlp_ratio: float = 4.0,
                 qkv_bias: bool = True, qk_norm: bool = False, init_values: Optional[float] = None,
                 class_token: bool = True, no_embed_class: bool = False, pre_norm: bool = False,
                 fc_norm: Optional[bool] = None, drop_rate: float = 0.0, pos_drop_rate: float = 0.0,
                 patch_drop_rate: float = 0.0, proj_drop_rate: float = 0.0,
                 attn_drop_rate: float = 0.0, drop_path_rate: float = 0.0, weight_init: str = "",
                 embed_layer: Callable = PatchEmbed, norm_layer: Optional[Callable] = None,
                 act_layer: Optional[Callable] = None, block_fn: Optional[Callable] = None,
                 mlp_layer: Callable = Mlp, tuning_config=None, select_config=None):
        super().__init__()
        assert global_pool in ("", "avg", "token")
        assert class_token or global_pool != "token"
        self.tuning_config = tuning_config
        use_fc_norm = global_pool == "avg" if fc_norm is None else fc_norm
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        act_layer = act_layer or nn.GELU
        block_fn = block_fn or self.block_cls
        self.num_classes = num_classes
        self.global_pool = global_pool
        self.num_features = self.embed_dim = embed_dim
        self.num_prefix_tokens = 1 if class_token else 0
        self.no_embed_class = no_embed_class
        self.grad_checkpointing = False
        if not class_token or no_embed_class or pre_norm:
            raise NotImplementedError("dyt_b200 keeps the reference configuration: class token, "
                                      "embedded class position, no pre-norm")
        self.patch_embed = embed_layer(img_size=img_size, patch_size=patch_size, in_chans=in_chans,
                                       embed_dim=embed_dim, bias=not pre_norm)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, num_patches + 1, embed_dim) * 0.02)
        self.pos_drop = nn.Dropout(p=pos_drop_rate)
        self.patch_drop = PatchDropout(patch_drop_rate, 1) if patch_drop_rate > 0 else nn.Identity()
        self.norm_pre = nn.Identity()
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.Sequential(*[
            block_fn(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                     qk_norm=qk_norm, init_values=init_values, proj_drop=proj_drop_rate,
                     attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer,
                     act_layer=act_layer, mlp_layer=mlp_layer, tuning_config=tuning_config,
                     layer_id=i, select=select_config.open and i >= select_config.keep_layers)
            for i in range(depth)])
        self.norm = norm_layer(embed_dim) if not use_fc_norm else nn.Identity()
        self.fc_norm = norm_layer(embed_dim) if use_fc_norm else nn.Identity()
        self.head_drop = nn.Dropout(drop_rate)
        self.head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        nn.init.normal_(self.cls_token, std=1e-6)
        self.apply(self.init_weights)

    def init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif hasattr(m, "_init_weights"):
            m._init_weights()

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed", "cls_token", "dist_token"}

    # -- stem: patch embedding + cls + position through the sm_100a im2col / GEMM / assemble path --
    def _embed(self, x):
        if torch.is_grad_enabled() and (x.requires_grad or any(
                p.requires_grad for p in (self.cls_token, self.pos_embed, *self.patch_embed.parameters()))):
            raise NotImplementedError("dyt_b200 fine-tuning keeps the stem frozen (patch_embed, "
                                      "cls_token, pos_embed: main_image.py:242-256) and has no "
                                      "gradient w.r.t. the pixels")
        if not x.is_cuda:
            raise DytError("dyt_b200 VisionTransformer needs CUDA inputs (no CPU fallback)")
        pe = self.patch_embed
        if not isinstance(pe, PatchEmbed) or not isinstance(pe.norm, nn.Identity):
            raise NotImplementedError("dyt_b200 stem implements the plain PatchEmbed (Conv2d k=s=P)")
        if tuple(x.shape[-2:]) != tuple(pe.img_size):
            raise ValueError(f"input size {tuple(x.shape[-2:])} != model size {pe.img_size}")
        x = ops.patch_embed(x, pe.proj.weight, pe.proj.bias, self.cls_token, self.pos_embed,
                            pe.patch_size[0])
        return self.norm_pre(self.patch_drop(self.pos_drop(x)))

    def _pooled_norm(self, x):
        """norm() then pooling == pooling then norm() for the 'token' pool (LayerNorm is per row):
        normalise only the cls rows instead of all B*N tokens."""
        if self.global_pool == "token" and isinstance(self.fc_norm, nn.Identity):
            return self.head_drop(self.norm(x[:, 0]))
        return None

    def _cls_head(self, tokens):
        """Final LayerNorm of the cls rows + classifier head on the sm_100a kernels (reference
        vision_transformer_IN21K.py:371-380 / model_speed_test.py:483-491 under fp16 autocast):
        row-gathering LayerNorm kernel -> tcgen05 GEMM.  Returns None when the configuration is not
        the plain 'token' pool + nn.Linear head under fp16 autocast (then the caller falls back to
        the generic module path)."""
        if not (self.global_pool == "token" and isinstance(self.fc_norm, nn.Identity) and
                isinstance(self.norm, nn.LayerNorm) and isinstance(self.head, nn.Linear) and
                _act_dtype() == torch.float16 and self.head.bias is not None and
                not (self.training and self.head_drop.p > 0)):
            return None
        B, N, Cd = tokens.shape
        key = (self.head.weight.data_ptr(), self.head.weight._version, self.head.bias._version)
        cache = self.__dict__.get("_dyt_head")
        if cache is None or cache[0] != key:
            nc = self.head.out_features
            n8 = (nc + 7) // 8 * 8          # the GEMM wants N % 8 == 0: zero-padded classes
            w16 = torch.zeros((n8, Cd), dtype=torch.float16, device=tokens.device)
            b16 = torch.zeros((n8,), dtype=torch.float16, device=tokens.device)
            w16[:nc] = self.head.weight.detach().to(torch.float16)
            b16[:nc] = self.head.bias.detach().to(torch.float16)
            cache = (key, w16, b16, nc)
            self.__dict__["_dyt_head"] = cache
        _, w16, b16, nc = cache
        # row indices of the cls tokens, one small tensor per (B, N, device), never freed: a CUDA
        # graph captured for one batch size keeps reading its tensor after another size was seen
        # (a single-entry cache here was a use-after-free: replaying the 256-image graph after a
        # 128-image call faulted)
        icache = self.__dict__.setdefault("_dyt_cls_idx", {})
        ikey = (B, N, tokens.device)
        idx = icache.get(ikey)
        if idx is None:
            idx = torch.arange(B, device=tokens.device, dtype=torch.int32) * N
            icache[ikey] = idx
        cls_n = ops.layernorm_f16(tokens.float(), self.norm.weight.detach().float(),
                                  self.norm.bias.detach().float(), float(self.norm.eps), row_idx=idx)
        logits, _ = ops.linear_f16(cls_n, w16, b16)
        return logits[:, :nc]

    def _blocks(self, x, complete_model=False, consume_input=True):
        """x: the stem's output (a temporary of this forward: the blocks may overwrite it)."""
        _no_backward("VisionTransformer (inference path)", x, *self.blocks.parameters())
        for blk in self.blocks:
            if not hasattr(blk, "mlp_token_select"):
                raise AttributeError("'Block' object has no attribute 'mlp_token_select'")
        B, N, _ = x.shape
        forced = None
        if complete_model:
            ones = torch.ones(B, N, device=x.device)
            forced = [ones] * len(self.blocks)
        # complete_model: the reported masks are the selectors' decisions on the teacher's activations
        x, masks, logits, _ = engine.run_blocks(x.float(), list(self.blocks),
                                                eps=float(self.blocks[0].norm1.eps),
                                                forced_masks=forced, report_gate=complete_model,
                                                consume_input=consume_input)
        return x, masks, logits

    def forward_head(self, x, pre_logits: bool = False):
        if self.global_pool:
            x = x[:, self.num_prefix_tokens:].mean(dim=1) if self.global_pool == "avg" else x[:, 0]
        x = self.fc_norm(x)
        x = self.head_drop(x)
        return x if pre_logits else self.head(x)


class SpeedVisionTransformer(VisionTransformer):
    flavour = "speed"
    block_cls = SpeedBlock

    def forward_features(self, x):
        x, _, _ = self._blocks(self._embed(x))
        return self.norm(x)

    def forward(self, x):
        tokens, _, _ = self._blocks(self._embed(x))
        logits = self._cls_head(tokens)
        if logits is not None:
            return logits
        pooled = self._pooled_norm(tokens)
        if pooled is not None:
            return self.head(pooled)
        return self.forward_head(self.norm(tokens))


class TrainVisionTransformer(VisionTransformer):
    flavour = "train"
    block_cls = TrainBlock

    def forward_features(self, x, complete_model=False):
        if _wants_grad(self, x):
            raise NotImplementedError("dyt_b200: differentiable forward_features is not built; the "
                                      "fine-tuning step goes through forward()")
        x, masks, logits = self._blocks(self._embed(x), complete_model)
        dt = _act_dtype()
        # [L, B, N] -> [B, L, N-1, 1] without the cls slot (vision_transformer_IN21K.py:367-368)
        token_select = masks.permute(1, 0, 2)[:, :, 1:].unsqueeze(-1).to(dt)
        token_logits = logits.permute(1, 0, 2).unsqueeze(-1).to(dt)
        return self.norm(x), dict(token_select=token_select, token_logits=token_logits)

    def _forward_finetune(self, x, complete_model=False):
        """The differentiable path of the fine-tuning step (engine_finetune.py:47-76): student
        (complete_model=False) and teacher (True) passes both carry gradients to the adapters."""
        x = self._embed(x)
        sels, lgs = [], []
        B, N, _ = x.shape
        noises = mults = [None] * len(self.blocks)
        if self.training and not (train._fixed["noises"] or train._fixed["drop_mults"]):
            noises, mults = train.draw_pass_randomness(list(self.blocks), B, N, x.device)
        xn = None          # f16(norm1(x)) of the next block, emitted by the merge kernel of this one
        blocks = list(self.blocks)
        for i, blk in enumerate(blocks):
            nxt = blocks[i + 1] if i + 1 < len(blocks) else None
            x, sel, lg, xn = train.block_train(blk, x, complete_model, noises[i], mults[i], xn=xn,
                                               next_block=nxt, return_xn=True)
            sels.append(sel[:, 1:])
            lgs.append(lg)
        dt = _act_dtype()
        token_select = dict(token_select=torch.stack(sels, dim=1).to(dt),
                            token_logits=torch.stack(lgs, dim=1).to(dt))
        return x, token_select

    def forward(self, x, complete_model=False):
        if _wants_grad(self, x) or self.training:   # train() mode keeps its semantics without autograd
            if not (self.global_pool == "token" and isinstance(self.fc_norm, nn.Identity) and
                    isinstance(self.head, nn.Linear) and self.head_drop.p == 0):
                raise NotImplementedError("dyt_b200 fine-tuning implements the reference head: token "
                                          "pool, final norm, nn.Linear head, no head dropout")
            tokens, token_select = self._forward_finetune(x, complete_model)
            return train.head_train(self, tokens).to(_act_dtype()), token_select
        tokens, masks, logits = self._blocks(self._embed(x), complete_model)
        dt = _act_dtype()
        token_select = dict(token_select=masks.permute(1, 0, 2)[:, :, 1:].unsqueeze(-1).to(dt),
                            token_logits=logits.permute(1, 0, 2).unsqueeze(-1).to(dt))
        cls_logits = self._cls_head(tokens)
        if cls_logits is not None:
            return cls_logits, token_select
        pooled = self._pooled_norm(tokens)
        if pooled is not None:
            return self.head(pooled), token_select
        return self.forward_head(self.norm(tokens)), token_select


# ---------------------------------------------------------------------------------------------
# video model (reference video_models/video_vision_transformer_IN21K.py): the image blocks applied
# to every frame, then an attentive pooling head over all t*N tokens of a clip
# ---------------------------------------------------------------------------------------------
class CrossAttention(nn.Module):
    """Parameter container + kernel-backed forward of the reference CrossAttention (:52-110)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0,
                 proj_drop=0.0, attn_head_dim=None):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads if attn_head_dim is None else attn_head_dim
        if head_dim != 64:
            raise NotImplementedError("dyt_b200 pooling head implements head_dim 64")
        if attn_drop > 0 or proj_drop > 0:
            raise NotImplementedError("dyt_b200 pooling head: dropout is inference-dead (p = 0)")
        all_head_dim = head_dim * num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.q = nn.Linear(dim, all_head_dim, bias=False)
        self.k = nn.Linear(dim, all_head_dim, bias=False)
        self.v = nn.Linear(dim, all_head_dim, bias=False)
        if qkv_bias:
            self.q_bias = nn.Parameter(torch.zeros(all_head_dim))
            self.v_bias = nn.Parameter(torch.zeros(all_head_dim))
        else:
            self.q_bias = None
            self.k_bias = None
            self.v_bias = None
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(all_head_dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x, k=None, v=None):
        """x [b, 1, C] fp16/fp32 normalised query tokens, k / v [b, n_keys, C] fp16 normalised tokens."""
        _no_backward("CrossAttention", x, self.q.weight)
        h16 = torch.float16
        b = k.shape[0]
        if x.shape[1] != 1:
            raise NotImplementedError("dyt_b200 pooling head implements one query token per clip")
        qb = None if self.q_bias is None else self.q_bias.to(h16)
        vb = None if self.v_bias is None else self.v_bias.to(h16)
        zb = None if vb is None else torch.zeros_like(vb)
        q, _ = ops.linear_f16(x.reshape(b, -1).to(h16), self.q.weight.to(h16), qb)
        q = (q * self.scale).to(h16)                       # fp16 tensor * python float (:101)
        kk, _ = ops.linear_f16(k.to(h16), self.k.weight.to(h16), zb)
        vv, _ = ops.linear_f16(v.to(h16), self.v.weight.to(h16), vb)
        o = ops.query_attn(q, kk, vv, self.num_heads)
        out, _ = ops.linear_f16(o, self.proj.weight.to(h16), self.proj.bias.to(h16))
        out = out.reshape(b, 1, -1)
        return out if _act_dtype() == h16 else out.float()


class AttentiveBlock(nn.Module):
    """Reference AttentiveBlock (:27-49): norm_q / norm_k / norm_v + CrossAttention."""

    def __init__(self, dim, num_heads, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, norm_layer=nn.LayerNorm, attn_head_dim=None):
        super().__init__()
        self.norm_q = norm_layer(dim)
        self.norm_k = norm_layer(dim)
        self.norm_v = norm_layer(dim)
        self.cross_attn = CrossAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias,
                                         qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop,
                                         attn_head_dim=attn_head_dim)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x_q, x_kv, _pre_norm=None):
        """x_q [b, 1, C]; x_kv [b, n_keys, C] fp32 tokens.  `_pre_norm` = (weight, bias) of a
        LayerNorm to apply to x_kv first inside the same kernel (the model's final `norm`)."""
        if not x_kv.is_cuda:
            raise DytError("dyt_b200 AttentiveBlock needs CUDA tensors (no CPU fallback)")
        eps = float(self.norm_k.eps)
        if _pre_norm is None:
            raise NotImplementedError("dyt_b200 AttentiveBlock is driven by the video model "
                                      "(final norm fused with norm_k / norm_v)")
        xk, xv = ops.pool_layernorm_f16(x_kv.float(), _pre_norm, (self.norm_k.weight, self.norm_k.bias),
                                        (self.norm_v.weight, self.norm_v.bias), eps)
        q_in = ops.layernorm_f16(x_q.float().contiguous(), self.norm_q.weight.detach().float(),
                                 self.norm_q.bias.detach().float(), eps)
        return self.cross_attn(q_in, k=xk, v=xv)


class VideoVisionTransformer(TrainVisionTransformer):
    """Reference video VisionTransformer (:279-483): forward(x [b, c, t, h, w], complete_model) ->
    (logits, dict(token_select [b*t, L, N-1, 1], token_logits))."""
    flavour = "video"

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        embed_dim = self.embed_dim
        num_heads = self.blocks[0].attn.num_heads
        self.query_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        norm_layer = type(self.blocks[0].norm1)
        eps = float(self.blocks[0].norm1.eps)
        self.attentive_blocks = AttentiveBlock(
            dim=embed_dim, num_heads=num_heads, qkv_bias=self.blocks[0].attn.qkv.bias is not None,
            qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0,
            norm_layer=lambda d: norm_layer(d, eps=eps))
        for m in self.attentive_blocks.modules():
            if isinstance(m, nn.Linear):
                trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward_features(self, x, complete_model=False):
        b, c, t, h, w = x.shape
        frames = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)     # "b c t h w -> (b t) c h w"
        return super().forward_features(frames, complete_model)

    def forward(self, x, complete_model=False):
        b, c, t, h, w = x.shape
        frames = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
        tokens, masks, logits = self._blocks(self._embed(frames), complete_model)
        dt = _act_dtype()
        token_select = dict(token_select=masks.permute(1, 0, 2)[:, :, 1:].unsqueeze(-1).to(dt),
                            token_logits=logits.permute(1, 0, 2).unsqueeze(-1).to(dt))
        if not isinstance(self.norm, nn.LayerNorm) or not isinstance(self.fc_norm, nn.Identity):
            raise NotImplementedError("dyt_b200 video head implements norm = LayerNorm, fc_norm = Identity")
        n_tok = tokens.shape[1]
        kv = tokens.reshape(b, t * n_tok, tokens.shape[2])            # "(b t) tokens c -> b (t tokens) c"
        pooled = self.attentive_blocks(self.query_token.expand(b, -1, -1), kv,
                                       _pre_norm=(self.norm.weight, self.norm.bias))[:, 0, :]
        h16 = torch.float16
        if isinstance(self.head, nn.Linear):
            nc = self.head.out_features
            n8 = (nc + 7) // 8 * 8
            w16 = torch.zeros((n8, self.embed_dim), dtype=h16, device=pooled.device)
            b16 = torch.zeros((n8,), dtype=h16, device=pooled.device)
            w16[:nc] = self.head.weight.detach().to(h16)
            if self.head.bias is not None:
                b16[:nc] = self.head.bias.detach().to(h16)
            out, _ = ops.linear_f16(pooled.to(h16).contiguous(), w16, b16)
            out = out[:, :nc]
            return (out if dt == h16 else out.float()), token_select
        return self.head(pooled), token_select
