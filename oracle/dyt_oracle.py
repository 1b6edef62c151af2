"""CPU oracle for the DyT token-dispatched ViT block forward.   *** TEST INFRASTRUCTURE ONLY ***

A plain-PyTorch (CPU) restatement of the reference algorithm, function by function, each citing
the reference file:line it follows (paths relative to NUS-HPC-AI-Lab/Dynamic-Tuning @ d1744f0).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package (dynamic-tuning_b200/) never does.

Parity pin: PINNED.  tests/golden/*.pt hold outputs of the *unmodified reference modules*
(imported from /root/reference through oracle/ref_shim, script: oracle/gen_golden.py) and
tests/test_oracle_golden.py checks every function below against them.

Two arithmetic policies:
  "fp32"  - what the reference computes on CPU (torch.cuda.amp.autocast is a no-op there).
  "amp16" - emulation of what torch.cuda.amp.autocast() does to the reference on a GPU
            (speed.py:254): Linear / attention operands rounded to fp16, products accumulated in
            fp32, results rounded once to fp16; LayerNorm and the residual stream in fp32.
            This is the arithmetic the CUDA kernels implement, so kernel-vs-oracle tolerances are
            tight (accumulation-order noise only).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# arithmetic helpers
# ----------------------------------------------------------------------------------------------
def _r16(x: Tensor) -> Tensor:
    """value held by an fp16 tensor, returned as fp32"""
    return x.to(torch.float16).to(torch.float32)


def linear(x: Tensor, w: Tensor, b: Optional[Tensor], policy: str = "fp32") -> Tensor:
    """nn.Linear forward.  amp16: fp16 operands, fp32 accumulate, bias added before the single
    rounding to fp16 (cuBLASLt bias epilogue).  Returns fp32 tensors holding fp16 values."""
    if policy == "fp32":
        return F.linear(x, w, b)
    y = F.linear(_r16(x), _r16(w), None if b is None else _r16(b))
    return _r16(y)


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    """nn.LayerNorm(eps=1e-6) (reference models/vision_transformer_IN21K.py:262); autocast keeps
    layer_norm in fp32, so both policies agree."""
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), eps)


def gelu(x: Tensor, policy: str = "fp32") -> Tensor:
    """nn.GELU() exact-erf form (timm Mlp act_layer, models/vision_transformer_IN21K.py:261)."""
    y = F.gelu(x)
    return y if policy == "fp32" else _r16(y)


# ----------------------------------------------------------------------------------------------
# a4: _gumbel_sigmoid (reference models/dynamic_adapter.py:25-54; eval copy
#     models/model_speed_test.py:27-37)
# ----------------------------------------------------------------------------------------------
def gumbel_sigmoid_hard(logits: Tensor, tau: float = 5.0, threshold: float = 0.5,
                        training: bool = False,
                        noise: Optional[Tuple[Tensor, Tensor]] = None) -> Tensor:
    """Forward value of the hard straight-through gate, computed in logits.dtype exactly as the
    reference does: y_soft = sigmoid(l) (eval) or sigmoid((l + g1 - g2) / tau) (train, g = Gumbel
    noise drawn by the caller so the test can share it with the kernel);
    y_hard = zeros.masked_fill(y_soft > threshold, 1).  `ret = y_hard - y_soft.detach() + y_soft`
    has the forward value y_hard (up to the rounding of that expression, which the eval-only copy
    model_speed_test.py:27-37 drops)."""
    if training:
        assert noise is not None, "training-mode oracle needs the Gumbel draws"
        g1, g2 = noise
        y_soft = ((logits + g1 - g2) / tau).sigmoid()
    else:
        y_soft = logits.sigmoid()
    return torch.zeros_like(logits).masked_fill(y_soft > threshold, 1.0)


def min_kept_logit(dtype: torch.dtype, threshold: float = 0.5) -> float:
    """Smallest logit (as a python float) that the gate keeps in `dtype`: the gate is monotone, so
    `sigmoid(l) > threshold` evaluated in `dtype` is equivalent to `l >= min_kept_logit`.
    SURVEY.md section 0.4: fp16 -> 2^-10 * (1 + 2^-10), bf16 -> just above 2^-7."""
    if dtype in (torch.float16, torch.bfloat16):
        bits = torch.arange(0, 1 << 16, dtype=torch.int32).to(torch.int16)
        vals = bits.view(dtype)
        keep = gumbel_sigmoid_hard(vals) > 0
        finite = torch.isfinite(vals)
        return float(vals[keep & finite].float().min())
    # fp32: bisection on the monotone gate
    lo, hi = -1.0, 1.0
    lo_t = torch.tensor(lo, dtype=torch.float32)
    hi_t = torch.tensor(hi, dtype=torch.float32)
    for _ in range(80):
        mid = ((lo_t.double() + hi_t.double()) / 2).float()
        if mid == lo_t or mid == hi_t:
            break
        if gumbel_sigmoid_hard(mid.view(1), threshold=threshold).item() > 0:
            hi_t = mid
        else:
            lo_t = mid
    return float(hi_t)


# ----------------------------------------------------------------------------------------------
# a3: TokenSelect.forward (reference models/dynamic_adapter.py:70-77 == model_speed_test.py:53-60)
# ----------------------------------------------------------------------------------------------
def token_select(x: Tensor, w: Tensor, b: Tensor, policy: str = "fp32", tau: float = 5.0,
                 threshold: float = 0.5, training: bool = False,
                 noise: Optional[Tuple[Tensor, Tensor]] = None) -> Tuple[Tensor, Tensor]:
    """x [B,N,C] -> (token_select [B,N,1] in {0,1} with the cls slot forced to 1, logits [B,N-1,1]).
    amp16: logits are an fp16 tensor (fp32-accumulated, rounded once) and the gate runs in fp16."""
    bsz = x.shape[0]
    if policy == "fp32":
        logits = F.linear(x[:, 1:, :], w, b)
    else:
        logits = linear(x[:, 1:, :], w, b, "amp16").to(torch.float16)
        if noise is not None:
            noise = (noise[0].to(torch.float16), noise[1].to(torch.float16))
    sel = gumbel_sigmoid_hard(logits, tau, threshold, training, noise)
    sel = torch.cat([sel.new_ones(bsz, 1, 1), sel], dim=1)
    return sel.float(), logits.float()


# ----------------------------------------------------------------------------------------------
# a9: gather / scatter glue (reference models/model_speed_test.py:297-305)
# ----------------------------------------------------------------------------------------------
def compact(mask: Tensor) -> Tuple[Tensor, Tensor]:
    """mask [B,N,1] -> (packed_idx int64 [T_kept] = flat (b*N+n) indices in ascending order, the
    order nonzero() returns (model_speed_test.py:300); cu_seqlens int32 [B+1])."""
    bsz, n = mask.shape[:2]
    flat = mask.reshape(bsz * n)
    packed_idx = flat.nonzero()[:, 0]
    counts = mask.reshape(bsz, n).sum(dim=1).to(torch.int32)
    cu = torch.zeros(bsz + 1, dtype=torch.int32, device=mask.device)
    cu[1:] = torch.cumsum(counts, 0)
    return packed_idx, cu


# ----------------------------------------------------------------------------------------------
# a5: Adapter.forward (reference models/dynamic_adapter.py:120-140; speed copy
#     models/model_speed_test.py:103-114).  LN option "none", add_residual=False, eval (no dropout)
# ----------------------------------------------------------------------------------------------
def adapter(x: Tensor, p: Dict[str, Tensor], prefix: str, scale: float,
            policy: str = "fp32") -> Tensor:
    down = linear(x, p[prefix + "down_proj.weight"], p[prefix + "down_proj.bias"], policy)
    down = F.relu(down)
    up = linear(down, p[prefix + "up_proj.weight"], p[prefix + "up_proj.bias"], policy)
    up = up * scale
    return up if policy == "fp32" else _r16(up)


# ----------------------------------------------------------------------------------------------
# MoE-adapter (BASELINE configs[3]).  NOT in the reference repository: this is
# a restatement of the published description (DyT paper, arXiv 2403.11808: routing weights from the
# token mean, experts mixed in WEIGHT space, so the cost stays that of one adapter).  PARITY UNPINNED:
# there is no reference implementation, golden vector or test to check it against; the `fp32` policy
# below is the definition, the `amp16` policy states the rounding points the kernels implement.
#   alpha = softmax(router(mean_tokens(x)));  W_mix = sum_i alpha_i W^i (down, up and their biases)
#   out = scale * (relu(x W_down_mix^T + b_down_mix) W_up_mix^T + b_up_mix)
# ----------------------------------------------------------------------------------------------
def moe_adapter(x: Tensor, p: Dict[str, Tensor], prefix: str, scale: float, num_experts: int,
                policy: str = "fp32") -> Tensor:
    E = num_experts
    wd = torch.stack([p[f"{prefix}down_proj.{i}.weight"] for i in range(E)])   # [E, K, C]
    bd = torch.stack([p[f"{prefix}down_proj.{i}.bias"] for i in range(E)])     # [E, K]
    wu = torch.stack([p[f"{prefix}up_proj.{i}.weight"] for i in range(E)])     # [E, C, K]
    bu = torch.stack([p[f"{prefix}up_proj.{i}.bias"] for i in range(E)])       # [E, C]
    mean = x.float().mean(dim=1)                                               # [B, C]
    alpha = torch.softmax(linear(mean, p[prefix + "router.weight"], p[prefix + "router.bias"], policy).float(), -1)
    if policy == "fp32":
        wd_mix = torch.einsum("be,ekc->bkc", alpha, wd)
        wu_mix = torch.einsum("be,eck->bck", alpha, wu)
        down = F.relu(torch.einsum("bnc,bkc->bnk", x, wd_mix) + (alpha @ bd)[:, None, :])
        up = torch.einsum("bnk,bck->bnc", down, wu_mix) + (alpha @ bu)[:, None, :]
        return up * scale
    # amp16: the same mixture applied to the experts' OUTPUTS (both projections are linear in their
    # weights); rounding points of dyt_moe_adapter_fwd: every expert's down output rounded to fp16,
    # mixture + mixed bias in fp32 -> fp16 -> ReLU, operand alpha_i * d rounded to fp16, fp32
    # accumulation of the up projection incl. the bias terms f16(alpha_i) * f16(b_up^i), fp16, * scale
    hid = torch.stack([linear(x, wd[i], None, policy) for i in range(E)])                 # [E, B, N, K]
    a = alpha.t()[:, :, None, None]                                                       # [E, B, 1, 1]
    down = F.relu(_r16((a * hid).sum(0) + (alpha @ _r16(bd))[:, None, :]))
    aup = _r16(a * down[None])                                                            # [E, B, N, K]
    acc = sum(F.linear(aup[i], _r16(wu[i])) for i in range(E))
    acc = acc + (_r16(alpha) @ _r16(bu))[:, None, :]
    return _r16(_r16(acc) * scale)


def block_sparse_moe(x: Tensor, p: Dict[str, Tensor], prefix: str, num_heads: int, scale: float,
                     num_experts: int, policy: str = "fp32",
                     forced_mask: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """block_sparse with the MoE-adapter in place of the plain Adapter (no reference parity)."""
    bsz, n, c = x.shape
    x1 = x + attention(layer_norm(x, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"]), p,
                       prefix + "attn.", num_heads, policy)
    mask, logits = token_select(x1, p[prefix + "mlp_token_select.mlp_head.weight"],
                                p[prefix + "mlp_token_select.mlp_head.bias"], policy)
    if forced_mask is not None:
        mask = forced_mask.float()
    adapt_x = moe_adapter(x1, p, prefix + "adaptmlp.", scale, num_experts, policy)
    packed_idx, cu = compact(mask)
    flat = x1.reshape(bsz * n, c)
    kept = mlp(layer_norm(flat[packed_idx, :], p[prefix + "norm2.weight"], p[prefix + "norm2.bias"]), p,
               prefix + "mlp.", policy)
    mlp_x = torch.zeros(flat.shape, dtype=kept.dtype, device=flat.device)
    mlp_x[packed_idx, :] = kept
    out = adapt_x + (x1 + mlp_x.reshape(bsz, n, c))
    return dict(out=out, x1=x1, mask=mask, logits=logits, adapt=adapt_x)


# ----------------------------------------------------------------------------------------------
# a6: Attention.forward (reference models/vision_transformer_IN21K.py:54-75 ==
#     models/model_speed_test.py:145-166); q_norm/k_norm Identity, dropout 0
# ----------------------------------------------------------------------------------------------
def attention_core(qkv: Tensor, num_heads: int, policy: str = "fp32") -> Tensor:
    """softmax(q k^T / sqrt(d)) v for qkv [B, N, 3*C] laid out [3, H, d] along the last dim
    (models/model_speed_test.py:147-161); returns [B, N, C]."""
    bsz, n, c3 = qkv.shape
    c = c3 // 3
    d = c // num_heads
    qkv = qkv.reshape(bsz, n, 3, num_heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    s = (q @ k.transpose(-2, -1)) * (d ** -0.5)
    if policy == "fp32":
        o = s.softmax(dim=-1) @ v
    else:
        # fused fp16 attention: fp32 scores and statistics, P rounded to fp16 for the PV product,
        # fp32 accumulation, normalisation by the fp32 row sum, one rounding of O to fp16
        m = s.max(dim=-1, keepdim=True).values
        e = torch.exp(s - m)
        o = (_r16(e) @ v) / e.sum(dim=-1, keepdim=True)
        o = _r16(o)
    return o.transpose(1, 2).reshape(bsz, n, c)


def relative_position_bias(table: Tensor, index: Tensor) -> Tensor:
    """[num_heads, N, N] bias of the segmentation backbone's attention (reference dense_tasks/
    Segmentation/backbone/segmentation_vision_transformer_IN21K.py:192-197): table
    [num_relative_distance, heads] gathered by relative_position_index [N, N]."""
    n = index.shape[0]
    return table[index.reshape(-1)].reshape(n, n, -1).permute(2, 0, 1).contiguous()


def attention_bias_core(qkv: Tensor, num_heads: int, bias: Optional[Tensor] = None,
                        policy: str = "fp32") -> Tensor:
    """Eager attention with an additive bias (same file, :181-203): q * scale, q k^T, + bias,
    softmax, @ v.  amp16: q * scale and the scores are fp16 tensors, the bias add and the softmax
    run in fp32, the probabilities are cast to fp16 for the PV product (fp32 accumulate)."""
    bsz, n, c3 = qkv.shape
    c = c3 // 3
    d = c // num_heads
    qkv = qkv.reshape(bsz, n, 3, num_heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    q = q * d ** -0.5                                                   # :189
    if policy == "amp16":
        q = _r16(q)
    s = q @ k.transpose(-2, -1)                                         # :190
    if policy == "amp16":
        s = _r16(s)
    if bias is not None:
        s = s + bias.unsqueeze(0)                                       # :197
    pr = s.softmax(dim=-1)                                              # :199
    if policy == "amp16":
        o = _r16(_r16(pr) @ v)
    else:
        o = pr @ v                                                      # :202
    return o.transpose(1, 2).reshape(bsz, n, c)


def attention(x: Tensor, p: Dict[str, Tensor], prefix: str, num_heads: int,
              policy: str = "fp32") -> Tensor:
    qkv = linear(x, p[prefix + "qkv.weight"], p[prefix + "qkv.bias"], policy)
    o = attention_core(qkv, num_heads, policy)
    return linear(o, p[prefix + "proj.weight"], p[prefix + "proj.bias"], policy)


# ----------------------------------------------------------------------------------------------
# a7: timm.layers.Mlp (timm==0.9.12, not vendored in the reference; call site
#     models/vision_transformer_IN21K.py:124-129): fc2(GELU(fc1(x))), dropouts p=0
# ----------------------------------------------------------------------------------------------
def mlp(x: Tensor, p: Dict[str, Tensor], prefix: str, policy: str = "fp32") -> Tensor:
    h = gelu(linear(x, p[prefix + "fc1.weight"], p[prefix + "fc1.bias"], policy), policy)
    return linear(h, p[prefix + "fc2.weight"], p[prefix + "fc2.bias"], policy)


# ----------------------------------------------------------------------------------------------
# a1: Block.batch_forward, the sparse inference block (reference models/model_speed_test.py:274-310)
# ----------------------------------------------------------------------------------------------
def block_sparse(x: Tensor, p: Dict[str, Tensor], prefix: str, num_heads: int, scale: float,
                 policy: str = "fp32", forced_mask: Optional[Tensor] = None,
                 threshold: float = 0.5) -> Dict[str, Tensor]:
    """Returns dict(out, x1, mask [B,N,1], logits [B,N-1,1], packed_idx, cu_seqlens).
    forced_mask (config 1 of BASELINE.json) overrides the selector's decision."""
    bsz, n, c = x.shape
    x1 = x + attention(layer_norm(x, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"]), p,
                       prefix + "attn.", num_heads, policy)                       # :278
    mask, logits = token_select(x1, p[prefix + "mlp_token_select.mlp_head.weight"],
                                p[prefix + "mlp_token_select.mlp_head.bias"], policy,
                                threshold=threshold)                             # :284
    if forced_mask is not None:
        mask = forced_mask.float()
    adapt_x = adapter(x1, p, prefix + "adaptmlp.", scale, policy)                 # :291
    packed_idx, cu = compact(mask)                                                # :297-300
    flat = x1.reshape(bsz * n, c)
    gather_x = flat[packed_idx, :]                                                # :301
    kept = mlp(layer_norm(gather_x, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"]), p,
               prefix + "mlp.", policy)                                           # :303
    # :302 allocates zeros in the mask's dtype (fp16 under autocast, like the MLP output)
    mlp_x = torch.zeros(flat.shape, dtype=kept.dtype, device=flat.device)
    mlp_x[packed_idx, :] = kept                                                   # :304
    out = adapt_x + (x1 + mlp_x.reshape(bsz, n, c))                               # :305-308
    return dict(out=out, x1=x1, mask=mask, logits=logits, packed_idx=packed_idx, cu_seqlens=cu)


# ----------------------------------------------------------------------------------------------
# a2: Block.forward, the dense masked block in eval mode
#     (reference models/vision_transformer_IN21K.py:144-165)
# ----------------------------------------------------------------------------------------------
def block_dense(x: Tensor, p: Dict[str, Tensor], prefix: str, num_heads: int, scale: float,
                policy: str = "fp32", complete_model: bool = False,
                forced_mask: Optional[Tensor] = None) -> Dict[str, Tensor]:
    x1 = x + attention(layer_norm(x, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"]), p,
                       prefix + "attn.", num_heads, policy)                       # :148
    mask, logits = token_select(x1, p[prefix + "mlp_token_select.mlp_head.weight"],
                                p[prefix + "mlp_token_select.mlp_head.bias"], policy)  # :150-152
    if forced_mask is not None:
        mask = forced_mask.float()
    adapt_x = adapter(x1, p, prefix + "adaptmlp.", scale, policy)                 # :157
    mlp_x = mlp(layer_norm(x1, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"]), p,
                prefix + "mlp.", policy)                                          # :159
    if not complete_model:
        mlp_x = mask * mlp_x                                                      # :161-162
    out = x1 + mlp_x + adapt_x                                                    # :163
    return dict(out=out, x1=x1, mask=mask, logits=logits)


# ----------------------------------------------------------------------------------------------
# a11: Block.forward_count_flops, the FLOP probe block_flops_dict.get_block_flops traces
#      (reference models/vision_transformer_IN21K.py:167-185): MLP on the FIRST token_select_num
#      tokens whatever the selector says
# ----------------------------------------------------------------------------------------------
def block_count_flops(x: Tensor, p: Dict[str, Tensor], prefix: str, num_heads: int, scale: float,
                      token_select_num: int, policy: str = "fp32") -> Tensor:
    x1 = x + attention(layer_norm(x, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"]), p,
                       prefix + "attn.", num_heads, policy)                       # :168
    adapt_x = adapter(x1, p, prefix + "adaptmlp.", scale, policy)                 # :177
    t = token_select_num
    mlp_x = mlp(layer_norm(x1[:, :t, :], p[prefix + "norm2.weight"], p[prefix + "norm2.bias"]), p,
                prefix + "mlp.", policy)                                          # :180
    out = x1 + adapt_x                                                            # :182
    out[:, :t, :] = out[:, :t, :] + mlp_x                                         # :183
    return out


# ----------------------------------------------------------------------------------------------
# a2 / a4 / a5 in train mode, differentiable (fp32): the dense masked block with the hard
# straight-through Gumbel gate and adapter dropout (reference models/vision_transformer_IN21K.py:
# 144-165, models/dynamic_adapter.py:25-54, :127-130).  Gradients come from torch.autograd on this
# restatement; the random draws are inputs so the kernels can share them.
# ----------------------------------------------------------------------------------------------
def gumbel_sigmoid_st(logits: Tensor, tau: float = 5.0, threshold: float = 0.5,
                      training: bool = True, noise: Optional[Tuple[Tensor, Tensor]] = None,
                      hard_override: Optional[Tensor] = None) -> Tensor:
    """models/dynamic_adapter.py:25-54 with hard=True: ret = y_hard - y_soft.detach() + y_soft.
    hard_override (tests only) imposes y_hard, e.g. the decisions of a lower-precision run whose
    borderline tokens rounded the other way."""
    if training:
        g1, g2 = noise
        y_soft = ((logits + g1 - g2) / tau).sigmoid()
    else:
        y_soft = logits.sigmoid()
    y_hard = torch.zeros_like(logits).masked_fill(y_soft > threshold, 1.0)
    if hard_override is not None:
        y_hard = hard_override.to(y_hard.dtype).reshape(y_hard.shape)
    return y_hard - y_soft.detach() + y_soft


def block_train(x: Tensor, p: Dict[str, Tensor], prefix: str, num_heads: int, scale: float,
                noise: Optional[Tuple[Tensor, Tensor]] = None, drop_mult: Optional[Tensor] = None,
                complete_model: bool = False, training: bool = True, tau: float = 5.0,
                threshold: float = 0.5, hard_override: Optional[Tensor] = None,
                relu_override: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """Block.forward of the train model (fp32).  drop_mult = keep / (1 - p) multiplier of the adapter
    dropout (None = no dropout).  Returns dict(out, mask [B,N,1] (differentiable), logits)."""
    x1 = x + attention(layer_norm(x, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"]), p,
                       prefix + "attn.", num_heads, "fp32")                        # :148
    logits = F.linear(x1[:, 1:, :], p[prefix + "mlp_token_select.mlp_head.weight"],
                      p[prefix + "mlp_token_select.mlp_head.bias"])                # dynamic_adapter.py:72
    sel = gumbel_sigmoid_st(logits, tau, threshold, training, noise, hard_override)  # :74
    sel = torch.cat([sel.new_ones(x.shape[0], 1, 1), sel], dim=1)                  # :75
    pre = prefix + "adaptmlp."
    down_pre = F.linear(x1, p[pre + "down_proj.weight"], p[pre + "down_proj.bias"])
    if relu_override is None:
        down = F.relu(down_pre)
    else:
        # tests only: the ReLU's on/off pattern of a lower-precision run (the derivative is a step, so
        # a pre-activation that rounds across zero changes the gradient by a whole term)
        down = down_pre * relu_override.to(down_pre.dtype)
    if drop_mult is not None:
        down = down * drop_mult                                                    # dynamic_adapter.py:129
    adapt_x = F.linear(down, p[pre + "up_proj.weight"], p[pre + "up_proj.bias"]) * scale
    mlp_x = mlp(layer_norm(x1, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"]), p,
                prefix + "mlp.", "fp32")                                           # :159
    if not complete_model:
        mlp_x = sel * mlp_x                                                        # :161-162
    out = x1 + mlp_x + adapt_x                                                     # :163
    return dict(out=out, mask=sel, logits=logits, x1=x1, down_pre=down_pre)


def vit_train_forward(img: Tensor, p: Dict[str, Tensor], depth: int, num_heads: int, scale: float,
                      noises=None, drop_mults=None, complete_model: bool = False,
                      training: bool = True, patch: int = 16) -> Dict[str, Tensor]:
    """Train model forward (reference models/vision_transformer_IN21K.py:343-385), fp32."""
    x = patch_embed(img, p, patch, "fp32")
    x = torch.cat((p["cls_token"].expand(x.shape[0], -1, -1), x), dim=1) + p["pos_embed"]
    sels, logs = [], []
    for i in range(depth):
        r = block_train(x, p, f"blocks.{i}.", num_heads, scale,
                        None if noises is None else noises[i],
                        None if drop_mults is None else drop_mults[i], complete_model, training)
        x = r["out"]
        sels.append(r["mask"])
        logs.append(r["logits"])
    xn = layer_norm(x, p["norm.weight"], p["norm.bias"])
    logits = F.linear(xn[:, 0], p["head.weight"], p["head.bias"])
    return dict(logits=logits, token_select=torch.stack(sels, dim=1)[:, :, 1:, :],
                token_logits=torch.stack(logs, dim=1))


def finetune_loss(student_logits: Tensor, token_select: Tensor, teacher_logits: Tensor,
                  targets: Tensor, token_target_ratio: float = 0.5, token_loss_ratio: float = 2.0,
                  token_minimal: float = 0.0, token_minimal_weight: float = 0.0) -> Tensor:
    """The loss of the fine-tuning step (reference engine_finetune.py:47-65 with models/losses.py:
    50-82 AdaLoss over a cross-entropy base criterion).  Defaults = the entry scripts' recipe
    (main_image.py:206-209: token_minimal 0, token_minimal_weight 0); the golden fixture was made
    with the AdaLoss class defaults (0.1, 1.0), which the tests pass explicitly."""
    kl = F.kl_div(F.log_softmax(student_logits, dim=-1),
                  F.log_softmax(teacher_logits.detach(), dim=-1), reduction="batchmean",
                  log_target=True)
    teacher = F.cross_entropy(teacher_logits, targets)
    base = F.cross_entropy(student_logits, targets)
    flops = ((token_select.mean() - token_target_ratio) ** 2).mean()              # losses.py:71-74
    token_loss = flops
    if token_minimal_weight > 0:                                                  # :76-80
        token_loss = token_loss + token_minimal_weight * (
            token_minimal - token_select.mean(-1)).clamp(min=0.0).sum()
    return base + token_loss_ratio * token_loss + teacher + kl


# ----------------------------------------------------------------------------------------------
# a10: VisionTransformer.forward (reference models/model_speed_test.py:467-496 and
#      models/vision_transformer_IN21K.py:343-385)
# ----------------------------------------------------------------------------------------------
def patch_embed(img: Tensor, p: Dict[str, Tensor], patch: int, policy: str = "fp32") -> Tensor:
    """timm PatchEmbed: Conv2d(3, C, k=patch, s=patch) -> flatten(2).transpose(1,2); as a GEMM over
    unfolded patches so the amp16 rounding points are those of a Linear."""
    w = p["patch_embed.proj.weight"]
    b = p["patch_embed.proj.bias"]
    cols = F.unfold(img, kernel_size=patch, stride=patch).transpose(1, 2)  # [B, L, 3*patch*patch]
    return linear(cols, w.reshape(w.shape[0], -1), b, policy)


def vit_forward(img: Tensor, p: Dict[str, Tensor], depth: int, num_heads: int, scale: float,
                patch: int = 16, policy: str = "fp32", sparse: bool = True,
                complete_model: bool = False, forced_masks=None) -> Dict[str, Tensor]:
    """Whole model.  Returns dict(logits [B,classes], token_select [B,depth,N-1,1],
    token_logits [B,depth,N-1,1], x_final)."""
    x = patch_embed(img, p, patch, policy)
    bsz = x.shape[0]
    x = torch.cat((p["cls_token"].expand(bsz, -1, -1), x), dim=1)   # model_speed_test.py:470-471
    x = x + p["pos_embed"]                                          # :472
    sels, logs = [], []
    for i in range(depth):
        fm = None if forced_masks is None else forced_masks[i]
        if sparse:
            r = block_sparse(x, p, f"blocks.{i}.", num_heads, scale, policy, forced_mask=fm)
        else:
            r = block_dense(x, p, f"blocks.{i}.", num_heads, scale, policy,
                            complete_model=complete_model, forced_mask=fm)
        x = r["out"]
        sels.append(r["mask"])
        logs.append(r["logits"])
    xn = layer_norm(x, p["norm.weight"], p["norm.bias"])            # :483
    logits = linear(xn[:, 0], p["head.weight"], p["head.bias"], policy)  # :486-491
    return dict(logits=logits, token_select=torch.stack(sels, dim=1)[:, :, 1:, :],
                token_logits=torch.stack(logs, dim=1), x_final=x)


# ----------------------------------------------------------------------------------------------
# segmentation backbone, eval mode (reference dense_tasks/Segmentation/backbone/
# segmentation_vision_transformer_IN21K.py: Block.forward :275-298, forward_features :526-560)
# ----------------------------------------------------------------------------------------------
def block_seg(x: Tensor, p: Dict[str, Tensor], prefix: str, num_heads: int, scale: float,
              policy: str = "fp32", forced_mask: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """Dense masked block whose attention is the eager path with an optional relative-position bias."""
    bias = None
    if prefix + "attn.relative_position_bias_table" in p:
        bias = relative_position_bias(p[prefix + "attn.relative_position_bias_table"],
                                      p[prefix + "attn.relative_position_index"])
    xn = layer_norm(x, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"])
    qkv = linear(xn, p[prefix + "attn.qkv.weight"], p[prefix + "attn.qkv.bias"], policy)
    o = attention_bias_core(qkv, num_heads, bias, policy)
    x1 = x + linear(o, p[prefix + "attn.proj.weight"], p[prefix + "attn.proj.bias"], policy)   # :276
    mask, logits = token_select(x1, p[prefix + "mlp_token_select.mlp_head.weight"],
                                p[prefix + "mlp_token_select.mlp_head.bias"], policy)          # :280-282
    if forced_mask is not None:      # test aid: continue with an imposed decision
        mask = forced_mask.float()
    adapt_x = adapter(x1, p, prefix + "adaptmlp.", scale, policy)                              # :287
    mlp_x = mlp(layer_norm(x1, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"]), p,
                prefix + "mlp.", policy)                                                       # :289
    out = x1 + mask * mlp_x + adapt_x                                                          # :291-293
    return dict(out=out, mask=mask, logits=logits)


def seg_forward(img: Tensor, p: Dict[str, Tensor], depth: int, num_heads: int, scale: float,
                out_indices, patch: int = 16, policy: str = "fp32", token_target_ratio: float = 0.5,
                token_ratio: float = 2.0, token_minimal: float = 0.1,
                token_minimal_weight: float = 1.0, forced_masks=None) -> Dict[str, Tensor]:
    """forward_features of the segmentation backbone: tokens -> blocks -> maps after `out_indices`
    -> FPN heads (fpn1: deconv-GELU-deconv, fpn2: deconv, fpn3: identity, fpn4: max-pool) + the
    token-rate loss."""
    bsz, _, h, w = img.shape
    hp, wp = h // patch, w // patch
    x = patch_embed(img, p, patch, policy)
    x = torch.cat((p["cls_token"].expand(bsz, -1, -1), x), dim=1) + p["pos_embed"]
    feats, sels, logs = [], [], []
    for i in range(depth):
        r = block_seg(x, p, f"blocks.{i}.", num_heads, scale, policy,
                      forced_mask=None if forced_masks is None else forced_masks[i])
        x = r["out"]
        sels.append(r["mask"])
        logs.append(r["logits"])
        if i in out_indices:
            feats.append(x[:, 1:, :].permute(0, 2, 1).reshape(bsz, -1, hp, wp).contiguous())   # :549-551
    token_sel = torch.stack(sels, dim=1)[:, :, 1:, :]
    heads = [
        lambda f: F.conv_transpose2d(F.gelu(F.conv_transpose2d(f, p["fpn1.0.weight"], p["fpn1.0.bias"],
                                                             stride=2)),
                                     p["fpn1.2.weight"], p["fpn1.2.bias"], stride=2),
        lambda f: F.conv_transpose2d(f, p["fpn2.0.weight"], p["fpn2.0.bias"], stride=2),
        lambda f: f,
        lambda f: F.max_pool2d(f, kernel_size=2, stride=2),
    ]
    feats = [heads[i](f) for i, f in enumerate(feats)]
    loss = ((token_sel.mean() - token_target_ratio) ** 2).mean()
    if token_minimal_weight > 0:
        loss = loss + token_minimal_weight * (token_minimal - token_sel.mean(-1)).clamp(min=0.0).sum()
    return dict(features=feats, token_select=token_sel, token_logits=torch.stack(logs, dim=1),
                loss=token_ratio * loss)


# ----------------------------------------------------------------------------------------------
# evaluation analytics (reference block_flops_dict.py:57-83, engine_finetune.py:341-352)
# ----------------------------------------------------------------------------------------------
def batch_select_flops(flops_dict: Tensor, token_select: Tensor, block_num: int = 12,
                       base_flops: float = 0.116) -> Tensor:
    """token_select [B, L, N-1, 1] -> per-image GFLOPs [B]: base + sum over the block_num layers of
    flops_dict[kept patch tokens + 1] (layers without a selector count all tokens); fp32 additions
    in layer order, like the reference's loop over torch scalars."""
    ts = token_select.squeeze(-1).float()
    out = []
    for img in ts:                                                  # :80-81
        t = img.shape[1]
        counts = [t] * (block_num - img.shape[0]) + img.sum(-1).int().tolist()   # :63-64
        f = torch.tensor(base_flops, dtype=torch.float32)
        for c in counts:
            f = f + flops_dict[c + 1].float()                       # :66-70
        out.append(f)
    return torch.stack(out)


def layer_keep_rates(token_select: Tensor) -> Tensor:
    """engine_finetune.py:349-351: mean of token_select[:, layer] per layer."""
    return token_select.float().mean(dim=(0, 2, 3))


# ----------------------------------------------------------------------------------------------
# deterministic synthetic parameters (shared by the golden generator, the tests and bench.py)
# ----------------------------------------------------------------------------------------------
def attentive_pool(tokens: Tensor, p: Dict[str, Tensor], num_heads: int,
                   policy: str = "fp32") -> Tensor:
    """Video pooling head: AttentiveBlock + CrossAttention with one learned query per clip
    (reference video_models/video_vision_transformer_IN21K.py:27-110, called at :479-480).
    tokens [b, t*N, C] = norm(x) of all frames (fp32).  Returns [b, C]."""
    pre = "attentive_blocks."
    b, nk, c = tokens.shape
    d = c // num_heads
    xq = layer_norm(p["query_token"].expand(b, -1, -1), p[pre + "norm_q.weight"], p[pre + "norm_q.bias"])
    xk = layer_norm(tokens, p[pre + "norm_k.weight"], p[pre + "norm_k.bias"])
    xv = layer_norm(tokens, p[pre + "norm_v.weight"], p[pre + "norm_v.bias"])
    qb = p.get(pre + "cross_attn.q_bias")
    vb = p.get(pre + "cross_attn.v_bias")
    q = linear(xq, p[pre + "cross_attn.q.weight"], qb, policy)                      # :92
    k = linear(xk, p[pre + "cross_attn.k.weight"], None if qb is None else torch.zeros_like(vb), policy)
    v = linear(xv, p[pre + "cross_attn.v.weight"], vb, policy)
    q = q.reshape(b, 1, num_heads, d).permute(0, 2, 1, 3)
    k = k.reshape(b, nk, num_heads, d).permute(0, 2, 1, 3)
    v = v.reshape(b, nk, num_heads, d).permute(0, 2, 1, 3)
    q = q * d ** -0.5                                                                # :101
    if policy == "amp16":
        q = _r16(q)
        attn = _r16(q @ k.transpose(-2, -1)).softmax(dim=-1)                        # fp16 matmul out, fp32 softmax
        o = _r16(_r16(attn) @ v)
    else:
        attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
        o = attn @ v
    o = o.transpose(1, 2).reshape(b, 1, c)
    return linear(o, p[pre + "cross_attn.proj.weight"], p[pre + "cross_attn.proj.bias"], policy)[:, 0]


def video_forward(clip: Tensor, p: Dict[str, Tensor], depth: int, num_heads: int, scale: float,
                  policy: str = "fp32", patch: int = 16, forced_masks=None) -> Dict[str, Tensor]:
    """Video model, eval mode (reference video_models/video_vision_transformer_IN21K.py:435-483):
    every frame goes through the image blocks independently (dense masked block == the sparse
    block in eval), then norm -> [b, t*N, C] -> attentive pooling -> head."""
    b, ch, t, h, w = clip.shape
    frames = clip.permute(0, 2, 1, 3, 4).reshape(b * t, ch, h, w)
    r = vit_forward(frames, p, depth, num_heads, scale, policy=policy, sparse=True, patch=patch,
                    forced_masks=forced_masks)
    xn = layer_norm(r["x_final"], p["norm.weight"], p["norm.bias"])
    tokens = xn.reshape(b, t * xn.shape[1], xn.shape[2])
    pooled = attentive_pool(tokens, p, num_heads, policy)
    logits = linear(pooled, p["head.weight"], p["head.bias"], policy)
    return dict(logits=logits, token_select=r["token_select"], token_logits=r["token_logits"],
                pooled=pooled)


def synthetic_state_dict(embed_dim: int = 768, depth: int = 12, num_heads: int = 12,
                         mlp_ratio: float = 4.0, bottleneck: int = 64, num_classes: int = 100,
                         img_size: int = 224, patch: int = 16, seed: int = 0) -> Dict[str, Tensor]:
    """Random-init weights with the reference's state_dict keys drawn key by key from
    a generator seeded with (seed, key), so the result does not depend on module construction
    order.  Scales follow SURVEY.md section 8d: trunc-normal(0.02)-like Linear weights, non-degenerate
    adapters (up_proj ~ N(0, 0.02^2)) and selectors (mlp_head ~ N(0, 0.5^2)); selector biases are
    left at zero here and calibrated to the target keep-rate by `calibrate_selector_bias`."""
    c, hid = embed_dim, int(embed_dim * mlp_ratio)
    n_tok = (img_size // patch) ** 2 + 1
    shapes = {
        "cls_token": ((1, 1, c), 0.02), "pos_embed": ((1, n_tok, c), 0.02),
        "patch_embed.proj.weight": ((c, 3, patch, patch), 0.02), "patch_embed.proj.bias": ((c,), 0.02),
        "norm.weight": ((c,), None), "norm.bias": ((c,), 0.02),
        "head.weight": ((num_classes, c), 0.02), "head.bias": ((num_classes,), 0.02),
    }
    for i in range(depth):
        b = f"blocks.{i}."
        shapes.update({
            b + "norm1.weight": ((c,), None), b + "norm1.bias": ((c,), 0.02),
            b + "attn.qkv.weight": ((3 * c, c), 0.02), b + "attn.qkv.bias": ((3 * c,), 0.02),
            b + "attn.proj.weight": ((c, c), 0.02), b + "attn.proj.bias": ((c,), 0.02),
            b + "norm2.weight": ((c,), None), b + "norm2.bias": ((c,), 0.02),
            b + "mlp.fc1.weight": ((hid, c), 0.02), b + "mlp.fc1.bias": ((hid,), 0.02),
            b + "mlp.fc2.weight": ((c, hid), 0.02), b + "mlp.fc2.bias": ((c,), 0.02),
            b + "adaptmlp.down_proj.weight": ((bottleneck, c), 0.02),
            b + "adaptmlp.down_proj.bias": ((bottleneck,), 0.02),
            b + "adaptmlp.up_proj.weight": ((c, bottleneck), 0.02),
            b + "adaptmlp.up_proj.bias": ((c,), 0.02),
            b + "mlp_token_select.mlp_head.weight": ((1, c), 0.5),
            b + "mlp_token_select.mlp_head.bias": ((1,), 0.0),
        })
    sd = {}
    for idx, key in enumerate(sorted(shapes)):
        shape, std = shapes[key]
        g = torch.Generator().manual_seed(seed * 1000003 + idx)
        if std is None:      # LayerNorm weight: around one
            sd[key] = 1.0 + 0.05 * torch.randn(shape, generator=g)
        elif std == 0.0:
            sd[key] = torch.zeros(shape)
        else:
            sd[key] = std * torch.randn(shape, generator=g)
    return sd


def calibrate_selector_bias(sd: Dict[str, Tensor], img: Tensor, depth: int, num_heads: int,
                            scale: float, rate: float, patch: int = 16,
                            policy: str = "fp32") -> Dict[str, Tensor]:
    """Set each layer's selector bias to minus the (1-rate)-quantile of that layer's logits on `img`
    so the realised keep-rate is ~rate.  Layer by layer, since layer i's mask changes
    the input of layer i+1."""
    sd = dict(sd)
    x = patch_embed(img, sd, patch, policy)
    x = torch.cat((sd["cls_token"].expand(x.shape[0], -1, -1), x), dim=1) + sd["pos_embed"]
    for i in range(depth):
        key = f"blocks.{i}.mlp_token_select.mlp_head.bias"
        sd[key] = torch.zeros(1)
        r = block_sparse(x, sd, f"blocks.{i}.", num_heads, scale, policy)
        q = torch.quantile(r["logits"].flatten().float(), 1.0 - rate)
        sd[key] = (-q).reshape(1).clone()
        r = block_sparse(x, sd, f"blocks.{i}.", num_heads, scale, policy)
        x = r["out"]
    return sd


def checkerboard_mask(bsz: int, n_tok: int) -> Tensor:
    """BASELINE.json config 1: keep cls + even-indexed patches (exactly half of the patches)."""
    m = torch.zeros(bsz, n_tok, 1)
    m[:, 0] = 1.0
    m[:, 1::2] = 1.0   # token index 1 is patch 0 (even-indexed patches)
    return m
