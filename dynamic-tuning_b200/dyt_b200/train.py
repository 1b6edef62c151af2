"""Train-mode forward and backward of the DyT block for parameter-efficient fine-tuning.

Reference behaviour:
  * models/vision_transformer_IN21K.py:144-165  Block.forward in train mode: dense MLP on every
    token, `x = residual + token_select * mlp_x + adapt_x`; `complete_model=True` (the teacher pass of
    engine_finetune.py:49) drops the mask.
  * models/dynamic_adapter.py:25-54  hard Gumbel-sigmoid gate with the straight-through estimator.
  * models/dynamic_adapter.py:127-130  Adapter with train-time dropout (p = 0.1).
  * main_image.py:242-256  only adaptmlp.*, mlp_token_select.* and head.* are trainable: the backward
    is data gradients through the frozen backbone plus weight gradients of those parameters.

Everything runs on the sm_100a kernels through the C ABI (ops.*); torch.autograd only routes the
gradients between the per-block Functions.  There is no CPU / eager fallback.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib, ops
from ._lib import DytError

_FROZEN = {
    "qkv": "attn.qkv", "proj": "attn.proj", "fc1": "mlp.fc1", "fc2": "mlp.fc2",
}
_TRAINABLE_PREFIXES = ("adaptmlp.", "mlp_token_select.")


def _get(module, dotted):
    obj = module
    for part in dotted.split("."):
        obj = getattr(obj, part)
    return obj


class FrozenWeights:
    """fp16 copies (and transposes, for the data-gradient GEMMs) of a block's frozen Linears and the
    fp32 LayerNorm parameters; rebuilt only if a frozen parameter changes."""

    def __init__(self, block):
        self.block = block
        self.key = None
        self.t: Dict[str, torch.Tensor] = {}

    def get(self) -> Dict[str, torch.Tensor]:
        blk = self.block
        params = [_get(blk, d).weight for d in _FROZEN.values()] + [blk.norm1.weight, blk.norm2.weight]
        key = tuple((p.data_ptr(), p._version) for p in params)
        if key != self.key:
            with torch.no_grad():
                h16 = torch.float16
                for short, dotted in _FROZEN.items():
                    lin = _get(blk, dotted)
                    if not lin.weight.is_cuda:
                        raise DytError("dyt_b200: block parameters must live on a CUDA device")
                    w = lin.weight.detach().to(h16).contiguous()
                    self.t[short + "_w"] = w
                    self.t[short + "_wT"] = w.t().contiguous()
                    self.t[short + "_b"] = None if lin.bias is None else lin.bias.detach().to(h16).contiguous()
                for short, ln in (("ln1", blk.norm1), ("ln2", blk.norm2)):
                    self.t[short + "_w"] = ln.weight.detach().float().contiguous()
                    self.t[short + "_b"] = ln.bias.detach().float().contiguous()
            self.key = key
        return self.t


def _check_peft(block) -> None:
    """The backward has data gradients through the frozen backbone and weight gradients only for
    adaptmlp.* / mlp_token_select.* (+ head): anything else that wants a gradient is refused."""
    for name, p in block.named_parameters():
        if p.requires_grad and not name.startswith(_TRAINABLE_PREFIXES):
            raise NotImplementedError(
                f"dyt_b200 fine-tuning implements the reference's PEFT setting (frozen "
                f"backbone, main_image.py:242-256); '{name}' requires grad")


def _frozen(block) -> Dict[str, torch.Tensor]:
    fw = block.__dict__.get("_dyt_frozen")
    if fw is None:
        fw = FrozenWeights(block)
        block.__dict__["_dyt_frozen"] = fw
    return fw.get()


def _adapter_f16(block, down_w, down_b, up_w, up_b):
    """fp16 copies and transposes of the trainable adapter parameters: rebuilt when a parameter
    changes (once per optimizer step), shared by the student and the teacher pass of a step."""
    key = tuple((p.data_ptr(), p._version) for p in (down_w, down_b, up_w, up_b))
    c = block.__dict__.get("_dyt_adapter_f16")
    capturing = torch.cuda.is_current_stream_capturing()
    if c is None or c[0] != key or c[2] != capturing:
        h16 = torch.float16
        with torch.no_grad():
            dw = down_w.detach().to(h16).contiguous()
            uw = up_w.detach().to(h16).contiguous()
            t = dict(dw=dw, db=down_b.detach().to(h16).contiguous(), uw=uw,
                     ub=up_b.detach().to(h16).contiguous(), dwT=dw.t().contiguous(),
                     uwT=uw.t().contiguous())
        c = (key, t, capturing)
        block.__dict__["_dyt_adapter_f16"] = c
    return c[1]


_arange_cache: Dict[Tuple, torch.Tensor] = {}


def _arange_i32(n: int, device) -> torch.Tensor:
    key = (n, device)
    t = _arange_cache.get(key)
    if t is None:
        t = torch.arange(n, device=device, dtype=torch.int32)
        _arange_cache[key] = t
    return t


def _scale_of(block) -> float:
    s = block.adaptmlp.scale
    if torch.is_tensor(s):
        if s.requires_grad:
            raise NotImplementedError("dyt_b200: adapter_scalar='learnable_scalar' is not used by any "
                                      "reference entry script and has no backward here")
        return float(s.item())
    return float(s)


class DytBlockFn(torch.autograd.Function):
    """One DyT block in train mode.  Inputs: x fp32 [B, N, C] and the six trainable tensors of the
    block; outputs: (x_out fp32, token_select [B, N, 1] fp32, token_logits [B, N-1, 1] fp32)."""

    @staticmethod
    def forward(ctx, x, down_w, down_b, up_w, up_b, sel_w, sel_b, block, complete_model, noise,
                drop_mult, eps, training_gate, xn_in=None, next_ln=None):
        ctx.set_materialize_grads(False)
        fz = _frozen(block)
        h16 = torch.float16
        B, N, Cd = x.shape
        H = block.attn.num_heads
        if Cd != 64 * H:
            raise DytError(f"dyt_b200 attention kernels need head_dim 64 (C={Cd}, heads={H})")
        x = x.detach().to(torch.float32).contiguous()
        scale = _scale_of(block)
        tau = float(getattr(block.mlp_token_select, "tau", 5.0))
        thr = float(getattr(block.mlp_token_select, "threshold", 0.5))
        ad = _adapter_f16(block, down_w, down_b, up_w, up_b)
        dw16, db16, uw16, ub16 = ad["dw"], ad["db"], ad["uw"], ad["ub"]

        # norm1: handed over by the previous block's merge kernel when the caller chains the blocks
        xn = xn_in if xn_in is not None else ops.layernorm_f16(x, fz["ln1_w"], fz["ln1_b"], eps)
        qkv, _ = ops.linear_f16(xn, fz["qkv_w"], fz["qkv_b"])
        o = ops.attn_varlen(qkv.reshape(B, N, -1), H)
        x1, x1h = ops.linear_f16(o, fz["proj_w"], fz["proj_b"], epilogue=_lib.EPI_BIAS_RESID,
                                 resid=x, want_f16_copy=True)
        T = B * N
        sparse_bwd = (not complete_model) and T >= SPARSE_STUDENT_MIN_TOKENS
        if sparse_bwd:   # the kept-rows-only backward needs the compaction as well
            d = ops.dispatch(x1.reshape(B, N, Cd), sel_w.detach(), sel_b.detach(),
                             logit_dtype=torch.float16, threshold=thr,
                             noise=noise if training_gate else None, tau=tau, pack=False)
            mask, logits = d["mask"], d["logits"]
            compaction = (d["token_pos"], d["packed_idx"], d["n_kept"])
        else:
            mask, logits, row_of = ops.token_select(
                x1.reshape(B, N, Cd), sel_w.detach(), sel_b.detach(), logit_dtype=torch.float16,
                threshold=thr, noise=noise if training_gate else None, tau=tau, want_row_of=True)
            empty = torch.empty(0, dtype=torch.int32, device=x.device)
            compaction = (empty, empty, empty)
        ln2 = ops.layernorm_f16(x1, fz["ln2_w"], fz["ln2_b"], eps)
        if fz["fc1_w"].shape[0] > 64:      # fc1 + GELU, keeping the pre-activation (one kernel)
            hdn, pre = ops.linear_f16_aux(ln2, fz["fc1_w"], fz["fc1_b"], _lib.EPI_BIAS_GELU_KEEP)
        else:
            pre, _ = ops.linear_f16(ln2, fz["fc1_w"], fz["fc1_b"])
            hdn = ops.eltwise_f16(_lib.EW_GELU_FWD, pre)
        mlp_x, _ = ops.linear_f16(hdn, fz["fc2_w"], fz["fc2_b"])
        hd, _ = ops.linear_f16(x1h, dw16, db16, epilogue=_lib.EPI_BIAS_RELU)
        if drop_mult is not None:
            hd = ops.eltwise_f16(_lib.EW_MUL, hd, drop_mult.reshape(hd.shape))
        if complete_model:
            token_pos = _arange_i32(T, x.device)
        elif sparse_bwd:
            ar = _arange_i32(T, x.device)
            token_pos = torch.where(mask.reshape(-1) > 0, ar, torch.full_like(ar, -1))
        else:
            token_pos = row_of          # written by the selector kernel: t if kept else -1
        bott_f = hd.shape[-1]
        if Cd % 128 == 0 and Cd <= 1024 and bott_f <= 64 and bott_f % 8 == 0 and T * Cd < (1 << 31) - Cd:
            # adapter up-projection inside the merge kernel (+ the next block's norm1)
            out, xn_next = ops.merge_up(hd.reshape(B, N, bott_f), uw16, ub16, scale, x1.reshape(B, N, Cd),
                                        mlp_x.reshape(T, Cd), token_pos, next_ln=next_ln, eps=eps)
        else:
            up, _ = ops.linear_f16(hd, uw16, ub16, epilogue=_lib.EPI_BIAS, scale=scale)
            out, xn_next = ops.scatter_merge(x1.reshape(B, N, Cd), up.reshape(B, N, Cd),
                                             mlp_x.reshape(T, Cd), token_pos, next_ln=next_ln, eps=eps)
        ctx.block = block
        ctx.meta = (B, N, Cd, H, scale, tau, eps, bool(complete_model), bool(training_gate))
        ctx.noise = noise if training_gate else None
        ctx.drop_mult = drop_mult
        ctx.save_for_backward(x, qkv, o, x1, x1h, mask, logits, pre, mlp_x, hd, ad["dwT"], ad["uwT"],
                              sel_w.detach(), *compaction)
        if debug_keep is not None:   # tests: look at the forward intermediates
            debug_keep.update(x1=x1, x1h=x1h, pre=pre, mlp_x=mlp_x, hd=hd, qkv=qkv, o=o)
        if xn_next is None:
            xn_next = torch.empty(0, dtype=torch.float16, device=x.device)
        ctx.mark_non_differentiable(xn_next)
        return out, mask, logits, xn_next

    @staticmethod
    def backward(ctx, g_out, g_sel, g_logits, _g_xn=None):
        (x, qkv, o, x1, x1h, mask, logits, pre, mlp_x, hd, dwT, uwT, sel_w, token_pos, packed_idx,
         n_kept) = ctx.saved_tensors
        B, N, Cd, H, scale, tau, eps, complete_model, training_gate = ctx.meta
        fz = _frozen(ctx.block)
        T = B * N
        if g_out is None:
            g_out = torch.zeros((T, Cd), dtype=torch.float32, device=x.device)
        g_out = g_out.to(torch.float32).contiguous().reshape(T, Cd)
        # student pass: the masked MLP gradient is zero on dropped rows, so the frozen MLP's backward
        # can run on the kept rows only (packed, count on the device, like the inference forward).
        # Kept as an option (SPARSE_STUDENT_MIN_TOKENS): see the measurement next to that constant.
        sparse = (not complete_model) and token_pos.numel() > 0
        g16, gm16, g_l, g_x1 = ops.merge_bwd(
            g_out, N, mlp_x=mlp_x, mask=mask, logits=logits, noise=ctx.noise, tau=tau,
            g_token_select=None if complete_model else g_sel,
            g_token_logits=None if complete_model else g_logits, masked=not complete_model,
            token_pos=token_pos if sparse else None, sel_w=sel_w if sparse else None)
        if complete_model:
            gm16 = g16
        # ---- adapter: up dgrad, wgrads, ReLU / dropout, (down dgrad further below) ----
        g_hd, _ = ops.linear_f16(g16, uwT, None, epilogue=_lib.EPI_BIAS, scale=scale)
        bott = hd.shape[-1]
        n_w = Cd * bott
        zbuf = torch.zeros(2 * n_w + Cd + bott + Cd + 4, dtype=torch.float32, device=x.device)
        d_up_w, d_up_b = zbuf[:n_w].view(Cd, bott), zbuf[2 * n_w:2 * n_w + Cd]
        d_down_w = zbuf[n_w:2 * n_w].view(bott, Cd)
        d_down_b = zbuf[2 * n_w + Cd:2 * n_w + Cd + bott]
        ops.wgrad_f16(g16, hd.reshape(T, -1), alpha=scale, out=(d_up_w, d_up_b))
        dm = None if ctx.drop_mult is None else ctx.drop_mult.reshape(T, -1)
        g_hp = ops.eltwise_f16(_lib.EW_RELU_DROP_BWD, g_hd.reshape(T, -1), hd.reshape(T, -1), dm)
        ops.wgrad_f16(g_hp, x1h.reshape(T, Cd), out=(d_down_w, d_down_b))
        # ---- frozen MLP: fc2 dgrad, GELU', fc1 dgrad, LayerNorm2 backward ----
        # fc2 dgrad, then GELU' as an HBM-bound elementwise pass: measured faster than the fused
        # DYT_EPI_DGELU epilogue (76 us vs 35 + 36 us at 12.6k rows: the derivative costs two MUFU +
        # a polynomial per element and holds the accumulator stage, and the plain GEMM can take
        # the 192-wide tile)
        d_sel_w = d_sel_b = None
        if complete_model:
            g_h, _ = ops.linear_f16(gm16, fz["fc2_wT"], None)
            g_pre = ops.eltwise_f16(_lib.EW_GELU_BWD, g_h.reshape(pre.shape), pre)
            g_ln2, _ = ops.linear_f16(g_pre, fz["fc1_wT"], None)
            g_x1, _ = ops.layernorm_bwd(g_ln2.reshape(T, Cd), x1.reshape(T, Cd), fz["ln2_w"], eps,
                                        resid=g_out)
        else:
            if sparse:
                g_h, _ = ops.linear_f16(gm16, fz["fc2_wT"], None, m_dev=n_kept)
                g_pre = ops.gelu_bwd_rows(g_h, pre.reshape(T, -1), packed_idx, n_kept)
                g_ln2, _ = ops.linear_f16(g_pre, fz["fc1_wT"], None, m_dev=n_kept)
                # g_x1 = g_out + g_l * w_sel (from merge_bwd) + dLN2 on the kept rows, in place
                ops.layernorm_bwd(g_ln2.reshape(T, Cd), x1.reshape(T, Cd), fz["ln2_w"], eps,
                                  resid=g_x1, out=g_x1, row_idx=packed_idx, n_rows_dev=n_kept)
            else:
                g_h, _ = ops.linear_f16(gm16, fz["fc2_wT"], None)
                g_pre = ops.eltwise_f16(_lib.EW_GELU_BWD, g_h.reshape(pre.shape), pre)
                g_ln2, _ = ops.linear_f16(g_pre, fz["fc1_wT"], None)
                # selector: data gradient g_l * w folded into the LayerNorm-backward pass
                g_x1, _ = ops.layernorm_bwd(g_ln2.reshape(T, Cd), x1.reshape(T, Cd), fz["ln2_w"],
                                            eps, resid=g_out, row_scale=g_l, axpy=sel_w)
            off = 2 * n_w + Cd + bott
            d_sel_w, d_sel_b = zbuf[off:off + Cd], zbuf[off + Cd:off + Cd + 1]
            ops.rowscale_colsum(g_l, x1h.reshape(T, Cd), d_sel_w, d_sel_b)
            d_sel_w = d_sel_w.reshape(1, Cd)
        need_x = ctx.needs_input_grad[0]
        g_x = None
        if need_x:
            g_x1, g_x1h = ops.linear_f16(g_hp, dwT, None,
                                         epilogue=_lib.EPI_BIAS_RESID, resid=g_x1, want_f16_copy=True)
            g_o, _ = ops.linear_f16(g_x1h, fz["proj_wT"], None)
            g_qkv = ops.attn_varlen_bwd(qkv.reshape(B, N, -1), o.reshape(B, N, Cd),
                                        g_o.reshape(B, N, Cd), H)
            g_xn, _ = ops.linear_f16(g_qkv.reshape(T, -1), fz["qkv_wT"], None)
            g_x, _ = ops.layernorm_bwd(g_xn.reshape(T, Cd), x.reshape(T, Cd), fz["ln1_w"], eps,
                                       resid=g_x1.reshape(T, Cd))
            g_x = g_x.reshape(B, N, Cd)
        return (g_x, d_down_w, d_down_b, d_up_w, d_up_b, d_sel_w, d_sel_b,
                None, None, None, None, None, None, None, None)


_fixed = {"noises": None, "drop_mults": None}
# B * N from which the student pass runs the frozen MLP's backward on the kept rows only.  Off by
# default: measured on one B200 it is a wash (64 images 21.19 vs 21.10 ms per step, 256 images
# 72.96 vs 72.47 ms): the packed GEMMs save half the tiles, but at these sizes that is at most one
# round of the persistent grid, and the packed form needs an extra fp32 pass over the stream.
SPARSE_STUDENT_MIN_TOKENS = 1 << 30
debug_keep: Optional[dict] = None   # set to a dict to receive the last block forward's intermediates


class fixed_randomness:
    """Context manager that feeds recorded Gumbel draws / dropout multipliers to successive
    block_train calls (one entry per block call, in call order) instead of drawing them: lets a test
    share the randomness with the reference."""

    def __init__(self, noises=None, drop_mults=None):
        self.new = {"noises": None if noises is None else list(noises),
                    "drop_mults": None if drop_mults is None else list(drop_mults)}

    def __enter__(self):
        self.old = dict(_fixed)
        _fixed.update(self.new)
        return self

    def __exit__(self, *exc):
        _fixed.update(self.old)
        return False


def draw_pass_randomness(blocks, B: int, N: int, device):
    """All Gumbel draws and dropout multipliers of one pass (every block) in a handful of launches:
    returns per-block lists (noise pairs as fp32 [B, N-1, 1] holding fp16 values, the dtype the
    reference draws them in under autocast; multipliers fp16 [B, N, bottleneck])."""
    L = len(blocks)
    from .modules import draw_gumbel_pair
    g1, g2 = draw_gumbel_pair((L, B, N - 1, 1), torch.float16, device)
    g1, g2 = g1.float(), g2.float()
    noises = [(g1[i], g2[i]) for i in range(L)]
    p = float(blocks[0].adaptmlp.dropout)
    bott = blocks[0].adaptmlp.down_proj.out_features
    if p > 0 and all(float(b.adaptmlp.dropout) == p and b.adaptmlp.down_proj.out_features == bott
                     for b in blocks):
        keep = torch.rand((L, B, N, bott), device=device) >= p
        mult = keep.to(torch.float16) * (1.0 / (1.0 - p))
        mults = [mult[i] for i in range(L)]
    else:
        mults = [None] * L
    return noises, mults


def block_train(block, x: torch.Tensor, complete_model: bool = False,
                noise: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                drop_mult: Optional[torch.Tensor] = None, *, xn: Optional[torch.Tensor] = None,
                next_block=None, return_xn: bool = False):
    """Train-mode block with autograd.  `noise` = the two Gumbel draws [B, N-1, 1] (drawn here in the
    reference's order when None and the block is in train mode); `drop_mult` = the adapter dropout
    multiplier keep / (1 - p), fp16 [B, N, bottleneck] (drawn here when None and p > 0).
    Chaining (the model's block loop): `xn` = f16(norm1(x)) of THIS block as produced by the previous
    block's merge kernel, `next_block` = the block whose norm1 this block's merge should emit;
    with return_xn=True the call returns (out, mask, logits, xn_of_next_block or None)."""
    if not x.is_cuda:
        raise DytError("dyt_b200 needs CUDA tensors (no CPU fallback)")
    if not hasattr(block, "mlp_token_select"):
        raise AttributeError("'Block' object has no attribute 'mlp_token_select'")
    if torch.is_grad_enabled():      # (a no_grad train()-mode forward needs no such guarantee)
        _check_peft(block)
    B, N, _ = x.shape
    training_gate = bool(block.training)
    if noise is None and _fixed["noises"]:
        noise = tuple(t.to(x.device) for t in _fixed["noises"].pop(0))
    if drop_mult is None and _fixed["drop_mults"]:
        drop_mult = _fixed["drop_mults"].pop(0)
    if training_gate and noise is None:
        from .modules import draw_gumbel_pair
        noise = draw_gumbel_pair((B, N - 1, 1), torch.float16, x.device)
    p = float(block.adaptmlp.dropout)
    if block.training and p > 0 and drop_mult is None:
        keep = torch.rand((B, N, block.adaptmlp.down_proj.out_features), device=x.device) >= p
        drop_mult = keep.to(torch.float16) / (1.0 - p)
    if drop_mult is not None:
        drop_mult = drop_mult.to(device=x.device, dtype=torch.float16).contiguous()
    a, s = block.adaptmlp, block.mlp_token_select.mlp_head
    sel_b = s.bias if s.bias is not None else torch.zeros(1, device=x.device)
    next_ln = None
    if next_block is not None:
        fzn = _frozen(next_block)
        next_ln = (fzn["ln1_w"], fzn["ln1_b"])
    out, mask, logits, xn_next = DytBlockFn.apply(
        x, a.down_proj.weight, a.down_proj.bias, a.up_proj.weight, a.up_proj.bias, s.weight, sel_b,
        block, bool(complete_model), noise, drop_mult, float(block.norm1.eps), training_gate, xn,
        next_ln)
    if return_xn:
        return out, mask, logits, (xn_next if xn_next.numel() else None)
    return out, mask, logits


class DytHeadFn(torch.autograd.Function):
    """Final LayerNorm of the cls rows + classifier (reference vision_transformer_IN21K.py:371-380);
    `norm` is frozen, `head` is trainable (main_image.py:247)."""

    @staticmethod
    def forward(ctx, tokens, head_w, head_b, norm_w, norm_b, eps):
        B, N, Cd = tokens.shape
        dev = tokens.device
        tokens = tokens.detach().to(torch.float32).contiguous()
        nc = head_w.shape[0]
        n8 = (nc + 7) // 8 * 8
        w16 = torch.zeros((n8, Cd), dtype=torch.float16, device=dev)
        b16 = torch.zeros((n8,), dtype=torch.float16, device=dev)
        w16[:nc] = head_w.detach().to(torch.float16)
        if head_b is not None:
            b16[:nc] = head_b.detach().to(torch.float16)
        idx = torch.arange(B, device=dev, dtype=torch.int32) * N
        nw = norm_w.detach().float().contiguous()
        cls_n = ops.layernorm_f16(tokens, nw, norm_b.detach().float().contiguous(), eps, row_idx=idx)
        logits, _ = ops.linear_f16(cls_n, w16, b16)
        ctx.save_for_backward(tokens, cls_n, w16, idx, nw)
        ctx.meta = (B, N, Cd, nc, n8, eps, head_b is not None)
        return logits[:, :nc]

    @staticmethod
    def backward(ctx, g_logits):
        tokens, cls_n, w16, idx, nw = ctx.saved_tensors
        B, N, Cd, nc, n8, eps, has_bias = ctx.meta
        g16 = torch.zeros((B, n8), dtype=torch.float16, device=tokens.device)
        g16[:, :nc] = g_logits.to(torch.float16)
        d_w, d_b = ops.wgrad_f16(g16, cls_n, n_out=nc, want_bias=has_bias)
        g_tok = None
        if ctx.needs_input_grad[0]:
            g_cls, _ = ops.linear_f16(g16, w16.t().contiguous(), None)
            g_tok = torch.zeros((B * N, Cd), dtype=torch.float32, device=tokens.device)
            ops.layernorm_bwd(g_cls, tokens.reshape(B * N, Cd), nw, eps, row_idx=idx, out=g_tok)
            g_tok = g_tok.reshape(B, N, Cd)
        return g_tok, d_w, d_b, None, None, None


def head_train(model, tokens: torch.Tensor) -> torch.Tensor:
    for name, p in (("norm.weight", model.norm.weight), ("norm.bias", model.norm.bias)):
        if p.requires_grad:
            raise NotImplementedError(f"dyt_b200 fine-tuning keeps the final norm frozen ('{name}' "
                                      "requires grad)")
    return DytHeadFn.apply(tokens, model.head.weight, model.head.bias, model.norm.weight,
                           model.norm.bias, float(model.norm.eps))
