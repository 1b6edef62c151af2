"""torch-tensor front-ends of the C-ABI kernels (PyTorch here = device memory + streams only)."""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional, Tuple

import torch

from . import _lib
from ._wscache import StreamWorkspaces
from ._lib import DytError, check
from .gate import min_kept_logit


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors: Optional[torch.Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise DytError("dyt_b200 kernels run on CUDA tensors only (sm_100a); "
                           "there is no CPU fallback")


def _rows2d(t: torch.Tensor) -> torch.Tensor:
    t2 = t.reshape(-1, t.shape[-1])
    if t2.stride(-1) != 1:
        t2 = t2.contiguous()
    return t2


def linear_f16(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
               epilogue: int = _lib.EPI_BIAS, resid: Optional[torch.Tensor] = None,
               scale: float = 1.0, m_dev: Optional[torch.Tensor] = None,
               want_f16_copy: bool = False,
               out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """y = epilogue(x @ w.T + bias).  x [..., K] fp16, w [N, K] fp16, bias [N] fp16.
    Returns (out, f16_copy): out is fp16 [..., N], or fp32 for EPI_BIAS_RESID (resid fp32)."""
    _need_cuda(x, w, bias, resid, m_dev)
    if x.dtype != torch.float16 or w.dtype != torch.float16:
        raise DytError("linear_f16 expects fp16 operands")
    x2 = _rows2d(x)
    M, K = x2.shape
    N = w.shape[0]
    lead = x.shape[:-1]
    out_h = out_f = None
    r2 = None
    if epilogue == _lib.EPI_BIAS_RESID:
        if resid is None or resid.dtype != torch.float32:
            raise DytError("EPI_BIAS_RESID needs an fp32 residual")
        r2 = _rows2d(resid)
        out_f = out if out is not None else torch.empty((M, N), dtype=torch.float32, device=x.device)
        if want_f16_copy:
            out_h = torch.empty((M, N), dtype=torch.float16, device=x.device)
    else:
        out_h = out if out is not None else torch.empty((M, N), dtype=torch.float16, device=x.device)
    check(_lib.lib().dyt_linear_f16(
        x2.data_ptr(), x2.stride(0), w.data_ptr(), w.stride(0), M, N, K, _ptr(m_dev), epilogue,
        _ptr(bias), _ptr(out_h), N, _ptr(out_f), N, _ptr(r2), 0 if r2 is None else r2.stride(0),
        float(scale), _stream()), "dyt_linear_f16")
    if epilogue == _lib.EPI_BIAS_RESID:
        return out_f.reshape(*lead, N), (None if out_h is None else out_h.reshape(*lead, N))
    return out_h.reshape(*lead, N), None


def linear_f16_aux(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], epilogue: int,
                   aux: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Train-mode MLP GEMMs.  EPI_BIAS_GELU_KEEP: returns (gelu(x W^T + b), pre-activation);
    EPI_DGELU: returns ((x W^T + b) * gelu'(aux), aux).  All fp16."""
    _need_cuda(x, w, bias, aux)
    if x.dtype != torch.float16 or w.dtype != torch.float16:
        raise DytError("linear_f16_aux expects fp16 operands")
    x2 = _rows2d(x)
    M, K = x2.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=torch.float16, device=x.device)
    if epilogue == _lib.EPI_BIAS_GELU_KEEP:
        aux2 = torch.empty((M, N), dtype=torch.float16, device=x.device)
    else:
        if aux is None or aux.dtype != torch.float16:
            raise DytError("EPI_DGELU needs the fp16 pre-activation")
        aux2 = _rows2d(aux)
        if aux2.shape != (M, N):
            raise DytError("EPI_DGELU: pre-activation shape mismatch")
    check(_lib.lib().dyt_linear_f16_aux(
        x2.data_ptr(), x2.stride(0), w.data_ptr(), w.stride(0), M, N, K, None, epilogue, _ptr(bias),
        out.data_ptr(), N, aux2.data_ptr(), aux2.stride(0), _stream()), "dyt_linear_f16_aux")
    lead = x.shape[:-1]
    return out.reshape(*lead, N), aux2.reshape(*lead, N)


def attn_varlen(qkv: torch.Tensor, num_heads: int, cu_seqlens: Optional[torch.Tensor] = None,
                num_seqs: Optional[int] = None, max_seqlen: Optional[int] = None) -> torch.Tensor:
    """qkv: fp16 [B, N, 3*C] (uniform) or [T, 3*C] with int32 cu_seqlens [num_seqs+1].
    Returns fp16 [.., C]."""
    _need_cuda(qkv, cu_seqlens)
    if qkv.dtype != torch.float16:
        raise DytError("attn_varlen expects fp16 qkv")
    C3 = qkv.shape[-1]
    Cdim = C3 // 3
    q2 = _rows2d(qkv)
    T = q2.shape[0]
    if cu_seqlens is None:
        if qkv.dim() != 3:
            raise DytError("uniform attention expects qkv [B, N, 3C]")
        nseq, uni = qkv.shape[0], qkv.shape[1]
        mx = uni
    else:
        nseq = int(num_seqs if num_seqs is not None else cu_seqlens.numel() - 1)
        if max_seqlen is None:
            raise DytError("varlen attention needs max_seqlen (no host sync is done here)")
        uni, mx = 0, int(max_seqlen)
    out = torch.empty((T, Cdim), dtype=torch.float16, device=qkv.device)
    check(_lib.lib().dyt_attn_varlen_fwd(
        q2.data_ptr(), q2.stride(0), _ptr(cu_seqlens), nseq, uni, mx, T, num_heads,
        Cdim // num_heads, out.data_ptr(), Cdim, _stream()), "dyt_attn_varlen_fwd")
    return out.reshape(*qkv.shape[:-1], Cdim)


def pad_attn_bias(bias: torch.Tensor) -> torch.Tensor:
    """[heads, N, N] fp32 view of a buffer whose rows are padded to a multiple of 4 floats, the
    layout dyt_attn_bias_fwd reads in 16-byte vectors (N = 1025 -> pitch 1028).  Build it once per
    bias (the segmentation modules cache it per table version), not per forward."""
    H, N, N2 = bias.shape
    pitch = (N2 + 3) // 4 * 4
    if pitch % 64 == 0:          # rows a multiple of 256 bytes apart all fall into the same cache sets
        pitch += 4               # (measured at N = 1024: 233 us with pitch 1024)
    buf = torch.zeros((H, N, pitch), dtype=torch.float32, device=bias.device)
    buf[:, :, :N2] = bias
    return buf[:, :, :N2]


def _bias_pitch(bias: torch.Tensor, num_heads: int, N: int):
    """(tensor to keep alive, row pitch in floats) of an attention bias: an fp32 [heads, N, N] tensor
    whose rows are dense and evenly pitched is used in place, anything else is made contiguous."""
    if tuple(bias.shape) != (num_heads, N, N):
        raise DytError(f"attn_bias: bias must be [heads, N, N] = [{num_heads}, {N}, {N}]")
    if bias.dtype != torch.float32:
        bias = bias.to(torch.float32)
    if not (bias.stride(2) == 1 and bias.stride(1) >= N and bias.stride(0) == N * bias.stride(1)):
        bias = bias.contiguous()
    return bias, bias.stride(1)


def attn_bias(qkv: torch.Tensor, num_heads: int, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Long-sequence attention with an additive per-head bias.  qkv fp16 [B, N, 3*C]; bias fp32
    [num_heads, N, N] (dense or a pad_attn_bias view) or None.  Returns fp16 [B, N, C]."""
    _need_cuda(qkv, bias)
    if qkv.dtype != torch.float16 or qkv.dim() != 3:
        raise DytError("attn_bias expects fp16 qkv [B, N, 3C]")
    B, N, C3 = qkv.shape
    Cdim = C3 // 3
    q2 = _rows2d(qkv)
    ld_bias = 0
    if bias is not None:
        bias, ld_bias = _bias_pitch(bias, num_heads, N)
    out = torch.empty((B * N, Cdim), dtype=torch.float16, device=qkv.device)
    check(_lib.lib().dyt_attn_bias_fwd(q2.data_ptr(), q2.stride(0), _ptr(bias), ld_bias, B, N, num_heads,
                                       Cdim // num_heads, out.data_ptr(), Cdim, _stream()),
          "dyt_attn_bias_fwd")
    return out.reshape(B, N, Cdim)


def layernorm_f16(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-6,
                  row_idx: Optional[torch.Tensor] = None,
                  n_rows_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x, weight, bias, row_idx, n_rows_dev)
    if x.dtype != torch.float32:
        raise DytError("layernorm_f16 expects an fp32 input (the residual stream)")
    x2 = _rows2d(x)
    Cdim = x2.shape[1]
    n_rows = x2.shape[0] if row_idx is None else row_idx.numel()
    out = torch.empty((n_rows, Cdim), dtype=torch.float16, device=x.device)
    check(_lib.lib().dyt_layernorm_f16(
        x2.data_ptr(), x2.stride(0), _ptr(row_idx), _ptr(n_rows_dev), n_rows, Cdim,
        weight.data_ptr(), bias.data_ptr(), float(eps), out.data_ptr(), Cdim, _stream()),
        "dyt_layernorm_f16")
    return out if row_idx is not None else out.reshape(*x.shape[:-1], Cdim)


_dispatch_ws = StreamWorkspaces(zero_filled=True, min_bytes=4096)


def _dispatch_workspace(device: torch.device, B: int) -> torch.Tensor:
    need = int(_lib.lib().dyt_dispatch_workspace_bytes(B))
    return _dispatch_ws.get(device, need)


def dispatch(x1: torch.Tensor, sel_w: torch.Tensor, sel_b: torch.Tensor, *,
             logit_dtype: torch.dtype = torch.float16, threshold: float = 0.5,
             noise: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, tau: float = 5.0,
             ln_w: Optional[torch.Tensor] = None, ln_b: Optional[torch.Tensor] = None,
             eps: float = 1e-6, forced_mask: Optional[torch.Tensor] = None, pack: bool = True):
    """Fused dispatcher on x1 [B, N, C] fp32.  Returns dict(mask [B,N,1] f32, logits [B,N-1,1] f32,
    packed_idx i32 [B*N], token_pos i32 [B*N], cu_seqlens i32 [B+1], n_kept i32 [1],
    packed f16 [B*N, C] or None).  Counts stay on the device (no sync)."""
    _need_cuda(x1, sel_w, sel_b, ln_w, ln_b, forced_mask)
    if x1.dtype != torch.float32 or x1.dim() != 3:
        raise DytError("dispatch expects x1 [B, N, C] fp32")
    if logit_dtype not in (torch.float16, torch.float32):
        raise DytError("dispatch implements fp16 and fp32 logits")
    x1 = x1.contiguous()
    B, N, Cdim = x1.shape
    dev = x1.device
    mask = torch.empty((B, N, 1), dtype=torch.float32, device=dev)
    logits = torch.empty((B, N - 1, 1), dtype=torch.float32, device=dev)
    packed_idx = torch.empty(B * N, dtype=torch.int32, device=dev)
    token_pos = torch.empty(B * N, dtype=torch.int32, device=dev)
    cu = torch.empty(B + 1, dtype=torch.int32, device=dev)
    n_kept = torch.empty(1, dtype=torch.int32, device=dev)
    packed = None
    if pack:
        if ln_w is None or ln_b is None:
            raise DytError("dispatch(pack=True) needs the norm2 parameters")
        packed = torch.empty((B * N, Cdim), dtype=torch.float16, device=dev)
    n1 = n2 = None
    if noise is not None:
        n1 = noise[0].to(torch.float32).contiguous()
        n2 = noise[1].to(torch.float32).contiguous()
    fm = None if forced_mask is None else forced_mask.to(torch.float32).contiguous()
    sw = sel_w.reshape(-1).to(torch.float32).contiguous()
    sb = sel_b.reshape(-1).to(torch.float32).contiguous()
    gate = torch.empty((B, N, 1), dtype=torch.float32, device=dev) if fm is not None else None
    ws = _dispatch_workspace(dev, B)
    check(_lib.lib().dyt_dispatch_fwd(
        x1.data_ptr(), Cdim, sw.data_ptr(), sb.data_ptr(),
        1 if logit_dtype == torch.float16 else 0, float(min_kept_logit(logit_dtype, threshold)),
        _ptr(n1), _ptr(n2), float(tau), B, N, Cdim, _ptr(ln_w), _ptr(ln_b), float(eps), _ptr(fm),
        mask.data_ptr(), _ptr(gate), logits.data_ptr(), packed_idx.data_ptr(), token_pos.data_ptr(),
        cu.data_ptr(), n_kept.data_ptr(), _ptr(packed), Cdim, ws.data_ptr(), _stream()),
        "dyt_dispatch_fwd")
    return dict(mask=mask, gate=gate if gate is not None else mask, logits=logits, packed_idx=packed_idx, token_pos=token_pos,
                cu_seqlens=cu, n_kept=n_kept, packed=packed)


def token_select(x1: torch.Tensor, sel_w: torch.Tensor, sel_b: torch.Tensor, *,
                 logit_dtype: torch.dtype = torch.float16, threshold: float = 0.5,
                 noise: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, tau: float = 5.0,
                 want_row_of: bool = False):
    """Score + gate without compaction: (mask [B, N, 1] f32, logits [B, N-1, 1] f32[, row_of i32
    [B*N]: t for kept tokens, -1 for dropped ones])."""
    _need_cuda(x1, sel_w, sel_b)
    if x1.dtype != torch.float32 or x1.dim() != 3:
        raise DytError("token_select expects x1 [B, N, C] fp32")
    x1 = x1.contiguous()
    B, N, Cdim = x1.shape
    mask = torch.empty((B, N, 1), dtype=torch.float32, device=x1.device)
    logits = torch.empty((B, N - 1, 1), dtype=torch.float32, device=x1.device)
    n1 = n2 = None
    if noise is not None:
        n1 = noise[0].to(torch.float32).contiguous()
        n2 = noise[1].to(torch.float32).contiguous()
    sw = sel_w.reshape(-1).to(torch.float32).contiguous()
    sb = sel_b.reshape(-1).to(torch.float32).contiguous()
    row_of = torch.empty(B * N, dtype=torch.int32, device=x1.device) if want_row_of else None
    check(_lib.lib().dyt_token_select_fwd(
        x1.data_ptr(), Cdim, sw.data_ptr(), sb.data_ptr(),
        1 if logit_dtype == torch.float16 else 0, float(min_kept_logit(logit_dtype, threshold)),
        _ptr(n1), _ptr(n2), float(tau), B, N, Cdim, mask.data_ptr(), logits.data_ptr(),
        _ptr(row_of), _stream()), "dyt_token_select_fwd")
    if want_row_of:
        return mask, logits, row_of
    return mask, logits


def scatter_merge(x1: torch.Tensor, adapt: torch.Tensor, mlp_packed: torch.Tensor,
                  token_pos: torch.Tensor, next_ln: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                  eps: float = 1e-6):
    """out = adapt + (x1 + scatter(mlp_packed)); optionally also LayerNorm(out) in fp16."""
    _need_cuda(x1, adapt, mlp_packed, token_pos)
    x2 = _rows2d(x1)
    a2 = _rows2d(adapt)
    m2 = _rows2d(mlp_packed)
    T, Cdim = x2.shape
    out = torch.empty((T, Cdim), dtype=torch.float32, device=x1.device)
    nln_out = None
    nw = nb = None
    if next_ln is not None:
        nw, nb = next_ln
        nln_out = torch.empty((T, Cdim), dtype=torch.float16, device=x1.device)
    check(_lib.lib().dyt_scatter_merge_fwd(
        x2.data_ptr(), x2.stride(0), a2.data_ptr(), a2.stride(0), m2.data_ptr(), m2.stride(0),
        token_pos.data_ptr(), T, Cdim, out.data_ptr(), Cdim, _ptr(nw), _ptr(nb), float(eps),
        _ptr(nln_out), Cdim, _stream()), "dyt_scatter_merge_fwd")
    out = out.reshape(x1.shape)
    return (out, None if nln_out is None else nln_out.reshape(x1.shape))


def merge_up(down: torch.Tensor, up_w: torch.Tensor, up_b: Optional[torch.Tensor], scale: float,
             x1: torch.Tensor, mlp_packed: torch.Tensor, token_pos: torch.Tensor,
             next_ln: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, eps: float = 1e-6):
    """out = f16(f16(down up_w^T + up_b) * scale) + (x1 + scatter(mlp_packed)) in one kernel (the
    adapter output never reaches HBM); optionally also LayerNorm(out) in fp16.  Same results as
    linear_f16(down, up_w, up_b, scale=scale) followed by scatter_merge."""
    _need_cuda(down, up_w, x1, mlp_packed, token_pos)
    x2 = _rows2d(x1)
    d2 = _rows2d(down)
    m2 = _rows2d(mlp_packed)
    T, Cdim = x2.shape
    K = d2.shape[1]
    assert d2.dtype == torch.float16 and up_w.dtype == torch.float16 and up_w.shape == (Cdim, K)
    up_w = up_w.contiguous()
    out = torch.empty((T, Cdim), dtype=torch.float32, device=x1.device)
    nln_out = None
    nw = nb = None
    if next_ln is not None:
        nw, nb = next_ln
        nln_out = torch.empty((T, Cdim), dtype=torch.float16, device=x1.device)
    check(_lib.lib().dyt_merge_up_fwd(
        d2.data_ptr(), d2.stride(0), up_w.data_ptr(), up_w.stride(0), _ptr(up_b), float(scale), K,
        x2.data_ptr(), x2.stride(0), m2.data_ptr(), m2.stride(0), token_pos.data_ptr(), T, Cdim,
        out.data_ptr(), Cdim, _ptr(nw), _ptr(nb), float(eps), _ptr(nln_out), Cdim, _stream()),
        "dyt_merge_up_fwd")
    out = out.reshape(x1.shape)
    return (out, None if nln_out is None else nln_out.reshape(x1.shape))


def adapter_merge(down_w: torch.Tensor, down_b: Optional[torch.Tensor], up_w: torch.Tensor,
                  up_b: Optional[torch.Tensor], scale: float, x1: torch.Tensor,
                  mlp_packed: torch.Tensor, token_pos: torch.Tensor,
                  next_ln: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, eps: float = 1e-6):
    """The whole adapter branch inside the merge kernel:
    out = f16(f16(relu(f16(f16(x1) down_w^T + down_b)) up_w^T + up_b) * scale) + (x1 + scatter(mlp_packed)),
    optionally also LayerNorm(out) in fp16.  down_w [K, C], up_w [C, K] fp16."""
    _need_cuda(down_w, up_w, x1, mlp_packed, token_pos)
    x2 = _rows2d(x1)
    m2 = _rows2d(mlp_packed)
    T, Cdim = x2.shape
    K = down_w.shape[0]
    assert down_w.dtype == torch.float16 and up_w.dtype == torch.float16
    assert down_w.shape == (K, Cdim) and up_w.shape == (Cdim, K)
    down_w, up_w = down_w.contiguous(), up_w.contiguous()
    out = torch.empty((T, Cdim), dtype=torch.float32, device=x1.device)
    nln_out = None
    nw = nb = None
    if next_ln is not None:
        nw, nb = next_ln
        nln_out = torch.empty((T, Cdim), dtype=torch.float16, device=x1.device)
    check(_lib.lib().dyt_adapter_merge_fwd(
        down_w.data_ptr(), down_w.stride(0), _ptr(down_b), up_w.data_ptr(), up_w.stride(0), _ptr(up_b),
        float(scale), K, x2.data_ptr(), x2.stride(0), m2.data_ptr(), m2.stride(0),
        token_pos.data_ptr(), T, Cdim, out.data_ptr(), Cdim, _ptr(nw), _ptr(nb), float(eps),
        _ptr(nln_out), Cdim, _stream()), "dyt_adapter_merge_fwd")
    out = out.reshape(x1.shape)
    return (out, None if nln_out is None else nln_out.reshape(x1.shape))


def moe_adapter(x1: torch.Tensor, router_w: torch.Tensor, router_b: Optional[torch.Tensor],
                down_ws, down_bs, up_ws, up_bs, scale: float) -> torch.Tensor:
    """MoE-adapter branch on the kernels (dyt_moe_adapter_fwd; not in the reference, see the header):
    x1 [B, N, C] fp32; lists of E expert weights (down [K, C], up [C, K]) and biases.  Returns
    adapt [B, N, C] fp16."""
    _need_cuda(x1, router_w)
    B, N, Cd = x1.shape
    E, K = len(down_ws), down_ws[0].shape[0]
    h16 = torch.float16
    x32 = x1.to(torch.float32).contiguous()
    x16 = x32.to(h16)
    down_cat = torch.cat([w.detach() for w in down_ws], 0).to(h16).contiguous()
    down_b = torch.stack([b.detach() for b in down_bs], 0).to(h16).contiguous()
    kup = (E * K + E + 7) // 8 * 8
    up_cat = torch.zeros((Cd, kup), dtype=h16, device=x1.device)
    for i in range(E):
        up_cat[:, i * K:(i + 1) * K] = up_ws[i].detach().to(h16)
        up_cat[:, E * K + i] = up_bs[i].detach().to(h16)
    rw = router_w.detach().float().contiguous()
    rb = None if router_b is None else router_b.detach().float().contiguous()
    need = int(_lib.lib().dyt_moe_workspace_bytes(B, N, E, K))
    if need == 0:
        raise DytError("moe_adapter: unsupported shape")
    ws = torch.empty(need, dtype=torch.uint8, device=x1.device)
    out = torch.empty((B * N, Cd), dtype=h16, device=x1.device)
    check(_lib.lib().dyt_moe_adapter_fwd(
        x32.data_ptr(), Cd, x16.data_ptr(), Cd, B, N, Cd, E, K, rw.data_ptr(), _ptr(rb),
        down_cat.data_ptr(), down_b.data_ptr(), up_cat.data_ptr(), float(scale), out.data_ptr(), Cd,
        ws.data_ptr(), ws.numel(), _stream()), "dyt_moe_adapter_fwd")
    return out.reshape(B, N, Cd)


_stem_ws = StreamWorkspaces(zero_filled=False)
# fp16 / fp32 working copies of the stem parameters, keyed by the identity of the PARAMETER OBJECT
# (held weakly: the entry is dropped when the parameter dies), so a new model whose tensors land on
# recycled device addresses can never pick up another model's copies (a cache keyed by data_ptr
# could), and nothing accumulates.
_stem_params = {}


def _stem_get(param):
    ent = _stem_params.get(id(param))
    return ent[1] if ent is not None and ent[0]() is param else None


def _stem_put(param, value):
    pid = id(param)
    _stem_params[pid] = (weakref.ref(param, lambda _r, pid=pid: _stem_params.pop(pid, None)), value)


def patch_embed(img: torch.Tensor, conv_w: torch.Tensor, conv_b: Optional[torch.Tensor],
                cls_token: torch.Tensor, pos_embed: torch.Tensor, patch: int) -> torch.Tensor:
    """ViT stem: img [B, Cin, H, W] fp32 -> tokens [B, L+1, C] fp32 (patch GEMM + cls + pos).
    conv_w [C, Cin, P, P] / conv_b [C] are the PatchEmbed.proj parameters (fp16 copies made here)."""
    _need_cuda(img, conv_w, conv_b, cls_token, pos_embed)
    img = img.to(torch.float32).contiguous()
    B, Cin, H, W = img.shape
    Cdim = conv_w.shape[0]
    L = (H // patch) * (W // patch)
    # fp16 copies of the (frozen) stem parameters, rebuilt only when a parameter changes
    key = (conv_w.data_ptr(), conv_w._version, None if conv_b is None else conv_b._version,
           cls_token.data_ptr(), cls_token._version, pos_embed.data_ptr(), pos_embed._version)
    cached = _stem_get(conv_w)
    if cached is None or cached[0] != key:
        w16 = conv_w.detach().reshape(Cdim, -1).to(torch.float16).contiguous()
        b16 = None if conv_b is None else conv_b.detach().to(torch.float16).contiguous()
        cls = cls_token.detach().reshape(-1).to(torch.float32).contiguous()
        pos = pos_embed.detach().reshape(L + 1, Cdim).to(torch.float32).contiguous()
        cached = (key, w16, b16, cls, pos)
        _stem_put(conv_w, cached)
    _, w16, b16, cls, pos = cached
    need = int(_lib.lib().dyt_patch_embed_workspace_bytes(B, H, W, patch, Cin, Cdim))
    if need == 0:
        raise DytError(f"patch_embed: unsupported geometry H={H} W={W} P={patch}")
    ws = _stem_ws.get(img.device, need)
    x = torch.empty((B, L + 1, Cdim), dtype=torch.float32, device=img.device)
    check(_lib.lib().dyt_patch_embed_fwd(
        img.data_ptr(), B, Cin, H, W, patch, w16.data_ptr(), _ptr(b16), cls.data_ptr(),
        pos.data_ptr(), Cdim, x.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
        "dyt_patch_embed_fwd")
    return x


def invalidate_stem_cache() -> None:
    """Drop the cached stem parameter copies (see engine.invalidate_caches)."""
    _stem_params.clear()


def pool_layernorm_f16(x: torch.Tensor, norm0, norm_k, norm_v, eps: float = 1e-6):
    """(f16(LN_k(LN_0(x))), f16(LN_v(LN_0(x)))) for fp32 rows x [..., C]; norm* = (weight, bias)."""
    _need_cuda(x, *norm0, *norm_k, *norm_v)
    if x.dtype != torch.float32:
        raise DytError("pool_layernorm_f16 expects the fp32 token stream")
    x2 = _rows2d(x)
    T, Cdim = x2.shape
    outk = torch.empty((T, Cdim), dtype=torch.float16, device=x.device)
    outv = torch.empty((T, Cdim), dtype=torch.float16, device=x.device)
    f = lambda t: t.detach().to(torch.float32).contiguous()
    prm = [f(t) for t in (*norm0, *norm_k, *norm_v)]
    check(_lib.lib().dyt_pool_layernorm_f16(
        x2.data_ptr(), x2.stride(0), T, Cdim, *[t.data_ptr() for t in prm], float(eps),
        outk.data_ptr(), outv.data_ptr(), Cdim, _stream()), "dyt_pool_layernorm_f16")
    return outk.reshape(*x.shape[:-1], Cdim), outv.reshape(*x.shape[:-1], Cdim)


def query_attn(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, num_heads: int) -> torch.Tensor:
    """Single-query cross attention.  q fp16 [C] or [b, C] (pre-scaled), k / v fp16 [b, n_keys, C].
    Returns fp16 [b, C]."""
    _need_cuda(q, k, v)
    if any(t.dtype != torch.float16 for t in (q, k, v)):
        raise DytError("query_attn expects fp16 operands")
    b, nk, Cdim = k.shape
    k2, v2 = _rows2d(k), _rows2d(v)
    if k2.stride(0) != v2.stride(0):
        v2 = v2.contiguous()
        k2 = k2.contiguous()
    q = q.contiguous()
    ldq = 0 if q.dim() == 1 or q.shape[0] == 1 else q.stride(0)
    out = torch.empty((b, Cdim), dtype=torch.float16, device=k.device)
    check(_lib.lib().dyt_query_attn_fwd(
        q.data_ptr(), ldq, k2.data_ptr(), v2.data_ptr(), k2.stride(0), b, nk, num_heads,
        Cdim // num_heads, out.data_ptr(), Cdim, _stream()), "dyt_query_attn_fwd")
    return out


# --------------------------------------------------------------------------------------------
# backward kernels (include/dyt_b200.h "backward of the block"); used by dyt_b200.train
# --------------------------------------------------------------------------------------------
def eltwise_f16(op: int, a: torch.Tensor, b: Optional[torch.Tensor] = None,
                c: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Elementwise fp16 kernels (GELU forward / backward, ReLU+dropout backward, product)."""
    _need_cuda(a, b, c)
    for t in (a, b, c):
        if t is not None and (t.dtype != torch.float16 or not t.is_contiguous()):
            raise DytError("eltwise_f16 expects contiguous fp16 tensors")
        if t is not None and t.shape != a.shape:
            raise DytError("eltwise_f16 operands must have one shape")
    out = torch.empty_like(a)
    check(_lib.lib().dyt_eltwise_f16(op, a.data_ptr(), _ptr(b), _ptr(c), out.data_ptr(), a.numel(),
                                     _stream()), "dyt_eltwise_f16")
    return out


def layernorm_bwd(gy: torch.Tensor, x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6,
                  resid: Optional[torch.Tensor] = None, row_scale: Optional[torch.Tensor] = None,
                  axpy: Optional[torch.Tensor] = None, row_idx: Optional[torch.Tensor] = None,
                  out: Optional[torch.Tensor] = None, want_f16_copy: bool = False,
                  n_rows_dev: Optional[torch.Tensor] = None):
    """g_x = resid + dLN(g_y) (+ row_scale[:, None] * axpy).  gy fp16 [R, C]; x fp32 rows (all rows
    of the stream when row_idx selects R of them); returns (fp32 like x, fp16 copy or None)."""
    _need_cuda(gy, x, weight, resid, row_scale, axpy, row_idx, out)
    if gy.dtype != torch.float16 or x.dtype != torch.float32:
        raise DytError("layernorm_bwd expects an fp16 gradient and the fp32 forward input")
    g2, x2 = _rows2d(gy), _rows2d(x)
    R, Cdim = g2.shape
    if row_idx is None and x2.shape[0] != R:
        raise DytError("layernorm_bwd: gradient / input row counts differ")
    if out is None:
        if row_idx is not None:
            raise DytError("layernorm_bwd with row_idx writes into a caller-provided buffer")
        out = torch.empty_like(x2)
    o2 = _rows2d(out)
    r2 = None if resid is None else _rows2d(resid)
    oh = torch.empty(x2.shape, dtype=torch.float16, device=x.device) if want_f16_copy else None
    w = weight.detach().to(torch.float32).contiguous()
    rs = None if row_scale is None else row_scale.reshape(-1).to(torch.float32).contiguous()
    ax = None if axpy is None else axpy.reshape(-1).to(torch.float32).contiguous()
    check(_lib.lib().dyt_layernorm_bwd(
        g2.data_ptr(), g2.stride(0), x2.data_ptr(), x2.stride(0), _ptr(row_idx), _ptr(n_rows_dev), R,
        Cdim, w.data_ptr(), float(eps), _ptr(r2), 0 if r2 is None else r2.stride(0), _ptr(rs), _ptr(ax),
        o2.data_ptr(), o2.stride(0), _ptr(oh), Cdim, _stream()), "dyt_layernorm_bwd")
    return out.reshape(x.shape), (None if oh is None else oh.reshape(x.shape))


def merge_bwd(g_out: torch.Tensor, N: int, mlp_x: Optional[torch.Tensor] = None,
              mask: Optional[torch.Tensor] = None, logits: Optional[torch.Tensor] = None,
              noise: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, tau: float = 5.0,
              g_token_select: Optional[torch.Tensor] = None,
              g_token_logits: Optional[torch.Tensor] = None, masked: bool = True,
              token_pos: Optional[torch.Tensor] = None, sel_w: Optional[torch.Tensor] = None):
    """Backward of the merge + straight-through gate.  g_out fp32 [B*N, C].  Returns
    (g16, g_masked16 or None, g_logit [B*N] fp32 or None, gx_init or None).  With token_pos the
    masked gradient is packed (kept rows only, compacted order); with sel_w the kernel also returns
    gx_init = g_out + g_logit * sel_w (fp32)."""
    _need_cuda(g_out, mlp_x, mask, logits, g_token_select, g_token_logits)
    if g_out.dtype != torch.float32:
        raise DytError("merge_bwd expects the fp32 stream gradient")
    g2 = _rows2d(g_out)
    T, Cdim = g2.shape
    B = T // N
    dev = g_out.device
    g16 = torch.empty((T, Cdim), dtype=torch.float16, device=dev)
    gm16 = gl = m2 = None
    mk = lg = n1 = n2 = gs = gle = None
    if masked:
        if mlp_x is None or mask is None or logits is None:
            raise DytError("merge_bwd(masked=True) needs mlp_x, mask and logits")
        m2 = _rows2d(mlp_x)
        gm16 = torch.empty((T, Cdim), dtype=torch.float16, device=dev)
        gl = torch.empty((T,), dtype=torch.float32, device=dev)
        mk = mask.reshape(-1).to(torch.float32).contiguous()
        lg = logits.reshape(-1).to(torch.float32).contiguous()
        if noise is not None:
            n1 = noise[0].reshape(-1).to(torch.float32).contiguous()
            n2 = noise[1].reshape(-1).to(torch.float32).contiguous()
        if g_token_select is not None:
            gs = g_token_select.reshape(-1).to(torch.float32).contiguous()
        if g_token_logits is not None:
            gle = g_token_logits.reshape(-1).to(torch.float32).contiguous()
    gx = sw = None
    if masked and sel_w is not None:
        sw = sel_w.reshape(-1).to(torch.float32).contiguous()
        gx = torch.empty((T, Cdim), dtype=torch.float32, device=dev)
    tp = token_pos if masked else None
    check(_lib.lib().dyt_merge_bwd(
        g2.data_ptr(), g2.stride(0), _ptr(m2), 0 if m2 is None else m2.stride(0), _ptr(mk),
        _ptr(lg), _ptr(n1), _ptr(n2), float(tau), _ptr(gs), _ptr(gle), B, N, Cdim, g16.data_ptr(),
        Cdim, _ptr(gm16), Cdim, _ptr(gl), _ptr(tp), _ptr(sw), _ptr(gx), Cdim, _stream()),
        "dyt_merge_bwd")
    return g16, gm16, gl, gx


def gelu_bwd_rows(g16: torch.Tensor, pre: torch.Tensor, row_idx: torch.Tensor,
                  n_rows_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[r] = g16[r] * gelu'(pre[row_idx[r]]) for the packed rows (count on the device)."""
    _need_cuda(g16, pre, row_idx, n_rows_dev)
    g2, p2 = _rows2d(g16), _rows2d(pre)
    if g2.dtype != torch.float16 or p2.dtype != torch.float16 or g2.shape[1] != p2.shape[1]:
        raise DytError("gelu_bwd_rows expects fp16 [R, H] and [T, H]")
    out = torch.empty_like(g2)
    check(_lib.lib().dyt_gelu_bwd_rows(g2.data_ptr(), g2.stride(0), p2.data_ptr(), p2.stride(0),
                                       row_idx.data_ptr(), _ptr(n_rows_dev), g2.shape[0],
                                       g2.shape[1], out.data_ptr(), out.stride(0), _stream()),
          "dyt_gelu_bwd_rows")
    return out


def rowscale_colsum(s: torch.Tensor, x16: torch.Tensor, out_w: torch.Tensor,
                    out_b: Optional[torch.Tensor]) -> None:
    """out_w[:] += sum_t s[t] * x16[t, :], out_b += sum_t s[t] (fp32 accumulators, in place)."""
    _need_cuda(s, x16, out_w, out_b)
    x2 = _rows2d(x16)
    if x2.dtype != torch.float16 or s.dtype != torch.float32 or out_w.dtype != torch.float32:
        raise DytError("rowscale_colsum: s fp32, x fp16, out fp32")
    check(_lib.lib().dyt_rowscale_colsum(s.data_ptr(), x2.data_ptr(), x2.stride(0), x2.shape[0],
                                         x2.shape[1], out_w.data_ptr(), _ptr(out_b), _stream()),
          "dyt_rowscale_colsum")


def wgrad_f16(g16: torch.Tensor, x16: torch.Tensor, n_out: Optional[int] = None,
              alpha: float = 1.0, want_bias: bool = True, out=None):
    """(dW [n_out, K] fp32, db [n_out] fp32 or None) = alpha * (g^T x, colsum g); g [T, >= n_out].
    out = (dW, db): accumulate into caller-provided (zero-filled or carried) contiguous buffers."""
    _need_cuda(g16, x16)
    g2, x2 = _rows2d(g16), _rows2d(x16)
    if g2.dtype != torch.float16 or x2.dtype != torch.float16 or g2.shape[0] != x2.shape[0]:
        raise DytError("wgrad_f16 expects fp16 [T, Nout] and [T, K]")
    n_out = g2.shape[1] if n_out is None else n_out
    T, K = x2.shape
    if out is not None:
        dW, db = out
        if dW.shape != (n_out, K) or not dW.is_contiguous() or dW.dtype != torch.float32:
            raise DytError("wgrad_f16: out[0] must be a contiguous fp32 [n_out, K] buffer")
    else:
        dW = torch.zeros((n_out, K), dtype=torch.float32, device=g16.device)
        db = torch.zeros((n_out,), dtype=torch.float32, device=g16.device) if want_bias else None
    check(_lib.lib().dyt_wgrad_f16(g2.data_ptr(), g2.stride(0), x2.data_ptr(), x2.stride(0), T,
                                   n_out, K, float(alpha), dW.data_ptr(), K, _ptr(db), _stream()),
          "dyt_wgrad_f16")
    return dW, db


def attn_varlen_bwd(qkv: torch.Tensor, out: torch.Tensor, d_out: torch.Tensor, num_heads: int,
                    cu_seqlens: Optional[torch.Tensor] = None, num_seqs: Optional[int] = None,
                    max_seqlen: Optional[int] = None) -> torch.Tensor:
    """d_qkv (fp16, like qkv) of attn_varlen."""
    _need_cuda(qkv, out, d_out, cu_seqlens)
    if any(t.dtype != torch.float16 for t in (qkv, out, d_out)):
        raise DytError("attn_varlen_bwd expects fp16 tensors")
    C3 = qkv.shape[-1]
    Cdim = C3 // 3
    q2, o2, g2 = _rows2d(qkv), _rows2d(out), _rows2d(d_out)
    if cu_seqlens is None:
        if qkv.dim() != 3:
            raise DytError("uniform attention expects qkv [B, N, 3C]")
        nseq, uni, mx = qkv.shape[0], qkv.shape[1], qkv.shape[1]
    else:
        nseq = int(num_seqs if num_seqs is not None else cu_seqlens.numel() - 1)
        if max_seqlen is None:
            raise DytError("varlen attention needs max_seqlen (no host sync is done here)")
        uni, mx = 0, int(max_seqlen)
    dqkv = torch.empty_like(q2)
    check(_lib.lib().dyt_attn_varlen_bwd(
        q2.data_ptr(), q2.stride(0), o2.data_ptr(), o2.stride(0), g2.data_ptr(), g2.stride(0),
        _ptr(cu_seqlens), nseq, uni, mx, q2.shape[0], num_heads, Cdim // num_heads, dqkv.data_ptr(),
        dqkv.stride(0), _stream()), "dyt_attn_varlen_bwd")
    return dqkv.reshape(qkv.shape)
