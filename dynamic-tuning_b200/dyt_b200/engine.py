"""Host-side driver of the block forward: fp16 weight cache, workspace, one C-ABI call per block.

`run_blocks` is what the drop-in `models.*.VisionTransformer.forward_features` loop calls instead
of `for blk in self.blocks: x = blk(x)` (reference models/model_speed_test.py:478-479,
models/vision_transformer_IN21K.py:357-365).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._wscache import StreamWorkspaces
from ._lib import BlockBuffers, BlockOpts, BlockShape, BlockWeights, DytError, check
from .gate import min_kept_logit

_WEIGHT_FIELDS_F16 = {
    "qkv_w": "attn.qkv.weight", "qkv_b": "attn.qkv.bias",
    "proj_w": "attn.proj.weight", "proj_b": "attn.proj.bias",
    "fc1_w": "mlp.fc1.weight", "fc1_b": "mlp.fc1.bias",
    "fc2_w": "mlp.fc2.weight", "fc2_b": "mlp.fc2.bias",
    "down_w": "adaptmlp.down_proj.weight", "down_b": "adaptmlp.down_proj.bias",
    "up_w": "adaptmlp.up_proj.weight", "up_b": "adaptmlp.up_proj.bias",
}
_WEIGHT_FIELDS_F32 = {
    "ln1_w": "norm1.weight", "ln1_b": "norm1.bias", "ln2_w": "norm2.weight", "ln2_b": "norm2.bias",
    "sel_w": "mlp_token_select.mlp_head.weight", "sel_b": "mlp_token_select.mlp_head.bias",
}


def _get(module: torch.nn.Module, dotted: str) -> torch.Tensor:
    obj = module
    for part in dotted.split("."):
        obj = getattr(obj, part)
    return obj


class PreparedBlock:
    """Device-resident fp16 copies of a Block's Linear parameters + the ctypes weight struct.
    Rebuilt automatically when a parameter is modified in place or replaced (fine-tuning)."""

    def __init__(self, block: torch.nn.Module):
        self.block = block
        self.key = None
        self.tensors: Dict[str, torch.Tensor] = {}
        self.struct = BlockWeights()

    def _is_moe(self) -> bool:
        return getattr(self.block.adaptmlp, "num_experts", 0) > 1

    def _signature(self):
        sig = []
        moe = self._is_moe()
        for dotted in list(_WEIGHT_FIELDS_F16.values()) + list(_WEIGHT_FIELDS_F32.values()):
            if moe and dotted.startswith("adaptmlp."):
                continue
            p = _get(self.block, dotted)
            sig.append((p.data_ptr(), p._version, p.device))
        if moe:
            for p in self.block.adaptmlp.parameters():
                sig.append((p.data_ptr(), p._version, p.device))
        return tuple(sig)

    def _prepare_moe(self):
        """MoE-adapter: the experts' weights in the layout dyt_moe_adapter_fwd takes (see the header):
        down_cat [E*K, C], down_b [E, K], up_cat [C, round8(E*K + E)] with the biases as columns."""
        a = self.block.adaptmlp
        E, K, Cd = a.num_experts, a.down_size, a.n_embd
        h16 = torch.float16
        dev = a.router.weight.device
        down_cat = torch.cat([l.weight.detach() for l in a.down_proj], dim=0).to(h16).contiguous()
        down_b = torch.stack([l.bias.detach() for l in a.down_proj], dim=0).to(h16).contiguous()
        kup = (E * K + E + 7) // 8 * 8
        up_cat = torch.zeros((Cd, kup), dtype=h16, device=dev)
        for i, l in enumerate(a.up_proj):
            up_cat[:, i * K:(i + 1) * K] = l.weight.detach().to(h16)
            up_cat[:, E * K + i] = l.bias.detach().to(h16)
        self.tensors.update(down_w=down_cat, down_b=down_b, up_w=up_cat, up_b=down_b,
                            moe_router_w=a.router.weight.detach().float().contiguous(),
                            moe_router_b=a.router.bias.detach().float().contiguous())
        for field in ("down_w", "down_b", "up_w", "up_b"):
            setattr(self.struct, field, self.tensors[field].data_ptr())

    def get(self) -> BlockWeights:
        sig = self._signature()
        if sig != self.key:
            with torch.no_grad():
                moe = self._is_moe()
                if moe:
                    self._prepare_moe()
                for field, dotted in _WEIGHT_FIELDS_F16.items():
                    if moe and dotted.startswith("adaptmlp."):
                        continue
                    p = _get(self.block, dotted)
                    if not p.is_cuda:
                        raise DytError("dyt_b200: block parameters must live on a CUDA device")
                    t = p.detach().to(torch.float16).contiguous()
                    self.tensors[field] = t
                    setattr(self.struct, field, t.data_ptr())
                for field, dotted in _WEIGHT_FIELDS_F32.items():
                    p = _get(self.block, dotted)
                    t = p.detach().to(torch.float32).contiguous().reshape(-1)
                    self.tensors[field] = t
                    setattr(self.struct, field, t.data_ptr())
            scale = self.block.adaptmlp.scale
            self.struct.adapter_scale = float(scale.item() if torch.is_tensor(scale) else scale)
            self.key = sig
        return self.struct


_workspaces = StreamWorkspaces(zero_filled=True)    # zero-filled once (ABI contract)
_moe_workspaces = StreamWorkspaces(zero_filled=False)


def block_shape_of(block: torch.nn.Module, B: int, N: int) -> BlockShape:
    C_ = block.attn.qkv.in_features
    H = block.attn.num_heads
    if C_ != 64 * H:
        raise DytError(f"dyt_b200 attention kernel needs head_dim 64 (C={C_}, heads={H})")
    bott = (block.adaptmlp.down_size if getattr(block.adaptmlp, "num_experts", 0) > 1
            else block.adaptmlp.down_proj.out_features)
    return BlockShape(B, N, C_, H, block.mlp.fc1.out_features, bott)


def _workspace(shape: BlockShape, device: torch.device) -> torch.Tensor:
    need = int(_lib.lib().dyt_block_workspace_bytes(C.byref(shape)))
    if need == 0:
        raise DytError("dyt_block_workspace_bytes rejected the shape: " +
                       _lib.lib().dyt_last_error().decode())
    return _workspaces.get(device, need)


def release_stream_workspaces(device: torch.device, stream: torch.cuda.Stream) -> None:
    """Free the scratch buffers that calls on `stream` allocated (block workspace, dispatcher words,
    stem buffer), unless a CUDA-graph capture used them.  For wrappers that warm up on a throw-away
    stream before capturing."""
    from . import ops
    _workspaces.release_stream(device, stream)
    _moe_workspaces.release_stream(device, stream)
    ops._dispatch_ws.release_stream(device, stream)
    ops._stem_ws.release_stream(device, stream)


def workspace_buffers(shape: BlockShape, ws: torch.Tensor) -> BlockBuffers:
    bufs = BlockBuffers()
    check(_lib.lib().dyt_block_workspace_layout(C.byref(shape), ws.data_ptr(), C.byref(bufs)),
          "dyt_block_workspace_layout")
    return bufs


def _prepared(block: torch.nn.Module) -> PreparedBlock:
    prep = block.__dict__.get("_dyt_prepared")
    if prep is None:
        prep = PreparedBlock(block)
        block.__dict__["_dyt_prepared"] = prep
    return prep


def invalidate_caches(module: torch.nn.Module) -> None:
    """Forget every cached fp16 / fp32 working copy of `module`'s parameters (block weights, head,
    stem).  The caches notice a parameter that was replaced or modified in place through autograd-
    visible ops (`_version`), i.e. optimizer steps and load_state_dict; a write through `.data`
    (`p.data.copy_(...)`, `p.data.mul_(...)`) bumps no version counter: call this after such writes."""
    from . import ops
    for m in module.modules():
        for key in [k for k in m.__dict__ if k.startswith("_dyt_")]:
            m.__dict__.pop(key, None)
    ops.invalidate_stem_cache()


def run_blocks(x: torch.Tensor, blocks: Sequence[torch.nn.Module], *, eps: float = 1e-6,
               logit_dtype: torch.dtype = torch.float16, forced_masks=None,
               noises=None, final_ln: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
               fuse_next_ln: bool = True, report_gate: bool = False, attn_biases=None,
               consume_input: bool = False):
    """x [B, N, C] fp32 CUDA -> (x_out fp32 [B,N,C], masks [L,B,N] f32, logits [L,B,N-1] f32,
    final_ln_out f16 [B,N,C] or None).  report_gate=True returns the selectors' own decisions as
    `masks` even where a mask is forced (teacher pass, complete_model=True).  forced_masks[i] ([B,N] or [B,N,1]) imposes layer i's mask;
    noises[i] = (g1, g2) switches layer i's gate to the train-mode Gumbel form.  attn_biases[i]
    (fp32 [H, N, N] or None) adds a per-head bias to layer i's attention scores (segmentation
    backbone); sequences longer than 256 tokens or a bias run the flash-style attention kernel.
    consume_input=True lets the call overwrite `x` (the returned stream may alias it)."""
    if not x.is_cuda:
        raise DytError("dyt_b200.run_blocks needs a CUDA tensor: the sm_100a kernels are the only "
                       "implementation (no CPU fallback)")
    if x.dim() != 3:
        raise DytError("run_blocks expects x [B, N, C]")
    # the blocks update the stream in place: work on a copy unless the caller hands over a temporary
    # (consume_input=True: the model path passes the stem's fresh output -- saves a 155 MB copy per
    # forward at 256 images)
    x32 = x.to(torch.float32).contiguous()
    x = x32 if (consume_input or x32.data_ptr() != x.data_ptr()) else x32.clone()
    B, N, _ = x.shape
    L = len(blocks)
    dev = x.device
    masks = torch.empty((L, B, N), dtype=torch.float32, device=dev)
    logits = torch.empty((L, B, max(N - 1, 1)), dtype=torch.float32, device=dev)
    gates = torch.empty((L, B, N), dtype=torch.float32, device=dev) if report_gate else None
    lib = _lib.lib()
    stream = torch.cuda.current_stream().cuda_stream
    keep_alive = []
    xn_ready = 0
    shape = ws = None
    for i, blk in enumerate(blocks):
        shape = block_shape_of(blk, B, N)
        ws = _workspace(shape, dev)
        wt = _prepared(blk).get()
        opts = BlockOpts()
        opts.struct_size = C.sizeof(BlockOpts)
        opts.eps = eps
        opts.logit_fp16 = 1 if logit_dtype == torch.float16 else 0
        thr = float(getattr(blk.mlp_token_select, "threshold", 0.5))
        opts.min_kept = float(min_kept_logit(logit_dtype, thr))
        opts.tau = float(getattr(blk.mlp_token_select, "tau", 5.0))
        if noises is not None and noises[i] is not None:
            g1 = noises[i][0].to(torch.float32).contiguous()
            g2 = noises[i][1].to(torch.float32).contiguous()
            keep_alive += [g1, g2]
            opts.noise1, opts.noise2 = g1.data_ptr(), g2.data_ptr()
        if forced_masks is not None and forced_masks[i] is not None:
            fm = forced_masks[i].to(device=dev, dtype=torch.float32).reshape(B, N).contiguous()
            keep_alive.append(fm)
            opts.forced_mask = fm.data_ptr()
        if gates is not None:
            opts.gate_out = gates[i].data_ptr()
        if attn_biases is not None and attn_biases[i] is not None:
            from . import ops
            ab, ab_ld = ops._bias_pitch(attn_biases[i].to(device=dev), shape.H, N)
            keep_alive.append(ab)
            opts.attn_bias = ab.data_ptr()
            opts.attn_bias_ld = ab_ld
        if getattr(blk.adaptmlp, "num_experts", 0) > 1:      # MoE-adapter (not in the reference)
            prep = _prepared(blk)
            E_, K_ = blk.adaptmlp.num_experts, blk.adaptmlp.down_size
            need = int(lib.dyt_moe_workspace_bytes(B, N, E_, K_))
            mws = _moe_workspaces.get(dev, need)
            opts.moe_experts = E_
            opts.moe_router_w = prep.tensors["moe_router_w"].data_ptr()
            opts.moe_router_b = prep.tensors["moe_router_b"].data_ptr()
            opts.moe_workspace = mws.data_ptr()
            opts.moe_workspace_bytes = mws.numel()
        opts.xn_ready = xn_ready
        nxt = None
        if fuse_next_ln:
            if i + 1 < L:
                nprep = _prepared(blocks[i + 1])
                nprep.get()
                nxt = (nprep.tensors["ln1_w"], nprep.tensors["ln1_b"])
            elif final_ln is not None:
                nxt = (final_ln[0].detach().float().contiguous(), final_ln[1].detach().float().contiguous())
                keep_alive += list(nxt)
        if nxt is not None:
            opts.next_ln_w, opts.next_ln_b = nxt[0].data_ptr(), nxt[1].data_ptr()
        check(lib.dyt_block_fwd(C.byref(shape), C.byref(wt), C.byref(opts), x.data_ptr(),
                                masks[i].data_ptr(), logits[i].data_ptr(), ws.data_ptr(),
                                ws.numel(), stream), f"dyt_block_fwd(layer {i})")
        xn_ready = 1 if (nxt is not None and i + 1 < L) else 0
    final_out = None
    if fuse_next_ln and final_ln is not None and L > 0:
        bufs = workspace_buffers(shape, ws)
        C_ = shape.C
        off = bufs.xn - ws.data_ptr()
        final_out = ws[off:off + B * N * C_ * 2].view(torch.float16).reshape(B, N, C_)
    return x, (gates if gates is not None else masks), logits, final_out
