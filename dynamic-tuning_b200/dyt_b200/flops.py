"""Keep-rate / FLOPs accounting of the evaluation loop, on the device.

Drop-in for the reference's `block_flops_dict.batch_select_flops` (block_flops_dict.py:76-83; called
from engine_finetune.py:343) plus a per-layer keep-rate accumulator that replaces gathering every
mask to every rank (engine_finetune.py:245-252, :341-352).  `get_block_flops` / `get_base_flops`
(block_flops_dict.py:33-55, :209-218) trace the model with fvcore, which is not available offline
and cannot see into custom kernels; the analytic tables below follow fvcore's convention (one
multiply-accumulate = one flop, Linear / matmul / conv only) -- parity unpinned for these two
(no fvcore here), exact for everything computed from a given table.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import DytError, check


def block_flops_table(num_tokens: int = 197, dim: int = 768, hidden: int = 3072,
                      bottleneck: int = 64, heads: int = 12) -> torch.Tensor:
    """flops_dict[t], t = 0..num_tokens: GMACs of one DyT block whose MLP runs on t tokens
    (Block.forward_count_flops, vision_transformer_IN21K.py:167-185); index 0 unused (zero)."""
    n, c = num_tokens, dim
    attn = n * c * 3 * c + 2 * n * n * c + n * c * c          # qkv, QK^T + PV, proj
    adapter = 2 * n * c * bottleneck
    selector = (n - 1) * c
    out = torch.zeros(num_tokens + 1, dtype=torch.float32)
    for t in range(1, num_tokens + 1):
        out[t] = (attn + adapter + selector + 2 * t * c * hidden) / 1e9
    return out


def base_flops(num_classes: int = 100, dim: int = 768, patch: int = 16, img: int = 224) -> float:
    """GMACs outside the blocks: patch embedding + head (reference comment: 0.1164 for ViT-B/100)."""
    return ((img // patch) ** 2 * dim * 3 * patch * patch + dim * num_classes) / 1e9


def batch_select_flops(bs, flops_dict, token_select, block_num=12, base_flops=0.116):
    """Same signature and result as the reference function: token_select [B, L, N-1, 1] (or
    [B, L, N-1]) -> fp32 [B] GFLOPs per image, computed by one kernel launch (tensor stays on the
    device; the reference returns a CPU tensor built by a Python loop over images)."""
    ts = token_select
    if ts.dim() == 4:
        ts = ts.squeeze(-1)
    if not ts.is_cuda:
        raise DytError("dyt_b200 kernels run on CUDA tensors only; there is no CPU fallback")
    ts = ts.to(torch.float32).contiguous()
    B, L, Np = ts.shape
    table = torch.as_tensor(flops_dict, dtype=torch.float32, device=ts.device).contiguous()
    out = torch.empty(B, dtype=torch.float32, device=ts.device)
    check(_lib.lib().dyt_keep_stats(ts.data_ptr(), B, L, Np, table.data_ptr(), table.numel(),
                                    int(block_num), float(base_flops), out.data_ptr(), None,
                                    torch.cuda.current_stream().cuda_stream), "dyt_keep_stats")
    return out


class KeepStats:
    """Running per-layer kept-token counters over an evaluation run (device-resident uint64)."""

    def __init__(self, num_layers: int, tokens_per_layer: int, device):
        self.L, self.Np = num_layers, tokens_per_layer
        self.counters = torch.zeros(num_layers + 2, dtype=torch.int64, device=device)

    def update(self, token_select: torch.Tensor) -> None:
        ts = token_select.squeeze(-1) if token_select.dim() == 4 else token_select
        ts = ts.to(torch.float32).contiguous()
        B, L, Np = ts.shape
        if (L, Np) != (self.L, self.Np):
            raise DytError("KeepStats: mask shape changed")
        check(_lib.lib().dyt_keep_stats(ts.data_ptr(), B, L, Np, None, 0, L, 0.0, None,
                                        self.counters.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream), "dyt_keep_stats")

    def all_reduce(self, group=None) -> None:
        """Sum the [L + 2] counters over the ranks: the whole cross-rank exchange of the analytics."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.counters, group=group)

    def layer_rates(self) -> torch.Tensor:
        """Fraction of patch tokens kept per layer (engine_finetune.py:349-351)."""
        n = self.counters[self.L].clamp(min=1).double()
        return (self.counters[: self.L].double() / (n * self.Np)).float()

    def overall_rate(self) -> float:
        return float(self.layer_rates().mean())
