"""Data-parallel fine-tuning plumbing: one flat fp32 gradient arena, one all-reduce per step.

The reference wraps the model in DistributedDataParallel (main_image.py:262-266); with the backbone
frozen only 74 tensors / 1.28 M parameters (5.1 MB fp32) carry gradients, so the
whole exchange is a single NCCL all-reduce over NVLink / NVSwitch.  `.grad` of every trainable
parameter is a view into the arena, autograd accumulates in place, and `all_reduce_mean()` is the
only collective of the step (the forward / backward data path has none: images are independent).
The keep-rate term of the loss is evaluated per rank on the local batch, as in the reference
(models/losses.py:69-72 runs inside each DDP replica).
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


class GradArena:
    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradArena: no trainable parameters")
        dev = self.params[0].device
        if any(p.device != dev for p in self.params):
            raise ValueError("GradArena: parameters live on different devices")
        self.offsets = []
        total = 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + 3) // 4 * 4          # 16-byte aligned slices
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.attach()

    def attach(self) -> None:
        """(Re)bind every .grad to its arena slice (call again if something set grads to None)."""
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)

    def zero(self) -> None:
        self.flat.zero_()
        if any(p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * off
               for p, off in zip(self.params, self.offsets)):
            self.attach()

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def all_reduce_mean(self, group=None) -> None:
        """Average the arena over the data-parallel ranks (no-op without a process group)."""
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(group)
        if world == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.mul_(1.0 / world)


def trainable_parameters(model: torch.nn.Module) -> List[torch.nn.Parameter]:
    """The reference's freeze rule (main_image.py:242-256): everything that a backbone checkpoint
    does not provide, i.e. adaptmlp.*, mlp_token_select.* and head.*, is trained."""
    out = []
    for name, p in model.named_parameters():
        p.requires_grad = ("adaptmlp" in name) or ("mlp_token_select" in name) or name.startswith("head.")
        if p.requires_grad:
            out.append(p)
    return out
