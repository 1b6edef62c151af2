"""One fine-tuning step of the reference's recipe on the sm_100a path.

Mirrors engine_finetune.py:47-76 (train_one_epoch inner loop): student pass, teacher pass
(complete_model=True), AdaLoss (models/losses.py:50-82: cross-entropy + token_loss_ratio * keep-rate
loss) + teacher cross-entropy + KL(student || teacher), scaled backward, gradient all-reduce,
optimizer step.  The loss itself is a handful of torch ops on [B, classes] logits and the [B, L,
N-1, 1] masks (host-side glue in the reference too); forward / backward of the model are kernels.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from .ddp import GradArena


def ada_loss(student_logits, token_select, targets, token_target_ratio=0.5, token_loss_ratio=2.0,
             token_minimal=0.0, token_minimal_weight=0.0):
    """models/losses.py:50-82 with a cross-entropy base criterion.  The defaults are the recipe of
    the reference's entry scripts, NOT the AdaLoss class defaults: every main_*.py passes
    token_ratio 2, token_minimal 0, token_minimal_weight 0 (main_image.py:206-209,
    main_vtab.py:200-203, main_video.py:234-237), i.e. the per-token minimal-keep term is off."""
    base = F.cross_entropy(student_logits, targets)
    token_loss = ((token_select.mean() - token_target_ratio) ** 2).mean()
    if token_minimal_weight > 0:                                   # losses.py:76-80
        token_loss = token_loss + token_minimal_weight * (
            token_minimal - token_select.mean(-1)).clamp(min=0.0).sum()
    return base + token_loss_ratio * token_loss


def finetune_loss(student_logits, token_select, teacher_logits, targets, **kw):
    """engine_finetune.py:52-65."""
    kl = F.kl_div(F.log_softmax(student_logits, dim=-1),
                  F.log_softmax(teacher_logits.detach(), dim=-1), reduction="batchmean",
                  log_target=True)
    teacher = F.cross_entropy(teacher_logits, targets)
    return ada_loss(student_logits, token_select, targets, **kw) + teacher + kl


class FinetuneStep:
    """model: TrainVisionTransformer with the reference's freeze rule applied; optimizer over the
    trainable parameters; arena: their flat gradient buffer (all-reduced once per step).  Loss scaling
    is torch's dynamic GradScaler, as in the reference (util/misc.py NativeScalerWithGradNormCount):
    the keep-rate loss puts a gradient of token_loss_ratio * scale on every dropped token's fp16
    mask entry, which overflows at the initial scale 65536; the scaler skips those steps and backs
    off exactly as it does for the reference."""

    def __init__(self, model, optimizer, arena: GradArena, token_target_ratio: float = 0.5,
                 init_scale: float = 65536.0, cuda_graph: bool = False):
        self.model, self.opt, self.arena = model, optimizer, arena
        self.scaler = torch.amp.GradScaler("cuda", init_scale=init_scale)
        self.ratio = token_target_ratio
        self.cuda_graph = cuda_graph
        self._graph = None

    def _forward_backward(self, images, targets):
        self.arena.zero()
        with torch.autocast("cuda", dtype=torch.float16):
            out_s, ts = self.model(images)
            out_t, _ = self.model(images, complete_model=True)
            loss = finetune_loss(out_s.float(), ts["token_select"].float(), out_t.float(), targets,
                                 token_target_ratio=self.ratio)
        self.scaler.scale(loss).backward()
        return loss.detach()

    def _capture(self, images, targets):
        """Whole forward + backward of the step as ONE CUDA graph (the launch-bound part: ~1100 kernel
        launches per step); the random draws (Gumbel noise, dropout) advance with every replay
        through torch's graph-safe generator, the loss scale is read from device memory."""
        self._img, self._tgt = images.clone(), targets.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):                       # warm-up off the capture: caches, workspaces
                self._forward_backward(self._img, self._tgt)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import engine
        engine.release_stream_workspaces(self._img.device, side)   # the warm-up stream's scratch buffers
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._loss = self._forward_backward(self._img, self._tgt)

    def __call__(self, images: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        if self.cuda_graph:
            if self._graph is None:
                self._capture(images, targets)
            self._img.copy_(images, non_blocking=True)
            self._tgt.copy_(targets, non_blocking=True)
            self._graph.replay()
            loss = self._loss
        else:
            loss = self._forward_backward(images, targets)
        self.arena.all_reduce_mean()          # the step's only collective (sum of scaled grads / world)
        self.scaler.step(self.opt)            # unscale, skip on inf / nan
        self.scaler.update()
        return loss
