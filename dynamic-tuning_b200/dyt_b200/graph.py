"""CUDA-graph replay of the inference forward.

One forward of the 12-layer model is ~115 kernel launches of 10-140 us; launched eagerly from Python
the gaps between them depend on the host (measured 0.3-0.8 ms per 9 ms step across boxes).  The C ABI
never synchronises or allocates and every data-dependent count stays on the device, so the whole
`model(images)` call captures into one CUDA graph (tests/test_block_gpu.py::
test_forward_is_cuda_graph_capturable_and_replay_safe): `GraphedForward` does that once per input
shape and replays it.  Same role as `torch.compile(mode="reduce-overhead")` around the reference's
speed.py loop, without a compiler.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from ._lib import DytError


class GraphedForward:
    """gm = GraphedForward(model); logits = gm(images)   (inference, fp16 autocast, no grad).
    `images` may be any CUDA tensor of a shape seen before (one graph per shape and per static
    input slot); the result tensor is owned by the graph and overwritten by the next call on the
    same slot.  `slot` selects one of several static input buffers so that a caller can fill the
    next input (e.g. an H2D copy on another stream, via `input_buffer`) while a replay runs.
    All graphs of one GraphedForward share the library's scratch buffers (one block workspace per
    device): replay them on one stream at a time."""

    def __init__(self, model: torch.nn.Module, autocast_dtype=torch.float16):
        self.model = model
        self.dtype = autocast_dtype
        self._graphs: Dict[Tuple, Tuple[torch.cuda.CUDAGraph, torch.Tensor, object]] = {}
        self._capture_streams: Dict[torch.device, torch.cuda.Stream] = {}

    def _eager(self, x):
        with torch.no_grad(), torch.autocast("cuda", dtype=self.dtype):
            return self.model(x)

    def _entry(self, shape, dtype, device, slot):
        key = (tuple(shape), dtype, device, slot)
        ent = self._graphs.get(key)
        if ent is None:
            if self.model.training:
                raise DytError("GraphedForward is for inference: call model.eval() first")
            static_in = torch.zeros(shape, dtype=dtype, device=device)
            # Warm-up and capture run on ONE stream owned by this object: the scratch buffers are
            # keyed by stream, so the warm-up allocates (and zero-fills, once) the workspace the
            # capture then reuses.  Captured on a fresh stream instead, the allocation happened
            # inside the capture and its zero-fill (1.1 GB at 256 images) was replayed every step.
            side = self._capture_streams.get(device)
            if side is None:
                side = torch.cuda.Stream(device=device)
                self._capture_streams[device] = side
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for _ in range(2):                 # warm-up: weight casts, workspaces, lazy init
                    self._eager(static_in)
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize(device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                static_out = self._eager(static_in)
            ent = (graph, static_in, static_out)
            self._graphs[key] = ent
        return ent

    def input_buffer(self, shape, dtype=torch.float32, device=None, slot: int = 0) -> torch.Tensor:
        """The static input tensor of a slot (captures the graph on first use): write the next
        batch into it, then call `replay(slot)`."""
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        return self._entry(shape, dtype, device, slot)[1]

    def replay(self, shape, dtype=torch.float32, device=None, slot: int = 0):
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        graph, _, out = self._entry(shape, dtype, device, slot)
        graph.replay()
        return out

    def __call__(self, images: torch.Tensor, slot: int = 0):
        if not images.is_cuda:
            raise DytError("dyt_b200 needs CUDA tensors (no CPU fallback)")
        graph, static_in, out = self._entry(images.shape, images.dtype, images.device, slot)
        if images.data_ptr() != static_in.data_ptr():
            static_in.copy_(images, non_blocking=True)
        graph.replay()
        return out
