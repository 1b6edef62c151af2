"""Per-(device, stream) scratch buffers for the C-ABI calls (block workspace, dispatcher words, stem
im2col buffer).  Two rules keep CUDA graphs safe and memory bounded:

* a buffer that was handed out while its stream was being captured is never freed when a larger one
  replaces it (a captured graph replays against the old address): it is retired, not dropped;
* buffers of a stream that only ran warm-up iterations can be released explicitly
  (`release_stream`), so a graph wrapper does not pin one full workspace per warm-up stream.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch


class StreamWorkspaces:
    def __init__(self, zero_filled: bool, min_bytes: int = 0):
        self.zero_filled = zero_filled
        self.min_bytes = min_bytes
        self._live: Dict[Tuple[int, int], list] = {}      # key -> [tensor, used_under_capture]
        self._retired: List[torch.Tensor] = []

    def get(self, device: torch.device, need: int) -> torch.Tensor:
        key = (device.index, torch.cuda.current_stream(device).cuda_stream)
        ent = self._live.get(key)
        if ent is None or ent[0].numel() < need:
            if ent is not None and ent[1]:
                self._retired.append(ent[0])
            n = max(int(need), self.min_bytes)
            make = torch.zeros if self.zero_filled else torch.empty
            ent = [make(n, dtype=torch.uint8, device=device), False]
            self._live[key] = ent
        if torch.cuda.is_current_stream_capturing():
            ent[1] = True
        return ent[0]

    def release_stream(self, device: torch.device, stream: torch.cuda.Stream) -> None:
        """Drop the buffers of `stream` unless a capture used them."""
        key = (device.index, stream.cuda_stream)
        ent = self._live.get(key)
        if ent is not None and not ent[1]:
            del self._live[key]

    def retired_bytes(self) -> int:
        return sum(t.numel() for t in self._retired)
