"""The keep/drop gate as a monotone threshold on the logit.

The reference gate is `sigmoid(logits) > threshold` evaluated in the logits' dtype
(models/dynamic_adapter.py:44-50, models/model_speed_test.py:30-34).  sigmoid is monotone, so the
gate is equivalent to `logit >= min_kept(dtype, threshold)`.  min_kept is derived here from torch's
own sigmoid (exhaustively for 16-bit dtypes, by bisection for fp32), which makes the kernel's masks
bit-identical to the reference's gate by construction (fp16 keeps iff
logit > 2^-10).
"""
from __future__ import annotations

import functools

import torch


def _gate(vals: torch.Tensor, threshold: float) -> torch.Tensor:
    return vals.sigmoid() > threshold


@functools.lru_cache(maxsize=None)
def min_kept_logit(dtype: torch.dtype, threshold: float = 0.5) -> float:
    if dtype in (torch.float16, torch.bfloat16):
        bits = torch.arange(0, 1 << 16, dtype=torch.int32).to(torch.int16)
        vals = bits.view(dtype)
        keep = _gate(vals, threshold) & torch.isfinite(vals)
        if not bool(keep.any()):
            return float("inf")
        return float(vals[keep].float().min())
    lo = torch.tensor(-64.0, dtype=torch.float32)
    hi = torch.tensor(64.0, dtype=torch.float32)
    if not bool(_gate(hi.view(1), threshold)):
        return float("inf")
    if bool(_gate(lo.view(1), threshold)):
        return float("-inf")
    for _ in range(200):
        mid = ((lo.double() + hi.double()) / 2).float()
        if mid == lo or mid == hi:
            break
        if bool(_gate(mid.view(1), threshold)):
            hi = mid
        else:
            lo = mid
    return float(hi)
