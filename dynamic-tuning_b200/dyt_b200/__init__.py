"""dyt_b200: B200-native (sm_100a) token-dispatched ViT block forward for Dynamic-Tuning (DyT).

Python here is host plumbing only (device memory, streams, module surface); all block compute is
in libdyt_b200.so (hand-written CUDA: TMA + tcgen05 GEMM / attention, fused dispatcher, scatter
merge) reached through the C ABI declared in include/dyt_b200.h.  There is no CPU fallback.
"""
from ._lib import ABI_VERSION, DytError, LIB_PATH, lib  # noqa: F401
from .gate import min_kept_logit  # noqa: F401
from .graph import GraphedForward  # noqa: F401,E402
from .engine import invalidate_caches  # noqa: F401,E402
