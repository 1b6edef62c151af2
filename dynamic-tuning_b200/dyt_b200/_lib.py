"""ctypes binding of libdyt_b200.so (C ABI: include/dyt_b200.h).

The library is the product: there is no Python/CPU fallback.  Importing this module never needs a
GPU (so the CPU test-suite can check that the library loads and exports every declared symbol),
but every compute entry point requires CUDA tensors and raises DytError otherwise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdyt_b200.so")

ABI_VERSION = 3

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RELU, EPI_BIAS_RESID = 0, 1, 2, 3
EPI_BIAS_GELU_KEEP, EPI_DGELU = 4, 5
OPT_PDL = 1
OPT_GEMM_TAIL_SPLIT = 2
OPT_FUSE_ADAPTER_UP = 3
OPT_ATTN_SPLIT = 4
OPT_FUSE_ADAPTER_DOWN = 5
OPT_TILE_ORDER = 6
OPT_SIDE_PLAN = 7
OPT_SM_LIMIT = 8
EW_GELU_FWD, EW_GELU_BWD, EW_RELU_DROP_BWD, EW_MUL = 0, 1, 2, 3


class DytError(RuntimeError):
    pass


class BlockShape(C.Structure):
    _fields_ = [("B", C.c_int), ("N", C.c_int), ("C", C.c_int), ("H", C.c_int),
                ("hidden", C.c_int), ("bottleneck", C.c_int)]


class BlockWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln1_w", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ln2_w", "ln2_b",
        "fc1_w", "fc1_b", "fc2_w", "fc2_b", "down_w", "down_b", "up_w", "up_b",
        "sel_w", "sel_b")] + [("adapter_scale", C.c_float)]


class BlockOpts(C.Structure):
    _fields_ = [("struct_size", C.c_size_t),
                ("eps", C.c_float), ("logit_fp16", C.c_int), ("min_kept", C.c_float),
                ("noise1", C.c_void_p), ("noise2", C.c_void_p), ("tau", C.c_float),
                ("forced_mask", C.c_void_p), ("gate_out", C.c_void_p), ("xn_ready", C.c_int),
                ("next_ln_w", C.c_void_p), ("next_ln_b", C.c_void_p), ("attn_bias", C.c_void_p),
                ("attn_bias_ld", C.c_int), ("moe_experts", C.c_int), ("moe_router_w", C.c_void_p),
                ("moe_router_b", C.c_void_p), ("moe_workspace", C.c_void_p),
                ("moe_workspace_bytes", C.c_size_t)]


class BlockBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "xn", "attn_o", "qkv", "x1", "x1h", "packed", "hidden", "mlp", "down", "adapt",
        "packed_idx", "token_pos", "cu_seqlens", "n_kept")]


_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every function declared in include/dyt_b200.h
SIGNATURES = {
    "dyt_version": (_i, []),
    "dyt_last_error": (C.c_char_p, []),
    "dyt_configure": (_i, [_i, _i]),
    "dyt_linear_f16": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _vp, _i,
                            _f, _vp]),
    "dyt_linear_f16_aux": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _vp]),
    "dyt_attn_varlen_fwd": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "dyt_attn_bias_fwd": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "dyt_layernorm_f16": (_i, [_vp, _i, _vp, _vp, _i, _i, _vp, _vp, _f, _vp, _i, _vp]),
    "dyt_dispatch_workspace_bytes": (_sz, [_i]),
    "dyt_dispatch_fwd": (_i, [_vp, _i, _vp, _vp, _i, _f, _vp, _vp, _f, _i, _i, _i, _vp, _vp, _f,
                              _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "dyt_token_select_fwd": (_i, [_vp, _i, _vp, _vp, _i, _f, _vp, _vp, _f, _i, _i, _i, _vp, _vp, _vp,
                                  _vp]),
    "dyt_scatter_merge_fwd": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _vp, _i, _vp, _vp, _f,
                                   _vp, _i, _vp]),
    "dyt_merge_up_fwd": (_i, [_vp, _i, _vp, _i, _vp, _f, _i, _vp, _i, _vp, _i, _vp, _i, _i, _vp, _i,
                              _vp, _vp, _f, _vp, _i, _vp]),
    "dyt_adapter_merge_fwd": (_i, [_vp, _i, _vp, _vp, _i, _vp, _f, _i, _vp, _i, _vp, _i, _vp, _i, _i, _vp,
                                   _i, _vp, _vp, _f, _vp, _i, _vp]),
    "dyt_moe_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "dyt_moe_adapter_fwd": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _i,
                                 _vp, _sz, _vp]),
    "dyt_patch_embed_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "dyt_patch_embed_fwd": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "dyt_pool_layernorm_f16": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _i, _vp]),
    "dyt_query_attn_fwd": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "dyt_layernorm_bwd": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _vp, _f, _vp, _i, _vp, _vp, _vp,
                               _i, _vp, _i, _vp]),
    "dyt_merge_bwd": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _f, _vp, _vp, _i, _i, _i, _vp, _i,
                           _vp, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "dyt_gelu_bwd_rows": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _vp, _i, _vp]),
    "dyt_rowscale_colsum": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "dyt_eltwise_f16": (_i, [_i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dyt_wgrad_f16": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _f, _vp, _i, _vp, _vp]),
    "dyt_attn_varlen_bwd": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "dyt_keep_stats": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _f, _vp, _vp, _vp]),
    "dyt_block_workspace_bytes": (_sz, [C.POINTER(BlockShape)]),
    "dyt_block_workspace_layout": (_i, [C.POINTER(BlockShape), _vp, C.POINTER(BlockBuffers)]),
    "dyt_block_fwd": (_i, [C.POINTER(BlockShape), C.POINTER(BlockWeights), C.POINTER(BlockOpts),
                           _vp, _vp, _vp, _vp, _sz, _vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libdyt_b200.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DytError(
                f"{LIB_PATH} is missing: build it with `make -C dynamic-tuning_b200/csrc` "
                "(or __graft_entry__.build()). dyt_b200 has no fallback path.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError here == missing export: fail loudly
            fn.restype = res
            fn.argtypes = args
        got = handle.dyt_version()
        if got != ABI_VERSION:
            raise DytError(f"libdyt_b200.so ABI version {got}, binding expects {ABI_VERSION}")
        _lib = handle
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().dyt_last_error().decode("utf-8", "replace")
        kind = "CUDA error" if status > 0 else "argument error"
        raise DytError(f"{what or 'dyt_b200'} failed with status {status} ({kind}): {msg}")
