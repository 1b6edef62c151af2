"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/dyt_b200.h declares (no compute calls), argument errors surface as status codes + text,
and the Python binding lists the same functions as the header."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dyt_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dyt_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = _declared_functions()
    for must in ("dyt_version", "dyt_last_error", "dyt_linear_f16", "dyt_attn_varlen_fwd",
                 "dyt_layernorm_f16", "dyt_dispatch_fwd", "dyt_scatter_merge_fwd", "dyt_block_fwd",
                 "dyt_block_workspace_bytes"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from dyt_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build libdyt_b200.so first (__graft_entry__.build())"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == _declared_functions(), "binding and header out of sync"
    assert _lib.lib().dyt_version() == _lib.ABI_VERSION


def test_option_ids_match_the_header_and_configure_accepts_them():
    """Every DYT_OPT_* id of the header has the same value in the binding, dyt_configure accepts each
    of them (host-side switches: no GPU needed) and rejects an unknown id with a status + text."""
    from dyt_b200 import _lib
    ids = dict(re.findall(r"#define\s+DYT_OPT_([A-Z_]+)\s+(\d+)", open(HEADER).read()))
    assert len(ids) >= 8 and len(set(ids.values())) == len(ids), "duplicate option ids"
    defaults = {"PDL": 1, "GEMM_TAIL_SPLIT": 1, "FUSE_ADAPTER_UP": 1, "ATTN_SPLIT": 1,
                "FUSE_ADAPTER_DOWN": 0, "TILE_ORDER": 7, "SIDE_PLAN": 0, "SM_LIMIT": 0}
    assert set(defaults) == set(ids), "an option without a documented default (or the reverse)"
    lib = _lib.lib()
    for name, value in ids.items():
        assert getattr(_lib, "OPT_" + name) == int(value), name
        assert lib.dyt_configure(int(value), defaults[name]) == 0, name    # the default again
    assert lib.dyt_configure(999, 1) != 0
    assert b"unknown option" in lib.dyt_last_error()


def test_struct_layouts_match_header_field_order():
    from dyt_b200 import _lib
    src = open(HEADER).read()
    for cname, cls in (("dyt_block_shape", _lib.BlockShape), ("dyt_block_weights", _lib.BlockWeights),
                       ("dyt_block_opts", _lib.BlockOpts), ("dyt_block_buffers", _lib.BlockBuffers)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), src, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                fields.append(re.sub(r"[\*\s]", " ", decl).split()[-1])
        assert fields == [f[0] for f in cls._fields_], cname


def _header_fields(cname):
    src = open(HEADER).read()
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            fields.append(re.sub(r"[\*\s]", " ", decl).split()[-1])
    return fields


def test_integration_doc_stub_matches_the_header():
    """INTEGRATION.md shows the ctypes stub a reference maintainer would add; its structure field
    lists must be the header's, field for field (the round-1 stub had lost `attn_bias`: the library
    would have read past the end of the caller's struct)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for cls, cname in (("Shape", "dyt_block_shape"), ("Weights", "dyt_block_weights"),
                       ("Opts", "dyt_block_opts")):
        block = re.search(r"class %s\(ctypes\.Structure\):(.*?)\n(?=class |\n|_lib)" % cls, doc, flags=re.S)
        assert block is not None, cls
        names = re.findall(r'"([A-Za-z_0-9]+)"', block.group(1))
        assert names == _header_fields(cname), (cls, names)
    assert "struct_size=ctypes.sizeof(Opts)" in doc
    assert re.search(r"dyt_version\(\) == (\d+)", doc).group(1) == re.search(
        r"#define DYT_ABI_VERSION (\d+)", open(HEADER).read()).group(1)


def test_block_opts_struct_size_is_enforced():
    """dyt_block_fwd validates opts.struct_size before anything touches the GPU."""
    from dyt_b200 import _lib
    lib = _lib.lib()
    shape = _lib.BlockShape(2, 197, 768, 12, 3072, 64)
    wt, opts = _lib.BlockWeights(), _lib.BlockOpts()
    opts.struct_size = ctypes.sizeof(_lib.BlockOpts) - 8          # a binding without the last field
    st = lib.dyt_block_fwd(ctypes.byref(shape), ctypes.byref(wt), ctypes.byref(opts), 256, 256, 256,
                           256, 1 << 40, None)
    assert st < 0 and b"struct_size" in lib.dyt_last_error()
    assert ctypes.sizeof(_lib.BlockOpts) == 136                   # 64-bit layout of the header struct


def test_argument_errors_are_reported_not_raised_across_the_abi():
    from dyt_b200 import _lib
    lib = _lib.lib()
    # null pointers / bad shapes must come back as negative status codes with a message; nothing
    # here touches a GPU (validation happens before any CUDA call)
    st = lib.dyt_linear_f16(None, 8, None, 8, 4, 8, 8, None, 0, None, None, 8, None, 0, None, 0, 1.0, None)
    assert st < 0 and b"null" in lib.dyt_last_error()
    st = lib.dyt_attn_varlen_fwd(1, 96, None, 1, 16, 16, 16, 1, 32, 1, 32, None)
    assert st < 0 and b"head_dim" in lib.dyt_last_error()
    shape = _lib.BlockShape(2, 197, 100, 2, 400, 64)      # C != 64 * H
    assert lib.dyt_block_workspace_bytes(ctypes.byref(shape)) == 0
    shape = _lib.BlockShape(256, 197, 768, 12, 3072, 64)
    need = lib.dyt_block_workspace_bytes(ctypes.byref(shape))
    assert 1.0e9 < need < 1.3e9                            # ~1.1 GB at the BASELINE configuration
    with pytest.raises(_lib.DytError):
        _lib.check(-1, "unit test")
