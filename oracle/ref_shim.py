"""Import shim that lets the UNMODIFIED reference modules under /root/reference be imported in a
container without timm / easydict.   *** TEST INFRASTRUCTURE ONLY ***

Used by oracle/gen_golden.py (golden fixtures) and, when /root/reference is present, by
tests/test_oracle_vs_reference.py.  Never imported by the product package.

timm==0.9.12 (requirements.txt:142) is not vendored by the reference; the six symbols its hot path
really uses are restated here from timm's published behaviour: Mlp, PatchEmbed,
DropPath, PatchDropout, trunc_normal_, use_fused_attn.  All other imported names are dead in the
reference files and only need to exist.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

def _reference_root() -> str:
    """The reference checkout when it is there (this container), else the travelling copy of the
    hot-path modules made by oracle/build_ref.py (GPU box: baseline timing only)."""
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in (os.environ.get("DYT_REFERENCE_ROOT"), "/root/reference", os.path.join(here, "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "models", "model_speed_test.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _reference_root()


class _Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU,
                 norm_layer=None, bias=True, drop=0.0, use_conv=False):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.norm = nn.Identity() if norm_layer is None else norm_layer(hidden_features)
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


class _PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None,
                 flatten=True, bias=True, **kw):
        super().__init__()
        t2 = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v)
        self.img_size, self.patch_size = t2(img_size), t2(patch_size)
        self.grid_size = tuple(s // p for s, p in zip(self.img_size, self.patch_size))
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size,
                              bias=bias)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        x = self.proj(x)
        if self.flatten:
            x = x.flatten(2).transpose(1, 2)
        return self.norm(x)


class _DropPath(nn.Module):
    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        assert self.drop_prob == 0.0 or not self.training, "shim DropPath: eval / p=0 only"
        return x


class _PatchDropout(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("PatchDropout is never enabled by the reference")


class EasyDict(dict):
    """attribute-access dict (easydict.EasyDict stand-in)"""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _use_fused_attn():
    return hasattr(F, "scaled_dot_product_attention") and int(os.environ.get("TIMM_FUSED_ATTN", "1")) > 0


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install() -> None:
    """Register the fake third-party modules (idempotent)."""
    if "timm" in sys.modules and getattr(sys.modules["timm"], "_dyt_shim", False):
        return
    dummy = lambda *a, **k: None
    ident_deco = lambda fn: fn
    trunc = lambda t, mean=0.0, std=1.0, a=-2.0, b=2.0: nn.init.trunc_normal_(t, mean, std, a, b)
    t2 = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    layers = dict(PatchEmbed=_PatchEmbed, Mlp=_Mlp, DropPath=_DropPath, PatchDropout=_PatchDropout,
                  trunc_normal_=trunc, use_fused_attn=_use_fused_attn, lecun_normal_=dummy,
                  _assert=dummy, to_2tuple=t2)
    timm = _module("timm", _dyt_shim=True)
    timm.layers = _module("timm.layers", **layers)
    _module("timm.layers.format", Format=object, nchw_to=dummy)
    _module("timm.data", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406),
            IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225), IMAGENET_INCEPTION_MEAN=(0.5, 0.5, 0.5),
            IMAGENET_INCEPTION_STD=(0.5, 0.5, 0.5))
    _module("timm.models")
    _module("timm.models.helpers", build_model_with_cfg=dummy, named_apply=dummy,
            adapt_input_conv=dummy, resolve_pretrained_cfg=dummy, checkpoint_seq=dummy)
    _module("timm.models.layers", **layers)
    _module("timm.models.registry", register_model=ident_deco)
    _module("easydict", EasyDict=EasyDict)
    # models/losses.py:1-3 imports three names it never uses
    timm.loss = _module("timm.loss")
    _module("timm.data.transforms_factory", transforms_imagenet_train=dummy)
    if "numpy.lib.arraysetops" not in sys.modules:
        try:
            importlib.import_module("numpy.lib.arraysetops")
        except ImportError:
            import numpy as _np
            _module("numpy.lib.arraysetops", isin=_np.isin)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "model_speed_test.py"))


def import_reference(module: str):
    """import e.g. 'models.model_speed_test' from the reference checkout, isolated from any other
    package called `models` (the product ships a drop-in package of the same name)."""
    install()
    def _ours(k):
        return k in ("models", "video_models") or k.startswith(("models.", "video_models."))
    saved = {k: v for k, v in sys.modules.items() if _ours(k)}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        mod = importlib.import_module(module)
    finally:
        sys.path.remove(REFERENCE_ROOT)
        ref_mods = {k: v for k, v in sys.modules.items() if _ours(k)}
        for k in ref_mods:
            del sys.modules[k]
        sys.modules.update(saved)
    return mod


def reference_configs(ffn_num: int = 64, scalar: str = "0.1", d_model: int = 768,
                      ratio: float = 0.5):
    """tuning_config / select_config exactly as main_image.py:186-210 builds them."""
    tuning = EasyDict(ffn_adapt=True, ffn_option="parallel", ffn_adapter_layernorm_option="none",
                      ffn_adapter_init_option="lora", ffn_adapter_scalar=scalar, ffn_num=ffn_num,
                      d_model=d_model, vpt_on=False, vpt_num=0)
    select = EasyDict(open=True, keep_layers=0, token_target_ratio=ratio)
    return tuning, select
