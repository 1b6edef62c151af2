"""Import surface of reference models/model_speed_test.py (the sparse inference ViT used by
speed.py:36): Block.forward(x) -> x, VisionTransformer.forward(x) -> logits,
vit_base_patch16_224_in21k(**kwargs) (:524-532).  Compute: dyt_b200 sm_100a kernels."""
from dyt_b200.layers import DropPath, Mlp, PatchDropout, PatchEmbed, trunc_normal_, use_fused_attn  # noqa: F401
from dyt_b200.modules import Adapter, Attention, LayerScale, TokenSelect, _gumbel_sigmoid  # noqa: F401
from dyt_b200.modules import SpeedBlock as Block
from dyt_b200.modules import SpeedVisionTransformer as VisionTransformer


def convert_list_to_tensor(list_convert):
    import torch
    return torch.stack(list_convert, dim=1) if len(list_convert) else None


def vit_base_patch16_224_in21k(**kwargs):
    """ViT-B/16 (patch 16, dim 768, depth 12, 12 heads, qkv bias) with DyT blocks."""
    return VisionTransformer(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                             qkv_bias=True, **kwargs)


def vit_large_patch16_224_in21k(**kwargs):
    """ViT-L/16 re-parameterisation named by BASELINE.json config 4 (no reference constructor;
    SURVEY.md section 0.6): dim 1024, depth 24, 16 heads."""
    return VisionTransformer(patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4.0,
                             qkv_bias=True, **kwargs)
