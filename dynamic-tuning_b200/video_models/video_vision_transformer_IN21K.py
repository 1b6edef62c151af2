"""Import surface of reference video_models/video_vision_transformer_IN21K.py (used by
main_video.py:29, :244): VisionTransformer.forward(x [b, 3, t, H, W], complete_model) ->
(logits, dict(token_select [b*t, depth, N-1, 1], token_logits)), AttentiveBlock, CrossAttention,
vit_base_patch16_224_in21k(**kwargs) (:512-519).  Compute: dyt_b200 kernels (per-frame DyT blocks
+ fused double-LayerNorm, tcgen05 k/v GEMMs and the single-query attention kernel of the pooling
head)."""
from dyt_b200.layers import DropPath, Mlp, PatchDropout, PatchEmbed, trunc_normal_, use_fused_attn  # noqa: F401
from dyt_b200.modules import Adapter, Attention, AttentiveBlock, CrossAttention, LayerScale, TokenSelect  # noqa: F401
from dyt_b200.modules import TrainBlock as Block  # noqa: F401
from dyt_b200.modules import VideoVisionTransformer as VisionTransformer


def convert_list_to_tensor(list_convert):
    import torch
    return torch.stack(list_convert, dim=1) if len(list_convert) else None


def vit_base_patch16_224_in21k(**kwargs):
    """ViT-B/16 video model (patch 16, dim 768, depth 12, 12 heads, qkv bias) with DyT blocks."""
    return VisionTransformer(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                             qkv_bias=True, **kwargs)
