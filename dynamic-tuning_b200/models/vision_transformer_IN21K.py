"""Import surface of reference models/vision_transformer_IN21K.py (training / evaluation ViT used by
main_image.py:33, main_vtab.py:31, block_flops_dict.py:6): Block.forward(x, complete_model) ->
(x, dict), VisionTransformer.forward(x, complete_model) -> (logits, dict(token_select,
token_logits)), vit_base_patch16_224_in21k(**kwargs) (:414-421).  Compute: dyt_b200 kernels."""
from dyt_b200.layers import DropPath, Mlp, PatchDropout, PatchEmbed, trunc_normal_, use_fused_attn  # noqa: F401
from dyt_b200.modules import Adapter, Attention, LayerScale, TokenSelect  # noqa: F401
from dyt_b200.modules import TrainBlock as Block
from dyt_b200.modules import TrainVisionTransformer as VisionTransformer


def convert_list_to_tensor(list_convert):
    import torch
    return torch.stack(list_convert, dim=1) if len(list_convert) else None


def vit_base_patch16_224_in21k(**kwargs):
    """ViT-B/16 (patch 16, dim 768, depth 12, 12 heads, qkv bias) with DyT blocks."""
    return VisionTransformer(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                             qkv_bias=True, **kwargs)
