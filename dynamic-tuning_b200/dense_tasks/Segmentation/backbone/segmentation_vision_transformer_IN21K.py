"""Drop-in for the reference's dense_tasks/Segmentation/backbone/segmentation_vision_transformer_IN21K.py:
the same names, backed by dyt_b200.modules_seg (sm_100a kernels).  When mmseg is installed the backbone
is registered in its BACKBONES registry like the reference does (:302)."""
from dyt_b200.modules_seg import (SegAttention as Attention, SegBlock as Block,  # noqa: F401
                                  TokenRateLoss as AdaLoss, VisionTransformer21K, resize_pos_embed,
                                  vit_base_patch16_224_in21k)

try:  # optional: mmseg is not part of this environment
    from mmseg.models.builder import BACKBONES
    BACKBONES.register_module()(VisionTransformer21K)
except Exception:  # noqa: BLE001
    pass
