from .segmentation_vision_transformer_IN21K import VisionTransformer21K  # noqa: F401
