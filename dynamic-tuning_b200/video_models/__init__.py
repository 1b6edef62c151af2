"""Import surface of the reference `video_models` package (main_video.py:29): the per-frame DyT ViT
with the attentive pooling head, computed by the dyt_b200 sm_100a kernels."""
