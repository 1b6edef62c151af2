"""Import surface of reference models/dynamic_adapter.py: TokenSelect (:58-77), Adapter (:80-140),
_gumbel_sigmoid (:25-54) backed by the sm_100a kernels (see dyt_b200/modules.py)."""
from dyt_b200.modules import Adapter, TokenSelect, _gumbel_sigmoid  # noqa: F401
from dyt_b200.layers import DropPath, Mlp, PatchDropout, PatchEmbed, trunc_normal_, use_fused_attn  # noqa: F401
