"""Drop-in for the reference's top-level `block_flops_dict` module (imported by main_image.py:38 and
engine_finetune.py:14): same function names, computed by dyt_b200.flops (see there for what is
pinned and what is not)."""
from dyt_b200.flops import batch_select_flops, block_flops_table  # noqa: F401
from dyt_b200 import flops as _flops


def get_block_flops(args=None):
    cfg = getattr(args, "tuning_config", None)
    bott = int(getattr(cfg, "ffn_num", 64)) if cfg is not None else 64
    return block_flops_table(bottleneck=bott)


def get_base_flops(args=None):
    return _flops.base_flops(num_classes=int(getattr(args, "nb_classes", 100)))


def select_flops(flops_dict, token_select, block_num, base_flops=0.33):
    """One image: token_select [L, N-1]."""
    return batch_select_flops(1, flops_dict, token_select.unsqueeze(0), block_num, base_flops)[0]
