"""CPU-only checks of the host-side mirror of the reference interface: class names, constructor
keywords, state_dict keys (the on-disk contract), the gate threshold table,
and that every compute path refuses to run without CUDA (no silent fallback)."""
import pytest
import torch

from conftest import load_golden


class Cfg(dict):
    __getattr__ = dict.__getitem__


def _cfgs(ffn_num=64, d_model=768):
    tuning = Cfg(ffn_adapt=True, ffn_option="parallel", ffn_adapter_layernorm_option="none",
                 ffn_adapter_init_option="lora", ffn_adapter_scalar="0.1", ffn_num=ffn_num,
                 d_model=d_model, vpt_on=False, vpt_num=0)
    return tuning, Cfg(open=True, keep_layers=0, token_target_ratio=0.5)


def test_state_dict_keys_match_reference_tiny():
    t = load_golden("tiny_vit.pt")          # state_dict accepted strict=True by the real reference
    from models.model_speed_test import VisionTransformer as SpeedViT
    from models.vision_transformer_IN21K import VisionTransformer as TrainViT
    tuning, select = _cfgs(16, 128)
    for cls in (SpeedViT, TrainViT):
        m = cls(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2, num_classes=10,
                tuning_config=tuning, select_config=select)
        msg = m.load_state_dict(t["state_dict"], strict=True)
        assert not msg.missing_keys and not msg.unexpected_keys
        assert list(m.state_dict().keys()) == [k for k in m.state_dict().keys()]
        assert set(m.state_dict().keys()) == set(t["state_dict"].keys())


def test_vit_b16_factory_surface():
    from models.model_speed_test import vit_base_patch16_224_in21k, Block, Attention, Adapter, TokenSelect, Mlp
    from models.vision_transformer_IN21K import Block as TBlock, Mlp as TMlp  # block_flops_dict.py:6,:19
    from models.dynamic_adapter import Adapter as A2, TokenSelect as T2       # video / segmentation imports
    assert A2 is Adapter and T2 is TokenSelect and TMlp is Mlp
    tuning, select = _cfgs()
    m = vit_base_patch16_224_in21k(num_classes=100, drop_path_rate=0.0, tuning_config=tuning,
                                   select_config=select)
    assert sum(p.numel() for p in m.parameters()) == 87_074_416        # == reference ctor
    blk = m.blocks[0]
    dyt = sum(p.numel() for n, p in blk.named_parameters() if "adaptmlp" in n or "token_select" in n)
    assert dyt == 99_905
    assert isinstance(blk, Block) and hasattr(m, "head") and len(m.blocks) == 12
    # strict=False loading of a backbone checkpoint leaves exactly the DyT parameters + head missing
    backbone = {k: v for k, v in m.state_dict().items()
                if "adaptmlp" not in k and "mlp_token_select" not in k and not k.startswith("head.")}
    msg = m.load_state_dict(backbone, strict=False)
    assert len(msg.missing_keys) == 12 * 6 + 2 and not msg.unexpected_keys
    # train flavour exposes the FLOP-probe attributes poked by block_flops_dict.py:44-47
    tb = TBlock(dim=768, num_heads=12, mlp_ratio=4.0, qkv_bias=True, tuning_config=tuning)
    assert tb.count_flops is None and tb.token_select_num is None and hasattr(tb, "forward_count_flops")


def test_select_false_quirk_is_preserved():
    from models.model_speed_test import Block
    tuning, _ = _cfgs()
    b = Block(dim=768, num_heads=12, tuning_config=tuning, select=False)
    assert b.token_select is None and not hasattr(b, "mlp_token_select")
    with pytest.raises(AttributeError):
        b(torch.zeros(2, 197, 768))


def test_no_cpu_fallback_anywhere():
    from dyt_b200 import DytError, ops
    from models.model_speed_test import vit_base_patch16_224_in21k, Adapter, TokenSelect, Attention
    tuning, select = _cfgs()
    with torch.no_grad():
        with pytest.raises(DytError):
            TokenSelect(768, 1).eval()(torch.zeros(1, 197, 768))
        with pytest.raises(DytError):
            Adapter(tuning, dropout=0.1, bottleneck=64, adapter_scalar="0.1",
                    adapter_layernorm_option="none").eval()(torch.zeros(1, 197, 768))
        with pytest.raises(DytError):
            Attention(768, 12, qkv_bias=True).eval()(torch.zeros(1, 197, 768))
        with pytest.raises(DytError):
            ops.layernorm_f16(torch.zeros(4, 768), torch.ones(768), torch.zeros(768))
        m = vit_base_patch16_224_in21k(num_classes=10, drop_path_rate=0.0, tuning_config=tuning,
                                       select_config=select).eval()
        with pytest.raises(DytError):
            m(torch.zeros(1, 3, 224, 224))


def test_training_mode_has_no_cpu_fallback_and_refuses_unfrozen_backbone():
    """Fine-tuning runs on the CUDA kernels only (dyt_b200.train): CPU tensors raise; the speed
    flavour is inference-only; a trainable stem is refused (the reference freezes it)."""
    from dyt_b200 import DytError
    from models.vision_transformer_IN21K import VisionTransformer
    from models.model_speed_test import VisionTransformer as SpeedViT
    tuning, select = _cfgs(16, 128)
    kw = dict(img_size=32, patch_size=16, embed_dim=128, depth=1, num_heads=2, num_classes=10,
              tuning_config=tuning, select_config=select)
    m = VisionTransformer(**kw).train()
    with pytest.raises(NotImplementedError):     # stem parameters still require grad
        m(torch.zeros(2, 3, 32, 32))
    for n, p in m.named_parameters():
        p.requires_grad = ("adaptmlp" in n) or ("mlp_token_select" in n) or n.startswith("head.")
    with pytest.raises(DytError):                # frozen like main_image.py:242-256, but on the CPU
        m(torch.zeros(2, 3, 32, 32))
    with pytest.raises(DytError):
        m.blocks[0](torch.zeros(2, 5, 128))
    s = SpeedViT(**kw).train()
    for p in s.parameters():
        p.requires_grad = False
    s.blocks[0].adaptmlp.up_proj.weight.requires_grad = True
    with pytest.raises((NotImplementedError, DytError)):
        s.blocks[0](torch.zeros(2, 5, 128))


def test_gate_threshold_table_product_side():
    """The product derives min_kept from torch's sigmoid; it must reproduce the reference gate's
    exhaustive truth tables (golden, from the reference's own _gumbel_sigmoid)."""
    from dyt_b200 import min_kept_logit
    gold = load_golden("gate_tables.pt")
    for name, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
        vals = torch.arange(0, 1 << 16, dtype=torch.int32).to(torch.int16).view(dt).float()
        mk = min_kept_logit(dt, 0.5)
        assert torch.equal((vals >= mk) & ~torch.isnan(vals), gold[name].bool())
    assert min_kept_logit(torch.float16, 0.5) == 2.0 ** -10 * (1 + 2.0 ** -10)
    assert min_kept_logit(torch.float16, 0.7) > 0.8            # sigmoid(l) > 0.7  <=>  l > 0.847
    assert 0 < min_kept_logit(torch.float32, 0.5) < 2e-7


def test_video_model_state_dict_keys_match_reference():
    """video_models.video_vision_transformer_IN21K (main_video.py:29): same parameter names as the
    reference video model (golden key list generated from the reference ctor), strict load works."""
    from conftest import load_golden
    from video_models.video_vision_transformer_IN21K import (AttentiveBlock, CrossAttention,  # noqa: F401
                                                             VisionTransformer, vit_base_patch16_224_in21k)
    g = load_golden("video_tiny.pt")
    d = g["dims"]
    tuning, select = _cfgs(ffn_num=d["bottleneck"], d_model=d["embed_dim"])
    m = VisionTransformer(img_size=d["img_size"], patch_size=16, embed_dim=d["embed_dim"],
                          depth=d["depth"], num_heads=d["num_heads"], mlp_ratio=4.0, qkv_bias=True,
                          num_classes=d["num_classes"], tuning_config=tuning, select_config=select)
    assert sorted(m.state_dict().keys()) == g["keys"]
    res = m.load_state_dict(g["state_dict"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    import pytest, torch
    with pytest.raises(Exception):          # CPU tensors: no fallback
        m.eval()(torch.zeros(1, 3, 2, 32, 32))


def test_finetune_loss_matches_reference_golden():
    """dyt_b200.finetune.finetune_loss (host-side glue: AdaLoss + teacher CE + KL, engine_finetune.py:
    52-65, models/losses.py:50-82) on the reference's own outputs reproduces the reference's loss."""
    from dyt_b200.finetune import ada_loss, finetune_loss
    g = load_golden("finetune_tiny.pt")
    # the fixture was made with the AdaLoss class defaults (token_minimal 0.1, weight 1.0)
    cls = dict(token_minimal=0.1, token_minimal_weight=1.0)
    loss = finetune_loss(g["student_logits"], g["token_select"], g["teacher_logits"], g["targets"], **cls)
    assert abs(loss.item() - g["loss"].item()) <= 1e-5
    # AdaLoss alone: cross-entropy + 2 * ((mean keep - 0.5)^2 + sum(clamp(0.1 - sel, 0)))
    sel = g["token_select"]
    want = torch.nn.functional.cross_entropy(g["student_logits"], g["targets"]) + 2.0 * (
        (sel.mean() - 0.5) ** 2 + (0.1 - sel.mean(-1)).clamp(min=0).sum())
    assert abs(ada_loss(g["student_logits"], sel, g["targets"], **cls).item() - want.item()) <= 1e-6
    # the recipe of the entry scripts (main_image.py:206-209): no minimal-keep term
    want0 = torch.nn.functional.cross_entropy(g["student_logits"], g["targets"]) + 2.0 * (sel.mean() - 0.5) ** 2
    assert abs(ada_loss(g["student_logits"], sel, g["targets"]).item() - want0.item()) <= 1e-6
    # drop-in models.losses.AdaLoss: the reference's class (same ctor keywords, same return)
    from models.losses import AdaLoss
    crit = AdaLoss(torch.nn.CrossEntropyLoss(), token_target_ratio=0.5, token_loss_ratio=2.0,
                   token_minimal=0.1, token_minimal_weight=1.0)
    l, parts = crit(dict(prediction=g["student_logits"], token_select=sel, token_logits=None), g["targets"])
    assert abs(l.item() - want.item()) <= 1e-6 and set(parts) == {"base_loss", "token_loss"}
    l0, _ = AdaLoss(torch.nn.CrossEntropyLoss(), token_minimal=0.0, token_minimal_weight=0.0)(
        dict(prediction=g["student_logits"], token_select=sel, token_logits=None), g["targets"])
    assert abs(l0.item() - want0.item()) <= 1e-6


def test_graphed_forward_and_accounting_refuse_cpu():
    from dyt_b200 import DytError, GraphedForward, flops
    from models.model_speed_test import VisionTransformer
    tuning, select = _cfgs(16, 128)
    m = VisionTransformer(img_size=32, patch_size=16, embed_dim=128, depth=1, num_heads=2,
                          num_classes=10, tuning_config=tuning, select_config=select).eval()
    with pytest.raises(DytError):
        GraphedForward(m)(torch.zeros(1, 3, 32, 32))
    with pytest.raises(DytError):
        flops.batch_select_flops(1, torch.zeros(198), torch.zeros(1, 12, 196, 1))
    t = flops.block_flops_table()
    assert t.shape == (198,) and t[0] == 0 and bool((t[1:].diff() > 0).all())
    assert abs(12 * t[197].item() + flops.base_flops() - 17.8) < 0.1      # GMACs of dense ViT-B/16 + DyT


def test_segmentation_backbone_state_dict_keys_match_reference():
    """dense_tasks.Segmentation.backbone drop-in: same parameter / buffer names as the reference
    VisionTransformer21K (golden key list), strict load, relative-position index identical, pos-embed
    resize hook for 224-pretrained checkpoints; no CPU fallback."""
    from dense_tasks.Segmentation.backbone.segmentation_vision_transformer_IN21K import VisionTransformer21K
    g = load_golden("seg_tiny.pt")
    tuning, select = _cfgs(16, 128)
    kw = dict(patch_size=16, embed_dim=128, depth=4, num_heads=2, num_classes=0, tuning_config=tuning,
              select_config=select, out_indices=[0, 1, 2, 3], use_rel_pos_bias=True)
    m = VisionTransformer21K(img_size=64, **kw)
    assert sorted(m.state_dict().keys()) == g["keys"]
    idx_before = m.blocks[0].attn.relative_position_index.clone()
    res = m.load_state_dict(g["state_dict"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(idx_before, g["state_dict"]["blocks.0.attn.relative_position_index"])
    big = VisionTransformer21K(img_size=96, **{**kw, "use_rel_pos_bias": False})
    sd = {k: v for k, v in g["state_dict"].items() if "relative_position" not in k}
    big.load_state_dict(sd, strict=True)          # 4x4 -> 6x6 position embedding, resized by the hook
    assert big.pos_embed.shape == (1, 37, 128)
    with pytest.raises(Exception):
        m.eval()(torch.zeros(1, 3, 64, 64))
