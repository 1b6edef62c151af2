"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules (imported from
/root/reference through oracle/ref_shim.py).   *** TEST INFRASTRUCTURE ONLY ***

Run here (the container that has /root/reference):   python oracle/gen_golden.py
The fixtures travel with the repo; nothing at test/bench time reads /root/reference.

The reference ships no tests or golden vectors, so these outputs of the reference
itself are the pins of the oracle (oracle/dyt_oracle.py):
  gate_tables.pt    exhaustive fp16 / bf16 truth tables of the eval gate (both reference copies)
  gumbel_train.pt   train-mode hard Gumbel gate given the RNG draws (fp32 and fp16 logits)
  tiny_vit.pt       a 2-layer dim-128 ViT (reference generic ctor): full state_dict, inputs, every
                    model-level output of both reference models + per-block activations
  video_tiny.pt     a 2-layer dim-128 video model (reference video ctor): state_dict incl. the
                    attentive pooling head, a 3x2-frame clip, logits / masks / pooled features
  finetune_tiny.pt  one fine-tuning step (student + teacher pass, AdaLoss + teacher CE + KL, backward)
                    of the reference train model in train() mode on the tiny ViT: inputs, the
                    replayed Gumbel draws, dropout multipliers, loss, outputs and every trainable
                    parameter's gradient
  flops_accounting.pt  per-image GFLOPs of block_flops_dict.batch_select_flops and per-layer keep rates
                    for random masks / a random table
  seg_attention.pt  the segmentation backbone's Attention module (relative-position bias, eager path):
                    inputs, qkv, outputs, bias table / index for two window sizes
  seg_tiny.pt       a 4-layer dim-128 segmentation backbone with relative-position bias: state_dict,
                    image, FPN feature maps, masks, logits, token loss
  count_flops_tiny.pt  Block.forward_count_flops (FLOP probe) of the reference train Block on the tiny
                    ViT for token_select_num = 1, 3, 5
  vitb_b2.pt        ViT-B/16, synthetic seed-0 weights (regenerated from the seed, not stored),
                    calibrated selector biases, B=2: logits / masks / token logits of the speed model
                    and the train model (eval, complete_model on/off), config-1 imposed-mask logits
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import dyt_oracle as O  # noqa: E402
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def gate_tables(ref_speed, ref_dyn):
    out = {}
    for name, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
        bits = torch.arange(0, 1 << 16, dtype=torch.int32).to(torch.int16)
        vals = bits.view(dt)
        a = ref_speed._gumbel_sigmoid(vals, 5, True, threshold=0.5, training=False)
        b = ref_dyn._gumbel_sigmoid(vals, 5, True, threshold=0.5, training=False)
        # the train-capable copy returns y_hard - y_soft.detach() + y_soft: equal to y_hard up to
        # one rounding, so compare it through > 0.5
        assert torch.equal(a > 0.5, b.float() > 0.5), "reference copies of the gate disagree"
        out[name] = (a > 0.5).to(torch.uint8)
    return out


def gumbel_train(ref_dyn):
    out = {}
    for name, dt in (("fp32", torch.float32), ("fp16", torch.float16)):
        g = torch.Generator().manual_seed(7)
        logits = (torch.randn(4, 196, 1, generator=g) * 2).to(dt)
        torch.manual_seed(1234)
        ret = ref_dyn._gumbel_sigmoid(logits, 5, True, threshold=0.5, training=True)
        torch.manual_seed(1234)   # replay the two draws (models/dynamic_adapter.py:30-39)
        g1 = -torch.empty_like(logits).exponential_().log()
        g2 = -torch.empty_like(logits).exponential_().log()
        out[name] = dict(logits=logits, g1=g1, g2=g2, hard=(ret.float() > 0.5).to(torch.uint8))
    return out


def _build(ref_mod, sd, num_classes, embed_dim, depth, heads, img, ffn_num, scalar, generic):
    tuning, select = ref_shim.reference_configs(ffn_num=ffn_num, scalar=scalar, d_model=embed_dim)
    if generic:
        m = ref_mod.VisionTransformer(img_size=img, patch_size=16, embed_dim=embed_dim, depth=depth,
                                      num_heads=heads, mlp_ratio=4.0, qkv_bias=True,
                                      num_classes=num_classes, tuning_config=tuning,
                                      select_config=select)
    else:
        m = ref_mod.vit_base_patch16_224_in21k(num_classes=num_classes, drop_path_rate=0.0,
                                               tuning_config=tuning, select_config=select)
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.eval()


def tiny_vit(ref_speed, ref_train):
    dims = dict(embed_dim=128, depth=2, num_heads=2, bottleneck=16, num_classes=10, img_size=32)
    sd = O.synthetic_state_dict(seed=3, **dims)
    g = torch.Generator().manual_seed(11)
    img = torch.randn(3, 3, 32, 32, generator=g)
    sd = O.calibrate_selector_bias(sd, img, 2, 2, 0.1, 0.5)
    ms = _build(ref_speed, sd, 10, 128, 2, 2, 32, 16, "0.1", True)
    mt = _build(ref_train, sd, 10, 128, 2, 2, 32, 16, "0.1", True)
    out = dict(dims=dims, scale=0.1, state_dict=sd, img=img)
    with torch.no_grad():
        out["speed_logits"] = ms(img)
        lg, d = mt(img)
        out["train_logits"], out["token_select"], out["token_logits"] = lg, d["token_select"], d["token_logits"]
        lg, d = mt(img, complete_model=True)
        out["teacher_logits"], out["teacher_token_select"] = lg, d["token_select"]
        # per-block activations of the speed model
        x = ms.patch_embed(img)
        x = torch.cat((ms.cls_token.expand(x.shape[0], -1, -1), x), dim=1) + ms.pos_embed
        out["x0"] = x.clone()
        blk = ms.blocks[0]
        out["blk0_attn"] = blk.attn(blk.norm1(x))
        x1 = x + out["blk0_attn"]
        sel, logit = blk.mlp_token_select(x1)
        out["blk0_sel"], out["blk0_logit"] = sel, logit
        out["blk0_adapter"] = blk.adaptmlp(x1, add_residual=False)
        out["blk0_mlp_dense"] = blk.mlp(blk.norm2(x1))
        out["blk0_out"] = blk(x)
        out["blk0_out_b1"] = blk(x[:1])          # single_forward path (B == 1)
        out["blk0_train_out"] = mt.blocks[0](x)[0]
    return out


def count_flops_tiny(ref_train):
    """Block.forward_count_flops of the reference train Block (the FLOP probe of
    block_flops_dict.get_block_flops, vision_transformer_IN21K.py:167-185) on the tiny ViT's block 0
    for several token_select_num."""
    dims = dict(embed_dim=128, depth=2, num_heads=2, bottleneck=16, num_classes=10, img_size=32)
    sd = O.synthetic_state_dict(seed=3, **dims)
    mt = _build(ref_train, sd, 10, 128, 2, 2, 32, 16, "0.1", True)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 5, 128, generator=g)
    blk = mt.blocks[0]
    out = dict(dims=dims, scale=0.1, seed=3, x=x, outs={})   # weights: O.synthetic_state_dict(seed, **dims)
    blk.count_flops = True
    with torch.no_grad():
        for t in (1, 3, 5):
            blk.token_select_num = t
            out["outs"][t] = blk(x).clone()
    blk.count_flops = None
    return out


def vitb_b2(ref_speed, ref_train):
    sd = O.synthetic_state_dict(seed=0)
    g = torch.Generator().manual_seed(0)
    img = torch.randn(2, 3, 224, 224, generator=g)
    sd = O.calibrate_selector_bias(sd, img, 12, 12, 0.1, 0.5)
    biases = torch.stack([sd[f"blocks.{i}.mlp_token_select.mlp_head.bias"] for i in range(12)])
    ms = _build(ref_speed, sd, 100, 768, 12, 12, 224, 64, "0.1", False)
    mt = _build(ref_train, sd, 100, 768, 12, 12, 224, 64, "0.1", False)
    out = dict(seed=0, img_seed=0, scale=0.1, selector_bias=biases)
    with torch.no_grad():
        out["speed_logits"] = ms(img)
        lg, d = mt(img)
        out["train_logits"] = lg
        out["token_select"] = d["token_select"].to(torch.uint8)
        out["token_logits"] = d["token_logits"]
        out["teacher_logits"] = mt(img, complete_model=True)[0]
        # BASELINE.json config 1: imposed 50% mask (cls + even-indexed patches)
        forced = O.checkerboard_mask(2, 197)
        def imposed(self, x):
            logits = self.mlp_head(x[:, 1:, :])
            return forced.to(x.dtype), logits
        for mod, key in ((ref_speed, "config1_speed_logits"), (ref_train, "config1_train_logits")):
            orig = mod.TokenSelect.forward
            mod.TokenSelect.forward = imposed
            try:
                r = (ms if mod is ref_speed else mt)(img)
                out[key] = r if mod is ref_speed else r[0]
            finally:
                mod.TokenSelect.forward = orig
    return out


def video_state_dict(sd, embed_dim, seed):
    """Adds the pooling-head parameters of the video model (query_token, attentive_blocks.*) to an
    image-model state_dict, drawn like synthetic_state_dict draws the others."""
    c = embed_dim
    g = torch.Generator().manual_seed(1000 + seed)
    rn = lambda *shape, std=0.02: torch.randn(*shape, generator=g) * std
    pre = "attentive_blocks."
    sd = dict(sd)
    sd["query_token"] = rn(1, 1, c, std=0.5)
    for n in ("norm_q", "norm_k", "norm_v"):
        sd[pre + n + ".weight"] = 1.0 + rn(c, std=0.1)
        sd[pre + n + ".bias"] = rn(c)
    for n in ("q", "k", "v"):
        sd[pre + f"cross_attn.{n}.weight"] = rn(c, c, std=0.05)
    sd[pre + "cross_attn.q_bias"] = rn(c)
    sd[pre + "cross_attn.v_bias"] = rn(c)
    sd[pre + "cross_attn.proj.weight"] = rn(c, c, std=0.05)
    sd[pre + "cross_attn.proj.bias"] = rn(c)
    return sd


def tiny_video(ref_video):
    """2-layer dim-128 video model (reference video_models/video_vision_transformer_IN21K.py generic
    ctor), 3 clips of 2 frames of 32x32: full state_dict, clip, logits, masks, pooled features."""
    dims = dict(embed_dim=128, depth=2, num_heads=2, bottleneck=16, num_classes=10, img_size=32)
    sd = O.synthetic_state_dict(seed=5, **dims)
    g = torch.Generator().manual_seed(21)
    clip = torch.randn(3, 3, 2, 32, 32, generator=g)
    frames = clip.permute(0, 2, 1, 3, 4).reshape(6, 3, 32, 32)
    sd = O.calibrate_selector_bias(sd, frames, 2, 2, 0.1, 0.5)
    sd = video_state_dict(sd, 128, 5)
    tuning, select = ref_shim.reference_configs(ffn_num=16, scalar="0.1", d_model=128)
    m = ref_video.VisionTransformer(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2,
                                    mlp_ratio=4.0, qkv_bias=True, num_classes=10,
                                    tuning_config=tuning, select_config=select)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    m = m.eval()
    out = dict(dims=dims, scale=0.1, state_dict=sd, clip=clip, keys=sorted(m.state_dict().keys()))
    with torch.no_grad():
        lg, d = m(clip)
        out["logits"], out["token_select"], out["token_logits"] = lg, d["token_select"], d["token_logits"]
        x, _ = m.forward_features(clip)
        tokens = x.reshape(3, -1, 128)
        out["tokens"] = tokens
        out["pooled"] = m.attentive_blocks(m.query_token.expand(3, -1, -1), tokens)[:, 0, :]
    return out


def tiny_finetune(ref_train, ref_losses):
    """One fine-tuning step of the reference train model (engine_finetune.py:47-76: student pass,
    teacher pass, AdaLoss + teacher CE + KL, backward) on the 2-layer dim-128 model in train() mode.
    The reference draws its randomness from the global RNG; to share it with the kernels the two
    Gumbel draws per layer and pass are replayed from the seed (models/dynamic_adapter.py:30-39) and
    nn.functional.dropout is swapped for a multiplication by recorded keep/(1-p) multipliers while
    the reference runs (p = 0 calls pass through)."""
    dims = dict(embed_dim=128, depth=2, num_heads=2, bottleneck=16, num_classes=10, img_size=32)
    sd = O.synthetic_state_dict(seed=9, **dims)
    g = torch.Generator().manual_seed(31)
    img = torch.randn(4, 3, 32, 32, generator=g)
    targets = torch.tensor([1, 7, 3, 3])
    sd = O.calibrate_selector_bias(sd, img, 2, 2, 0.1, 0.5)
    tuning, select = ref_shim.reference_configs(ffn_num=16, scalar="0.1", d_model=128)
    m = ref_train.VisionTransformer(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2,
                                    mlp_ratio=4.0, qkv_bias=True, num_classes=10,
                                    tuning_config=tuning, select_config=select)
    m.load_state_dict(sd, strict=True)
    trainable = []
    for name, prm in m.named_parameters():      # main_image.py:242-256 freeze logic
        prm.requires_grad = ("adaptmlp" in name) or ("mlp_token_select" in name) or name.startswith("head.")
        if prm.requires_grad:
            trainable.append(name)
    m.train()
    B, N = 4, 5
    drop_mults = [(torch.rand(B, N, 16, generator=g) >= 0.1).float() / 0.9 for _ in range(4)]
    queue = list(drop_mults)
    import torch.nn.functional as F
    orig_dropout = F.dropout

    def fixed_dropout(x, p=0.5, training=True, inplace=False):
        if p == 0.0 or not training:
            return x
        return x * queue.pop(0)

    F.dropout = fixed_dropout
    torch.nn.functional.dropout = fixed_dropout
    try:
        torch.manual_seed(777)
        out_s, ts = m(img)
        out_t, _ = m(img, complete_model=True)
        crit = ref_losses.AdaLoss(torch.nn.CrossEntropyLoss(), token_target_ratio=0.5,
                                  token_loss_ratio=2.0, token_minimal=0.1, token_minimal_weight=1.0)
        kl = F.kl_div(F.log_softmax(out_s, dim=-1), F.log_softmax(out_t.detach(), dim=-1),
                      reduction="batchmean", log_target=True)
        teacher_loss = crit.base_criterion(out_t, targets)
        loss, _ = crit(dict(prediction=out_s, **ts), targets)
        loss = loss + teacher_loss + kl
        loss.backward()
    finally:
        F.dropout = orig_dropout
        torch.nn.functional.dropout = orig_dropout
    assert not queue, "dropout call count differs from the recorded multipliers"
    torch.manual_seed(777)     # replay: student layers 0..1, then teacher layers 0..1
    noises = []
    for _ in range(4):
        e = torch.empty(B, N - 1, 1)
        g1 = -e.clone().exponential_().log()
        g2 = -e.clone().exponential_().log()
        noises.append((g1, g2))
    grads = {n: prm.grad.clone() for n, prm in m.named_parameters() if prm.requires_grad}
    margin = min(float(((lg + n1 - n2) / 5).abs().min()) for lg, (n1, n2) in
                 zip(ts["token_logits"].unbind(1), noises[:2]))
    print("tiny_finetune: loss", float(loss), "keep", float(ts["token_select"].mean()),
          "gate margin", margin)
    assert margin > 5e-3, "a gate decision sits too close to the threshold for a cross-precision pin"
    return dict(dims=dims, scale=0.1, state_dict=sd, img=img, targets=targets, noises=noises,
                drop_mults=drop_mults, trainable=trainable, loss=loss.detach(),
                student_logits=out_s.detach(), teacher_logits=out_t.detach(),
                token_select=ts["token_select"].detach(), token_logits=ts["token_logits"].detach(),
                grads=grads)


def flops_accounting():
    """block_flops_dict.batch_select_flops of the reference (its module imports fvcore / timm.models
    at the top, stubbed here: neither is used by the two accounting functions) on random masks and a
    random table; plus the per-layer rates engine_finetune.py:349-351 logs."""
    import types
    ref_shim.install()
    for name, attrs in (("fvcore", {}), ("fvcore.nn", {"FlopCountAnalysis": object}),
                        ("timm.models", {"create_model": None})):
        m = sys.modules.get(name) or types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    sys.modules["fvcore"].nn = sys.modules["fvcore.nn"]
    ref = ref_shim.import_reference("block_flops_dict")
    g = torch.Generator().manual_seed(5)
    table = torch.rand(198, generator=g) * 1.5
    ts = (torch.rand(37, 12, 196, 1, generator=g) < torch.rand(1, 12, 1, 1, generator=g)).float()
    ts[3] = 0.0
    ts[4] = 1.0
    out = dict(table=table, token_select=ts.to(torch.uint8))
    out["flops_12"] = ref.batch_select_flops(37, table, ts, block_num=12, base_flops=0.116)
    out["flops_10of12"] = ref.batch_select_flops(37, table, ts[:, 2:], block_num=12, base_flops=0.25)
    out["layer_rates"] = torch.stack([ts[:, l].mean() for l in range(12)])
    return out


def _load_seg_module():
    """Import the reference segmentation backbone file.  It imports mmcv_custom / mmseg at the top
    (checkpoint loading, registry): stubbed, unused by the modules exercised here."""
    import importlib.util
    import types
    ref_shim.install()

    class _Reg:
        def register_module(self, *a, **k):
            return lambda cls: cls

    for name, attrs in (("mmcv_custom", {"load_checkpoint": None}), ("mmseg", {}),
                        ("mmseg.utils", {"get_root_logger": None}), ("mmseg.models", {}),
                        ("mmseg.models.builder", {"BACKBONES": _Reg()})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    ref_shim.import_reference("models.dynamic_adapter")   # warms the shim; the backbone imports it
    path = os.path.join(ref_shim.REFERENCE_ROOT, "dense_tasks", "Segmentation", "backbone",
                        "segmentation_vision_transformer_IN21K.py")
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
    for k in saved:
        del sys.modules[k]
    try:
        spec = importlib.util.spec_from_file_location("_ref_seg_backbone", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(ref_shim.REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved)
        for name in ("mmcv_custom", "mmseg", "mmseg.utils", "mmseg.models", "mmseg.models.builder"):
            sys.modules.pop(name, None)
    return mod


def seg_attention():
    """Attention module of the segmentation backbone (reference dense_tasks/Segmentation/backbone/
    segmentation_vision_transformer_IN21K.py:120-203) with its relative-position bias: eager path
    (the module always takes the bias branch)."""
    mod = _load_seg_module()
    out = {}
    for tag, (win, bsz) in (("w4", ((4, 4), 3)), ("w9x7", ((9, 7), 2))):
        torch.manual_seed(17)
        attn = mod.Attention(128, num_heads=2, qkv_bias=True, window_size=win).eval()
        g = torch.Generator().manual_seed(23)
        with torch.no_grad():
            attn.relative_position_bias_table.copy_(
                torch.randn(attn.relative_position_bias_table.shape, generator=g) * 0.5)
            n = win[0] * win[1] + 1
            x = torch.randn(bsz, n, 128, generator=g)
            y = attn(x)
            qkv = attn.qkv(x)
        out[tag] = dict(x=x, y=y, qkv=qkv, table=attn.relative_position_bias_table.detach().clone(),
                        index=attn.relative_position_index.clone(), window=win,
                        qkv_w=attn.qkv.weight.detach().clone(), qkv_b=attn.qkv.bias.detach().clone(),
                        proj_w=attn.proj.weight.detach().clone(), proj_b=attn.proj.bias.detach().clone())
    return out


def seg_tiny():
    """A 4-layer dim-128 segmentation backbone (reference VisionTransformer21K, use_rel_pos_bias=True,
    64 x 64 input = 17 tokens, a feature map after every block): state_dict, image, the four FPN
    feature maps, masks, logits and the token loss, eval mode, fp32."""
    mod = _load_seg_module()
    tuning, select = ref_shim.reference_configs(ffn_num=16, scalar="0.1", d_model=128)
    select.update(layer_target_ratio=0.5, layer_loss_ratio=2.0, layer_diverse_ratio=0.0,
                  layer_entropy_weight=0.0, layer_minimal_weight=0.0, layer_minimal=0.0,
                  token_ratio=2.0, token_minimal=0.1, token_minimal_weight=1.0)
    torch.manual_seed(29)
    m = mod.VisionTransformer21K(img_size=64, patch_size=16, embed_dim=128, depth=4, num_heads=2,
                                 num_classes=0, tuning_config=tuning, select_config=select,
                                 out_indices=[0, 1, 2, 3], use_rel_pos_bias=True).eval()
    g = torch.Generator().manual_seed(37)
    with torch.no_grad():
        for name, prm in m.named_parameters():      # non-degenerate adapters / selectors / bias tables
            if name.endswith("up_proj.weight"):
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.02)
            elif "mlp_token_select" in name and name.endswith("weight"):
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.5)
            elif name.endswith("relative_position_bias_table"):
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.5)
            elif name in ("cls_token",):
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.02)
        img = torch.randn(3, 3, 64, 64, generator=g)
        feats, d = m(img)
    return dict(state_dict={k: v.clone() for k, v in m.state_dict().items()}, img=img,
                features=[f.clone() for f in feats], token_select=d["token_select"].clone(),
                token_logits=d["token_logits"].clone(), loss=d["loss"].clone(),
                keys=sorted(m.state_dict().keys()))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--only":      # regenerate a single fixture
        assert ref_shim.reference_available()
        if sys.argv[2] == "count_flops_tiny":
            ref_train = ref_shim.import_reference("models.vision_transformer_IN21K")
            torch.save(count_flops_tiny(ref_train), os.path.join(OUT, "count_flops_tiny.pt"))
            return
        raise SystemExit("unknown fixture " + sys.argv[2])
    assert ref_shim.reference_available(), "needs the reference checkout at " + ref_shim.REFERENCE_ROOT
    torch.set_num_threads(os.cpu_count())
    ref_dyn = ref_shim.import_reference("models.dynamic_adapter")
    ref_speed = ref_shim.import_reference("models.model_speed_test")
    ref_train = ref_shim.import_reference("models.vision_transformer_IN21K")
    os.makedirs(OUT, exist_ok=True)
    torch.save(gate_tables(ref_speed, ref_dyn), os.path.join(OUT, "gate_tables.pt"))
    torch.save(gumbel_train(ref_dyn), os.path.join(OUT, "gumbel_train.pt"))
    torch.save(tiny_vit(ref_speed, ref_train), os.path.join(OUT, "tiny_vit.pt"))
    torch.save(vitb_b2(ref_speed, ref_train), os.path.join(OUT, "vitb_b2.pt"))
    torch.save(count_flops_tiny(ref_train), os.path.join(OUT, "count_flops_tiny.pt"))
    ref_video = ref_shim.import_reference("video_models.video_vision_transformer_IN21K")
    torch.save(tiny_video(ref_video), os.path.join(OUT, "video_tiny.pt"))
    torch.save(flops_accounting(), os.path.join(OUT, "flops_accounting.pt"))
    torch.save(seg_attention(), os.path.join(OUT, "seg_attention.pt"))
    torch.save(seg_tiny(), os.path.join(OUT, "seg_tiny.pt"))
    ref_losses = ref_shim.import_reference("models.losses")
    torch.save(tiny_finetune(ref_train, ref_losses), os.path.join(OUT, "finetune_tiny.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
