"""GPU parity tests of the whole hot path: dyt_block_fwd (one block) and the drop-in models, against
the CPU oracle ("amp16" = fp16-autocast arithmetic) and the committed golden outputs of the
unmodified reference (tests/golden/, fp32 CPU).

Tolerances (stated per BASELINE.json: masks bit-exact, block outputs within 1e-3 rel):
  * block output vs the amp16 oracle on identical inputs: max|err| <= 1e-3 * max|ref|
  * masks: bit-equal; a mismatch is tolerated only for a token whose oracle logit lies within one
    fp16 ulp of the gate threshold (accumulation-order noise), and is counted
  * whole 12-layer model vs the fp32 reference golden: fp16 arithmetic + borderline mask flips in
    later layers bound the logits to 3e-2 * max|ref| (the autocast reference itself is this far from
    its own fp32 run); per-layer mask agreement >= 99 %
"""
import ctypes as C

import pytest
import torch

import dyt_oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu


class Cfg(dict):
    __getattr__ = dict.__getitem__


def configs(ffn_num=64, scalar="0.1", d_model=768):
    tuning = Cfg(ffn_adapt=True, ffn_option="parallel", ffn_adapter_layernorm_option="none",
                 ffn_adapter_init_option="lora", ffn_adapter_scalar=scalar, ffn_num=ffn_num,
                 d_model=d_model, vpt_on=False, vpt_num=0)
    select = Cfg(open=True, keep_layers=0, token_target_ratio=0.5)
    return tuning, select


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def vitb_sd():
    g = load_golden("vitb_b2.pt")
    sd = O.synthetic_state_dict(seed=g["seed"])
    for i in range(12):
        sd[f"blocks.{i}.mlp_token_select.mlp_head.bias"] = g["selector_bias"][i].clone()
    img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(g["img_seed"]))
    return g, sd, img


def _rel(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return ((got - ref).abs().max() / ref.abs().max()).item()


def _speed_model(sd, dev, num_classes=100):
    from models.model_speed_test import vit_base_patch16_224_in21k
    tuning, select = configs()
    m = vit_base_patch16_224_in21k(num_classes=num_classes, drop_path_rate=0.0,
                                   tuning_config=tuning, select_config=select)
    m.load_state_dict(sd, strict=True)
    return m.eval().to(dev)


def _train_model(sd, dev, num_classes=100):
    from models.vision_transformer_IN21K import vit_base_patch16_224_in21k
    tuning, select = configs()
    m = vit_base_patch16_224_in21k(num_classes=num_classes, drop_path_rate=0.0,
                                   tuning_config=tuning, select_config=select)
    m.load_state_dict(sd, strict=True)
    return m.eval().to(dev)


def _rel_rows(got, ref, rows):
    """max|err| over the selected token rows [B, N] relative to max|ref| of the whole tensor: a token
    whose gate decision sits on the threshold tie changes its own row by a full MLP output, every
    other row of the block is still compared."""
    got, ref = got.float().cpu(), ref.float().cpu()
    assert bool(rows.any())
    return ((got - ref)[rows].abs().max() / ref.abs().max()).item()


def _check_masks(mask, mask_ref, logit_ref, max_flips, near_tol=1.2e-5):
    """Gate decisions must be bit-equal; a mismatch is tolerated (and counted) only for a token whose
    reference logit lies within near_tol of the gate threshold: the 768-term fp32 dot product of the
    logit (|logit| ~ 10) carries ~1e-6 relative accumulation-order noise."""
    mism = mask != mask_ref
    n = int(mism.sum())
    if n:
        thr = O.min_kept_logit(torch.float16)
        dist = (logit_ref[mism[:, 1:]] - thr).abs()
        assert bool((dist <= near_tol).all()), f"mask mismatch away from the gate threshold: {dist}"
    assert n <= max_flips
    return n


# ---------------------------------------------------------------------------------------------
# one block, identical inputs
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B", [1, 3])
def test_block_matches_oracle(dev, vitb_sd, B):
    g, sd, img = vitb_sd
    m = _speed_model(sd, dev)
    x = torch.randn(B, 197, 768, generator=torch.Generator().manual_seed(B)) * 0.7
    for layer in (0, 7):
        ref = O.block_sparse(x, sd, f"blocks.{layer}.", 12, g["scale"], "amp16")
        with torch.no_grad():
            out = m.blocks[layer](x.to(dev))
        assert out.dtype == torch.float32 and out.shape == x.shape
        from dyt_b200 import engine
        out2, masks, logits, _ = engine.run_blocks(x.to(dev), [m.blocks[layer]], fuse_next_ln=False)
        assert torch.equal(out, out2)
        _check_masks(masks[0].unsqueeze(-1).cpu(), ref["mask"], ref["logits"], max_flips=1)
        same = masks[0].cpu() == ref["mask"][..., 0]            # [B, N]: rows with the oracle's decision
        assert _rel_rows(out, ref["out"], same) <= 1e-3
        assert _rel(logits[0].unsqueeze(-1), ref["logits"]) <= 2e-3


def test_fused_selector_score_matches_standalone_dispatcher(dev, vitb_sd):
    """The block computes the selector logits inside the proj GEMM epilogue (column partial sums);
    the standalone dispatcher on the block's own x1 buffer must give the same logits (fp32
    summation order differs: <= 1 fp16 ulp) and the same mask except on threshold ties."""
    g, sd, img = vitb_sd
    m = _speed_model(sd, dev)
    from dyt_b200 import engine, ops
    import ctypes as C
    B = 4
    x = torch.randn(B, 197, 768, generator=torch.Generator().manual_seed(11)) * 0.7
    blk = m.blocks[5]
    out, masks, logits, _ = engine.run_blocks(x.to(dev), [blk], fuse_next_ln=False)
    shape = engine.block_shape_of(blk, B, 197)
    ws = engine._workspace(shape, dev)
    bufs = engine.workspace_buffers(shape, ws)
    off = bufs.x1 - ws.data_ptr()
    x1 = ws[off:off + B * 197 * 768 * 4].view(torch.float32).reshape(B, 197, 768).clone()
    d = ops.dispatch(x1, blk.mlp_token_select.mlp_head.weight.detach(),
                     blk.mlp_token_select.mlp_head.bias.detach(), pack=False)
    l_ref, l_got = d["logits"][..., 0], logits[0]
    assert float((l_ref - l_got).abs().max()) <= 2e-3 * float(l_ref.abs().max())
    mism = d["mask"][..., 0] != masks[0]
    if bool(mism.any()):
        thr = O.min_kept_logit(torch.float16)
        assert bool(((l_ref[mism[:, 1:]] - thr).abs() <= 2e-3).all())
    assert int(mism.sum()) <= 2


def test_block_imposed_mask_config1(dev, vitb_sd):
    """BASELINE configs[0] on the GPU path: fixed 50 % keep mask, block output vs oracle."""
    g, sd, img = vitb_sd
    m = _speed_model(sd, dev)
    from dyt_b200 import engine
    x = torch.randn(2, 197, 768, generator=torch.Generator().manual_seed(42))
    forced = O.checkerboard_mask(2, 197)
    ref = O.block_sparse(x, sd, "blocks.3.", 12, g["scale"], "amp16", forced_mask=forced)
    out, masks, _, _ = engine.run_blocks(x.to(dev), [m.blocks[3]], forced_masks=[forced],
                                         fuse_next_ln=False)
    assert torch.equal(masks[0].cpu(), forced[..., 0])
    assert _rel(out, ref["out"]) <= 1e-3
    # dropped tokens receive no MLP contribution: out - (x1 + adapter) == 0 there, so the dense
    # all-kept run must differ on kept tokens only through the MLP term
    ones = torch.ones(2, 197, 1)
    out_all, _, _, _ = engine.run_blocks(x.to(dev), [m.blocks[3]], forced_masks=[ones],
                                         fuse_next_ln=False)
    ref_all = O.block_dense(x, sd, "blocks.3.", 12, g["scale"], "amp16", complete_model=True)
    assert _rel(out_all, ref_all["out"]) <= 1e-3
    kept = forced[..., 0].bool()
    assert torch.equal(out_all.cpu()[kept], out.cpu()[kept])      # kept rows identical, bit-exact
    assert not torch.equal(out_all.cpu()[~kept], out.cpu()[~kept])


def test_block_fused_adapter_down_option(dev, vitb_sd):
    """DYT_OPT_FUSE_ADAPTER_DOWN: the whole adapter branch inside the merge kernel (no fp16 copy of x1,
    no down GEMM) against the default path on one ViT-B block: same gate decisions, outputs to
    rounding noise (the down accumulation runs in another tile shape)."""
    g, sd, img = vitb_sd
    m = _speed_model(sd, dev)
    from dyt_b200 import engine, _lib
    lib = _lib.lib()
    x = torch.randn(3, 197, 768, generator=torch.Generator().manual_seed(7)).to(dev)
    base = engine.run_blocks(x, list(m.blocks[1:3]), fuse_next_ln=True)
    try:
        assert lib.dyt_configure(_lib.OPT_FUSE_ADAPTER_DOWN, 1) == 0
        fused = engine.run_blocks(x, list(m.blocks[1:3]), fuse_next_ln=True)
    finally:
        assert lib.dyt_configure(_lib.OPT_FUSE_ADAPTER_DOWN, 0) == 0
    assert torch.equal(fused[1], base[1])
    assert _rel(fused[0], base[0]) <= 5e-4
    assert _rel(fused[2], base[2]) <= 2e-3


def _moe_params(C, K, E, seed):
    g = torch.Generator().manual_seed(seed)
    p = {"adaptmlp.router.weight": torch.randn(E, C, generator=g) * 0.3,
         "adaptmlp.router.bias": torch.randn(E, generator=g) * 0.2}
    for i in range(E):
        p[f"adaptmlp.down_proj.{i}.weight"] = torch.randn(K, C, generator=g) * 0.03
        p[f"adaptmlp.down_proj.{i}.bias"] = torch.randn(K, generator=g) * 0.1
        p[f"adaptmlp.up_proj.{i}.weight"] = torch.randn(C, K, generator=g) * 0.05
        p[f"adaptmlp.up_proj.{i}.bias"] = torch.randn(C, generator=g) * 0.1
    return p


@pytest.mark.parametrize("B,N,C,K,E", [(3, 197, 768, 64, 4), (2, 50, 128, 16, 2), (5, 197, 1024, 64, 8)])
def test_moe_adapter_kernels_vs_own_oracle(dev, B, N, C, K, E):
    """dyt_moe_adapter_fwd against this repository's own restatement (oracle.moe_adapter): the
    MoE-adapter is NOT in the reference, so there is no reference parity to claim.  amp16 rounding
    points to 2e-3, and the weight-space fp32 definition to fp16 noise."""
    from dyt_b200 import ops
    p = _moe_params(C, K, E, 50 + E)
    x1 = torch.randn(B, N, C, generator=torch.Generator().manual_seed(3)) * 2 + 0.2
    got = ops.moe_adapter(x1.to(dev), p["adaptmlp.router.weight"].to(dev), p["adaptmlp.router.bias"].to(dev),
                          [p[f"adaptmlp.down_proj.{i}.weight"].to(dev) for i in range(E)],
                          [p[f"adaptmlp.down_proj.{i}.bias"].to(dev) for i in range(E)],
                          [p[f"adaptmlp.up_proj.{i}.weight"].to(dev) for i in range(E)],
                          [p[f"adaptmlp.up_proj.{i}.bias"].to(dev) for i in range(E)], 0.5).float().cpu()
    amp = O.moe_adapter(x1, p, "adaptmlp.", 0.5, E, "amp16")
    ref = O.moe_adapter(x1, p, "adaptmlp.", 0.5, E, "fp32")
    assert float((got - amp).abs().max()) <= 2e-3 * float(amp.abs().max()) + 1e-3
    assert float((got - ref).abs().max()) <= 6e-3 * float(ref.abs().max()) + 2e-3


def test_block_with_moe_adapter_vs_own_oracle(dev, vitb_sd):
    """A ViT-B block whose adapter is the MoE-adapter (tuning_config.moe_experts = 4) through
    dyt_block_fwd against oracle.block_sparse_moe (own restatement: no reference parity)."""
    from models.model_speed_test import Block
    from dyt_b200 import engine
    g, sd, img = vitb_sd
    tuning, _ = configs()
    tuning["moe_experts"] = 4
    blk = Block(dim=768, num_heads=12, mlp_ratio=4.0, qkv_bias=True, tuning_config=tuning, select=True)
    own = {k[len("blocks.2."):]: v for k, v in sd.items()
           if k.startswith("blocks.2.") and "adaptmlp" not in k}
    own.update(_moe_params(768, 64, 4, 77))
    blk.load_state_dict(own, strict=True)
    blk = blk.eval().to(dev)
    x = torch.randn(3, 197, 768, generator=torch.Generator().manual_seed(8))
    out, masks, logits, _ = engine.run_blocks(x.to(dev), [blk], fuse_next_ln=False)
    p = {"blocks.2." + k: v for k, v in own.items()}
    ref = O.block_sparse_moe(x, p, "blocks.2.", 12, g["scale"], 4, "amp16",
                             forced_mask=masks[0].cpu().unsqueeze(-1))
    assert _rel(out, ref["out"]) <= 1e-3
    own_gate = O.block_sparse_moe(x, p, "blocks.2.", 12, g["scale"], 4, "amp16")
    flips = int((own_gate["mask"][..., 0] != masks[0].cpu()).sum())
    assert flips <= 2


def test_fused_next_layernorm_is_equivalent(dev, vitb_sd):
    g, sd, img = vitb_sd
    m = _speed_model(sd, dev)
    from dyt_b200 import engine
    x = torch.randn(2, 197, 768, generator=torch.Generator().manual_seed(5)).to(dev)
    from dyt_b200 import _lib
    lib = _lib.lib()
    try:
        # separate up GEMM + scatter-merge: the merge's LayerNorm is the stand-alone kernel's code
        assert lib.dyt_configure(_lib.OPT_FUSE_ADAPTER_UP, 0) == 0
        a = engine.run_blocks(x, list(m.blocks[:4]), fuse_next_ln=True)
        b = engine.run_blocks(x, list(m.blocks[:4]), fuse_next_ln=False)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    finally:
        assert lib.dyt_configure(_lib.OPT_FUSE_ADAPTER_UP, 1) == 0
    # adapter-up fused into the merge (default): same fp32 stream after one block bit for bit; its
    # LayerNorm sums in another order (one fp16 ulp on a few elements), so deeper outputs agree to
    # rounding noise
    c1 = engine.run_blocks(x, list(m.blocks[:1]), fuse_next_ln=True)
    b1 = engine.run_blocks(x, list(m.blocks[:1]), fuse_next_ln=False)
    assert torch.equal(c1[0], b1[0]) and torch.equal(c1[1], b1[1])
    c = engine.run_blocks(x, list(m.blocks[:4]), fuse_next_ln=True)
    assert torch.equal(c[1], b[1])
    assert _rel(c[0], b[0]) <= 1e-3


def test_block_fused_adapter_up_equals_separate_launches(dev, vitb_sd):
    """dyt_block_fwd with DYT_OPT_FUSE_ADAPTER_UP on / off: one block's output, mask and logits are
    bit-equal (the fused kernel performs the same roundings and fp32 additions)."""
    g, sd, img = vitb_sd
    m = _speed_model(sd, dev)
    from dyt_b200 import engine, _lib
    lib = _lib.lib()
    x = torch.randn(3, 197, 768, generator=torch.Generator().manual_seed(6)).to(dev)
    on = engine.run_blocks(x, list(m.blocks[2:3]), fuse_next_ln=False)
    try:
        assert lib.dyt_configure(_lib.OPT_FUSE_ADAPTER_UP, 0) == 0
        off = engine.run_blocks(x, list(m.blocks[2:3]), fuse_next_ln=False)
    finally:
        assert lib.dyt_configure(_lib.OPT_FUSE_ADAPTER_UP, 1) == 0
    for u, v in zip(on[:3], off[:3]):
        assert torch.equal(u, v)


# ---------------------------------------------------------------------------------------------
# whole model (drop-in surface) vs oracle and vs the reference golden
# ---------------------------------------------------------------------------------------------
def test_speed_model_vs_reference_golden(dev, vitb_sd):
    g, sd, img = vitb_sd
    m = _speed_model(sd, dev)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        logits = m(img.to(dev))
    assert logits.shape == (2, 100)
    assert logits.dtype == torch.float16                   # what the reference returns under autocast
    assert _rel(logits, g["speed_logits"]) <= 3e-2
    ref = O.vit_forward(img, sd, 12, 12, g["scale"], policy="amp16")
    assert _rel(logits, ref["logits"]) <= 3e-2


def test_train_model_eval_outputs(dev, vitb_sd):
    g, sd, img = vitb_sd
    m = _train_model(sd, dev)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        logits, d = m(img.to(dev))
        t_logits, t_d = m(img.to(dev), complete_model=True)
    assert d["token_select"].shape == (2, 12, 196, 1) and d["token_logits"].shape == (2, 12, 196, 1)
    agree = (d["token_select"].float().cpu() == g["token_select"].float()).float().mean(dim=(0, 2, 3))
    assert bool((agree >= 0.99).all()), f"per-layer mask agreement {agree}"
    # layer 0 sees identical inputs; only fp16-vs-fp32 logit noise at the threshold can flip a token
    assert agree[0] >= 0.995
    assert _rel(logits, g["train_logits"]) <= 3e-2
    assert _rel(t_logits, g["teacher_logits"]) <= 3e-2
    keep = d["token_select"].float().mean().item()
    assert 0.45 < keep < 0.55


def test_model_config1_imposed_mask_vs_golden(dev, vitb_sd):
    """With the mask imposed there are no data-dependent flips: the 12-layer fp16 forward must
    track the fp32 reference golden to fp16 accuracy."""
    g, sd, img = vitb_sd
    m = _speed_model(sd, dev)
    from dyt_b200 import engine
    forced = [O.checkerboard_mask(2, 197)] * 12
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        x = m._embed(img.to(dev))
        x, masks, _, _ = engine.run_blocks(x, list(m.blocks), forced_masks=forced)
        logits = m.forward_head(m.norm(x))
    assert _rel(logits, g["config1_speed_logits"]) <= 1e-2
    ref = O.vit_forward(img, sd, 12, 12, g["scale"], policy="amp16", forced_masks=forced)
    assert _rel(logits, ref["logits"]) <= 5e-3


def test_vit_large_block_matches_oracle(dev):
    """BASELINE configs[3] shape (ViT-L/16: C = 1024, 16 heads, hidden 4096; the MoE-adapter of that
    config does not exist in the reference): one block and a 2-layer model against
    the oracle, which is the reference Block re-parameterised (vision_transformer_IN21K.py:199-231)."""
    from models.model_speed_test import VisionTransformer
    from dyt_b200 import engine
    tuning, select = configs(d_model=1024)
    sd = O.synthetic_state_dict(embed_dim=1024, depth=2, num_heads=16, seed=3)
    img = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    sd = O.calibrate_selector_bias(sd, img, 2, 16, 0.1, 0.7)
    m = VisionTransformer(embed_dim=1024, depth=2, num_heads=16, num_classes=100,
                          tuning_config=tuning, select_config=select)
    m.load_state_dict(sd, strict=True)
    m = m.eval().to(dev)
    x = torch.randn(2, 197, 1024, generator=torch.Generator().manual_seed(5)) * 0.7
    ref = O.block_sparse(x, sd, "blocks.1.", 16, 0.1, "amp16")
    out, masks, logits, _ = engine.run_blocks(x.to(dev), [m.blocks[1]], fuse_next_ln=False)
    _check_masks(masks[0].unsqueeze(-1).cpu(), ref["mask"], ref["logits"], max_flips=1)
    same = masks[0].cpu() == ref["mask"][..., 0]
    assert _rel_rows(out, ref["out"], same) <= 1e-3
    keep = float(masks[0][:, 1:].mean())
    assert 0.4 < keep < 0.95
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        got = m(img.to(dev))
    full = O.vit_forward(img, sd, 2, 16, 0.1, policy="amp16", sparse=True)
    assert _rel(got.float(), full["logits"]) <= 2e-2


def test_tiny_dims_unsupported_are_loud(dev):
    """head_dim != 64 is outside the implemented kernels: must raise, never fall back."""
    from dyt_b200 import DytError
    from models.model_speed_test import VisionTransformer
    tuning, select = configs(ffn_num=16, d_model=96)
    m = VisionTransformer(img_size=32, patch_size=16, embed_dim=96, depth=1, num_heads=2,
                          num_classes=10, tuning_config=tuning, select_config=select).eval().to(dev)
    with pytest.raises(DytError), torch.no_grad():
        m(torch.randn(2, 3, 32, 32, device=dev))


def test_tiny_vit_golden_dim128(dev):
    """The 2-layer dim-128 reference model (heads 2 x 64) runs through the same kernels."""
    t = load_golden("tiny_vit.pt")
    from models.model_speed_test import VisionTransformer
    tuning, select = configs(ffn_num=16, d_model=128)
    m = VisionTransformer(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2,
                          num_classes=10, tuning_config=tuning, select_config=select)
    m.load_state_dict(t["state_dict"], strict=True)
    m = m.eval().to(dev)
    with torch.no_grad():
        logits = m(t["img"].to(dev))
        blk_out = m.blocks[0](t["x0"].to(dev))
    ref = O.vit_forward(t["img"], t["state_dict"], 2, 2, t["scale"], policy="amp16")
    assert _rel(logits, ref["logits"]) <= 5e-3
    assert _rel(logits, t["speed_logits"]) <= 2e-2
    assert _rel(blk_out, t["blk0_out"]) <= 2e-3


# ---------------------------------------------------------------------------------------------
# BASELINE sizes (batch 256): size-independent properties
# ---------------------------------------------------------------------------------------------
def test_full_batch_properties(dev, vitb_sd):
    g, sd, img = vitb_sd
    m = _speed_model(sd, dev)
    from dyt_b200 import engine
    B = 256
    x = torch.randn(B, 197, 768, generator=torch.Generator().manual_seed(1)).to(dev) * 0.7
    blocks = list(m.blocks[:2])
    out, masks, logits, _ = engine.run_blocks(x, blocks)
    assert bool(torch.isfinite(out).all())
    # (1) per-image independence: a permutation of the batch permutes the outputs, bit-exactly
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(2)).to(dev)
    out_p, masks_p, _, _ = engine.run_blocks(x[perm], blocks)
    assert torch.equal(out_p, out[perm]) and torch.equal(masks_p, masks[:, perm])
    # (2) compaction bookkeeping of the last layer run
    shape = engine.block_shape_of(blocks[-1], B, 197)
    ws = engine._workspace(shape, dev)
    bufs = engine.workspace_buffers(shape, ws)
    def view(ptr, n, dtype):
        off = ptr - ws.data_ptr()
        return ws[off:off + n * torch.empty((), dtype=dtype).element_size()].view(dtype)
    n_kept = int(view(bufs.n_kept, 1, torch.int32).item())
    cu = view(bufs.cu_seqlens, B + 1, torch.int32).cpu()
    idx = view(bufs.packed_idx, B * 197, torch.int32)[:n_kept].cpu().long()
    last = masks_p[-1].cpu()       # the workspace holds the most recent (permuted) run
    assert n_kept == int(last.sum().item()) == int(cu[-1])
    assert torch.equal(cu[1:] - cu[:-1], last.sum(dim=1).to(torch.int32))
    assert bool((idx[1:] > idx[:-1]).all())
    assert torch.equal(idx, last.reshape(-1).nonzero()[:, 0])
    # (3) an all-dropped (cls only) and an all-kept mask both run (ragged extremes)
    none = torch.zeros(B, 197, 1)
    none[:, 0] = 1
    o0, m0, _, _ = engine.run_blocks(x, blocks[:1], forced_masks=[none])
    o1, m1, _, _ = engine.run_blocks(x, blocks[:1], forced_masks=[torch.ones(B, 197, 1)])
    assert int(m0.sum()) == B and int(m1.sum()) == B * 197
    assert bool(torch.isfinite(o0).all()) and bool(torch.isfinite(o1).all())
    assert torch.equal(o0[:, 0], o1[:, 0])          # cls rows are kept in both


def test_full_batch_matches_oracle(dev, vitb_sd):
    """BASELINE batch size (256 images, T = 50 432 tokens): two layers against the amp16 oracle on
    identical inputs.  Exercises what the small batches do not: several rounds of the persistent
    GEMM tile loops, the device-side row count of the MLP GEMMs, the 256-image look-back of the
    dispatcher.  Layer by layer the oracle continues with the kernels' own gate decisions (after
    they were checked against its own: bit-equal except threshold ties), so every row is compared."""
    g, sd, img = vitb_sd
    m = _speed_model(sd, dev)
    from dyt_b200 import engine
    B = 256
    x = torch.randn(B, 197, 768, generator=torch.Generator().manual_seed(77)) * 0.7
    torch.set_num_threads(max(1, (torch.get_num_threads())))
    xg = x.to(dev)
    xo = x
    total_flips = 0
    for layer in (0, 1):
        outg, masks, logits, _ = engine.run_blocks(xg, [m.blocks[layer]], fuse_next_ln=False)
        free = O.block_sparse(xo, sd, f"blocks.{layer}.", 12, g["scale"], "amp16")
        # 50 432 tokens per layer with logits ~ N(0, 10^2): about a dozen lie within 5e-3 of the
        # threshold, and the logit inherits the (equally valid) fp16 rounding differences of x1
        # between two fp32-accumulating implementations: ~2e-4 per element x sqrt(768) x |w| ~ 3e-3
        total_flips += _check_masks(masks[0].unsqueeze(-1).cpu(), free["mask"], free["logits"],
                                    max_flips=12, near_tol=5e-3)
        assert _rel(logits[0].unsqueeze(-1), free["logits"]) <= 2e-3
        ref = O.block_sparse(xo, sd, f"blocks.{layer}.", 12, g["scale"], "amp16",
                             forced_mask=masks[0].unsqueeze(-1).cpu())
        assert _rel(outg, ref["out"]) <= 1e-3, layer
        assert int(masks[0].sum()) == int(ref["cu_seqlens"][-1])
        xg, xo = outg, ref["out"]                      # each side continues with its own stream
    assert total_flips <= 16


def test_forward_count_flops_matches_oracle(dev, vitb_sd):
    """a11: Block.forward_count_flops (block_flops_dict.get_block_flops' probe): MLP on the FIRST
    token_select_num tokens (vision_transformer_IN21K.py:167-185); oracle pinned to the reference by
    tests/golden/count_flops_tiny.pt."""
    g, sd, img = vitb_sd
    m = _train_model(sd, dev)
    x = torch.randn(2, 197, 768, generator=torch.Generator().manual_seed(8)) * 0.7
    blk = m.blocks[2]
    blk.count_flops = True
    try:
        for t in (1, 64, 197):
            blk.token_select_num = t
            with torch.no_grad():
                out = blk(x.to(dev))
            ref = O.block_count_flops(x, sd, "blocks.2.", 12, g["scale"], t, "amp16")
            assert out.shape == x.shape
            assert _rel(out, ref) <= 1e-3, t
    finally:
        blk.count_flops = None
        blk.token_select_num = None


# ---------------------------------------------------------------------------------------------
# video model: per-frame DyT blocks + attentive pooling head
# ---------------------------------------------------------------------------------------------
def test_pooling_head_kernels_match_oracle(dev):
    """Fused final-norm + norm_k / norm_v LayerNorm kernel and the single-query attention kernel
    against the oracle's restatement of AttentiveBlock / CrossAttention (amp16 policy)."""
    from dyt_b200 import ops
    g = torch.Generator().manual_seed(3)
    b, nk, C, H = 3, 2 * 197, 768, 12
    x = torch.randn(b, nk, C, generator=g) * 1.5 + 0.2
    n0 = (1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g))
    nk_ = (1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g))
    nv_ = (1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g))
    xn = O.layer_norm(x, n0[0], n0[1])
    ref_k, ref_v = O._r16(O.layer_norm(xn, *nk_)), O._r16(O.layer_norm(xn, *nv_))
    to = lambda t: tuple(u.to(dev) for u in t)
    outk, outv = ops.pool_layernorm_f16(x.to(dev), to(n0), to(nk_), to(nv_), 1e-6)
    assert float((outk.float().cpu() - ref_k).abs().max()) <= 4e-3      # one fp16 ulp at |x| ~ 4
    assert float((outv.float().cpu() - ref_v).abs().max()) <= 4e-3
    q = O._r16(torch.randn(C, generator=g) * 0.5)
    k = O._r16(torch.randn(b, nk, C, generator=g))
    v = O._r16(torch.randn(b, nk, C, generator=g))
    qh = q.reshape(1, H, 1, 64)
    kh = k.reshape(b, nk, H, 64).permute(0, 2, 1, 3)
    vh = v.reshape(b, nk, H, 64).permute(0, 2, 1, 3)
    attn = O._r16(qh @ kh.transpose(-2, -1)).softmax(dim=-1)
    ref = O._r16(O._r16(attn) @ vh).transpose(1, 2).reshape(b, C)
    out = ops.query_attn(q.half().to(dev), k.half().to(dev), v.half().to(dev), H)
    assert float((out.float().cpu() - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 1e-3


def test_tiny_video_model_vs_reference_golden(dev):
    """The drop-in video model on the reference's own tiny video model outputs (fp32 golden) and on
    the oracle (amp16): logits, per-frame token masks, pooled features."""
    from video_models.video_vision_transformer_IN21K import VisionTransformer
    g = load_golden("video_tiny.pt")
    d = g["dims"]
    tuning, select = configs(ffn_num=d["bottleneck"], d_model=d["embed_dim"])
    m = VisionTransformer(img_size=d["img_size"], patch_size=16, embed_dim=d["embed_dim"],
                          depth=d["depth"], num_heads=d["num_heads"], mlp_ratio=4.0, qkv_bias=True,
                          num_classes=d["num_classes"], tuning_config=tuning, select_config=select)
    m.load_state_dict(g["state_dict"], strict=True)
    m = m.eval().to(dev)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        logits, ts = m(g["clip"].to(dev))
    assert logits.shape == g["logits"].shape and logits.dtype == torch.float16
    assert ts["token_select"].shape == g["token_select"].shape
    ref16 = O.video_forward(g["clip"], g["state_dict"], d["depth"], d["num_heads"], g["scale"],
                            policy="amp16")
    got_sel = ts["token_select"].float().cpu()                    # [b*t, depth, N-1, 1]
    mism = int((got_sel != ref16["token_select"].float()).sum())
    assert mism <= 1
    # the oracle continued with the kernels' own gate decisions: compared whatever the flip count
    forced = [torch.cat([torch.ones(got_sel.shape[0], 1, 1), got_sel[:, i]], dim=1)
              for i in range(d["depth"])]
    ref_f = O.video_forward(g["clip"], g["state_dict"], d["depth"], d["num_heads"], g["scale"],
                            policy="amp16", forced_masks=forced)
    assert _rel(logits.float(), ref_f["logits"]) <= 1e-2
    if bool((got_sel == g["token_select"].float()).all()):        # same decisions as the fp32 reference
        assert _rel(logits.float(), g["logits"]) <= 3e-2


def test_forward_is_cuda_graph_capturable_and_replay_safe(dev, vitb_sd):
    """No host synchronisation, no allocation inside the C ABI: the whole forward captures into a
    CUDA graph; replays with new input contents reproduce the eager results (the dispatcher's
    launch epoch lives in the workspace, so replays do not see stale look-back state)."""
    g, sd, img = vitb_sd
    m = _speed_model(sd, dev)
    imgs = [torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(s)).to(dev)
            for s in (1, 2, 3)]

    def fwd(x):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            return m(x)

    eager = [fwd(x).clone() for x in imgs]
    static_in = imgs[0].clone()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            fwd(static_in)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_out = fwd(static_in)
    for x, ref in zip(imgs, eager):
        static_in.copy_(x)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(static_out, ref)


def test_graphed_forward_several_batch_sizes_interleaved(dev, vitb_sd):
    """Graphs of different batch sizes captured one after the other (larger first) and replayed
    interleaved, with allocator churn in between: no captured graph may read a cache entry or a
    workspace that a later call replaced (regression: the cls-row index tensor of the head was a
    single-entry cache -- replaying the larger graph after a smaller call faulted)."""
    from dyt_b200 import GraphedForward
    g, sd, img = vitb_sd
    m = _speed_model(sd, dev)
    gm = GraphedForward(m)
    xs = {b: torch.randn(b, 3, 224, 224, generator=torch.Generator().manual_seed(40 + b)).to(dev)
          for b in (6, 3, 8)}
    refs = {}
    for b, x in xs.items():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            refs[b] = m(x).clone()
    for b in (6, 3, 6, 8, 3, 6, 8):
        out = gm(xs[b])
        junk = [torch.full((1 << 18,), float(b), device=dev) for _ in range(8)]   # reuse freed blocks
        torch.cuda.synchronize()
        assert torch.equal(out, refs[b]), f"batch {b}"
        del junk


def test_programmatic_dependent_launch_is_bit_identical(dev, vitb_sd):
    """DYT_OPT_PDL (default on): every forward-path kernel is launched with programmatic stream
    serialization and waits (griddepcontrol.wait) before its first dependent global access.  The
    whole model, eager and as a CUDA graph, must give bit-identical logits with and without it,
    repeatedly (a missing wait would show up as a race)."""
    from dyt_b200 import GraphedForward, _lib
    lib = _lib.lib()
    g, sd, img = vitb_sd
    m = _speed_model(sd, dev)
    xs = [torch.randn(b, 3, 224, 224, generator=torch.Generator().manual_seed(60 + b)).to(dev) for b in (2, 24)]

    def fwd(x):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            return m(x).clone()

    try:
        assert lib.dyt_configure(_lib.OPT_PDL, 0) == 0
        ref = [fwd(x) for x in xs]
    finally:
        assert lib.dyt_configure(_lib.OPT_PDL, 1) == 0
    gm = GraphedForward(m)            # captured with PDL edges
    for rep in range(6):
        for x, r in zip(xs, ref):
            assert torch.equal(fwd(x), r), f"eager, repetition {rep}"
            assert torch.equal(gm(x), r), f"graph replay, repetition {rep}"


def test_gemm_tile_order_and_side_plan_are_bit_identical(dev, vitb_sd):
    """DYT_OPT_TILE_ORDER (default 7: the qkv, proj and fc2 GEMMs walk their row tiles from the last to
    the first so that each starts on the rows its producer left in the L2): only the order of the
    tiles changes, so the whole model must give bit-identical logits for every mask, including
    shapes with a partial last row pair and a split tail round."""
    from dyt_b200 import _lib
    lib = _lib.lib()
    g, sd, img = vitb_sd
    m = _speed_model(sd, dev)
    xs = [torch.randn(b, 3, 224, 224, generator=torch.Generator().manual_seed(70 + b)).to(dev) for b in (3, 40)]

    def fwd(x):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            return m(x).clone()

    try:
        assert lib.dyt_configure(_lib.OPT_TILE_ORDER, 0) == 0
        ref = [fwd(x) for x in xs]
        for mask in (1, 2, 4, 8, 7, 15):
            assert lib.dyt_configure(_lib.OPT_TILE_ORDER, mask) == 0
            for x, r in zip(xs, ref):
                assert torch.equal(fwd(x), r), f"tile order mask {mask}, batch {x.shape[0]}"
        assert lib.dyt_configure(_lib.OPT_TILE_ORDER, 7) == 0
        # DYT_OPT_SIDE_PLAN (where and how wide the adapter's down GEMM runs beside the dispatcher):
        # scheduling only
        for plan in (1, 2, 4, 5, 0):
            assert lib.dyt_configure(_lib.OPT_SIDE_PLAN, plan) == 0
            for x, r in zip(xs, ref):
                assert torch.equal(fwd(x), r), f"side plan {plan}, batch {x.shape[0]}"
    finally:
        assert lib.dyt_configure(_lib.OPT_TILE_ORDER, 7) == 0
        assert lib.dyt_configure(_lib.OPT_SIDE_PLAN, 0) == 0


def test_full_batch_kernels_run_on_full_grids(dev, vitb_sd, tmp_path):
    """At the BASELINE batch (256 images x 2 layers) every persistent kernel of the block is launched on
    one CTA per SM and every launch is one of the library's kernels (a scheduling flag that leaked into
    the wrong GEMM once halved fc2's grid without changing any result: results cannot catch that)."""
    import json
    from torch.profiler import profile, ProfilerActivity
    from dyt_b200 import engine
    g, sd, img = vitb_sd
    m = _speed_model(sd, dev)
    x = torch.randn(256, 197, 768, generator=torch.Generator().manual_seed(3)).to(dev) * 0.7
    blocks = list(m.blocks[:2])
    engine.run_blocks(x.clone(), blocks)
    torch.cuda.synchronize()
    path = str(tmp_path / "trace.json")
    try:
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            engine.run_blocks(x.clone(), blocks)
            torch.cuda.synchronize()
        prof.export_chrome_trace(path)
    except RuntimeError as e:   # CUPTI taken by another tool
        pytest.skip(f"torch profiler unavailable: {e}")
    events = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
    if not events:   # no CUPTI activity records (e.g. the tests themselves run under ncu / nsys)
        pytest.skip("the torch profiler recorded no kernels here")
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    seen = {}
    for e in events:
        name, grid = e["name"], e["args"]["grid"]
        if "dyt::" not in name:
            assert "elementwise" in name or "copy" in name.lower() or "fill" in name.lower(), name   # the clone
            continue
        for key in ("gemm_tn_kernel", "attn_split_kernel", "merge_up_kernel", "dispatch_kernel"):
            if key in name:
                seen.setdefault(key, []).append(grid[0])
    assert len(seen["gemm_tn_kernel"]) == 2 * 5 and set(seen["gemm_tn_kernel"]) == {sms}, seen
    assert set(seen["attn_split_kernel"]) == {sms} and set(seen["merge_up_kernel"]) == {sms}, seen
    assert set(seen["dispatch_kernel"]) == {256}, seen


def test_graphed_forward_public_wrapper(dev, vitb_sd):
    """dyt_b200.GraphedForward: same logits as the eager call for new input contents, per shape and
    per static-input slot; writing straight into a slot's input buffer + replay works (bench e2e)."""
    from dyt_b200 import GraphedForward
    g, sd, img = vitb_sd
    m = _speed_model(sd, dev)
    gm = GraphedForward(m)
    xs = [torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(s)).to(dev) for s in (5, 6)]
    for x in xs + [xs[0][:2].contiguous()]:
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            ref = m(x)
        assert torch.equal(gm(x), ref)
        assert torch.equal(gm(x, slot=1), ref)
    buf = gm.input_buffer(xs[1].shape, xs[1].dtype, dev, slot=2)
    buf.copy_(xs[1])
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        ref = m(xs[1])
    assert torch.equal(gm.replay(xs[1].shape, xs[1].dtype, dev, slot=2), ref)
    with pytest.raises(Exception):
        gm(torch.zeros(1, 3, 224, 224))                      # CPU tensor: no fallback


def _seg_cfg(ffn_num=16, d_model=128):
    tuning, select = configs(ffn_num=ffn_num, d_model=d_model)
    select.update(token_ratio=2.0, token_minimal=0.1, token_minimal_weight=1.0)
    return tuning, select


def test_segmentation_backbone_vs_reference_golden(dev):
    """The drop-in segmentation backbone (relative-position-bias attention through dyt_attn_bias_fwd,
    DyT blocks, FPN heads) against the unmodified reference (fp32 golden) and the amp16 oracle."""
    from dense_tasks.Segmentation.backbone.segmentation_vision_transformer_IN21K import VisionTransformer21K
    g = load_golden("seg_tiny.pt")
    tuning, select = _seg_cfg()
    m = VisionTransformer21K(img_size=64, patch_size=16, embed_dim=128, depth=4, num_heads=2,
                             num_classes=0, tuning_config=tuning, select_config=select,
                             out_indices=[0, 1, 2, 3], use_rel_pos_bias=True)
    m.load_state_dict(g["state_dict"], strict=True)
    m = m.eval().to(dev)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        feats, d = m(g["img"].to(dev))
    assert len(feats) == 4 and d["token_select"].shape == g["token_select"].shape
    ref16 = O.seg_forward(g["img"], g["state_dict"], 4, 2, 0.1, [0, 1, 2, 3], policy="amp16")
    got_sel = (d["token_select"].float().cpu() > 0.5).float()     # [B, depth, N-1, 1]
    same = (got_sel > 0.5) == (ref16["token_select"] > 0.5)
    assert same.float().mean() >= 0.98
    # the oracle continued with the kernels' own gate decisions: compared whatever the flip count
    forced = [torch.cat([torch.ones(got_sel.shape[0], 1, 1), got_sel[:, i]], dim=1) for i in range(4)]
    ref_f = O.seg_forward(g["img"], g["state_dict"], 4, 2, 0.1, [0, 1, 2, 3], policy="amp16",
                          forced_masks=forced)
    for a, b, c in zip(feats, ref_f["features"], g["features"]):
        assert a.shape == c.shape
        assert _rel(a, b) <= 1e-2
    assert abs(d["loss"].item() - ref_f["loss"].item()) <= 1e-3 * abs(ref_f["loss"].item())
    if bool(same.all()) and bool(((g["token_select"] > 0.5) == (got_sel > 0.5)).all()):
        for a, c in zip(feats, g["features"]):
            assert _rel(a, c) <= 3e-2
    with pytest.raises(NotImplementedError):
        m.train()(g["img"].to(dev))


def test_segmentation_backbone_512_tokens_1025(dev):
    """ViT-B width at 512 x 512 (1025 tokens per image, no relative bias = the reference default
    use_rel_pos_bias=False): two layers against the amp16 oracle."""
    from dense_tasks.Segmentation.backbone.segmentation_vision_transformer_IN21K import VisionTransformer21K
    tuning, select = _seg_cfg(ffn_num=64, d_model=768)
    torch.manual_seed(3)
    m = VisionTransformer21K(img_size=512, patch_size=16, embed_dim=768, depth=2, num_heads=12,
                             num_classes=0, tuning_config=tuning, select_config=select,
                             out_indices=[0, 1], use_rel_pos_bias=False)
    gen = torch.Generator().manual_seed(9)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("up_proj.weight"):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.02)
            elif "mlp_token_select" in n and n.endswith("weight"):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.5)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    img = torch.randn(2, 3, 512, 512, generator=gen)
    m = m.eval().to(dev)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        feats, d = m(img.to(dev))
    assert d["token_select"].shape == (2, 2, 1024, 1)
    assert feats[0].shape == (2, 768, 128, 128) and feats[1].shape == (2, 768, 64, 64)
    ref = O.seg_forward(img, sd, 2, 12, 0.1, [0, 1], policy="amp16")
    agree = ((d["token_select"].float().cpu() > 0.5) == (ref["token_select"] > 0.5)).float().mean(dim=(0, 2, 3))
    assert agree[0] >= 0.995 and agree[1] >= 0.98, agree
    assert _rel(feats[0], ref["features"][0]) <= 2e-2


def test_parameter_caches_are_per_model_and_can_be_invalidated(dev):
    """The fp16 working copies of the stem / block / head parameters belong to their model: a second
    model built after the first was freed (its tensors typically land on the recycled addresses,
    with identical version counters) must not see the first one's copies; a write through `.data`
    (no version bump) is picked up after dyt_b200.invalidate_caches()."""
    import dyt_b200
    from models.model_speed_test import VisionTransformer

    def stem_ref(m, img):
        w = m.patch_embed.proj.weight.detach().half().float()
        b = m.patch_embed.proj.bias.detach().half().float()
        t = torch.nn.functional.conv2d(img.half().float(), w, b, stride=16).flatten(2).transpose(1, 2)
        t = torch.cat((m.cls_token.detach().expand(img.shape[0], -1, -1), t), dim=1)
        return t + m.pos_embed.detach()

    img = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(0)).to(dev)
    tuning, select = configs(ffn_num=16, d_model=128)
    for seed in (1, 2, 3):
        torch.manual_seed(seed)
        m = VisionTransformer(img_size=32, patch_size=16, embed_dim=128, depth=1, num_heads=2,
                              num_classes=10, tuning_config=tuning, select_config=select).eval().to(dev)
        with torch.no_grad():
            torch.nn.init.normal_(m.patch_embed.proj.weight, std=0.05 * seed)
            torch.nn.init.normal_(m.cls_token, std=0.1 * seed)
            got = m._embed(img)
            assert _rel(got, stem_ref(m, img)) <= 2e-3, seed
            logits0 = m(img)
            # a write that bumps no version counter ...
            m.patch_embed.proj.weight.data.mul_(2.0)
            m.head.weight.data.mul_(0.5)
            dyt_b200.invalidate_caches(m)
            assert _rel(m._embed(img), stem_ref(m, img)) <= 2e-3, seed
            assert not torch.equal(m(img), logits0)
        del m
