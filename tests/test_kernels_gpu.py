"""GPU parity tests, kernel by kernel: every hand-written sm_100a kernel is called through the
C ABI (dyt_b200.ops -> libdyt_b200.so) and compared with the CPU oracle (oracle/dyt_oracle.py,
"amp16" policy = the fp16-autocast arithmetic) on the same seeded inputs.
Bar: masks / indices bit-exact; floating point within the tolerance stated in each test."""
import pytest
import torch

import dyt_oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from dyt_b200 import lib
    lib()
    return torch.device("cuda:0")


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _close(got, ref, atol, rtol):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    assert not bool(bad.any()), f"max err {err.max().item():.3e}, {int(bad.sum())} elements out of tolerance"


# ---------------------------------------------------------------------------------------------
# GEMM + epilogues (nn.Linear under fp16 autocast)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(1, 8, 8), (197, 64, 768), (394, 2304, 768), (300, 768, 3072),
                                   (130, 1000, 200), (257, 768, 64), (64, 16, 768)])
@pytest.mark.parametrize("epi", ["bias", "gelu", "relu"])
def test_linear_epilogues(dev, M, N, K, epi):
    from dyt_b200 import ops, _lib
    g = _gen(M * 7 + N)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) * 0.5
    ref = O.linear(x, w, b, "amp16")
    if epi == "gelu":
        ref = O.gelu(ref, "amp16")
    elif epi == "relu":
        ref = torch.relu(ref)
    code = dict(bias=_lib.EPI_BIAS, gelu=_lib.EPI_BIAS_GELU, relu=_lib.EPI_BIAS_RELU)[epi]
    out, _ = ops.linear_f16(x.half().to(dev), w.half().to(dev), b.half().to(dev), epilogue=code)
    assert out.dtype == torch.float16 and out.shape == (M, N)
    # fp32 accumulation order differs from the CPU GEMM: allow one fp16 ulp
    _close(out, ref, atol=1e-3, rtol=2e-3)


def test_linear_residual_and_device_row_count(dev):
    from dyt_b200 import ops, _lib
    g = _gen(5)
    M, N, K = 500, 768, 768
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g) * 3
    v = O.linear(x, w, b, "amp16")
    for scale in (1.0, 0.1):
        vv = v if scale == 1.0 else O._r16(v * scale)
        out, out_h = ops.linear_f16(x.half().to(dev), w.half().to(dev), b.half().to(dev),
                                    epilogue=_lib.EPI_BIAS_RESID, resid=res.to(dev), scale=scale,
                                    want_f16_copy=True)
        assert out.dtype == torch.float32
        # one fp16 ulp of the GEMM term v (the residual add itself is exact fp32 arithmetic)
        _close(out.cpu() - res, vv, atol=1.5e-3, rtol=2e-3)
        assert torch.equal(out_h, out.half())
    # device-resident row count: rows >= m_dev are left untouched
    m_dev = torch.tensor([123], dtype=torch.int32, device=dev)
    buf = torch.full((M, N), -7.0, dtype=torch.float16, device=dev)
    ops.linear_f16(x.half().to(dev), w.half().to(dev), b.half().to(dev), m_dev=m_dev, out=buf)
    _close(buf[:123], v[:123], atol=1e-3, rtol=2e-3)
    assert bool((buf[123:] == -7.0).all())


def test_gelu_epilogue_exhaustive_fp16(dev):
    """Every finite fp16 value through the fc1 epilogue (identity weight, zero bias): the Gaussian-tail
    polynomial GELU must land on the correctly rounded fp16 GELU (float64 erf) or its fp16 neighbour.
    (torch's own fp32 formula 0.5*x*(1+erf(x/sqrt2)) cancels catastrophically for x < -3 and is
    itself several fp16 ulps off there, so it is only checked with an absolute bound.)"""
    from dyt_b200 import ops, _lib
    bits = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(torch.float16)
    x = bits[torch.isfinite(bits)]
    pad = (-x.numel()) % 8
    x = torch.cat([x, torch.zeros(pad, dtype=torch.float16)]).reshape(-1, 8)
    w = torch.eye(8, dtype=torch.float16)
    out, _ = ops.linear_f16(x.to(dev), w.to(dev), torch.zeros(8, dtype=torch.float16, device=dev),
                            epilogue=_lib.EPI_BIAS_GELU)
    ref = torch.nn.functional.gelu(x.double()).half()
    got = out.cpu()
    ref32 = torch.nn.functional.gelu(x.float()).half()        # what autocast does to nn.GELU
    fin = torch.isfinite(ref32.float()) & torch.isfinite(got.float())
    err32 = (got.float() - ref32.float()).abs()[fin]
    tol32 = 1.5e-6 + 2e-3 * ref32.float().abs()[fin]
    assert bool((err32 <= tol32).all()), "GELU differs from torch's fp32 GELU beyond its own noise"
    d = (got.view(torch.int16).int() - ref.view(torch.int16).int()).abs()
    d = torch.where((got == 0) & (ref == 0), torch.zeros_like(d), d)   # +0 / -0
    assert int(d.max()) <= 1, f"GELU off by {int(d.max())} fp16 ulps"
    assert int((d > 0).sum()) <= 400, f"{int((d > 0).sum())} of {x.numel()} values differ"


@pytest.mark.parametrize("M,N,K", [(129, 256, 64), (255, 512, 128), (1000, 768, 768), (257, 2304, 64),
                                   (12608, 768, 3072), (5000, 2304, 768), (12608, 3072, 768)])
def test_linear_cta_pair_edges(dev, M, N, K):
    """Tile rows that are odd in count (the CTA pair's second tile is a dummy), rows past M, the
    short-K shapes that take the sixteen-warp epilogue, and the shapes whose wave count picks the
    192-wide tile (N % 192 == 0, row count known on the host)."""
    from dyt_b200 import ops, _lib
    g = _gen(M + N + K)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    ref = O.linear(x, w, b, "amp16")
    out, _ = ops.linear_f16(x.half().to(dev), w.half().to(dev), b.half().to(dev))
    _close(out, ref, atol=1e-3, rtol=2e-3)
    out2, _ = ops.linear_f16(x.half().to(dev), w.half().to(dev), b.half().to(dev), scale=0.1)
    _close(out2, O._r16(ref * 0.1), atol=1e-3, rtol=3e-3)


def test_linear_rejects_bad_arguments(dev):
    from dyt_b200 import ops, DytError
    with pytest.raises(DytError):
        ops.linear_f16(torch.zeros(4, 12, dtype=torch.float16, device=dev),
                       torch.zeros(8, 12, dtype=torch.float16, device=dev))       # K % 8 != 0
    with pytest.raises(DytError):
        ops.linear_f16(torch.zeros(4, 16), torch.zeros(8, 16))                      # CPU tensors


# ---------------------------------------------------------------------------------------------
# LayerNorm
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C", [128, 768, 1024])
def test_layernorm(dev, C):
    from dyt_b200 import ops
    g = _gen(C)
    x = torch.randn(3, 197, C, generator=g) * 2 + 0.3
    w = 1 + 0.1 * torch.randn(C, generator=g)
    b = 0.1 * torch.randn(C, generator=g)
    ref = O._r16(O.layer_norm(x, w, b, 1e-6))
    out = ops.layernorm_f16(x.to(dev), w.to(dev), b.to(dev), 1e-6)
    _close(out, ref, atol=1e-3, rtol=1e-3)      # one fp16 ulp
    idx = torch.tensor([5, 0, 590, 17, 17], dtype=torch.int32)
    out = ops.layernorm_f16(x.to(dev), w.to(dev), b.to(dev), 1e-6, row_idx=idx.to(dev))
    _close(out, ref.reshape(-1, C)[idx.long()], atol=1e-3, rtol=1e-3)


# ---------------------------------------------------------------------------------------------
# attention (TMA + tcgen05, P in TMEM)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,N", [(1, 1, 16), (2, 12, 197), (3, 2, 128), (2, 16, 256), (4, 3, 1)])
def test_attention_uniform(dev, B, H, N):
    from dyt_b200 import ops
    g = _gen(B * 100 + N)
    qkv = torch.randn(B, N, 3 * H * 64, generator=g) * 1.5
    ref = O.attention_core(O._r16(qkv), H, "amp16")
    out = ops.attn_varlen(qkv.half().to(dev), H)
    assert out.shape == (B, N, H * 64)
    _close(out, ref, atol=2e-3, rtol=3e-3)


def test_attention_varlen(dev):
    from dyt_b200 import ops
    H = 4
    lens = [1, 17, 128, 129, 197, 256, 64]
    cu = torch.tensor([0] + lens).cumsum(0).to(torch.int32)
    T = int(cu[-1])
    qkv = torch.randn(T, 3 * H * 64, generator=_gen(9))
    out = ops.attn_varlen(qkv.half().to(dev), H, cu_seqlens=cu.to(dev), max_seqlen=max(lens))
    for i, n in enumerate(lens):
        s = int(cu[i])
        ref = O.attention_core(O._r16(qkv[s:s + n]).unsqueeze(0), H, "amp16")[0]
        _close(out[s:s + n], ref, atol=2e-3, rtol=3e-3)


def test_attention_unsupported_is_loud(dev):
    from dyt_b200 import ops, DytError
    with pytest.raises(DytError):
        ops.attn_varlen(torch.zeros(1, 300, 3 * 64, dtype=torch.float16, device=dev), 1)   # > 256 keys
    with pytest.raises(DytError):
        ops.attn_varlen(torch.zeros(1, 16, 3 * 32, dtype=torch.float16, device=dev), 1)    # head_dim 32


# ---------------------------------------------------------------------------------------------
# fused dispatcher: score + gate + compaction + LN2 pack
# ---------------------------------------------------------------------------------------------
def _dispatch_case(dev, B, N, C, seed, forced=None, policy="amp16"):
    from dyt_b200 import ops
    g = _gen(seed)
    x1 = torch.randn(B, N, C, generator=g)
    w = torch.randn(1, C, generator=g) * 0.5
    b = torch.randn(1, generator=g) * 0.1
    lw = 1 + 0.1 * torch.randn(C, generator=g)
    lb = 0.1 * torch.randn(C, generator=g)
    r = ops.dispatch(x1.to(dev), w.to(dev), b.to(dev), ln_w=lw.to(dev), ln_b=lb.to(dev),
                     logit_dtype=torch.float16 if policy == "amp16" else torch.float32,
                     forced_mask=None if forced is None else forced.to(dev))
    torch.cuda.synchronize()
    mask_ref, logit_ref = O.token_select(x1, w, b, policy)
    return x1, (lw, lb), r, mask_ref, logit_ref


@pytest.mark.parametrize("B,N,C", [(2, 197, 768), (5, 197, 1024), (3, 5, 128), (1, 2, 128),
                                   (300, 50, 128), (2, 600, 384), (700, 5, 128)])
@pytest.mark.parametrize("policy", ["amp16", "fp32"])
def test_dispatch_matches_oracle(dev, B, N, C, policy):
    x1, (lw, lb), r, mask_ref, logit_ref = _dispatch_case(dev, B, N, C, B * N + C, policy=policy)
    logits = r["logits"].cpu()
    mask = r["mask"].cpu()
    if policy == "amp16":
        # the logit is rounded once to fp16: kernel and oracle may differ by accumulation order
        # only for values within half an ulp of a rounding boundary -> allow one fp16 ulp
        _close(logits, logit_ref, atol=2e-5, rtol=1.1e-3)
    else:
        _close(logits, logit_ref, atol=1e-4, rtol=1e-5)
    # the gate itself is exact given the logits the kernel produced ...
    dt = torch.float16 if policy == "amp16" else torch.float32
    gate_of_kernel_logits = O.gumbel_sigmoid_hard(logits.to(dt)).float()
    assert torch.equal(mask[:, 1:], gate_of_kernel_logits)
    assert bool((mask[:, 0] == 1).all())
    # ... and bit-equal to the oracle's mask except where the oracle's logit sits on the threshold
    mism = mask != mask_ref
    if bool(mism.any()):
        thr = O.min_kept_logit(dt)
        slack = 1.2e-5 if policy == "amp16" else 2e-4
        near = (logit_ref[mism[:, 1:]] - thr).abs() <= slack
        assert bool(near.all()), "mask mismatch away from the gate threshold"
        assert int(mism.sum()) <= 2
    # compaction: nonzero() order, inverse map, cu_seqlens, count
    n_kept = int(r["n_kept"].item())
    idx_ref, cu_ref = O.compact(mask)
    assert n_kept == idx_ref.numel()
    assert torch.equal(r["packed_idx"][:n_kept].cpu().long(), idx_ref)
    assert torch.equal(r["cu_seqlens"].cpu(), cu_ref)
    pos = r["token_pos"].cpu().long()
    flat = mask.reshape(-1) != 0
    assert bool((pos[~flat] == -1).all())
    assert torch.equal(pos[flat], torch.arange(n_kept))
    # packed buffer = LayerNorm2 of the kept rows, fp16
    ref_rows = O._r16(O.layer_norm(x1.reshape(B * N, C)[idx_ref], lw, lb))
    _close(r["packed"][:n_kept], ref_rows, atol=1e-3, rtol=1e-3)


def test_dispatch_many_images_lookback(dev):
    """More images than CTAs can be co-resident (video batches: b*t frame-sequences): the per-image
    CTAs take tickets in launch order and chain their packed bases by decoupled look-back over
    several 32-image windows; bookkeeping must equal torch's nonzero() order, twice in a row on the
    same workspace (epoch reuse)."""
    from dyt_b200 import ops
    B, N, C = 1500, 9, 128
    g = _gen(77)
    x = torch.randn(B, N, C, generator=g)
    w = torch.randn(1, C, generator=g) * 0.3
    b = torch.zeros(1)
    lw, lb = torch.ones(C), torch.zeros(C)
    for _ in range(2):
        d = ops.dispatch(x.to(dev), w.to(dev), b.to(dev), ln_w=lw.to(dev), ln_b=lb.to(dev))
        mask = d["mask"][..., 0].cpu()
        nz = torch.nonzero(mask.reshape(-1))[:, 0].to(torch.int32)
        n_kept = int(d["n_kept"].cpu())
        assert n_kept == nz.numel() and 0 < n_kept < B * N
        assert torch.equal(d["packed_idx"].cpu()[:n_kept], nz)
        cu = torch.cat([torch.zeros(1), mask.sum(1).cumsum(0)]).to(torch.int32)
        assert torch.equal(d["cu_seqlens"].cpu(), cu)
        pos = d["token_pos"].cpu()
        assert torch.equal(pos[nz.long()], torch.arange(n_kept, dtype=torch.int32))
        assert bool((pos[mask.reshape(-1) == 0] == -1).all())
        ref = O._r16(O.layer_norm(x.reshape(-1, C)[nz.long()], lw, lb))
        _close(d["packed"].cpu()[:n_kept], ref, atol=2e-3, rtol=2e-3)


def test_dispatch_gate_exhaustive_fp16(dev):
    """Every fp16 bit pattern as a logit: x1 rows are one-hot * value, w = e_0, bias 0, so the
    kernel's logit IS the pattern; the mask must equal the reference gate's truth table (golden,
    from the reference's own _gumbel_sigmoid) and torch's CUDA sigmoid."""
    from dyt_b200 import ops
    gold = load_golden("gate_tables.pt")["fp16"].bool()
    C, N = 128, 129
    B = 65536 // (N - 1)
    vals = torch.arange(0, 1 << 16, dtype=torch.int32).to(torch.int16).view(torch.float16)
    x1 = torch.zeros(B, N, C)
    x1[:, 1:, 0] = vals.float().reshape(B, N - 1)
    w = torch.zeros(1, C)
    w[0, 0] = 1.0
    r = ops.dispatch(x1.to(dev), w.to(dev), torch.zeros(1, device=dev), pack=False)
    mask = r["mask"][:, 1:, 0].reshape(-1).cpu() != 0
    finite = torch.isfinite(vals)
    nan = torch.isnan(vals)
    assert torch.equal(mask[~nan], gold[~nan])
    assert not bool(mask[nan].any())              # NaN -> dropped
    cuda_gate = (vals.to(dev).sigmoid() > 0.5).cpu()
    assert torch.equal(mask[finite], cuda_gate[finite])
    lg = r["logits"].reshape(-1).cpu()
    assert torch.equal(lg[~nan], vals.float()[~nan])


@pytest.mark.parametrize("name", ["fp16", "fp32"])
def test_dispatch_train_mode_gumbel(dev, name):
    """Train-mode gate with the reference's own RNG draws (golden fixture)."""
    from dyt_b200 import ops
    gfix = load_golden("gumbel_train.pt")[name]
    logits = gfix["logits"].float()            # [4, 196, 1]
    B, N, C = 4, 197, 128
    x1 = torch.zeros(B, N, C)
    x1[:, 1:, 0] = logits[..., 0]
    w = torch.zeros(1, C)
    w[0, 0] = 1.0
    r = ops.dispatch(x1.to(dev), w.to(dev), torch.zeros(1, device=dev), pack=False,
                     logit_dtype=torch.float16 if name == "fp16" else torch.float32,
                     noise=(gfix["g1"].float().to(dev), gfix["g2"].float().to(dev)), tau=5.0)
    assert torch.equal(r["mask"][:, 1:].cpu() != 0, gfix["hard"].bool())


def test_dispatch_forced_mask_and_gate_report(dev):
    forced = O.checkerboard_mask(3, 197)
    x1, _, r, mask_ref, _ = _dispatch_case(dev, 3, 197, 768, 77, forced=forced)
    assert torch.equal(r["mask"].cpu(), forced)
    assert int(r["n_kept"].item()) == 3 * 99
    assert torch.equal(r["gate"].cpu(), mask_ref)      # selector's own decision still reported


def test_dispatch_workspace_is_reusable(dev):
    """The grid-barrier words must be left zeroed: run twice back to back with different B."""
    for B in (7, 300, 7):
        x1, _, r, mask_ref, _ = _dispatch_case(dev, B, 50, 128, 1000 + B)
        assert int(r["n_kept"].item()) == int(r["mask"].sum().item())


# ---------------------------------------------------------------------------------------------
# scatter-merge
# ---------------------------------------------------------------------------------------------
def test_scatter_merge(dev):
    from dyt_b200 import ops
    g = _gen(3)
    B, N, C = 3, 197, 768
    x1 = torch.randn(B, N, C, generator=g)
    adapt = (torch.randn(B, N, C, generator=g) * 0.1).half()
    mask = (torch.rand(B, N, 1, generator=g) > 0.5).float()
    idx, _ = O.compact(mask)
    mlp = torch.randn(idx.numel(), C, generator=g).half()
    pos = torch.full((B * N,), -1, dtype=torch.int32)
    pos[idx] = torch.arange(idx.numel(), dtype=torch.int32)
    full = torch.zeros(B * N, C)
    full[idx] = mlp.float()
    ref = adapt.float() + (x1 + full.reshape(B, N, C))          # model_speed_test.py:305-308
    lw, lb = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    out, ln = ops.scatter_merge(x1.to(dev), adapt.to(dev), mlp.to(dev), pos.to(dev),
                                next_ln=(lw.to(dev), lb.to(dev)))
    assert torch.equal(out.cpu(), ref)                          # same fp32 additions: bit-exact
    _close(ln, O._r16(O.layer_norm(ref, lw, lb)), atol=1e-3, rtol=1e-3)


@pytest.mark.parametrize("B,N,C,K,scale", [(3, 197, 768, 64, 0.1), (2, 50, 128, 8, 1.0),
                                           (5, 197, 1024, 64, 0.5), (1, 17, 384, 32, 0.1),
                                           (300, 197, 768, 64, 0.1)])
def test_merge_up_fused_matches_linear_plus_scatter_merge(dev, B, N, C, K, scale):
    """dyt_merge_up_fwd (adapter up-projection inside the merge kernel) against the two separate
    launches it replaces and against the oracle's arithmetic (model_speed_test.py:106-111, :302-308):
    the fp32 stream bit-equal, the fused LayerNorm within one fp16 ulp of the two-pass kernel."""
    from dyt_b200 import ops, _lib
    g = _gen(11 + B)
    x1 = torch.randn(B, N, C, generator=g) * 3 + 0.5
    down = torch.relu(torch.randn(B, N, K, generator=g)).half()
    up_w = (torch.randn(C, K, generator=g) * 0.05).half()
    up_b = (torch.randn(C, generator=g) * 0.1).half()
    mask = (torch.rand(B, N, 1, generator=g) > 0.5).float()
    idx, _ = O.compact(mask)
    mlp = torch.randn(idx.numel(), C, generator=g).half()
    pos = torch.full((B * N,), -1, dtype=torch.int32)
    pos[idx] = torch.arange(idx.numel(), dtype=torch.int32)
    lw, lb = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    d = lambda t: t.to(dev)
    adapt, _ = ops.linear_f16(d(down).reshape(B * N, K), d(up_w), d(up_b), scale=scale)
    out0, ln0 = ops.scatter_merge(d(x1), adapt.reshape(B, N, C), d(mlp), d(pos), next_ln=(d(lw), d(lb)))
    out1, ln1 = ops.merge_up(d(down), d(up_w), d(up_b), scale, d(x1), d(mlp), d(pos),
                             next_ln=(d(lw), d(lb)))
    out2, none = ops.merge_up(d(down), d(up_w), d(up_b), scale, d(x1), d(mlp), d(pos))
    assert none is None
    assert torch.equal(out1, out0) and torch.equal(out2, out0)
    # LayerNorm: same mean / variance up to fp32 rounding of a different summation order
    diff = (ln1.float() - ln0.float()).abs()
    ulp = torch.maximum(ln0.float().abs(), torch.tensor(2.0 ** -14, device=dev)) * 2.0 ** -10
    assert bool((diff <= ulp).all())
    assert float((diff > 0).float().mean()) < 0.01
    if B <= 5:   # oracle arithmetic on the CPU
        a_ref = O._r16(O._r16(down.float().reshape(-1, K) @ up_w.float().t() + up_b.float()) * scale)
        full = torch.zeros(B * N, C)
        full[idx] = mlp.float()
        ref = a_ref.reshape(B, N, C) + (x1 + full.reshape(B, N, C))
        _close(out1, ref, atol=2e-3, rtol=1e-3)
        _close(ln1, O._r16(O.layer_norm(ref, lw, lb)), atol=2e-3, rtol=2e-3)


@pytest.mark.parametrize("B,N,C,K,scale", [(3, 197, 768, 64, 0.1), (2, 50, 128, 8, 1.0),
                                           (5, 197, 1024, 64, 0.5), (1, 17, 384, 32, 0.1),
                                           (300, 197, 768, 64, 0.1), (7, 197, 768, 16, 1.0)])
def test_adapter_merge_fused_down_and_up(dev, B, N, C, K, scale):
    """dyt_adapter_merge_fwd (down projection + ReLU + up projection + merge + LayerNorm in one kernel,
    x1 read as fp32) against the separate launches: down GEMM on the fp16 copy of x1, then
    dyt_merge_up_fwd.  The down accumulation runs in another tile shape, so `down` may differ in a
    last fp16 bit on a few elements: outputs are compared to rounding noise, and against the oracle's
    arithmetic (model_speed_test.py:106-111, :302-308)."""
    from dyt_b200 import ops, _lib
    g = _gen(31 + B)
    x1 = torch.randn(B, N, C, generator=g) * 2 + 0.3
    down_w = (torch.randn(K, C, generator=g) * 0.03).half()
    down_b = (torch.randn(K, generator=g) * 0.1).half()
    up_w = (torch.randn(C, K, generator=g) * 0.05).half()
    up_b = (torch.randn(C, generator=g) * 0.1).half()
    mask = (torch.rand(B, N, 1, generator=g) > 0.5).float()
    idx, _ = O.compact(mask)
    mlp = torch.randn(idx.numel(), C, generator=g).half()
    pos = torch.full((B * N,), -1, dtype=torch.int32)
    pos[idx] = torch.arange(idx.numel(), dtype=torch.int32)
    lw, lb = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    d = lambda t: t.to(dev)
    x1h = d(x1).half()
    down, _ = ops.linear_f16(x1h.reshape(B * N, C), d(down_w), d(down_b), epilogue=_lib.EPI_BIAS_RELU)
    out0, ln0 = ops.merge_up(down.reshape(B, N, K), d(up_w), d(up_b), scale, d(x1), d(mlp), d(pos),
                             next_ln=(d(lw), d(lb)))
    out1, ln1 = ops.adapter_merge(d(down_w), d(down_b), d(up_w), d(up_b), scale, d(x1), d(mlp), d(pos),
                                  next_ln=(d(lw), d(lb)))
    out2, none = ops.adapter_merge(d(down_w), d(down_b), d(up_w), d(up_b), scale, d(x1), d(mlp), d(pos))
    assert none is None and torch.equal(out2, out1)
    _close(out1, out0.cpu(), atol=2e-3, rtol=1e-3)
    assert float((out1 != out0).float().mean()) < 0.02       # the same up to rare last-bit flips of `down`
    _close(ln1, ln0.float().cpu(), atol=4e-3, rtol=2e-3)
    if B <= 7:
        dn = torch.relu(O._r16(O._r16(x1).reshape(-1, C) @ down_w.float().t() + down_b.float()))
        a_ref = O._r16(O._r16(dn @ up_w.float().t() + up_b.float()) * scale)
        full = torch.zeros(B * N, C)
        full[idx] = mlp.float()
        ref = a_ref.reshape(B, N, C) + (x1 + full.reshape(B, N, C))
        _close(out1, ref, atol=3e-3, rtol=1e-3)


def test_merge_up_refuses_unsupported_shapes(dev):
    from dyt_b200 import ops, _lib
    x1 = torch.zeros(1, 4, 192, device=dev)
    with pytest.raises(_lib.DytError):
        ops.merge_up(torch.zeros(1, 4, 64, device=dev).half(), torch.zeros(192, 64, device=dev).half(),
                     None, 1.0, x1, torch.zeros(4, 192, device=dev).half(),
                     torch.zeros(4, dtype=torch.int32, device=dev))


@pytest.mark.parametrize("B,H,N", [(3, 12, 197), (2, 4, 161), (1, 2, 256), (40, 12, 197), (2, 3, 224)])   # 256: falls back to the two-stream kernel
def test_attention_four_stream_kernel(dev, B, H, N):
    """DYT_OPT_ATTN_SPLIT (default on): the four-stream kernel (query tile x key half, partial softmax per half,
    exact combine) against the oracle's attention and against the two-stream kernel."""
    from dyt_b200 import ops, _lib
    lib = _lib.lib()
    g = _gen(100 + N)
    qkv = (torch.randn(B, N, 3 * H * 64, generator=g) * 1.5).half()
    # one head with peaked scores late in the sequence: the halves' maxima differ widely
    qkv[:, N - 7, H * 64:H * 64 + 64] *= 6.0
    out = ops.attn_varlen(qkv.to(dev), H)            # the default for these lengths
    try:
        assert lib.dyt_configure(_lib.OPT_ATTN_SPLIT, 0) == 0
        base = ops.attn_varlen(qkv.to(dev), H)       # the two-stream kernel
    finally:
        assert lib.dyt_configure(_lib.OPT_ATTN_SPLIT, 1) == 0
    ref = O.attention_core(qkv.float(), H, "amp16")
    _close(out, ref, atol=2e-3, rtol=3e-3)
    # same rounding points as the two-stream kernel (row maximum shared by the key halves): only the
    # fp32 summation order of the PV product differs
    d = (out.float() - base.float()).abs()
    assert float((d > 0).float().mean()) < 0.05 and float(d.max()) <= 2e-3


def test_keep_stats_kernel_matches_reference_accounting():
    """dyt_keep_stats / dyt_b200.flops.batch_select_flops against the reference's
    block_flops_dict.batch_select_flops (golden, bit-equal: same fp32 additions in the same order) and
    the per-layer keep rates of engine_finetune.py:349-351; the drop-in module keeps the names."""
    from conftest import load_golden
    from dyt_b200 import flops
    import block_flops_dict as drop_in
    g = load_golden("flops_accounting.pt")
    dev = torch.device("cuda:0")
    ts = g["token_select"].float().to(dev)
    a = flops.batch_select_flops(37, g["table"], ts, block_num=12, base_flops=0.116)
    assert torch.equal(a.cpu(), g["flops_12"])
    b = drop_in.batch_select_flops(37, g["table"].to(dev), ts[:, 2:].half(), block_num=12, base_flops=0.25)
    assert torch.equal(b.cpu(), g["flops_10of12"])
    st = flops.KeepStats(12, 196, dev)
    st.update(ts[:20])
    st.update(ts[20:])
    assert int(st.counters[12]) == 37
    assert torch.equal(st.counters[:12].cpu(), g["token_select"].long().sum(dim=(0, 2, 3)))
    assert torch.allclose(st.layer_rates().cpu(), g["layer_rates"], atol=1e-6)
    assert drop_in.get_block_flops().shape == (198,) and abs(drop_in.get_base_flops() - 0.1157) < 2e-3
    with pytest.raises(Exception):
        flops.batch_select_flops(1, g["table"], g["token_select"].float())      # CPU tensor: no fallback


@pytest.mark.parametrize("B,N,H,with_bias", [(2, 1025, 12, True), (3, 197, 2, False), (1, 64, 1, True),
                                              (2, 65, 3, True), (4, 5, 2, True), (2, 257, 2, True),
                                              (1, 384, 1, False), (2, 513, 2, True), (1, 128, 1, True),
                                              (1, 129, 2, False), (20, 300, 4, True), (9, 1025, 12, False)])
def test_attention_with_bias_long_sequences(dev, B, N, H, with_bias):
    """dyt_attn_bias_fwd (any sequence length, additive per-head bias) against the oracle's amp16
    restatement of the segmentation backbone's eager attention; 1025 tokens = 512 x 512 images."""
    from dyt_b200 import ops
    g = _gen(B * N + H)
    C = 64 * H
    qkv = (torch.randn(B, N, 3 * C, generator=g) * 1.2).half()
    bias = (torch.randn(H, N, N, generator=g) * 1.5) if with_bias else None
    ref = O.attention_bias_core(qkv.float(), H, bias, "amp16")
    got = ops.attn_bias(qkv.to(dev), H, None if bias is None else bias.to(dev))
    assert got.shape == (B, N, C) and got.dtype == torch.float16
    _close(got, ref, atol=2e-3, rtol=4e-3)
    if with_bias:                        # rows padded to a multiple of 4 floats: the 16-byte load path
        padded = ops.pad_attn_bias(bias.to(dev))
        assert padded.stride(1) % 4 == 0 and padded.stride(1) % 64 != 0 and padded.stride(1) >= N
        assert torch.equal(padded.cpu(), bias)
        got_p = ops.attn_bias(qkv.to(dev), H, padded)
        assert torch.equal(got_p, got)
    if not with_bias and N <= 256:      # same answer as the tcgen05 kernel on its own domain
        _close(got, ops.attn_varlen(qkv.to(dev), H), atol=2e-3, rtol=4e-3)


def test_segmentation_attention_module_golden(dev):
    """The reference's segmentation Attention module (golden, fp32): qkv Linear -> attention with the
    gathered relative-position bias -> proj, on the kernels."""
    from conftest import load_golden
    from dyt_b200 import ops
    g = load_golden("seg_attention.pt")
    for tag, d in g.items():
        bias = O.relative_position_bias(d["table"], d["index"])
        qkv, _ = ops.linear_f16(d["x"].half().to(dev), d["qkv_w"].half().to(dev), d["qkv_b"].half().to(dev))
        o = ops.attn_bias(qkv, 2, bias.to(dev))
        y, _ = ops.linear_f16(o, d["proj_w"].half().to(dev), d["proj_b"].half().to(dev))
        ref = d["y"]
        assert ((y.float().cpu() - ref).abs().max() / ref.abs().max()).item() <= 1e-2, tag
