"""The backward prologue of a fake-quant Linear (qv_gp_planes: gradient x STE mask of the output fake-quant [x gelu'] x
per-channel weight scale -> bf16 hi/lo planes + bias-grad partial sums; the ATen nodes FusedMovingAvgObsFqHelperBackward0 /
GeluBackward0 / the bias reduction of AddmmBackward0) fused into the kernels that PRODUCE the gradient:
the attention backward's output stage (qv_attn_bwd_gp), the LayerNorm backward (qv_ln_bwd_gp) and the dgrad GEMM epilogue
(act = 2, tests/test_gemm_gpu.py).  Each fused form must reproduce the unfused chain: planes bit for bit (same expression on
the same fp32 values), bias sums to fp32 summation-order noise; and the engine must produce the same gradients either way."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("B,H,T", [(3, 6, 197), (4, 2, 17), (2, 3, 129), (150, 2, 33), (90, 6, 197), (2, 2, 224), (1, 2, 64), (2, 1, 193)])
def test_attn_bwd_gp_matches_unfused(cuda_dev, B, H, T):
    from qatvit_b200 import ops
    dev = cuda_dev
    g = torch.Generator().manual_seed(11 * B + H + T)
    D = H * 64
    sval = 0.0437
    zpv = 60
    y_raw = ((torch.randint(-70, 78, (B * T, 3 * D), generator=g).float() + 0.3 * torch.randn(B * T, 3 * D, generator=g)) * sval).to(dev)
    s = torch.tensor([sval], device=dev)
    zp = torch.tensor([zpv], dtype=torch.int32, device=dev)
    fq = (s, zp, 0, 127)                                   # codes -60 .. 67, a few per cent of y_raw clipped
    cp = torch.empty(1, B * T, 3 * D, dtype=torch.bfloat16, device=dev)
    ops.act_planes(y_raw, fq, False, cp, codes_only=True)
    out = torch.empty(2, B * T, D, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B * H * T, device=dev)
    ops.attn_fwd(cp, B, T, H, 0.125, out, qk_scale=s, v_scale=s, lse=lse)
    dOp = ops.split_planes(torch.randn(B * T, D, generator=g).to(dev))
    wsc = (torch.rand(3 * D, generator=g) * 0.02 + 0.001).to(dev)
    # unfused: fp32 dQ|dK|dV -> gp_planes -> colsum_reduce
    g_qkv = torch.empty(B * T, 3 * D, device=dev)
    ops.attn_bwd(cp, s, out, dOp, lse, B, T, H, 0.125, g_qkv)
    rpb = 64
    nblk = -(-(B * T) // rpb)
    part = torch.empty(nblk, 3 * D, device=dev)
    ref_planes = torch.empty(2, B * T, 3 * D, dtype=torch.bfloat16, device=dev)
    ops.gp_planes(g_qkv, y_raw, fq, wsc, True, False, B * T, 3 * D, ref_planes, part, rpb)
    ref_bias = torch.empty(3 * D, device=dev)
    ops.colsum_reduce(part, nblk, 3 * D, ref_bias)
    # fused
    planes = torch.full((2, B * T, 3 * D), float("nan"), dtype=torch.bfloat16, device=dev)
    nslab = B * (-(-T // 128)) * 4
    slab = torch.full((nslab, 3 * D), float("nan"), device=dev)
    for _ in range(2):
        assert ops.attn_bwd_gp(cp, s, out, dOp, lse, B, T, H, 0.125, y_raw, fq, wsc, planes, slab) == nslab
    bias = torch.empty(3 * D, device=dev)
    ops.colsum_reduce(slab, nslab, 3 * D, bias)
    torch.cuda.synchronize()
    assert torch.equal(planes.view(torch.int16), ref_planes.view(torch.int16))
    masked = (ref_planes.float().abs().sum(0) == 0).float().mean()
    assert 0.005 < float(masked) < 0.5
    assert _rel(bias, ref_bias) < 1e-5


@pytest.mark.parametrize("R,D,with_res", [(197 * 4, 384, True), (37 * 3, 128, True), (197 * 2 + 3, 384, False), (70, 768, True)])
def test_ln_bwd_gp_matches_unfused(cuda_dev, R, D, with_res):
    from qatvit_b200 import ops
    dev = cuda_dev
    g = torch.Generator().manual_seed(R + D)
    x = torch.randn(R, D, generator=g).to(dev)
    g_h = torch.randn(R, D, generator=g).to(dev)
    g_res = torch.randn(R, D, generator=g).to(dev) if with_res else None
    gamma = (1.0 + 0.1 * torch.randn(D, generator=g)).to(dev)
    mean = x.mean(1).contiguous()
    rstd = (x.var(1, unbiased=False) + 1e-6).rsqrt().contiguous()
    y_raw = (torch.randn(R, D, generator=g) * 2.0).to(dev)
    fq = (torch.tensor([4.0 / 255], device=dev), torch.tensor([128], dtype=torch.int32, device=dev), 0, 255)
    wsc = (torch.rand(D, generator=g) * 0.02 + 0.001).to(dev)
    rpb = 64
    nblk = -(-R // rpb)
    # unfused
    gx_ref = torch.empty(R, D, device=dev)
    lnp_ref = torch.empty(nblk, 2, D, device=dev)
    ops.ln_bwd(g_h, x, mean, rstd, gamma, g_res, R, D, gx_ref, lnp_ref, rpb)
    part = torch.empty(nblk, D, device=dev)
    ref_planes = torch.empty(2, R, D, dtype=torch.bfloat16, device=dev)
    ops.gp_planes(gx_ref, y_raw, fq, wsc, True, False, R, D, ref_planes, part, rpb)
    ref_bias = torch.empty(D, device=dev)
    ops.colsum_reduce(part, nblk, D, ref_bias)
    # fused
    gx = torch.empty(R, D, device=dev)
    lnp = torch.empty(nblk, 2, D, device=dev)
    planes = torch.full((2, R, D), float("nan"), dtype=torch.bfloat16, device=dev)
    part2 = torch.full((nblk, D), float("nan"), device=dev)
    ops.ln_bwd(g_h, x, mean, rstd, gamma, g_res, R, D, gx, lnp, rpb, gp=(y_raw, fq, wsc, planes, part2))
    bias = torch.empty(D, device=dev)
    ops.colsum_reduce(part2, nblk, D, bias)
    torch.cuda.synchronize()
    assert torch.equal(gx, gx_ref) and torch.equal(lnp, lnp_ref)
    assert torch.equal(planes.view(torch.int16), ref_planes.view(torch.int16))
    assert _rel(bias, ref_bias) < 1e-5
    # and against autograd of LayerNorm in fp64
    xd = x.double().cpu().requires_grad_(True)
    h = torch.nn.functional.layer_norm(xd, (D,), gamma.double().cpu(), None, 1e-6)
    h.backward(g_h.double().cpu())
    want = xd.grad + (g_res.double().cpu() if with_res else 0)
    assert _rel(gx, want) < 1e-4


@pytest.mark.parametrize("backend,img,B,fused_attn,ln_variant", [("fbgemm", 64, 4, True, "subclass"), ("qnnpack", 96, 3, True, "plain"),
                                                                 ("fbgemm", 96, 3, False, "subclass")])
def test_engine_fused_gp_equals_standalone_gp(cuda_dev, backend, img, B, fused_attn, ln_variant):
    """Same model, same batch, two steps: gradients with the prologue fused into its producers == with the standalone
    kernel (weights: identical planes feed identical GEMMs -> bit-equal; biases: summation order only)."""
    from parity_utils import build_models
    from qatvit_b200.engine import QATDistillStep
    vr, prepared, teacher = build_models(backend, "vit_test_tiny", "vit_test_teacher", img, ln_variant=ln_variant)
    hp = dict(vr.DEFAULT_HPARAMS)
    grads = []
    for fused_gp in (True, False):
        student = copy.deepcopy(prepared).to(cuda_dev)
        step = QATDistillStep(student, copy.deepcopy(teacher).to(cuda_dev), B, hp, fused_attention=fused_attn, fused_gp=fused_gp)
        assert step.student_engine.fused_gp == fused_gp
        for it in range(2):
            images, labels = vr.synthetic_batch(B, seed=3 + it, img=img)
            out3 = step(images.to(cuda_dev), labels.to(cuda_dev))
        torch.cuda.synchronize()
        grads.append(({n: p.grad.clone() for n, p in student.named_parameters()}, out3.clone()))
    (ga, la), (gb, lb) = grads
    assert torch.equal(la, lb)
    for n in ga:
        if n.endswith("bias") and ga[n].dim() == 1 and "norm" not in n:
            assert _rel(ga[n], gb[n]) < 1e-5, n
        else:
            assert torch.equal(ga[n], gb[n]), n


def test_attn_bwd_gp_is_race_free(cuda_dev):
    """The output warps stage the y tile in and the gradient tile out through the SAME shared-memory buffers (TMA both ways) and
    drain one accumulator set while the tensor pipe fills the other: 60 launches of the bench-shaped problem (8 images x 6 heads x
    197 tokens) must be bit-identical, planes and column sums."""
    from qatvit_b200 import ops
    dev = cuda_dev
    B, H, T = 8, 6, 197
    D = H * 64
    g = torch.Generator().manual_seed(99)
    sval = 0.0437
    y_raw = ((torch.randint(-70, 78, (B * T, 3 * D), generator=g).float() + 0.3 * torch.randn(B * T, 3 * D, generator=g)) * sval).to(dev)
    s = torch.tensor([sval], device=dev)
    fq = (s, torch.tensor([60], dtype=torch.int32, device=dev), 0, 127)
    cp = torch.empty(1, B * T, 3 * D, dtype=torch.bfloat16, device=dev)
    ops.act_planes(y_raw, fq, False, cp, codes_only=True)
    out = torch.empty(2, B * T, D, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B * H * T, device=dev)
    ops.attn_fwd(cp, B, T, H, 0.125, out, qk_scale=s, v_scale=s, lse=lse)
    dOp = ops.split_planes(torch.randn(B * T, D, generator=g).to(dev))
    wsc = (torch.rand(3 * D, generator=g) * 0.02 + 0.001).to(dev)
    nslab = B * 2 * 4
    runs = []
    for _ in range(60):
        planes = torch.full((2, B * T, 3 * D), float("nan"), dtype=torch.bfloat16, device=dev)
        slab = torch.full((nslab, 3 * D), float("nan"), device=dev)
        ops.attn_bwd_gp(cp, s, out, dOp, lse, B, T, H, 0.125, y_raw, fq, wsc, planes, slab)
        runs.append((planes, slab))
    torch.cuda.synchronize()
    p0, s0 = runs[0]
    assert torch.isfinite(p0.float()).all() and torch.isfinite(s0).all()
    bad = sum(0 if (torch.equal(p.view(torch.int16), p0.view(torch.int16)) and torch.equal(sl, s0)) else 1 for p, sl in runs[1:])
    assert bad == 0, f"{bad} of 59 launches differ from the first"
