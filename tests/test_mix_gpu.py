"""GPU parity of the mixed fp16 + fp8 operand format of the frozen teacher's Linears (qv_gemm_args.mix, qv_split_planes_mix and
the producers that write it: GEMM plane-output epilogue, qv_resid_ln_fwd, qv_attn_fwd) against fp64 torch on the true fp32
inputs -- the op being replaced is the teacher's nn.Linear / LayerNorm / scaled_dot_product_attention (ref
qat_trainer.py:337-338).  A product is fp16.fp16 + hi8.lo8 + lo8.hi8 with fp32 accumulation: ~2^-16 per element product.
Tolerance 1e-4 relative to the largest output (north_star: 1e-3 on fp32 logits)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def decode_mix(planes, weight=False):
    """[2, rows, cols] 2-byte stack -> (value from fp16 + residual, value from the fp8 copy), both fp64."""
    rows, cols = planes.shape[1], planes.shape[2]
    s16, sh8, sl8 = 128.0, 128.0, 1.0
    h16 = planes[0].contiguous().view(torch.float16).double()
    b = planes[1].contiguous().view(torch.uint8).reshape(rows, cols // 64, 2, 64)
    h8 = b[:, :, 0, :].contiguous().view(torch.float8_e4m3fn if weight else torch.float8_e5m2).double().reshape(rows, cols)
    l8 = b[:, :, 1, :].contiguous().view(torch.float8_e5m2).double().reshape(rows, cols)
    return (h16 + l8 / sl8) / s16, h8 / sh8


@pytest.mark.parametrize("weight", [False, True])
@pytest.mark.parametrize("rows,cols,mag", [(197 * 4, 768, 1.0), (33, 64, 30.0), (5, 3072, 0.02), (130, 128, 1e-3), (64, 64, 100.0)])
def test_split_planes_mix_roundtrip(cuda_dev, rows, cols, mag, weight):
    from qatvit_b200 import ops
    g = torch.Generator().manual_seed(rows + cols)
    x = (torch.randn(rows, cols, generator=g) * mag).to(cuda_dev)
    p = ops.split_planes_mix(x, weight=weight)
    torch.cuda.synchronize()
    rec, h8 = decode_mix(p, weight)
    xd = x.double()
    # fp16 (11 bits) + e5m2 residual (3 more): 2^-14 relative for every element whose residual is a normal e5m2 number; smaller
    # elements keep (at least) the fp16 value, i.e. an absolute error far below that of the typical element
    floor = 2.0 ** -9            # x * 128 = 0.25: fp16 ulp 2^-12, residual <= 2^-13 -- a normal e5m2 number (>= 2^-14)
    assert float(((rec - xd).abs() / xd.abs().clamp_min(floor)).max()) < 2.0 ** -13
    assert float(((rec - xd).abs() - 2.0 ** -11 * xd.abs()).max()) <= 2.0 ** -25 / 128.0   # never worse than fp16
    big = xd.abs() > (2.0 ** -13 if weight else 2.0 ** -21)     # fp8 copy of x * 128 is a normal number
    if weight:
        big &= xd.abs() <= 3.5           # e4m3(w * 128) saturates at 448 (the cross term then under-corrects; no ViT weight is that big)
    if bool(big.any()):
        assert float(((h8 - xd).abs() / xd.abs().clamp_min(1e-30))[big].max()) <= (2.0 ** -4 if weight else 2.0 ** -3) * 1.001


def test_split_planes_mix_rejects_bad_shapes(cuda_dev):
    from qatvit_b200 import ops
    with pytest.raises(RuntimeError, match="multiple of 64"):
        ops.split_planes_mix(torch.randn(8, 96, device=cuda_dev))


SHAPES = [(197 * 8, 2304, 768), (197 * 8, 768, 3072), (197 * 2, 3072, 768), (197 * 8, 768, 768), (130, 128, 64), (1, 64, 64),
          (300, 192, 128), (333, 64, 128), (200, 96, 192)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_mix_forward(cuda_dev, M, N, K):
    from qatvit_b200 import ops
    from qatvit_b200.ops import Op, PAIRS_FP32
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, generator=g) * 1.5).to(cuda_dev)
    w = (torch.randn(N, K, generator=g) * 0.03).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev) * 0.2
    ref = a.double() @ w.double().t() + bias.double()[None]
    out = ops.gemm(Op.full(ops.split_planes_mix(a)), Op.full(ops.split_planes_mix(w, weight=True)), M, N, K, PAIRS_FP32, bias=bias,
                   mix=True)
    torch.cuda.synchronize()
    assert out.shape == (M, N) and _rel(out, ref) < 1e-4
    # the three-pass bf16 product of the same operands: the mixed one stays within one order of magnitude of it
    out3 = ops.gemm(Op.full(ops.split_planes(a)), Op.full(ops.split_planes(w)), M, N, K, PAIRS_FP32, bias=bias)
    torch.cuda.synchronize()
    assert _rel(out, ref) < max(10 * _rel(out3, ref), 3e-5)


@pytest.mark.parametrize("M,N,K", [(197 * 8, 3072, 768), (197 * 4, 768, 768), (70, 128, 64), (333, 192, 192)])
@pytest.mark.parametrize("gelu", [False, True])
@pytest.mark.parametrize("out_mix", [False, True])
def test_gemm_mix_plane_output(cuda_dev, M, N, K, gelu, out_mix):
    """fc1 (+GELU) -> fc2 operand in the mixed format; qkv -> attention operand as bf16 hi/lo planes."""
    from qatvit_b200 import ops
    from qatvit_b200.ops import Op, PAIRS_FP32
    g = torch.Generator().manual_seed(M + N + K + int(gelu))
    a = (torch.randn(M, K, generator=g) * 1.2).to(cuda_dev)
    w = (torch.randn(N, K, generator=g) * 0.04).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev) * 0.3
    ref = a.double() @ w.double().t() + bias.double()[None]
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    planes = torch.full((2, M, N), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    ops.gemm(Op.full(ops.split_planes_mix(a)), Op.full(ops.split_planes_mix(w, weight=True)), M, N, K, PAIRS_FP32, bias=bias,
             out_planes=planes, gelu=gelu, mix=True, out_mix=out_mix)
    torch.cuda.synchronize()
    if out_mix:
        rec, h8 = decode_mix(planes)
        assert _rel(rec, ref) < 1e-4
        assert float((h8.cpu() - ref.cpu()).abs().max() / ref.abs().max()) < 2.0 ** -3
        # bit-identical to splitting the fp32 result of the same GEMM (same epilogue expression)
        y = ops.gemm(Op.full(ops.split_planes_mix(a)), Op.full(ops.split_planes_mix(w, weight=True)), M, N, K, PAIRS_FP32, bias=bias,
                     mix=True)
        if not gelu:
            assert torch.equal(planes.view(torch.int16), ops.split_planes_mix(y).view(torch.int16))
    else:
        assert _rel(planes[0].double() + planes[1].double(), ref) < 1e-4


@pytest.mark.parametrize("R,D", [(197 * 4, 768), (37, 128), (300, 384)])
def test_resid_ln_fwd_mixed_planes(cuda_dev, R, D):
    from qatvit_b200 import ops
    g = torch.Generator().manual_seed(R + D)
    x = torch.randn(R, D, generator=g).to(cuda_dev)
    y = torch.randn(R, D, generator=g).to(cuda_dev) * 0.5
    gamma = (torch.rand(D, generator=g) + 0.5).to(cuda_dev)
    beta = torch.randn(D, generator=g).to(cuda_dev) * 0.1
    hp = torch.full((2, R, D), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    hf = torch.empty(R, D, device=cuda_dev)
    xo = torch.empty(R, D, device=cuda_dev)
    ops.resid_ln_fwd(x, y, None, gamma, beta, 1e-6, R, D, x_out=xo, h_planes=hp, h_f32=hf, planes_mix=True)
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm((x + y).double(), (D,), gamma.double(), beta.double(), 1e-6)
    assert _rel(hf, ref) < 1e-5 and torch.equal(xo, x + y)
    assert torch.equal(hp.view(torch.int16), ops.split_planes_mix(hf).view(torch.int16))      # same split of the same fp32 values
    assert _rel(decode_mix(hp)[0], ref) < 1e-4


@pytest.mark.parametrize("B,H,T", [(3, 12, 197), (2, 2, 37), (1, 1, 128)])
def test_attn_fwd_mixed_output(cuda_dev, B, H, T):
    from qatvit_b200 import ops
    g = torch.Generator().manual_seed(B * 100 + H * 10 + T)
    D = H * 64
    qkv = (torch.randn(B * T, 3 * D, generator=g) * 1.5).to(cuda_dev)
    planes = ops.split_planes(qkv)
    out_b = torch.empty(2, B * T, D, dtype=torch.bfloat16, device=cuda_dev)
    out_m = torch.full((2, B * T, D), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    out_f = torch.empty(B * T, D, device=cuda_dev)
    ops.attn_fwd(planes, B, T, H, 0.125, out_b)
    ops.attn_fwd(planes, B, T, H, 0.125, out_m, out_f32=out_f, out_mix=True)
    torch.cuda.synchronize()
    x = qkv.double().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax((x[0] @ x[1].transpose(-1, -2)) * 0.125, dim=-1) @ x[2]).permute(0, 2, 1, 3).reshape(B * T, D)
    assert _rel(decode_mix(out_m)[0], ref) < 1e-4
    assert _rel(out_b[0].double() + out_b[1].double(), ref) < 1e-4
    assert torch.equal(out_m.view(torch.int16), ops.split_planes_mix(out_f).view(torch.int16))


@pytest.mark.parametrize("img,B", [(64, 4), (96, 3)])
def test_teacher_engine_mixed_vs_three_pass(cuda_dev, img, B):
    """Whole frozen-teacher forward: mixed-format Linears vs the bf16 three-pass ones vs fp64 torch."""
    import copy
    from parity_utils import build_models
    from qatvit_b200.engine import TeacherEngine
    vr, _, teacher = build_models("fbgemm", "vit_test_tiny", "vit_test_teacher", img)
    images, _ = vr.synthetic_batch(B, seed=3, img=img)
    t_gpu = copy.deepcopy(teacher).to(cuda_dev)
    with torch.no_grad():
        ref = copy.deepcopy(teacher).double()(images.double())
    mixed = TeacherEngine(t_gpu, B, mixed=True)
    assert mixed.mixed
    lm = mixed.forward(images.to(cuda_dev)).clone()
    l3 = TeacherEngine(t_gpu, B, mixed=False).forward(images.to(cuda_dev)).clone()
    torch.cuda.synchronize()
    assert _rel(lm, ref) < 1e-4 and _rel(l3, ref) < 1e-4
