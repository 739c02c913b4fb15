"""GPU parity of the tcgen05 GEMM family (qv_gemm_bf16 through the C-ABI) against an fp64 torch matmul of the same
operands: the fake-quant Linear shapes of BASELINE.json configs[3] (tokens 197xB, dims 384/768/1536/3072), forward / dgrad
(K-major) and wgrad (MN-major, split-K), ragged edges, batched per-head addressing, fused epilogue terms (per-channel scale,
bias, observer min/max) and the plane-output epilogue (+GELU).  fp32 operands travel as bf16 hi/lo planes (3 products,
~2^-16 relative); integer-code operands are exact.  Tolerance 1e-4 relative to the largest output (north_star: 1e-3)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _planes(x):
    from qatvit_b200 import ops
    return ops.split_planes(x.contiguous())


def _codes(shape, dev, g):
    return torch.randint(-128, 128, shape, generator=g).float().to(dev)


SHAPES = [  # (M, N, K): student / teacher Linear shapes at B = 8, plus ragged ones
    (197 * 8, 1152, 384), (197 * 8, 384, 384), (197 * 8, 1536, 384), (197 * 8, 384, 1536),
    (197 * 8, 2304, 768), (197 * 8, 768, 3072), (197 * 2, 3072, 768), (130, 96, 40), (1, 32, 8), (300, 160, 72),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("pairs", ["fp32", "exact_b", "single"])
def test_gemm_forward_epilogue(cuda_dev, M, N, K, pairs):
    from qatvit_b200 import ops
    from qatvit_b200.ops import Op, PAIRS_EXACT_B, PAIRS_FP32, PAIRS_SINGLE
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    dev = cuda_dev
    if pairs == "single":
        a, b = _codes((M, K), dev, g), _codes((N, K), dev, g)
        ap, bp, pr = a.bfloat16()[None].contiguous(), b.bfloat16()[None].contiguous(), PAIRS_SINGLE
    elif pairs == "exact_b":
        a, b = torch.randn(M, K, generator=g).to(dev), _codes((N, K), dev, g)
        ap, bp, pr = _planes(a), b.bfloat16()[None].contiguous(), PAIRS_EXACT_B
    else:
        a, b = torch.randn(M, K, generator=g).to(dev), torch.randn(N, K, generator=g).to(dev)
        ap, bp, pr = _planes(a), _planes(b), PAIRS_FP32
    scale = (torch.rand(N, generator=g) * 0.02 + 0.001).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    alpha = torch.tensor([0.37], device=dev)
    acc = ops.new_minmax(dev)
    out = ops.gemm(Op.full(ap), Op.full(bp), M, N, K, pr, col_scale=scale, bias=bias, alpha=alpha, minmax=acc)
    ref = (a.double() @ b.double().t()) * scale.double() * 0.37 + bias.double()
    assert _rel(out, ref) < 1e-4
    mn = torch.full((1,), float("inf"), device=dev)
    mx = torch.full((1,), float("-inf"), device=dev)
    sc, zp = torch.ones(1, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
    on = torch.ones(1, dtype=torch.int64, device=dev)
    ops.obs_update(acc, on, on, mn, mx, sc, zp, 0.01, 0, 127, False)
    assert float(mn) == float(out.min()) and float(mx) == float(out.max())      # fused observer min/max is exact


@pytest.mark.parametrize("M,N,K", [(197 * 8, 2304, 768), (197 * 8, 3072, 768), (197 * 4, 768, 768), (70, 128, 64), (333, 192, 136)])
@pytest.mark.parametrize("gelu", [False, True])
def test_gemm_plane_output(cuda_dev, M, N, K, gelu):
    from qatvit_b200 import ops
    from qatvit_b200.ops import Op, PAIRS_FP32
    g = torch.Generator().manual_seed(M + N + K)
    a, b = torch.randn(M, K, generator=g).to(cuda_dev), (torch.randn(N, K, generator=g) * 0.05).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev)
    outp = torch.full((2, M, N), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    ops.gemm(Op.full(_planes(a)), Op.full(_planes(b)), M, N, K, PAIRS_FP32, bias=bias, out_planes=outp, gelu=gelu)
    ref = a.double() @ b.double().t() + bias.double()
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    got = outp[0].double() + outp[1].double()
    assert _rel(got, ref) < 1e-4
    # the planes are a hi/lo split: |lo| is at most half an ulp (2^-8 relative) of hi
    assert bool((outp[1].float().abs() <= outp[0].float().abs() * 2.0 ** -8 + 1e-38).all())


@pytest.mark.parametrize("N,K,Mtok", [(1152, 384, 197 * 8), (384, 1536, 197 * 8), (96, 40, 130), (384, 768, 196 * 4)])
def test_gemm_wgrad_splitk(cuda_dev, N, K, Mtok):
    """weight.grad[N,K] = mask * (gy^T @ x) / scale: MN-major operands, split-K over tokens, deterministic reduce."""
    from qatvit_b200 import ops
    from qatvit_b200.engine import wgrad_splits
    from qatvit_b200.ops import Op, PAIRS_FP32
    g = torch.Generator().manual_seed(N + K)
    gy, x = torch.randn(Mtok, N, generator=g).to(cuda_dev), torch.randn(Mtok, K, generator=g).to(cuda_dev)
    rscale = (torch.rand(N, generator=g) + 0.5).to(cuda_dev)
    mask = (torch.rand(N, K, generator=g) > 0.1).to(torch.uint8).to(cuda_dev)
    sms = torch.cuda.get_device_properties(cuda_dev).multi_processor_count
    s = wgrad_splits(N, K, Mtok, sms)
    gyp, xp = _planes(gy), _planes(x)
    out = torch.empty(N, K, device=cuda_dev)
    res = []
    for _ in range(2):
        if s > 1:
            ws = ops.gemm(Op.full(gyp, mn_major=True), Op.full(xp, mn_major=True), N, K, Mtok, PAIRS_FP32, splits=s)
        else:
            ws = ops.gemm(Op.full(gyp, mn_major=True), Op.full(xp, mn_major=True), N, K, Mtok, PAIRS_FP32)
        ops.splitk_reduce(ws, s, N, K, out, row_rscale=rscale, mask=mask)
        res.append(out.clone())
    ref = (gy.double().t() @ x.double()) / rscale.double()[:, None] * mask.double()
    assert _rel(out, ref) < 1e-4
    assert torch.equal(res[0], res[1])                     # bitwise reproducible


@pytest.mark.parametrize("B,H,T", [(3, 6, 197), (2, 2, 37), (1, 12, 197)])
def test_gemm_batched_per_head(cuda_dev, B, H, T):
    """softmax-attention matmuls addressed per (image, head) inside a [tokens, 3*D] tensor (TMA zero-fills ragged T)."""
    from qatvit_b200 import ops
    from qatvit_b200.ops import Op, Out, PAIRS_FP32
    D = H * 64
    g = torch.Generator().manual_seed(B * H * T)
    qkv = torch.randn(B * T, 3 * D, generator=g).to(cuda_dev)
    qkvp = _planes(qkv)
    ldS = -(-T // 4) * 4
    S = torch.zeros(B * H * T, ldS, device=cuda_dev)
    ops.gemm(Op.tokens(qkvp, B, T, 0, 64), Op.tokens(qkvp, B, T, D, 64), T, T, 64, PAIRS_FP32,
             out=Out.per_head(S, B * H, H, T, T), nbatch=B * H, batch_inner=H)
    q = qkv[:, :D].view(B, T, H, 64).permute(0, 2, 1, 3).double()
    k = qkv[:, D:2 * D].view(B, T, H, 64).permute(0, 2, 1, 3).double()
    ref = (q @ k.transpose(-1, -2)).reshape(B * H * T, T)
    assert _rel(S[:, :T], ref) < 1e-4


@pytest.mark.parametrize("M,N,K,gelu", [(197 * 8, 1536, 384, True), (197 * 8, 384, 384, False), (197 * 8 + 5, 1152, 384, False),
                                        (130, 128, 72, True), (33, 192, 64, True), (197 * 32, 1536, 384, True)])
@pytest.mark.parametrize("qrange", [(0, 127), (0, 255)])
def test_gemm_gradient_planes_epilogue(cuda_dev, M, N, K, gelu, qrange):
    """act = 2: the dgrad GEMM applies gelu'(FQ(y)) * STE mask * w_scale in its epilogue and emits hi/lo planes plus bias-grad
    slab sums.  Must equal the unfused chain (fp32 GEMM -> qv_gp_planes -> qv_colsum_reduce) bit for bit on the planes (same
    accumulator, same expression) and to fp32 summation-order noise on the column sums."""
    from qatvit_b200 import ops
    from qatvit_b200.ops import Op, PAIRS_EXACT_B
    g = torch.Generator().manual_seed(M + N + K + qrange[1])
    dev = cuda_dev
    a = torch.randn(M, K, generator=g).to(dev)
    b = _codes((N, K), dev, g)
    ap, bp = _planes(a), b.bfloat16()[None].contiguous()
    y = (torch.randn(M, N, generator=g) * 2.0).to(dev)
    scale = torch.tensor([4.0 / (qrange[1] - qrange[0])], device=dev)       # clips ~5 % of y on either side
    zp = torch.tensor([(qrange[0] + qrange[1]) // 2], dtype=torch.int32, device=dev)
    fq = (scale, zp, qrange[0], qrange[1])
    wsc = (torch.rand(N, generator=g) * 0.02 + 0.001).to(dev)
    # unfused
    gmat = ops.gemm(Op.full(ap), Op.full(bp), M, N, K, PAIRS_EXACT_B)
    rpb = 64
    nblk = -(-M // rpb)
    part = torch.empty(nblk, N, device=dev)
    ref_planes = torch.empty(2, M, N, dtype=torch.bfloat16, device=dev)
    ops.gp_planes(gmat, y, fq, wsc, True, gelu, M, N, ref_planes, part, rpb)
    ref_bias = torch.empty(N, device=dev)
    ops.colsum_reduce(part, nblk, N, ref_bias)
    # fused
    out_planes = torch.full((2, M, N), float("nan"), dtype=torch.bfloat16, device=dev)
    nslab = -(-M // 32)
    slab = torch.full((nslab, N), float("nan"), device=dev)
    ops.gemm(Op.full(ap), Op.full(bp), M, N, K, PAIRS_EXACT_B, out_planes=out_planes, col_scale=wsc, grad_of=(y, fq, gelu, slab))
    bias = torch.empty(N, device=dev)
    ops.colsum_reduce(slab, nslab, N, bias)
    torch.cuda.synchronize()
    assert torch.equal(out_planes.view(torch.int16), ref_planes.view(torch.int16))
    assert (ref_planes.float().abs().sum(0) == 0).float().mean() > 0.02      # the STE mask really zeroed something
    assert _rel(bias, ref_bias) < 1e-5


def test_gemm_gradient_planes_epilogue_is_race_free(cuda_dev):
    """Regression: with a single y staging buffer per warp the TMA write of the next unit could land while ld.shared reads of
    the current unit were still in flight (a few stale 16-byte chunks in ~3 % of launches).  100 launches must all match."""
    from qatvit_b200 import ops
    from qatvit_b200.ops import Op, PAIRS_EXACT_B
    dev = cuda_dev
    M, N, K = 1576, 384, 384
    g = torch.Generator().manual_seed(1)
    a = torch.randn(M, K, generator=g).to(dev)
    b = _codes((N, K), dev, g)
    ap, bp = _planes(a), b.bfloat16()[None].contiguous()
    y = (torch.randn(M, N, generator=g) * 2.0).to(dev)
    fq = (torch.tensor([4.0 / 127], device=dev), torch.tensor([63], dtype=torch.int32, device=dev), 0, 127)
    wsc = (torch.rand(N, generator=g) * 0.02 + 0.001).to(dev)
    gmat = ops.gemm(Op.full(ap), Op.full(bp), M, N, K, PAIRS_EXACT_B)
    part = torch.empty(-(-M // 64), N, device=dev)
    ref = torch.empty(2, M, N, dtype=torch.bfloat16, device=dev)
    ops.gp_planes(gmat, y, fq, wsc, True, False, M, N, ref, part, 64)
    slab = torch.empty(-(-M // 32), N, device=dev)
    outs = [torch.empty(2, M, N, dtype=torch.bfloat16, device=dev) for _ in range(100)]
    for out in outs:
        ops.gemm(Op.full(ap), Op.full(bp), M, N, K, PAIRS_EXACT_B, out_planes=out, col_scale=wsc, grad_of=(y, fq, False, slab))
    torch.cuda.synchronize()
    bad = sum(0 if torch.equal(o.view(torch.int16), ref.view(torch.int16)) else 1 for o in outs)
    assert bad == 0, f"{bad} of 100 launches differ"


@pytest.mark.parametrize("M,N,K", [(197 * 8, 1152, 384), (130, 96, 40), (197 * 64, 384, 1536)])
def test_gemm_fused_observer_update(cuda_dev, M, N, K):
    """obs_ticket: the output observer's EMA + qparams run in the GEMM's tail (last epilogue warp of the grid).  Three launches
    (first call, EMA, EMA) must leave exactly the state that GEMM + qv_obs_update leave, and the ticket back at zero."""
    from qatvit_b200 import ops
    from qatvit_b200.ops import Op, PAIRS_EXACT_B
    dev = cuda_dev
    g = torch.Generator().manual_seed(M + N)
    b = _codes((N, K), dev, g)
    bp = b.bfloat16()[None].contiguous()
    scale = (torch.rand(N, generator=g) * 0.02 + 0.001).to(dev)
    bias = torch.randn(N, generator=g).to(dev)

    def fresh():
        return (torch.full((1,), float("inf"), device=dev).squeeze(0), torch.full((1,), float("-inf"), device=dev).squeeze(0),
                torch.ones(1, device=dev), torch.zeros(1, dtype=torch.int32, device=dev))
    on = torch.ones(1, dtype=torch.int64, device=dev)
    ticket = torch.zeros(1, dtype=torch.int32, device=dev)
    st_f, st_r = fresh(), fresh()
    for it in range(3):
        a = (torch.randn(M, K, generator=g) * (1.0 + it)).to(dev)
        ap = _planes(a)
        acc_f, acc_r = ops.new_minmax(dev), ops.new_minmax(dev)
        out_f = ops.gemm(Op.full(ap), Op.full(bp), M, N, K, PAIRS_EXACT_B, col_scale=scale, bias=bias, minmax=acc_f,
                         observer=(st_f[0], st_f[1], st_f[2], st_f[3], on, on, 0.01, 0, 127, False, ticket))
        out_r = ops.gemm(Op.full(ap), Op.full(bp), M, N, K, PAIRS_EXACT_B, col_scale=scale, bias=bias, minmax=acc_r)
        ops.obs_update(acc_r, on, on, st_r[0], st_r[1], st_r[2], st_r[3], 0.01, 0, 127, False)
        torch.cuda.synchronize()
        assert torch.equal(out_f, out_r) and torch.equal(acc_f, acc_r)
        for x, y in zip(st_f, st_r):
            assert torch.equal(x, y)
        assert int(ticket) == 0
    assert float(st_f[2]) != 1.0
