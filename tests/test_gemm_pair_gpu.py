"""CTA-pair GEMM (tcgen05 cta_group::2, clusters of two CTAs: a 256 x 192 tile per SM pair, each SM staging its 128 rows of A
and HALF of the B tile) against the one-CTA-per-tile kernel on identical operands.  Same products, same accumulation order
along K, same epilogue code: the outputs must be bit-identical -- fp32, bf16 hi/lo planes, mixed planes, the gradient-planes
epilogue with its bias-grad slab sums and the fused observer update.  M is chosen ragged against the 256-row pair tile
(the last pair's second CTA is entirely out of range) and large enough for the pair path to be taken (>= 256 x 74 rows).
QV_GEMM_PAIR (bit 0 mixed-format GEMMs with fp32 output, bit 4 with plane output, bit 1 (2,2) bf16-plane GEMMs, bit 2 gradient-planes
dgrad, bit 3 (2,1) GEMMs) is
read per launch."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

M_RAGGED = 197 * 128 + 3          # 25 219 rows: 98 full pair tiles + 131 rows (second CTA of the last pair: 3 rows)
M_HALF = 256 * 98 + 100           # last pair's second CTA has no rows at all


class pair_mode:
    def __init__(self, v):
        self.v = str(v)

    def __enter__(self):
        self.old = os.environ.get("QV_GEMM_PAIR")
        os.environ["QV_GEMM_PAIR"] = self.v

    def __exit__(self, *exc):
        if self.old is None:
            os.environ.pop("QV_GEMM_PAIR", None)
        else:
            os.environ["QV_GEMM_PAIR"] = self.old


def _both(fn):
    from qatvit_b200 import ops
    with pair_mode(0):
        n0 = ops.gemm_pair_launches()
        ref = fn()
        torch.cuda.synchronize()
        assert ops.gemm_pair_launches() == n0                 # one CTA per tile
    with pair_mode(63):
        got = fn()
        torch.cuda.synchronize()
        assert ops.gemm_pair_launches() == n0 + 1             # the pair kernel really ran
    return ref, got


def _same(a, b):
    if a.dtype == torch.bfloat16:
        return torch.equal(a.view(torch.int16), b.view(torch.int16))
    return torch.equal(a, b)


@pytest.mark.parametrize("M", [M_RAGGED, M_HALF])
@pytest.mark.parametrize("N,K", [(768, 768), (2304, 768), (768, 3072)])
def test_pair_mixed_fp32_out(cuda_dev, M, N, K):
    from qatvit_b200 import ops
    from qatvit_b200.ops import Op, PAIRS_FP32
    g = torch.Generator().manual_seed(N + K)
    a = (torch.randn(M, K, generator=g) * 1.3).to(cuda_dev)
    w = (torch.randn(N, K, generator=g) * 0.02).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev)
    am, wm = ops.split_planes_mix(a), ops.split_planes_mix(w, weight=True)
    ref, got = _both(lambda: ops.gemm(Op.full(am), Op.full(wm), M, N, K, PAIRS_FP32, bias=bias, mix=True))
    assert torch.isfinite(got).all() and _same(ref, got)
    exact = a[:2048].double() @ w.double().t() + bias.double()
    assert float((got[:2048].double() - exact).abs().max() / exact.abs().max()) < 1e-4


@pytest.mark.parametrize("N,K,gelu,out_mix", [(2304, 768, False, False), (3072, 768, True, True)])
def test_pair_mixed_plane_out(cuda_dev, N, K, gelu, out_mix):
    from qatvit_b200 import ops
    from qatvit_b200.ops import Op, PAIRS_FP32
    M = M_RAGGED
    g = torch.Generator().manual_seed(N + K + 1)
    a = (torch.randn(M, K, generator=g) * 1.3).to(cuda_dev)
    w = (torch.randn(N, K, generator=g) * 0.02).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev)
    am, wm = ops.split_planes_mix(a), ops.split_planes_mix(w, weight=True)

    def run():
        planes = torch.full((2, M, N), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
        ops.gemm(Op.full(am), Op.full(wm), M, N, K, PAIRS_FP32, bias=bias, out_planes=planes, gelu=gelu, mix=True, out_mix=out_mix)
        return planes
    ref, got = _both(run)
    assert _same(ref, got)


@pytest.mark.parametrize("N,K", [(1152, 384), (384, 384), (384, 1536)])
def test_pair_student_forward_with_observer(cuda_dev, N, K):
    """(2,1) planes: activations x exact weight codes, per-channel scale + bias epilogue, fused min/max."""
    from qatvit_b200 import ops
    from qatvit_b200.ops import Op, PAIRS_EXACT_B
    M = M_RAGGED
    g = torch.Generator().manual_seed(N + K + 2)
    a = torch.randn(M, K, generator=g).to(cuda_dev)
    codes = torch.randint(-128, 128, (N, K), generator=g).float().to(cuda_dev).bfloat16()[None].contiguous()
    wsc = (torch.rand(N, generator=g) * 0.02 + 0.001).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev)
    ap = ops.split_planes(a)

    def run():
        acc = ops.new_minmax(cuda_dev)
        out = ops.gemm(Op.full(ap), Op.full(codes), M, N, K, PAIRS_EXACT_B, col_scale=wsc, bias=bias, minmax=acc)
        return out, acc.clone()
    (ref, racc), (got, gacc) = _both(run)
    assert _same(ref, got) and torch.equal(racc, gacc)
    exact = (a[:1024].double() @ codes[0].double().t()) * wsc.double() + bias.double()
    assert float((got[:1024].double() - exact).abs().max() / exact.abs().max()) < 1e-4


@pytest.mark.parametrize("N,K,gelu", [(1536, 384, True), (1152, 384, False)])
def test_pair_gradient_planes_epilogue(cuda_dev, N, K, gelu):
    from qatvit_b200 import ops
    from qatvit_b200.ops import Op, PAIRS_EXACT_B
    M = M_RAGGED
    dev = cuda_dev
    g = torch.Generator().manual_seed(N + K + 3)
    a = torch.randn(M, K, generator=g).to(dev)
    codes = torch.randint(-128, 128, (N, K), generator=g).float().to(dev).bfloat16()[None].contiguous()
    ap = ops.split_planes(a)
    y = (torch.randn(M, N, generator=g) * 2.0).to(dev)
    fq = (torch.tensor([4.0 / 127], device=dev), torch.tensor([63], dtype=torch.int32, device=dev), 0, 127)
    wsc = (torch.rand(N, generator=g) * 0.02 + 0.001).to(dev)
    nslab = -(-M // 32)

    def run():
        planes = torch.full((2, M, N), float("nan"), dtype=torch.bfloat16, device=dev)
        slab = torch.full((nslab, N), float("nan"), device=dev)
        ops.gemm(Op.full(ap), Op.full(codes), M, N, K, PAIRS_EXACT_B, out_planes=planes, col_scale=wsc, grad_of=(y, fq, gelu, slab))
        return planes, slab
    (rp, rs), (gp, gs) = _both(run)
    assert _same(rp, gp) and torch.equal(rs, gs)


def test_pair_plain_fp32_planes(cuda_dev):
    """(2,2) bf16 hi/lo planes (three passes): the pre-QAT student / three-pass teacher."""
    from qatvit_b200 import ops
    from qatvit_b200.ops import Op, PAIRS_FP32
    M, N, K = M_HALF, 768, 768
    g = torch.Generator().manual_seed(5)
    a = torch.randn(M, K, generator=g).to(cuda_dev)
    w = (torch.randn(N, K, generator=g) * 0.05).to(cuda_dev)
    ap, wp = ops.split_planes(a), ops.split_planes(w)
    ref, got = _both(lambda: ops.gemm(Op.full(ap), Op.full(wp), M, N, K, PAIRS_FP32))
    assert _same(ref, got)

    def run():
        planes = torch.full((2, M, N), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
        ops.gemm(Op.full(ap), Op.full(wp), M, N, K, PAIRS_FP32, out_planes=planes, gelu=True)
        return planes
    ref, got = _both(run)
    assert _same(ref, got)


def test_pair_is_deterministic_over_many_launches(cuda_dev):
    from qatvit_b200 import ops
    from qatvit_b200.ops import Op, PAIRS_FP32
    M, N, K = M_RAGGED, 768, 768
    g = torch.Generator().manual_seed(9)
    a = (torch.randn(M, K, generator=g) * 1.3).to(cuda_dev)
    w = (torch.randn(N, K, generator=g) * 0.02).to(cuda_dev)
    am, wm = ops.split_planes_mix(a), ops.split_planes_mix(w, weight=True)
    with pair_mode(63):
        outs = [ops.gemm(Op.full(am), Op.full(wm), M, N, K, PAIRS_FP32, mix=True) for _ in range(30)]
        torch.cuda.synchronize()
    assert all(torch.equal(o, outs[0]) for o in outs[1:])
