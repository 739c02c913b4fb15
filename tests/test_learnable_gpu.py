"""Learnable per-channel fake-quant (north_star kernel 2: "the per-channel scale gradient done as a warp-shuffle reduction").
The reference has no learnable scale, so this is opt-in; its oracle is the live CPU op
torch._fake_quantize_learnable_per_channel_affine (torch/ao/quantization/_learnable_fake_quantize.py:158-196) and the C
restatement oracle/fq_oracle.c (pinned to it in tests/test_oracle.py).  Bar: y and dx bit-exact (they are element-wise, no
summation order); dscale / dzero_point -- sums of up to 3 072 fp32 terms whose order ATen leaves unspecified -- within
5e-7 x the channel's sum of |g| (|xq - zr| + 1) grad_factor (the magnitude fp32 rounding acts on, oracle/fq_oracle.c) of both
the live op and the double-precision oracle (tolerance stated here), and
bit-reproducible run to run."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(N, K, qmin, qmax, seed, ties=True):
    g = torch.Generator().manual_seed(seed)
    sc = torch.rand(N, generator=g) * 0.01 + 0.002
    x = torch.randn(N, K, generator=g) * 0.5
    zp = torch.randn(N, generator=g) * 3
    if ties and N >= 4:                 # rows on exact .5 ties (forward rounds before adding zp, backward after), far-out zero points
        sc[0], sc[1] = 2.0 ** -7, 2.0 ** -6
        k = torch.randint(-140, 140, (K,), generator=g).float()
        x[0], x[1] = (k + 0.5) * sc[0], (k + 0.5) * sc[1]
        zp[0], zp[1], zp[2], zp[3] = 1.0, 2.5, 300.0, -300.0
    gy = torch.randn(N, K, generator=g)
    return x, sc, zp, gy


@pytest.mark.parametrize("N,K,qmin,qmax,gf", [(384, 384, -128, 127, 1.0), (1536, 384, -128, 127, 0.0131), (384, 1536, -128, 127, 0.5),
                                               (2304, 768, -128, 127, 1.0), (10, 384, -128, 127, 1.0), (7, 33, 0, 255, 0.37),
                                               (5, 1, -128, 127, 1.0), (384, 768, 0, 127, 1.0)])
def test_learnable_per_channel_matches_live_cpu_op_and_oracle(cuda_dev, N, K, qmin, qmax, gf):
    from qatvit_b200 import ops
    from oracle import fq_oracle as fo
    x, sc, zp, gy = _case(N, K, qmin, qmax, seed=N * 7 + K)
    xr, sr, zr = x.clone().requires_grad_(), sc.clone().requires_grad_(), zp.clone().requires_grad_()
    y_ref = torch._fake_quantize_learnable_per_channel_affine(xr, sr, zr, 0, qmin, qmax, gf)
    y_ref.backward(gy)
    d = cuda_dev
    y = ops.fq_learnable_fwd(x.to(d), sc.to(d), zp.to(d), qmin, qmax)
    dx, ds, dz = ops.fq_learnable_bwd(gy.to(d), x.to(d), sc.to(d), zp.to(d), qmin, qmax, gf)
    assert torch.equal(y.cpu(), y_ref.detach())
    assert torch.equal(dx.cpu(), xr.grad)
    dx_o, ds_o, dz_o, da = fo.fq_learnable_bwd(gy.numpy(), x.numpy(), sc.numpy(), zp.numpy(), qmin, qmax, gf)
    tol_s = 5e-7 * da + 1e-30
    tol_z = 2e-6 * (np.abs(gy.numpy()).sum(1) * sc.numpy() * gf) + 1e-30
    assert np.all(np.abs(ds.cpu().numpy() - sr.grad.numpy()) <= tol_s)
    assert np.all(np.abs(dz.cpu().numpy() - zr.grad.numpy()) <= tol_z)
    # the C oracle (double sums) agrees too
    assert np.array_equal(dx.cpu().numpy(), dx_o)
    assert np.all(np.abs(ds.cpu().numpy().astype(np.float64) - ds_o) <= tol_s)
    assert np.all(np.abs(dz.cpu().numpy().astype(np.float64) - dz_o) <= tol_z)
    # shuffle reduction in a fixed order: bit-reproducible
    dx2, ds2, dz2 = ops.fq_learnable_bwd(gy.to(d), x.to(d), sc.to(d), zp.to(d), qmin, qmax, gf)
    assert torch.equal(ds, ds2) and torch.equal(dz, dz2)


def test_learnable_module_dropin_is_opt_in_and_matches_cpu_module(cuda_dev):
    """_LearnableFakeQuantize as a per-channel weight fake-quant: install(learnable=True) routes its forward / backward to the
    kernels; parameters' gradients equal the CPU module's; install() alone leaves the class untouched."""
    import copy
    from torch.ao.quantization._learnable_fake_quantize import _LearnableFakeQuantize
    from torch.ao.quantization.observer import MovingAveragePerChannelMinMaxObserver
    from qatvit_b200 import dropin
    stock = _LearnableFakeQuantize.forward
    dropin.install()
    assert _LearnableFakeQuantize.forward is stock                    # default: off
    try:
        dropin.install(learnable=True)
        assert _LearnableFakeQuantize.forward is not stock
        torch.manual_seed(3)
        N, K = 384, 1536
        fq = _LearnableFakeQuantize(MovingAveragePerChannelMinMaxObserver, quant_min=-128, quant_max=127, scale=0.01, zero_point=0.0,
                                    channel_len=N, use_grad_scaling=True, dtype=torch.qint8, qscheme=torch.per_channel_symmetric,
                                    ch_axis=0)
        with torch.no_grad():
            fq.scale.copy_(torch.rand(N) * 0.004 + 0.001)
        fq.enable_param_learning()                      # learning on, static observation off, fake-quant on
        fq_gpu = copy.deepcopy(fq).to(cuda_dev)
        w = (torch.randn(N, K) * 0.2)
        gy = torch.randn(N, K)
        wc = w.clone().requires_grad_()
        fq(wc).backward(gy)
        wg = w.to(cuda_dev).requires_grad_()
        y = fq_gpu(wg)
        y.backward(gy.to(cuda_dev))
        assert torch.equal(y.detach().cpu(), fq(w).detach())
        assert torch.equal(wg.grad.cpu(), wc.grad)
        assert float((fq_gpu.scale.grad.cpu() - fq.scale.grad).abs().max()) <= 1e-5 * float(fq.scale.grad.abs().max())
        assert fq_gpu.zero_point.grad is not None
    finally:
        dropin.uninstall()
    assert _LearnableFakeQuantize.forward is stock
