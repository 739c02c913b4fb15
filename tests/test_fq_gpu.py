"""GPU parity (tier 1, bit-exact): observer + fake-quant kernels through the C-ABI vs
(a) the C oracle (oracle/fq_oracle.c) and (b) the live torch CPU op the reference runs
(torch.fused_moving_avg_obs_fake_quant, torch/ao/quantization/fake_quantize.py:423-438)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CONFIGS = [  # (qmin, qmax, symmetric)  -- fbgemm act, qnnpack act, per-tensor symmetric weights
    (0, 127, False), (0, 255, False), (-128, 127, True),
]


def _inputs(kind, shape, gen):
    x = torch.randn(shape, generator=gen)
    if kind == "normal":
        return [x * 2 + 0.3, x * 3 - 1, x * 0.5]
    if kind == "tiny":            # hits the 6.1e-5 small-scale cut-off
        return [x * 1e-4, x * 2e-4, x * 1e-5]
    if kind == "positive":
        return [x.abs() + 0.1, x.abs() * 2, x.abs()]
    if kind == "negative":
        return [-x.abs() - 0.1, -x.abs() * 2, -x.abs()]
    if kind == "zeros_then_ties":
        t = (torch.randint(-300, 300, shape, generator=gen).float() + 0.5) * 0.05
        return [torch.zeros(shape), t, x]
    if kind == "ties":            # values exactly on .5 rounding boundaries of a power-of-two scale
        t = (torch.randint(-100, 100, shape, generator=gen).float() + 0.5) * 0.125
        t.view(-1)[0] = -15.875
        if t.numel() > 1:
            t.view(-1)[1] = 16.0
        return [t, t * 2, t]
    raise ValueError(kind)


def _torch_cpu(xs, qmin, qmax, sym, obs_on=1, fq_on=1):
    mn, mx = torch.tensor(float("inf")), torch.tensor(float("-inf"))
    s, z = torch.ones(1), torch.zeros(1, dtype=torch.int32)
    outs = []
    for x in xs:
        xt = x.clone().requires_grad_(True)
        y = torch.fused_moving_avg_obs_fake_quant(xt, torch.tensor([obs_on]), torch.tensor([fq_on]), mn, mx, s, z, 0.01,
                                                  qmin, qmax, 0, False, sym)
        g = torch.ones_like(y)
        y.backward(g)
        outs.append((y.detach().clone(), xt.grad.clone(), float(mn), float(mx), float(s), int(z)))
    return outs


@pytest.mark.parametrize("kind", ["normal", "tiny", "positive", "negative", "zeros_then_ties", "ties"])
@pytest.mark.parametrize("cfg", CONFIGS)
@pytest.mark.parametrize("shape", [(8, 197, 384), (3, 5, 7), (1,)])
def test_act_fq_bit_exact(cuda_dev, kind, cfg, shape):
    from qatvit_b200 import ops
    from oracle import fq_oracle as fo
    qmin, qmax, sym = cfg
    gen = torch.Generator().manual_seed(hash((kind, cfg, shape)) % (2 ** 31))
    xs = _inputs(kind, shape, gen)
    ref = _torch_cpu(xs, qmin, qmax, sym)
    st = fo.FQState(qmin, qmax, sym)
    dev = cuda_dev
    min_val = torch.tensor(float("inf"), device=dev)
    max_val = torch.tensor(float("-inf"), device=dev)
    scale = torch.ones(1, device=dev)
    zp = torch.zeros(1, dtype=torch.int32, device=dev)
    on = torch.ones(1, dtype=torch.int64, device=dev)
    acc = ops.new_minmax(dev)
    for x, (y_ref, mask_ref, mn_ref, mx_ref, s_ref, z_ref) in zip(xs, ref):
        yo, mo, _ = fo.fused_obs_fq(x.numpy(), st)
        xd = x.to(dev)
        ops.minmax_reset(acc)
        ops.minmax_accumulate(xd, acc)
        ops.obs_update(acc, on, on, min_val, max_val, scale, zp, 0.01, qmin, qmax, sym)
        y, mask = ops.fq_apply(xd, scale, zp, on, qmin, qmax)
        gx = ops.fq_bwd(torch.ones_like(xd), mask)
        # state: bit-exact against the live torch CPU op and the C oracle
        assert float(min_val) == mn_ref == float(st.min_val[0])
        assert float(max_val) == mx_ref == float(st.max_val[0])
        assert float(scale) == s_ref == float(st.scale[0])
        assert int(zp) == z_ref == int(st.zero_point[0])
        # fake-quantised values (hence integer codes) and STE mask: bit-exact
        assert torch.equal(y.cpu(), y_ref)
        assert np.array_equal(y.cpu().numpy(), yo)
        assert torch.equal(gx.cpu(), mask_ref)
        assert np.array_equal(mask.cpu().numpy(), mo)


@pytest.mark.parametrize("per_channel", [True, False])
@pytest.mark.parametrize("shape", [(1152, 384), (10, 384), (384, 3, 16, 16), (7, 33)])
def test_weight_fq_bit_exact(cuda_dev, per_channel, shape):
    from qatvit_b200 import ops
    dev = cuda_dev
    gen = torch.Generator().manual_seed(1234)
    w0 = torch.randn(shape, generator=gen) * 0.02
    w0[0] = w0[0].abs() + 1e-3          # all-positive row  -> zp = -128 (per-channel)
    if shape[0] > 2:
        w0[1] = -w0[1].abs() - 1e-3     # all-negative row  -> zp = 127
        w0[2] = 0.0                     # all-zero row      -> scale 0.1
    ws = [w0, w0 + 1e-3 * torch.randn(shape, generator=gen), w0 * 1.5]
    C = shape[0]
    qmin, qmax, sym = -128, 127, True
    # reference: live torch CPU op
    mn = torch.empty(0) if per_channel else torch.tensor(float("inf"))
    mx = torch.empty(0) if per_channel else torch.tensor(float("-inf"))
    s, z = torch.ones(1), torch.zeros(1, dtype=torch.int32)
    n_state = C if per_channel else 1
    min_val = torch.full((n_state,), float("inf"), device=dev)
    max_val = torch.full((n_state,), float("-inf"), device=dev)
    scale = torch.ones(n_state, device=dev)
    zp = torch.zeros(n_state, dtype=torch.int32, device=dev)
    on = torch.ones(1, dtype=torch.int64, device=dev)
    scratch = torch.zeros(2, dtype=torch.int32, device=dev)
    for w in ws:
        wt = w.clone().requires_grad_(True)
        y_ref = torch.fused_moving_avg_obs_fake_quant(wt, torch.tensor([1]), torch.tensor([1]), mn, mx, s, z, 0.01, qmin,
                                                      qmax, 0, per_channel, sym)
        y_ref.backward(torch.ones_like(y_ref))
        wd = w.to(dev).contiguous()
        rows, cols = C, w.numel() // C
        y = torch.empty_like(wd)
        mask = torch.empty(wd.shape, dtype=torch.uint8, device=dev)
        codes = torch.empty(rows, cols, dtype=torch.bfloat16, device=dev)
        codes_t = torch.empty(cols, rows, dtype=torch.bfloat16, device=dev)
        ops.fq_weight(wd, per_channel, on, on, min_val, max_val, scale, zp, 0.01, qmin, qmax, sym, y=y, mask=mask,
                      codes=codes, codes_t=codes_t, scratch=scratch)
        assert torch.equal(min_val.cpu(), mn.reshape(-1))
        assert torch.equal(max_val.cpu(), mx.reshape(-1))
        assert torch.equal(scale.cpu(), s.reshape(-1))
        assert torch.equal(zp.cpu(), z.reshape(-1))
        assert torch.equal(y.cpu(), y_ref.detach())
        assert torch.equal(mask.cpu().float(), wt.grad)
        # integer codes: (q - zp) * scale must reproduce y exactly, and the transposed copy must agree
        sc = scale.cpu().reshape(-1, 1) if per_channel else scale.cpu()
        assert torch.equal((codes.float().cpu() * sc).reshape(shape), y_ref.detach())
        assert torch.equal(codes_t.cpu().t().contiguous(), codes.cpu())
        q = codes.float().cpu() + (zp.cpu().reshape(-1, 1).float() if per_channel else zp.cpu().float())
        assert q.min() >= qmin and q.max() <= qmax


def test_flags_gate_on_device(cuda_dev):
    """observer off => state frozen but still quantises; fake-quant off => identity, min/max move, scale/zp do not."""
    from qatvit_b200 import ops
    dev = cuda_dev
    x1, x2 = torch.randn(1000) * 2, torch.randn(1000) * 5 + 1
    for obs_on, fq_on in [(0, 1), (1, 0), (0, 0)]:
        mn, mx = torch.tensor(float("inf")), torch.tensor(float("-inf"))
        s, z = torch.ones(1), torch.zeros(1, dtype=torch.int32)
        torch.fused_moving_avg_obs_fake_quant(x1, torch.tensor([1]), torch.tensor([1]), mn, mx, s, z, 0.01, 0, 127, 0, False, False)
        y_ref = torch.fused_moving_avg_obs_fake_quant(x2, torch.tensor([obs_on]), torch.tensor([fq_on]), mn, mx, s, z, 0.01,
                                                      0, 127, 0, False, False)
        min_val = torch.tensor(float("inf"), device=dev)
        max_val = torch.tensor(float("-inf"), device=dev)
        scale = torch.ones(1, device=dev)
        zp = torch.zeros(1, dtype=torch.int32, device=dev)
        one = torch.ones(1, dtype=torch.int64, device=dev)
        acc = ops.new_minmax(dev)
        ops.minmax_accumulate(x1.to(dev), acc)
        ops.obs_update(acc, one, one, min_val, max_val, scale, zp, 0.01, 0, 127, False)
        o = torch.tensor([obs_on], dtype=torch.int64, device=dev)
        f = torch.tensor([fq_on], dtype=torch.int64, device=dev)
        ops.minmax_reset(acc)
        ops.minmax_accumulate(x2.to(dev), acc)
        ops.obs_update(acc, o, f, min_val, max_val, scale, zp, 0.01, 0, 127, False)
        y, _ = ops.fq_apply(x2.to(dev), scale, zp, f, 0, 127)
        assert float(min_val) == float(mn) and float(max_val) == float(mx)
        assert float(scale) == float(s) and int(zp) == int(z)
        assert torch.equal(y.cpu(), y_ref)


def test_empty_input_is_noop(cuda_dev):
    from qatvit_b200 import ops
    dev = cuda_dev
    acc = ops.new_minmax(dev)
    ops.minmax_accumulate(torch.empty(0, device=dev), acc)
    min_val = torch.tensor(float("inf"), device=dev)
    max_val = torch.tensor(float("-inf"), device=dev)
    scale = torch.ones(1, device=dev)
    zp = torch.zeros(1, dtype=torch.int32, device=dev)
    one = torch.ones(1, dtype=torch.int64, device=dev)
    ops.obs_update(acc, one, one, min_val, max_val, scale, zp, 0.01, 0, 127, False)
    assert float(scale) == 1.0 and int(zp) == 0 and float(min_val) == float("inf")


@pytest.mark.parametrize("B,C", [(8, 10), (256, 10), (64, 1000)])
@pytest.mark.parametrize("fq", [False, True])
def test_kd_ce_loss(cuda_dev, B, C, fq):
    """Loss + gradient vs the reference's own torch expression (qat_trainer.py:343-349) on CPU; 1e-5 rel."""
    from qatvit_b200 import ops
    from oracle import vit_ref as vr
    dev = cuda_dev
    gen = torch.Generator().manual_seed(5)
    s = torch.randn(B, C, generator=gen) * 3
    t = torch.randn(B, C, generator=gen) * 6
    y = torch.randint(0, C, (B,), generator=gen)
    hp = dict(vr.DEFAULT_HPARAMS)
    if fq:
        scale, zp = torch.tensor([0.0731]), torch.tensor([61], dtype=torch.int32)
        sq = s.clone().requires_grad_(True)
        s_in = torch.fake_quantize_per_tensor_affine(sq, 0.0731, 61, 0, 127)
    else:
        sq = s.clone().requires_grad_(True)
        s_in = sq
    loss, kd, ce = vr.distill_loss(s_in, t, y, hp)
    loss.backward()
    out3, grad = ops.kd_ce_loss(s.to(dev), t.to(dev), y.to(dev), hp["kd_temp"], hp["kd_alpha"], hp["label_smoothing"],
                                s_scale=scale.to(dev) if fq else None, s_zp=zp.to(dev) if fq else None, qmin=0, qmax=127)
    out3 = out3.cpu()
    assert abs(out3[0] - loss.item()) <= 1e-5 * abs(loss.item())
    assert abs(out3[1] - kd.item()) <= 1e-5 * abs(kd.item()) + 1e-7
    assert abs(out3[2] - ce.item()) <= 1e-5 * abs(ce.item())
    assert torch.allclose(grad.cpu(), sq.grad, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("B,C", [(513, 10), (300, 1000), (2048, 1002), (4096, 1000)])
@pytest.mark.parametrize("fq", [False, True])
def test_kd_ce_loss_rows(cuda_dev, B, C, fq):
    """The grid form (qv_kd_ce_loss_rows: one warp per row, 16-byte loads when C % 4 == 0, scalar otherwise) against the same
    CPU torch expression, launched three times on one workspace (the ticket word must come back to zero) with bit-identical
    results (fixed-order reduction), and against the one-block kernel."""
    from qatvit_b200 import ops
    from oracle import vit_ref as vr
    dev = cuda_dev
    gen = torch.Generator().manual_seed(6)
    s = torch.randn(B, C, generator=gen) * 3
    t = torch.randn(B, C, generator=gen) * 6
    y = torch.randint(0, C, (B,), generator=gen)
    hp = dict(vr.DEFAULT_HPARAMS)
    scale, zp = torch.tensor([0.0731]), torch.tensor([61], dtype=torch.int32)
    sq = s.clone().requires_grad_(True)
    s_in = torch.fake_quantize_per_tensor_affine(sq, 0.0731, 61, 0, 127) if fq else sq
    loss, kd, ce = vr.distill_loss(s_in, t, y, hp)
    loss.backward()
    kw = dict(s_scale=scale.to(dev) if fq else None, s_zp=zp.to(dev) if fq else None, qmin=0, qmax=127)
    args = (s.to(dev), t.to(dev), y.to(dev), hp["kd_temp"], hp["kd_alpha"], hp["label_smoothing"])
    runs = []
    for _ in range(3):
        out3, grad = ops.kd_ce_loss(*args, rows=True, **kw)
        runs.append((out3.clone(), grad.clone()))
    torch.cuda.synchronize()
    for o, g in runs[1:]:
        assert torch.equal(o, runs[0][0]) and torch.equal(g, runs[0][1])
    out3, grad = runs[0][0].cpu(), runs[0][1].cpu()
    assert abs(out3[0] - loss.item()) <= 1e-5 * abs(loss.item())
    assert abs(out3[1] - kd.item()) <= 1e-5 * abs(kd.item()) + 1e-7
    assert abs(out3[2] - ce.item()) <= 1e-5 * abs(ce.item())
    assert torch.allclose(grad, sq.grad, rtol=1e-4, atol=1e-7)
    one3, one_grad = ops.kd_ce_loss(*args, rows=False, **kw)
    assert torch.allclose(one3.cpu(), out3, rtol=2e-6, atol=0)
    assert torch.allclose(one_grad.cpu(), grad, rtol=1e-5, atol=1e-9)     # same expression, different lane -> column assignment


def test_weight_fq_grouped_bit_exact(cuda_dev):
    """qv_fq_weight_grouped: several per-channel weights (ViT-S shapes, a ragged row count and edge-case rows) in ONE launch,
    three EMA steps; state, fake-quantised values (codes * scale), STE mask and the transposed code plane must equal the live
    torch CPU op applied to each weight separately."""
    from qatvit_b200 import ops
    dev = cuda_dev
    gen = torch.Generator().manual_seed(77)
    shapes = [(1152, 384), (384, 384), (1536, 384), (384, 1536), (384, 768), (40, 64), (8, 4)]
    qmin, qmax, sym = -128, 127, True
    w0s = []
    for shp in shapes:
        w = torch.randn(shp, generator=gen) * 0.02
        w[0] = w[0].abs() + 1e-3
        w[1] = -w[1].abs() - 1e-3
        w[2] = 0.0
        w[3] = w[3] * 1e-4              # below the 6.1e-5 scale cut-off
        w0s.append(w)
    ref_state = [dict(mn=torch.empty(0), mx=torch.empty(0), s=torch.ones(1), z=torch.zeros(1, dtype=torch.int32)) for _ in shapes]
    on = torch.ones(1, dtype=torch.int64, device=dev)
    ent = []
    for shp in shapes:
        n, k = shp
        ent.append(dict(w=torch.empty(n, k, device=dev), min_val=torch.full((n,), float("inf"), device=dev),
                        max_val=torch.full((n,), float("-inf"), device=dev), scale=torch.ones(n, device=dev),
                        zero_point=torch.zeros(n, dtype=torch.int32, device=dev), observer_enabled=on, fake_quant_enabled=on,
                        mask=torch.empty(n, k, dtype=torch.uint8, device=dev), codes=torch.empty(n, k, dtype=torch.bfloat16, device=dev),
                        codes_t=torch.empty(k, n, dtype=torch.bfloat16, device=dev)))
    table, blocks, max_cols = ops.fq_weight_group_table(ent, dev)
    assert blocks == sum(-(-s[0] // ops.FQW_ROWS) for s in shapes) and max_cols == 1536
    for step in range(3):
        for w0, e in zip(w0s, ent):
            e["w"].copy_(w0 * (1.0 + 0.3 * step) + 1e-3 * step)
        ops.fq_weight_grouped(table, len(ent), blocks, max_cols, 0.01, qmin, qmax, sym)
        torch.cuda.synchronize()
        for w0, e, r in zip(w0s, ent, ref_state):
            wt = (w0 * (1.0 + 0.3 * step) + 1e-3 * step).clone().requires_grad_(True)
            y_ref = torch.fused_moving_avg_obs_fake_quant(wt, torch.tensor([1]), torch.tensor([1]), r["mn"], r["mx"], r["s"], r["z"],
                                                          0.01, qmin, qmax, 0, True, sym)
            y_ref.backward(torch.ones_like(y_ref))
            assert torch.equal(e["min_val"].cpu(), r["mn"]) and torch.equal(e["max_val"].cpu(), r["mx"])
            assert torch.equal(e["scale"].cpu(), r["s"]) and torch.equal(e["zero_point"].cpu(), r["z"])
            assert torch.equal(e["codes"].float().cpu() * e["scale"].cpu().reshape(-1, 1), y_ref.detach())
            assert torch.equal(e["mask"].cpu().float(), wt.grad)
            assert torch.equal(e["codes_t"].cpu().t().contiguous(), e["codes"].cpu())
