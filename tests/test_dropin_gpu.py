"""GPU parity of the module-level drop-ins (qatvit_b200.dropin.install): the stock torch.ao modules the reference's
prepare_qat call creates (ref/src/training/qat_trainer.py:304-308), run under stock autograd on CUDA tensors after
install(), against the same modules on CPU (the reference path).

* FusedMovingAvgObsFakeQuantize.forward  -> bit-exact y, STE gradient and observer state
* torch.ao.nn.qat.Linear.forward         -> weight observer bit-exact; y / gx / gW / gb within 1e-4 (tolerance:
                                            north_star 1e-3; bf16 hi/lo tensor-core products are ~2^-16)
* dropin.distill_loss                    -> loss and d loss / d logits vs the inline reference loss (ref :343-349)
* the reference's unmodified loop body on a prepared model, module types / state_dict / convert() unchanged
"""
import copy
import warnings

import pytest
import torch

from parity_utils import build_models, rel_l2, rel_max

pytestmark = pytest.mark.gpu


@pytest.fixture()
def installed():
    from qatvit_b200 import dropin
    dropin.install()
    yield dropin
    dropin.uninstall()


def _qconfig(backend):
    from torch.ao.quantization import get_default_qat_qconfig
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return get_default_qat_qconfig(backend)


@pytest.mark.parametrize("backend", ["fbgemm", "qnnpack"])
@pytest.mark.parametrize("which", ["activation", "weight"])
def test_fake_quantize_module_bit_exact(cuda_dev, installed, backend, which):
    qc = _qconfig(backend)
    cpu_mod = qc.activation() if which == "activation" else qc.weight()
    gpu_mod = copy.deepcopy(cpu_mod).to(cuda_dev)
    g = torch.Generator().manual_seed(7)
    shape = (4, 197, 384) if which == "activation" else (1152, 384)
    for it in range(3):
        x = torch.randn(shape, generator=g) * (0.5 + it) + 0.1 * it
        xc = x.clone().requires_grad_(True)
        xg = x.to(cuda_dev).requires_grad_(True)
        yc, yg = cpu_mod(xc), gpu_mod(xg)
        gy = torch.randn(shape, generator=g)
        yc.backward(gy)
        yg.backward(gy.to(cuda_dev))
        assert torch.equal(yg.cpu(), yc), (backend, which, it)
        assert torch.equal(xg.grad.cpu(), xc.grad)
        for k, v in cpu_mod.state_dict().items():
            assert torch.equal(gpu_mod.state_dict()[k].cpu(), v), k
    assert type(gpu_mod) is type(cpu_mod)


@pytest.mark.parametrize("backend", ["fbgemm", "qnnpack"])
@pytest.mark.parametrize("shape", [(4, 197, 384, 1152), (2, 50, 1536, 384), (3, 7, 64, 96)])
def test_qat_linear_module(cuda_dev, installed, backend, shape):
    import torch.ao.nn.qat as nnqat
    B, T, K, N = shape
    torch.manual_seed(0)
    cpu_mod = nnqat.Linear(K, N, bias=True, qconfig=_qconfig(backend))
    with torch.no_grad():
        cpu_mod.bias.normal_(std=0.1)
    gpu_mod = copy.deepcopy(cpu_mod).to(cuda_dev)
    for it in range(2):
        x = torch.randn(B, T, K) * (1.0 + it)
        gy = torch.randn(B, T, N)
        xc = x.clone().requires_grad_(True)
        xg = x.to(cuda_dev).requires_grad_(True)
        yc, yg = cpu_mod(xc), gpu_mod(xg)
        yc.backward(gy)
        yg.backward(gy.to(cuda_dev))
        assert rel_max(yg, yc) < 1e-4
        assert rel_max(xg.grad, xc.grad) < 1e-4
        assert rel_max(gpu_mod.weight.grad, cpu_mod.weight.grad) < 1e-4
        assert rel_max(gpu_mod.bias.grad, cpu_mod.bias.grad) < 1e-4
        # the STE mask zeroes exactly the same weight-gradient entries
        assert torch.equal(gpu_mod.weight.grad.cpu() == 0, cpu_mod.weight.grad == 0)
        for k, v in cpu_mod.weight_fake_quant.state_dict().items():
            assert torch.equal(gpu_mod.weight_fake_quant.state_dict()[k].cpu(), v), k
        cpu_mod.zero_grad(set_to_none=True)
        gpu_mod.zero_grad(set_to_none=True)
        with torch.no_grad():                      # move the weights so the EMA branch sees new min/max
            delta = 0.01 * torch.randn_like(cpu_mod.weight)
            cpu_mod.weight.add_(delta)
            gpu_mod.weight.add_(delta.to(cuda_dev))


@pytest.mark.parametrize("backend", ["fbgemm", "qnnpack"])
def test_qat_conv2d_patch_embedding_module(cuda_dev, installed, backend):
    """nnqat.Conv2d with kernel == stride (timm PatchEmbed.proj, ref model_registry.py:167-172 -> torch/ao/nn/qat/modules/conv.py:55-56)
    as im2col + the fake-quant GEMM: y, all three gradients and the weight observer vs the stock module on CPU."""
    import torch.ao.nn.qat as nnqat
    torch.manual_seed(1)
    cpu_mod = nnqat.Conv2d(3, 64, kernel_size=16, stride=16, bias=True, qconfig=_qconfig(backend))
    gpu_mod = copy.deepcopy(cpu_mod).to(cuda_dev)
    n0 = dict(installed.stats)
    for it in range(2):
        x = torch.randn(4, 3, 64, 64) * (1.0 + it)
        gy = torch.randn(4, 64, 4, 4)
        xc = x.clone().requires_grad_(True)
        xg = x.to(cuda_dev).requires_grad_(True)
        yc, yg = cpu_mod(xc), gpu_mod(xg)
        assert yg.shape == yc.shape
        yc.backward(gy)
        yg.backward(gy.to(cuda_dev))
        assert rel_max(yg, yc) < 1e-4
        assert rel_max(xg.grad, xc.grad) < 1e-4
        assert rel_max(gpu_mod.weight.grad, cpu_mod.weight.grad) < 1e-4
        assert rel_max(gpu_mod.bias.grad, cpu_mod.bias.grad) < 1e-4
        assert torch.equal(gpu_mod.weight.grad.cpu() == 0, cpu_mod.weight.grad == 0)
        for k, v in cpu_mod.weight_fake_quant.state_dict().items():
            assert torch.equal(gpu_mod.weight_fake_quant.state_dict()[k].cpu(), v), k
        cpu_mod.zero_grad(set_to_none=True)
        gpu_mod.zero_grad(set_to_none=True)
    assert installed.stats["conv_gemm"] - n0["conv_gemm"] == 2 and installed.stats["conv_stock"] == n0["conv_stock"]
    # a convolution that is not a patch embedding keeps the stock route (still fake-quantised by our kernel)
    other = nnqat.Conv2d(3, 32, kernel_size=3, stride=1, padding=1, qconfig=_qconfig(backend)).to(cuda_dev)
    other(torch.randn(2, 3, 8, 8, device=cuda_dev))
    assert installed.stats["conv_stock"] == n0["conv_stock"] + 1


@pytest.mark.parametrize("backend", ["fbgemm", "qnnpack"])
@pytest.mark.parametrize("B", [8, 256])
def test_qat_small_linear_head_module(cuda_dev, installed, backend, B):
    """The 10-class head (N = 10 is not a tensor-core tile): fake-quantised fp32 weight + the small-Linear kernels."""
    import torch.ao.nn.qat as nnqat
    torch.manual_seed(2)
    cpu_mod = nnqat.Linear(384, 10, bias=True, qconfig=_qconfig(backend))
    gpu_mod = copy.deepcopy(cpu_mod).to(cuda_dev)
    n0 = dict(installed.stats)
    for it in range(2):
        x = torch.randn(B, 384) * (1.0 + it)
        gy = torch.randn(B, 10)
        xc = x.clone().requires_grad_(True)
        xg = x.to(cuda_dev).requires_grad_(True)
        yc, yg = cpu_mod(xc), gpu_mod(xg)
        yc.backward(gy)
        yg.backward(gy.to(cuda_dev))
        assert rel_max(yg, yc) < 1e-5
        assert rel_max(xg.grad, xc.grad) < 1e-5
        assert rel_max(gpu_mod.weight.grad, cpu_mod.weight.grad) < 1e-5
        assert rel_max(gpu_mod.bias.grad, cpu_mod.bias.grad) < 1e-5
        assert torch.equal(gpu_mod.weight.grad.cpu() == 0, cpu_mod.weight.grad == 0)
        for k, v in cpu_mod.weight_fake_quant.state_dict().items():
            assert torch.equal(gpu_mod.weight_fake_quant.state_dict()[k].cpu(), v), k
        cpu_mod.zero_grad(set_to_none=True)
        gpu_mod.zero_grad(set_to_none=True)
    assert installed.stats["linear_small"] - n0["linear_small"] == 2 and installed.stats["linear_stock"] == n0["linear_stock"]


def test_fake_quant_disabled_and_reused_modules(cuda_dev, installed):
    """(a) torch.ao.quantization.disable_fake_quant: weights are then not integer codes -- the drop-in must notice the flag
    (cached, re-read when the buffer changes) and compute x W^T with the RAW weight like the stock module; (b) a module
    applied twice before backward (weight sharing) must not have the first call's saved operands overwritten by the second."""
    import torch.ao.nn.qat as nnqat
    from torch.ao.quantization import disable_fake_quant, enable_fake_quant
    torch.manual_seed(3)
    cpu_mod = nnqat.Linear(64, 96, bias=True, qconfig=_qconfig("fbgemm"))
    gpu_mod = copy.deepcopy(cpu_mod).to(cuda_dev)
    x = torch.randn(5, 7, 64)
    for step in range(3):
        if step == 1:
            cpu_mod.apply(disable_fake_quant)
            gpu_mod.apply(disable_fake_quant)
        if step == 2:
            cpu_mod.apply(enable_fake_quant)
            gpu_mod.apply(enable_fake_quant)
        n0 = dict(installed.stats)
        yc, yg = cpu_mod(x), gpu_mod(x.to(cuda_dev))
        assert rel_max(yg, yc) < 1e-4, step
        took_stock = installed.stats["linear_stock"] - n0["linear_stock"]
        assert took_stock == (1 if step == 1 else 0)
        if step == 1:       # raw weights: the result must differ from the fake-quantised one by more than the tolerance
            w = cpu_mod.weight.detach()
            assert rel_max(yc, torch.nn.functional.linear(x, w, cpu_mod.bias.detach())) < 1e-6
    # (b) y = f(f(x)) with one module (square so that it composes)
    torch.manual_seed(4)
    cpu_sq = nnqat.Linear(64, 64, bias=True, qconfig=_qconfig("fbgemm"))
    gpu_sq = copy.deepcopy(cpu_sq).to(cuda_dev)
    xc = torch.randn(33, 64, requires_grad=True)
    xg = xc.detach().to(cuda_dev).requires_grad_(True)
    cpu_sq(cpu_sq(xc)).square().sum().backward()
    gpu_sq(gpu_sq(xg)).square().sum().backward()
    assert rel_max(xg.grad, xc.grad) < 2e-4
    assert rel_max(gpu_sq.weight.grad, cpu_sq.weight.grad) < 2e-4
    assert rel_max(gpu_sq.bias.grad, cpu_sq.bias.grad) < 2e-4


@pytest.mark.parametrize("B,C", [(256, 10), (8, 10), (5, 1000)])
def test_distill_loss_dropin(cuda_dev, installed, B, C):
    from oracle import vit_ref as vr
    hp = dict(vr.DEFAULT_HPARAMS)
    g = torch.Generator().manual_seed(B + C)
    s = torch.randn(B, C, generator=g) * 3
    t = torch.randn(B, C, generator=g) * 5
    y = torch.randint(0, C, (B,), generator=g)
    sc = s.clone().requires_grad_(True)
    ref, _, _ = vr.distill_loss(sc, t, y, hp)
    ref.backward()
    sg = s.to(cuda_dev).requires_grad_(True)
    ours = installed.distill_loss(sg, t.to(cuda_dev), y.to(cuda_dev), hp["kd_temp"], hp["kd_alpha"], hp["label_smoothing"])
    (ours * 2.0).backward()
    assert abs(float(ours) - float(ref)) <= 1e-5 * abs(float(ref))
    assert rel_max(sg.grad * 0.5, sc.grad) < 1e-4


def test_reference_loop_body_with_dropins(cuda_dev, installed):
    """The reference's own loop body (restated verbatim in oracle.vit_ref.distill_step) over the prepared student, on CUDA
    with the drop-ins installed, vs the same body on CPU.  Free-running, so the bound is the eager-QAT noise floor
    (tests/parity_utils.py); module types, state_dict keys and stock convert() must be unchanged."""
    from torch.ao.quantization import convert
    from qatvit_b200 import ops
    vr, prepared, teacher = build_models("fbgemm", "vit_test_tiny", "vit_test_teacher", 64)
    images, labels = vr.synthetic_batch(8, seed=5, img=64)
    hp = dict(vr.DEFAULT_HPARAMS)
    gpu_student = copy.deepcopy(prepared).to(cuda_dev)
    gpu_teacher = copy.deepcopy(teacher).to(cuda_dev)
    n0 = ops.launch_count()
    st0 = dict(installed.stats)
    l_gpu, s_gpu, _ = vr.distill_step(gpu_student, gpu_teacher, images.to(cuda_dev), labels.to(cuda_dev), None, hp, clip=False)
    torch.cuda.synchronize()
    assert ops.launch_count() - n0 > 100            # the sm_100a kernels ran, not ATen's
    st = {k: installed.stats[k] - st0[k] for k in st0}
    # every fake-quant module of the prepared student took a native route: block Linears on the tensor-core GEMM, the patch
    # embedding as im2col + GEMM, the 10-class head on the small-Linear kernels -- nothing left on cuDNN / cuBLAS
    assert st["linear_stock"] == 0 and st["conv_stock"] == 0, st
    assert st["conv_gemm"] == 1 and st["linear_small"] == 1 and st["linear_gemm"] == 4 * len(prepared.model.blocks), st
    l_cpu, s_cpu, _ = vr.distill_step(prepared, teacher, images, labels, None, hp, clip=False)
    assert abs(float(l_gpu) - float(l_cpu)) < 2e-2 * abs(float(l_cpu))
    g_cpu = torch.cat([p.grad.flatten() for p in prepared.parameters()])
    g_gpu = torch.cat([p.grad.flatten().cpu() for p in gpu_student.parameters()])
    assert rel_l2(g_gpu, g_cpu) < 0.15
    # weight observers see identical weights -> bit-exact state; module types untouched
    ref_sd, sd = prepared.state_dict(), gpu_student.state_dict()
    assert list(sd.keys()) == list(ref_sd.keys())
    for k in ref_sd:
        if "weight_fake_quant" in k or k.startswith("quant."):
            assert torch.equal(sd[k].cpu(), ref_sd[k]), k
    for (n1, m1), (n2, m2) in zip(prepared.named_modules(), gpu_student.named_modules()):
        assert type(m1) is type(m2)
    converted = convert(copy.deepcopy(gpu_student).cpu().eval(), inplace=False)
    assert any(k.endswith("_packed_params._packed_params") for k in converted.state_dict())
