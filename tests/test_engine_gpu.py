"""GPU parity (tier 2, <= 1e-3 relative): the fused QAT-distillation step vs the reference path on CPU
(stock torch.ao prepare_qat + restated timm ViT + restated step body, oracle/vit_ref.py) on identical
inputs and weights.  Weight observers / codes are compared bit-exactly (identical inputs), activation
observers to 1e-5 relative (GEMM summation order differs)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _build(backend, batch, student_name, teacher_name, img, seed=0):
    from oracle import vit_ref as vr
    torch.manual_seed(seed)
    kw = dict(img_size=img) if img != 224 else {}
    student = vr.qat_wrapper_cls(prefer_reference=False)(vr.create_model(student_name, num_classes=10, **kw))
    torch.manual_seed(seed + 1)
    teacher = vr.create_model(teacher_name, num_classes=10, **kw)
    with torch.no_grad():
        teacher.head.weight.mul_(8.0)
        # a random-init ViT has almost no signal: widen a few things so every path carries gradient
        for p in student.parameters():
            if p.dim() == 1:
                p.add_(0.02 * torch.randn_like(p))
    teacher.eval()
    for p in teacher.parameters():
        p.requires_grad = False
    prepared = vr.enable_qat(student, backend)
    images, labels = vr.synthetic_batch(batch, seed=3, img=img)
    return vr, prepared, teacher, images, labels


@pytest.mark.parametrize("backend", ["fbgemm", "qnnpack"])
def test_tiny_step_matches_reference(cuda_dev, backend):
    from qatvit_b200.engine import QATDistillStep
    vr, prepared, teacher, images, labels = _build(backend, 4, "vit_test_tiny", "vit_test_teacher", 64)
    hp = dict(vr.DEFAULT_HPARAMS)
    gpu_student = copy.deepcopy(prepared).to(cuda_dev)
    gpu_teacher = copy.deepcopy(teacher).to(cuda_dev)
    step = QATDistillStep(gpu_student, gpu_teacher, 4, hp)
    for it in range(2):     # second iteration exercises the EMA branch of every observer
        loss_ref, s_ref, t_ref = vr.distill_step(prepared, teacher, images, labels, None, hp, clip=False)
        ref_grads = {n: p.grad.clone() for n, p in prepared.named_parameters()}
        prepared.zero_grad(set_to_none=True)
        out3 = step(images.to(cuda_dev), labels.to(cuda_dev))
        torch.cuda.synchronize()
        # teacher logits (fp32 path on bf16x3 tensor cores)
        assert _rel(step.teacher_engine.logits, t_ref) < 1e-3
        assert abs(float(out3[0]) - float(loss_ref)) <= 1e-3 * abs(float(loss_ref))
        # student logits = fake-quantised head output
        hd = step.student_engine.head
        from qatvit_b200 import ops
        s_gpu, _ = ops.fq_apply(step.student_logits_raw, hd.afq.scale, hd.afq.zero_point, hd.afq.fake_quant_enabled,
                                hd.afq.qmin, hd.afq.qmax)
        assert _rel(s_gpu, s_ref) < 1e-3 + 1.01 * float(hd.afq.scale) / float(s_ref.abs().max())   # <= one code step
        # every parameter gradient
        worst = ("", 0.0)
        for n, p in gpu_student.named_parameters():
            r = _rel(p.grad, ref_grads[n])
            if r > worst[1]:
                worst = (n, r)
        assert worst[1] < 2e-3, f"iteration {it}: gradient mismatch {worst}"
        # observer state
        ref_sd, gpu_sd = prepared.state_dict(), gpu_student.state_dict()
        assert list(ref_sd.keys()) == list(gpu_sd.keys())
        for k in ref_sd:
            a, b = gpu_sd[k].cpu(), ref_sd[k]
            assert a.shape == b.shape and a.dtype == b.dtype, k
            if "weight_fake_quant" in k:
                assert torch.equal(a, b), f"weight observer state differs: {k}"
            elif k.endswith(("min_val", "max_val", "scale")):
                assert _rel(a, b) < 1e-4, k
            elif k.endswith("zero_point"):
                assert (a.long() - b.long()).abs().max() <= 1, k


def test_state_dict_and_convert_flow_unchanged(cuda_dev):
    """best_qat.pth / convert() flow (ref qat_trainer.py:376-388) still works on a model trained by the engine."""
    from torch.ao.quantization import convert
    from qatvit_b200.engine import QATDistillStep
    vr, prepared, teacher, images, labels = _build("fbgemm", 2, "vit_test_tiny", "vit_test_teacher", 64)
    ref_keys = list(prepared.state_dict().keys())
    gpu_student = copy.deepcopy(prepared).to(cuda_dev)
    step = QATDistillStep(gpu_student, copy.deepcopy(teacher).to(cuda_dev), 2, dict(vr.DEFAULT_HPARAMS))
    step(images.to(cuda_dev), labels.to(cuda_dev))
    torch.cuda.synchronize()
    sd = gpu_student.state_dict()
    assert list(sd.keys()) == ref_keys
    base = copy.deepcopy(gpu_student).cpu().eval()
    converted = convert(base, inplace=False)
    csd = converted.state_dict()
    assert any(k.endswith("_packed_params._packed_params") for k in csd)
    # the converted int8 weights come from the observer state our kernels produced
    w, _ = csd["model.blocks.0.attn.qkv._packed_params._packed_params"]
    assert w.dtype == torch.qint8 and w.int_repr().abs().max() > 0
