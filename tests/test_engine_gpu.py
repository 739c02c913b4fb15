"""GPU parity of the fused QAT-distillation step (teacher fwd + student fwd + KL/CE + hand-written backward)
against the reference path on CPU: stock torch.ao prepare_qat + restated timm ViT + restated step body
(oracle/vit_ref.py, following ref/src/training/qat_trainer.py:300-316,337-361).

Tolerances (north_star): bit-exact observer scale / zero-point / integer codes on identical inputs,
<= 1e-3 relative on loss, logits and every gradient.  See tests/parity_utils.py for the forcing protocol."""
import copy

import pytest
import torch

from parity_utils import build_models, engine_raw_tensors, install_forcing_hooks, rel_l2, rel_max

pytestmark = pytest.mark.gpu

CASES = [
    # (backend, student, teacher, img, batch, fused attention (integer-code tcgen05 kernels) or the unfused fallback,
    #  LayerNorm variant: "subclass" = timm.layers.LayerNorm, not observed (101 fake-quant modules);
    #                     "plain" = nn.LayerNorm, observed by prepare_qat (126 fake-quant modules) -- SURVEY.md §0.6)
    ("fbgemm", "vit_test_tiny", "vit_test_teacher", 64, 4, True, "subclass"),
    ("qnnpack", "vit_test_tiny", "vit_test_teacher", 64, 3, True, "subclass"),
    ("fbgemm", "vit_test_tiny", "vit_test_teacher", 96, 5, True, "subclass"),          # 37 tokens: ragged everything
    ("fbgemm", "vit_test_tiny", "vit_test_teacher", 96, 3, False, "subclass"),
    ("fbgemm", "vit_test_tiny", "vit_test_teacher", 64, 4, True, "plain"),
    ("qnnpack", "vit_test_tiny", "vit_test_teacher", 96, 3, True, "plain"),
    ("fbgemm", "vit_small_patch16_224", "vit_base_patch16_224", 224, 8, True, "subclass"),   # BASELINE.json config 1
    ("fbgemm", "vit_small_patch16_224", "vit_base_patch16_224", 224, 4, True, "plain"),
]


@pytest.mark.parametrize("backend,sname,tname,img,B,fused,ln_variant", CASES)
def test_forced_parity_with_reference(cuda_dev, backend, sname, tname, img, B, fused, ln_variant):
    from qatvit_b200.engine import QATDistillStep
    vr, prepared, teacher = build_models(backend, sname, tname, img, ln_variant=ln_variant)
    n_fq = sum(1 for m in prepared.modules() if type(m).__name__ == "FusedMovingAvgObsFakeQuantize")
    L = len(prepared.model.blocks)
    assert n_fq == (8 * L + 5 if ln_variant == "subclass" else 10 * L + 6)       # 101 / 126 at depth 12
    images, labels = vr.synthetic_batch(B, seed=3, img=img)
    hp = dict(vr.DEFAULT_HPARAMS)
    gpu_student = copy.deepcopy(prepared).to(cuda_dev)
    step = QATDistillStep(gpu_student, copy.deepcopy(teacher).to(cuda_dev), B, hp, fused_attention=fused)
    assert step.student_engine.fused_attn == fused and step.student_engine.ln_obs == (ln_variant == "plain")
    for it in range(2):                       # 2nd iteration: EMA branch of every observer, new images
        if it == 1:
            images, labels = vr.synthetic_batch(B, seed=11, img=img)
            images = images * 1.3 + 0.2
        out3 = step(images.to(cuda_dev), labels.to(cuda_dev))
        torch.cuda.synchronize()
        se = step.student_engine
        stage_err = {}
        handles = install_forcing_hooks(prepared, engine_raw_tensors(se), stage_err)
        loss_ref, s_ref, t_ref = vr.distill_step(prepared, teacher, images, labels, None, hp, clip=False)
        for h in handles:
            h.remove()
        # teacher: plain fp32 forward (bf16x3 tensor-core GEMMs on our side)
        assert rel_max(step.teacher_engine.logits, t_ref) < 1e-3
        # every stage of the student forward, given identical upstream codes
        worst = max(stage_err, key=stage_err.get)
        assert stage_err[worst] < 1e-4, (worst, stage_err[worst])
        assert abs(float(out3[0]) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
        # student logits after the head's fake-quant: identical codes -> identical values up to the scale's last bit
        from qatvit_b200 import ops
        hd = se.head
        s_gpu, _ = ops.fq_apply(se.logits_raw, hd.afq.scale, hd.afq.zero_point, hd.afq.fake_quant_enabled, hd.afq.qmin,
                                hd.afq.qmax)
        assert rel_max(s_gpu, s_ref) < 1e-5
        # every parameter gradient (max-norm AND l2) within 1e-3
        for n, p in gpu_student.named_parameters():
            ref_g = dict(prepared.named_parameters())[n].grad
            assert rel_max(p.grad, ref_g) < 1e-3 and rel_l2(p.grad, ref_g) < 1e-3, (it, n, rel_max(p.grad, ref_g))
        prepared.zero_grad(set_to_none=True)
        # observer state: weights bit-exact (identical inputs); activations decided on (near-)identical raw tensors
        ref_sd, gpu_sd = prepared.state_dict(), gpu_student.state_dict()
        assert list(ref_sd.keys()) == list(gpu_sd.keys())
        for k in ref_sd:
            a, b = gpu_sd[k].cpu(), ref_sd[k]
            assert a.shape == b.shape and a.dtype == b.dtype, k
            # every observer of the reference saw exactly the tensor ours saw (weights: identical parameters; activations: the
            # value-exact forcing hook), so running min / max, scale and zero_point must agree BIT FOR BIT (north_star)
            if k.endswith(("min_val", "max_val", "scale", "zero_point", "observer_enabled", "fake_quant_enabled")):
                assert torch.equal(a, b), f"observer state differs: {k}: {a.flatten()[:4]} vs {b.flatten()[:4]}"


def test_free_running_divergence_is_at_reference_noise_floor(cuda_dev):
    """Without forcing, our path vs torch-CPU must not diverge more than torch-CUDA (the reference's own GPU path)
    diverges from torch-CPU on the same model -- that is the reproducibility floor of eager-mode QAT."""
    from qatvit_b200.engine import QATDistillStep
    vr, prepared, teacher = build_models("fbgemm", "vit_test_tiny", "vit_test_teacher", 64)
    B = 8
    images, labels = vr.synthetic_batch(B, seed=5, img=64)
    hp = dict(vr.DEFAULT_HPARAMS)
    ref_cpu, ref_gpu, ours = copy.deepcopy(prepared), copy.deepcopy(prepared).to(cuda_dev), copy.deepcopy(prepared).to(cuda_dev)
    t_gpu = copy.deepcopy(teacher).to(cuda_dev)
    l_cpu, s_cpu, _ = vr.distill_step(ref_cpu, teacher, images, labels, None, hp, clip=False)
    l_gpu, s_gpu, _ = vr.distill_step(ref_gpu, t_gpu, images.to(cuda_dev), labels.to(cuda_dev), None, hp, clip=False)
    step = QATDistillStep(ours, t_gpu, B, hp)
    out3 = step(images.to(cuda_dev), labels.to(cuda_dev))
    torch.cuda.synchronize()
    g_cpu = torch.cat([p.grad.flatten() for p in ref_cpu.parameters()])
    g_gpu = torch.cat([p.grad.flatten().cpu() for p in ref_gpu.parameters()])
    g_ours = torch.cat([p.grad.flatten().cpu() for p in ours.parameters()])
    floor = rel_l2(g_gpu, g_cpu)
    mine = rel_l2(g_ours, g_cpu)
    assert mine < max(3.0 * floor, 2e-2), (mine, floor)
    assert abs(float(out3[0]) - float(l_cpu)) < max(3.0 * abs(float(l_gpu) - float(l_cpu)), 1e-2 * abs(float(l_cpu)))


def test_state_dict_and_convert_flow_unchanged(cuda_dev):
    """best_qat.pth / convert() flow (ref qat_trainer.py:376-388) still works on a model trained by the engine."""
    from torch.ao.quantization import convert
    from qatvit_b200.engine import QATDistillStep
    vr, prepared, teacher = build_models("fbgemm", "vit_test_tiny", "vit_test_teacher", 64)
    images, labels = vr.synthetic_batch(2, seed=3, img=64)
    ref_keys = list(prepared.state_dict().keys())
    gpu_student = copy.deepcopy(prepared).to(cuda_dev)
    step = QATDistillStep(gpu_student, copy.deepcopy(teacher).to(cuda_dev), 2, dict(vr.DEFAULT_HPARAMS))
    opt = vr.make_optimizer(gpu_student.parameters(), vr.DEFAULT_HPARAMS, 0.5)
    w0 = gpu_student.model.blocks[0].attn.qkv.weight.detach().clone()
    for _ in range(2):
        step(images.to(cuda_dev), labels.to(cuda_dev))
        torch.nn.utils.clip_grad_norm_(gpu_student.parameters(), 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        step.student_engine.attach_grads()
    torch.cuda.synchronize()
    assert not torch.equal(w0, gpu_student.model.blocks[0].attn.qkv.weight.detach())
    sd = gpu_student.state_dict()
    assert list(sd.keys()) == ref_keys
    base = copy.deepcopy(gpu_student).cpu().eval()
    converted = convert(base, inplace=False)
    csd = converted.state_dict()
    assert any(k.endswith("_packed_params._packed_params") for k in csd)
    w, _ = csd["model.blocks.0.attn.qkv._packed_params._packed_params"]
    assert w.dtype == torch.qint8 and w.int_repr().abs().max() > 0


@pytest.mark.parametrize("backend,ln_variant", [("fbgemm", "subclass"), ("qnnpack", "plain")])
def test_predict_matches_reference_forward(cuda_dev, backend, ln_variant):
    """Forward only (the reference's evaluate_fp32 loop, qat_trainer.py:49-61): engine.predict(images) == prepared(images) on CPU
    with identical codes forced; two calls, so the EMA branch of every observer is covered; no gradient is touched."""
    from qatvit_b200.engine import QATDistillStep
    vr, prepared, teacher = build_models(backend, "vit_test_tiny", "vit_test_teacher", 64, ln_variant=ln_variant)
    B = 5
    gpu_student = copy.deepcopy(prepared).to(cuda_dev)
    step = QATDistillStep(gpu_student, copy.deepcopy(teacher).to(cuda_dev), B, dict(vr.DEFAULT_HPARAMS))
    step.grad_arena.fill_(7.0)
    for it in range(2):
        images, _ = vr.synthetic_batch(B, seed=20 + it, img=64)
        y = step.predict(images.to(cuda_dev))
        torch.cuda.synchronize()
        stage_err = {}
        handles = install_forcing_hooks(prepared, engine_raw_tensors(step.student_engine), stage_err)
        with torch.no_grad():
            y_ref = prepared(images)
        for h in handles:
            h.remove()
        assert max(stage_err.values()) < 1e-4
        assert rel_max(y, y_ref) < 1e-5
    assert bool((step.grad_arena == 7.0).all())
    ref_sd, gpu_sd = prepared.state_dict(), gpu_student.state_dict()
    for k in ref_sd:
        if k.endswith(("min_val", "max_val", "scale", "zero_point")):
            assert torch.equal(gpu_sd[k].cpu(), ref_sd[k]), k


@pytest.mark.parametrize("sname,tname,img,B,steps", [("vit_test_tiny", "vit_test_teacher", 64, 4, 12),
                                                     ("vit_small_patch16_224", "vit_base_patch16_224", 224, 8, 3)])
def test_side_streams_change_nothing(cuda_dev, monkeypatch, sname, tname, img, B, steps):
    """The teacher forward and the weight-gradient GEMMs run on side streams (event-ordered).  Every kernel is deterministic, so
    a run with the overlap on must equal a run with everything on one stream BIT FOR BIT -- loss, every gradient and every
    observer buffer, step after step with the optimizer in the loop (a missing wait shows up as a difference)."""
    from qatvit_b200.engine import QATDistillStep
    from qatvit_b200.optim import FusedClipAdamW
    vr, prepared, teacher = build_models("fbgemm", sname, tname, img)
    hp = dict(vr.DEFAULT_HPARAMS)
    runs = []
    for overlap in ("1", "0"):
        monkeypatch.setenv("QV_OVERLAP_TEACHER", overlap)
        monkeypatch.setenv("QV_OVERLAP_WGRAD", overlap)
        student = copy.deepcopy(prepared).to(cuda_dev)
        step = QATDistillStep(student, copy.deepcopy(teacher).to(cuda_dev), B, hp)
        assert step.overlap_teacher == (overlap == "1") and (step.student_engine._wstream is not None) == (overlap == "1")
        opt = FusedClipAdamW(student.parameters(), step.grad_arena, lr=1e-3, weight_decay=hp["weight_decay"], max_norm=1.0)
        trace = []
        for it in range(steps):
            images, labels = vr.synthetic_batch(B, seed=100 + it, img=img)
            out3 = step(images.to(cuda_dev), labels.to(cuda_dev))
            trace.append((out3.clone(), step.grad_arena.clone()))
            opt.step()
        torch.cuda.synchronize()
        runs.append((trace, {k: v.clone() for k, v in student.state_dict().items()}))
    (ta, sa), (tb, sb) = runs
    for it, ((la, ga), (lb, gb)) in enumerate(zip(ta, tb)):
        assert torch.equal(la, lb), it
        assert torch.equal(ga, gb), (it, int((ga != gb).sum()))
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
