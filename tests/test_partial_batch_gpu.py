"""Ragged batches: the reference's DataLoaders have no drop_last (ref/src/training/qat_trainer.py:227-235 train, :237-254
eval), so the last batch of every epoch is smaller (50 000 % 256 = 80 train images, 10 000 % 256 = 16 test images).  An
engine built for batch B must take any b <= B and compute, BIT FOR BIT, what an engine built for b computes: same kernels,
same tile schedules and split-K factors, the buffers re-viewed as the contiguous tensors of the smaller engine."""
import copy

import pytest
import torch

from parity_utils import build_models

pytestmark = pytest.mark.gpu


def _state(student):
    return {k: v.detach().clone() for k, v in student.state_dict().items()}


def _assert_same_state(a, b):
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("backend,ln_variant,fused,img,B,b", [
    ("fbgemm", "subclass", True, 64, 8, 5),
    ("qnnpack", "plain", True, 96, 6, 1),           # observed LayerNorm, 37 tokens, a single-image tail
    ("fbgemm", "subclass", False, 64, 7, 3),        # unfused attention fallback (zero-padded probability planes)
])
def test_tail_batch_equals_engine_built_for_it(cuda_dev, backend, ln_variant, fused, img, B, b):
    """[B images, then b images] on an engine built for B  ==  [B images] on an engine built for B, its state handed to an
    engine built for b, [b images] there: loss, every gradient and every observer buffer bit for bit.  The first (full) step
    leaves stale rows of a larger batch behind every tail view."""
    from qatvit_b200.engine import QATDistillStep
    vr, prepared, teacher = build_models(backend, "vit_test_tiny", "vit_test_teacher", img, ln_variant=ln_variant)
    hp = dict(vr.DEFAULT_HPARAMS)
    t_gpu = copy.deepcopy(teacher).to(cuda_dev)
    im_full, lb_full = vr.synthetic_batch(B, seed=31, img=img)
    im_tail, lb_tail = vr.synthetic_batch(b, seed=32, img=img)
    im_tail = im_tail * 1.4 - 0.3
    im_full, lb_full, im_tail, lb_tail = (t.to(cuda_dev) for t in (im_full, lb_full, im_tail, lb_tail))

    s_a = copy.deepcopy(prepared).to(cuda_dev)
    step_a = QATDistillStep(s_a, t_gpu, B, hp, fused_attention=fused)
    step_a(im_full, lb_full)
    loss_a = step_a(im_tail, lb_tail).clone()
    grads_a = step_a.grad_arena.clone()
    logits_a = step_a.student_logits_raw.clone()
    t_logits_a = step_a.teacher_engine.logits.clone()
    torch.cuda.synchronize()
    assert logits_a.shape[0] == b and t_logits_a.shape[0] == b

    s_b = copy.deepcopy(prepared).to(cuda_dev)
    step_b = QATDistillStep(s_b, t_gpu, B, hp, fused_attention=fused)
    step_b(im_full, lb_full)
    torch.cuda.synchronize()
    s_c = copy.deepcopy(s_b)                                 # state after the full step -> an engine built for the tail size
    step_c = QATDistillStep(s_c, t_gpu, b, hp, fused_attention=fused)
    loss_c = step_c(im_tail, lb_tail).clone()
    torch.cuda.synchronize()

    assert torch.isfinite(loss_a).all() and float(grads_a.abs().max()) > 0
    assert torch.equal(loss_a, loss_c)
    assert torch.equal(logits_a, step_c.student_logits_raw)
    assert torch.equal(t_logits_a, step_c.teacher_engine.logits)
    assert torch.equal(grads_a, step_c.grad_arena), int((grads_a != step_c.grad_arena).sum())
    _assert_same_state(_state(s_a), _state(s_c))
    # ... and the engine goes back to full batches afterwards
    loss_a2 = step_a(im_full, lb_full).clone()
    step_c2 = QATDistillStep(copy.deepcopy(s_c), t_gpu, B, hp, fused_attention=fused)
    loss_c2 = step_c2(im_full, lb_full).clone()
    torch.cuda.synchronize()
    assert torch.equal(loss_a2, loss_c2) and torch.equal(step_a.grad_arena, step_c2.grad_arena)


def test_epoch_with_ragged_tail_through_the_reference_loop(cuda_dev):
    """INTEGRATION.md section 2(b)'s loop, unchanged, over an 'epoch' of 8 + 5 images with the reference's own optimizer calls
    (optimizer.zero_grad(set_to_none=True) every iteration, ref qat_trainer.py:351; clip_grad_norm_, :360; AdamW.step, :361),
    then the validation loop (ref :49-61) over 8 + 3 images through predict()."""
    from qatvit_b200.engine import QATDistillStep
    vr, prepared, teacher = build_models("fbgemm", "vit_test_tiny", "vit_test_teacher", 64)
    base = copy.deepcopy(prepared).to(cuda_dev)
    hp = dict(vr.DEFAULT_HPARAMS)
    step = QATDistillStep(base, copy.deepcopy(teacher).to(cuda_dev), 8, hp)
    optimizer = vr.make_optimizer(base.parameters(), hp, 0.5)
    images, labels = vr.synthetic_batch(13, seed=41, img=64)
    loader = [(images[:8], labels[:8]), (images[8:], labels[8:])]
    w_prev = base.model.blocks[1].mlp.fc1.weight.detach().clone()
    for im, lb in loader:
        im, lb = im.to(cuda_dev, non_blocking=True), lb.to(cuda_dev, non_blocking=True)
        optimizer.zero_grad(set_to_none=True)
        loss3 = step(im, lb)
        assert all(p.grad is not None for p in base.parameters())          # zero_grad(set_to_none) dropped the views; re-attached
        total = torch.nn.utils.clip_grad_norm_(base.parameters(), 1.0)
        optimizer.step()
        torch.cuda.synchronize()
        assert torch.isfinite(loss3).all() and float(total) > 0
        w_now = base.model.blocks[1].mlp.fc1.weight.detach().clone()
        assert not torch.equal(w_now, w_prev)                                # the optimizer really stepped on this batch
        w_prev = w_now
    val_images, _ = vr.synthetic_batch(11, seed=42, img=64)
    outs = [step.predict(chunk.to(cuda_dev)).clone() for chunk in (val_images[:8], val_images[8:])]
    torch.cuda.synchronize()
    assert outs[0].shape == (8, 10) and outs[1].shape == (3, 10) and all(torch.isfinite(o).all() for o in outs)
    with pytest.raises(RuntimeError, match="built for batch 8"):
        step.predict(torch.zeros(9, 3, 64, 64, device=cuda_dev))


def test_pre_qat_step_takes_a_tail_batch(cuda_dev):
    """Same property for the pre-QAT engine (epochs before qat_start_epoch, ref qat_trainer.py:320)."""
    from qatvit_b200.plain import PlainDistillStep
    from oracle import vit_ref as vr
    torch.manual_seed(0)
    student = vr.qat_wrapper_cls(prefer_reference=False)(vr.create_model("vit_test_tiny", num_classes=10, img_size=64)).train()
    teacher = vr.create_model("vit_test_teacher", num_classes=10, img_size=64).eval()
    hp = dict(vr.DEFAULT_HPARAMS)
    im_full, lb_full = vr.synthetic_batch(6, seed=51, img=64)
    im_tail, lb_tail = vr.synthetic_batch(4, seed=52, img=64)
    im_full, lb_full, im_tail, lb_tail = (t.to(cuda_dev) for t in (im_full, lb_full, im_tail, lb_tail))
    t_gpu = copy.deepcopy(teacher).to(cuda_dev)
    step_a = PlainDistillStep(copy.deepcopy(student).to(cuda_dev), t_gpu, 6, hp)
    step_a(im_full, lb_full)
    loss_a = step_a(im_tail, lb_tail).clone()
    step_c = PlainDistillStep(copy.deepcopy(student).to(cuda_dev), t_gpu, 4, hp)
    loss_c = step_c(im_tail, lb_tail).clone()
    torch.cuda.synchronize()
    assert torch.equal(loss_a, loss_c) and torch.equal(step_a.grad_arena, step_c.grad_arena)
    assert torch.equal(step_a.student_engine.logits, step_c.student_engine.logits)
