"""BASELINE.json configs[1] at FULL size (ViT-B/16 teacher -> ViT-S/16 QAT student, batch 256, 224x224) through size-independent
properties -- the CPU oracle cannot run this size in test time, so: determinism (two runs from the same state are bit-identical),
conservation laws of the loss gradient, the bias-gradient = column-sum identity, fake-quant idempotence on a full activation
tensor, and agreement of the big step with the same images run as two half batches wherever the arithmetic is batch-local."""
import copy
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def big(cuda_dev):
    import bench
    from qatvit_b200.engine import QATDistillStep
    B = 256
    student, teacher = bench.build_models(B, cuda_dev)
    g = torch.Generator().manual_seed(5)
    images = torch.randn(B, 3, 224, 224, generator=g).to(cuda_dev)
    labels = torch.randint(0, 10, (B,), generator=g).to(cuda_dev)
    return dict(B=B, student=student, teacher=teacher, images=images, labels=labels, hp=bench.HP, Step=QATDistillStep)


def test_full_size_step_is_deterministic_and_conservative(cuda_dev, big):
    B = big["B"]
    runs = []
    for _ in range(2):
        s = copy.deepcopy(big["student"])
        step = big["Step"](s, big["teacher"], B, big["hp"])
        out3 = step(big["images"], big["labels"])
        torch.cuda.synchronize()
        runs.append((out3.clone(), step.grad_arena.clone(), step, s))
    (la, ga, step, s), (lb, gb, _, _) = runs
    assert torch.isfinite(la).all() and torch.isfinite(ga).all()
    assert torch.equal(la, lb) and torch.equal(ga, gb)                       # same state, same inputs -> same bits
    se = step.student_engine
    # dL/ds = a T (p_s - p_t)/B + (1-a)(softmax(s) - q)/B: every row sums to zero (before the head's STE mask zeroes entries)
    hd = se.head
    from qatvit_b200 import ops
    _, mask = ops.fq_apply(se.logits_raw, hd.afq.scale, hd.afq.zero_point, hd.afq.fake_quant_enabled, hd.afq.qmin, hd.afq.qmax)
    rows_unmasked = mask.bool().all(dim=1)
    assert int(rows_unmasked.sum()) > B // 2
    assert float(se.g_logits[rows_unmasked].sum(dim=1).abs().max()) < 1e-6
    # loss3 = [total, kd, ce] with total = a kd + (1 - a) ce
    a = big["hp"]["kd_alpha"]
    assert abs(float(la[0]) - (a * float(la[1]) + (1 - a) * float(la[2]))) < 1e-5 * abs(float(la[0]))
    # head: weight.grad = mask_w * g^T xn, bias.grad = column sums of g  (fp64 check of the two small reductions)
    gW = dict(s.named_parameters())["model.head.weight"].grad
    gb_ = dict(s.named_parameters())["model.head.bias"].grad
    want_b = se.g_logits.double().sum(0)
    assert float((gb_.double() - want_b).abs().max()) <= 1e-5 * float(want_b.abs().max())
    want_W = (se.g_logits.double().t() @ se.xn.double()) * hd.wmask.double()
    assert float((gW.double() - want_W).abs().max()) <= 1e-5 * float(want_W.abs().max())
    # a block's qkv bias gradient = column sums of its gradient planes divided by the weight scale (fused in attention backward)
    ql = se.lin[0]["qkv"]
    planes = se.gp3.double().sum(0) / ql.wscale_vec.double()                 # hi + lo, un-fold the per-channel scale
    got = dict(s.named_parameters())["model.blocks.0.attn.qkv.bias"].grad.double()
    assert float((got - planes.sum(0)).abs().max()) <= 2e-4 * float(planes.sum(0).abs().max())


def test_full_size_fake_quant_is_idempotent(cuda_dev, big):
    """FQ(FQ(x)) == FQ(x) with the same (scale, zero_point): checked on a full qkv activation (50 432 x 1 152) through the code
    plane the attention kernels consume."""
    from qatvit_b200 import ops
    s = copy.deepcopy(big["student"])
    step = big["Step"](s, big["teacher"], big["B"], big["hp"])
    step(big["images"], big["labels"])
    se = step.student_engine
    fq = se.lin[5]["qkv"].afq
    raw = se.qkv_raw[5]
    codes = torch.empty(1, *raw.shape, dtype=torch.bfloat16, device=cuda_dev)
    ops.act_planes(raw, fq.q, False, codes, codes_only=True)
    y = codes[0].float() * fq.scale                                          # FQ(x) = code * scale
    codes2 = torch.empty_like(codes)
    ops.act_planes(y.contiguous(), fq.q, False, codes2, codes_only=True)
    torch.cuda.synchronize()
    assert torch.equal(codes, codes2)
    lo, hi = fq.qmin - int(fq.zero_point), fq.qmax - int(fq.zero_point)
    assert int(codes.float().min()) >= lo and int(codes.float().max()) <= hi
    assert torch.equal(codes, se.qkvc[5])                                    # and it is what the step itself produced


def test_full_size_teacher_is_batch_local(cuda_dev, big):
    """The frozen teacher has no batch statistics: logits of the 256-image batch == logits of its two halves run separately
    (different tile schedules, same per-row arithmetic -> bit-identical)."""
    from qatvit_b200.engine import TeacherEngine
    B = big["B"]
    full = TeacherEngine(big["teacher"], B).forward(big["images"]).clone()
    half = TeacherEngine(big["teacher"], B // 2)
    lo = half.forward(big["images"][:B // 2].contiguous()).clone()
    hi = half.forward(big["images"][B // 2:].contiguous()).clone()
    torch.cuda.synchronize()
    assert torch.equal(full, torch.cat([lo, hi]))
