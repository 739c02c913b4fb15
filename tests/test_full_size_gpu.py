"""BASELINE.json configs[1] at FULL size (ViT-B/16 teacher -> ViT-S/16 QAT student, batch 256, 224x224).

Against the oracle where the CPU can follow in seconds: the teacher's logits on the first 8 images of the 256-image batch vs the
CPU fp32 teacher on those images (the teacher is batch-local), and a FORWARD-ONLY forced-parity pass of the student at batch
256 (every stage, the fake-quantised logits and every observer buffer; the CTA-pair GEMMs, taken only for M >= 18 944 rows,
are thereby compared with the oracle inside the engine).  Through size-independent properties for the full step: determinism
(two runs from the same state are bit-identical), conservation laws of the loss gradient, the bias-gradient = column-sum
identity, fake-quant idempotence on a full activation tensor, batch-locality of the teacher."""
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


def _oracle_teacher(big):
    """The CPU side: oracle/vit_ref.py's restated timm ViT-B (stock ATen ops) holding the bench teacher's weights."""
    from oracle import vit_ref as vr
    ref = vr.create_model("vit_base_patch16_224", num_classes=10).eval()
    ref.load_state_dict({k: v.detach().cpu() for k, v in big["teacher"].state_dict().items()})
    return ref


def test_full_size_teacher_rows_match_cpu_fp32(cuda_dev, big):
    """Rows 0..7 of the batch-256 teacher logits (mixed-format CTA-pair GEMMs at M = 50 432) vs the CPU fp32 teacher on those 8
    images: north_star tolerance 1e-3 relative, expected ~1e-4."""
    from parity_utils import rel_max
    from qatvit_b200.engine import TeacherEngine
    from qatvit_b200 import ops
    n0 = ops.gemm_pair_launches()
    eng = TeacherEngine(big["teacher"], big["B"])
    full = eng.forward(big["images"]).clone()
    torch.cuda.synchronize()
    assert ops.gemm_pair_launches() > n0                       # this size really takes the cta_group::2 kernels
    assert eng.saturation_events == []
    with torch.no_grad():
        ref = _oracle_teacher(big)(big["images"][:8].cpu())
    err = rel_max(full[:8], ref)
    assert err < 1e-3, err
    print(f"teacher B=256 rows 0..7 vs CPU fp32: rel_max {err:.2e}")


def test_full_size_student_forward_forced_parity(cuda_dev, big):
    """Forward of the prepared student at batch 256 vs the reference path on the host CPU (stock prepare_qat + live ATen ops,
    oracle/vit_ref.py) with our raw tensors forced value-exactly into every activation fake-quant: stage error < 1e-4,
    fake-quantised logits < 1e-5, and EVERY observer buffer (weights and activations: running min / max, scale, zero_point)
    bit for bit."""
    from parity_utils import engine_raw_tensors, install_forcing_hooks, rel_max
    from oracle import vit_ref as vr
    B = big["B"]
    s = copy.deepcopy(big["student"])
    sd0 = {k: v.detach().cpu().clone() for k, v in s.state_dict().items()}        # before the engine resizes per-channel buffers
    ref = vr.enable_qat(vr.make_student(prefer_reference=False), "fbgemm")
    ref.load_state_dict(sd0)
    step = big["Step"](s, big["teacher"], B, big["hp"])
    y = step.predict(big["images"]).clone()
    torch.cuda.synchronize()
    stage_err = {}
    handles = install_forcing_hooks(ref, engine_raw_tensors(step.student_engine, lazy=True), stage_err)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        y_ref = ref(big["images"].cpu())
    for h in handles:
        h.remove()
    worst = max(stage_err, key=stage_err.get)
    assert len(stage_err) == 50 and stage_err[worst] < 1e-4, (worst, stage_err[worst])
    assert rel_max(y, y_ref) < 1e-5
    ref_sd, gpu_sd = ref.state_dict(), s.state_dict()
    assert list(ref_sd.keys()) == list(gpu_sd.keys())
    n_obs = 0
    for k in ref_sd:
        if k.endswith(("min_val", "max_val", "scale", "zero_point")):
            assert torch.equal(gpu_sd[k].cpu(), ref_sd[k]), k
            n_obs += 1
    assert n_obs == 4 * 101
    print(f"student B=256 forward forced parity: worst stage {worst} {stage_err[worst]:.2e}, {n_obs} observer buffers bit-exact")
