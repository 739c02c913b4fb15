"""Out-of-bounds WRITE check of the whole hot path (the pool's compute-sanitizer is closed, so the bounds check is ours).

Every device buffer an engine allocates (activations, planes, saved tensors, workspaces, the gradient arena, observer
accumulators) is carved out of a larger allocation with a 4 KB band of a known byte pattern on either side (plus the
round-up slack behind the payload).  After complete steps -- ragged shapes included: 17 / 37 tokens, odd batches, the
ViT-S / ViT-B widths -- every band must still hold the pattern: no kernel (TMA store boxes, vectorised tails, split-K
workspaces, per-slab partials) wrote a byte outside the tensor it was given.  Results must also be unchanged by the
re-homing of the buffers (same loss bits as an engine on ordinary allocations)."""
import copy
import math
import warnings

import pytest
import torch

from parity_utils import build_models

pytestmark = pytest.mark.gpu

GUARD = 4096
PATTERN = 0xA5


class GuardedAlloc:
    """Replaces torch.empty / torch.zeros for CUDA tensors while an engine is being constructed."""

    def __init__(self):
        self.bufs = []
        self._empty, self._zeros = torch.empty, torch.zeros

    def _carve(self, size, dtype, device, zero):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            size = tuple(size[0])
        dtype = dtype or torch.float32
        n = math.prod(size) * torch.empty((), dtype=dtype).element_size()
        n_pad = -(-max(n, 1) // 512) * 512
        raw = self._empty(GUARD + n_pad + GUARD, dtype=torch.uint8, device=device)
        raw.fill_(PATTERN)
        self.bufs.append((raw, n, tuple(size), dtype))
        t = raw[GUARD:GUARD + n].view(dtype).view(size)
        if zero:
            t.zero_()
        assert t.data_ptr() % (512 if t.is_cuda else 16) == 0 and t.is_contiguous()
        return t

    def _wrap(self, orig, zero):
        def fn(*size, dtype=None, device=None, **kw):
            if device is not None and torch.device(device).type == "cuda" and not kw:
                return self._carve(size, dtype, device, zero)
            return orig(*size, dtype=dtype, device=device, **kw)
        return fn

    def __enter__(self):
        torch.empty, torch.zeros = self._wrap(self._empty, False), self._wrap(self._zeros, True)
        return self

    def __exit__(self, *exc):
        torch.empty, torch.zeros = self._empty, self._zeros

    def check(self):
        torch.cuda.synchronize()
        bad = []
        for raw, n, size, dtype in self.bufs:
            front_ok = bool((raw[:GUARD] == PATTERN).all())
            back = raw[GUARD + n:]
            back_ok = bool((back == PATTERN).all())
            if not (front_ok and back_ok):
                first = int((back != PATTERN).nonzero()[0]) if not back_ok else -1
                bad.append((size, str(dtype), "front" if not front_ok else f"back +{first} B"))
        assert not bad, f"{len(bad)} of {len(self.bufs)} buffers written out of bounds: {bad[:8]}"
        return len(self.bufs)


QAT_CASES = [
    # backend, student, teacher, img, batch, fused attention, fused backward prologue, LayerNorm variant
    ("fbgemm", "vit_test_tiny", "vit_test_teacher", 64, 3, True, True, "subclass"),       # 17 tokens, odd batch
    ("qnnpack", "vit_test_tiny", "vit_test_teacher", 96, 5, True, True, "plain"),         # 37 tokens, observed LayerNorm
    ("fbgemm", "vit_test_tiny", "vit_test_teacher", 96, 3, False, False, "subclass"),     # unfused fallbacks
    ("fbgemm", "vit_small_patch16_224", "vit_base_patch16_224", 224, 3, True, True, "subclass"),
]


@pytest.mark.parametrize("backend,sname,tname,img,B,fused_attn,fused_gp,ln_variant", QAT_CASES)
def test_qat_step_writes_nothing_out_of_bounds(cuda_dev, backend, sname, tname, img, B, fused_attn, fused_gp, ln_variant):
    from qatvit_b200.engine import QATDistillStep
    vr, prepared, teacher = build_models(backend, sname, tname, img, ln_variant=ln_variant)
    hp = dict(vr.DEFAULT_HPARAMS)
    images, labels = vr.synthetic_batch(B, seed=5, img=img)
    images, labels = images.to(cuda_dev), labels.to(cuda_dev)
    plain = QATDistillStep(copy.deepcopy(prepared).to(cuda_dev), copy.deepcopy(teacher).to(cuda_dev), B, hp,
                           fused_attention=fused_attn, fused_gp=fused_gp)
    gs, gt = copy.deepcopy(prepared).to(cuda_dev), copy.deepcopy(teacher).to(cuda_dev)
    with GuardedAlloc() as ga:
        step = QATDistillStep(gs, gt, B, hp, fused_attention=fused_attn, fused_gp=fused_gp)
    assert len(ga.bufs) > 50
    for it in range(2):                                 # second step: EMA branch of the observers
        ref3 = plain(images, labels).clone()
        out3 = step(images, labels).clone()
        ga.check()
        assert torch.equal(out3, ref3), (it, out3, ref3)
    assert torch.equal(step.student_engine.grad_arena, plain.student_engine.grad_arena)
    step.predict(images)
    ga.check()


@pytest.mark.parametrize("sname,tname,img,B", [("vit_test_tiny", "vit_test_teacher", 96, 3),
                                               ("vit_small_patch16_224", "vit_base_patch16_224", 224, 2)])
def test_pre_qat_step_writes_nothing_out_of_bounds(cuda_dev, sname, tname, img, B):
    from oracle import vit_ref as vr
    from qatvit_b200.plain import PlainDistillStep
    torch.manual_seed(0)
    kw = dict(img_size=img) if img != 224 else {}
    student = vr.qat_wrapper_cls(prefer_reference=False)(vr.create_model(sname, num_classes=10, **kw)).train()
    teacher = vr.create_model(tname, num_classes=10, **kw).eval()
    for p in teacher.parameters():
        p.requires_grad = False
    hp = dict(vr.DEFAULT_HPARAMS)
    images, labels = vr.synthetic_batch(B, seed=5, img=img)
    gs, gt = student.to(cuda_dev), teacher.to(cuda_dev)
    with GuardedAlloc() as ga:
        step = PlainDistillStep(gs, gt, B, hp)
    assert len(ga.bufs) > 30
    for _ in range(2):
        out3 = step(images.to(cuda_dev), labels.to(cuda_dev))
        ga.check()
    assert torch.isfinite(out3).all()


def test_converted_executor_writes_nothing_out_of_bounds(cuda_dev):
    from torch.ao.quantization import convert
    from qatvit_b200.int8 import ConvertedStudent
    B, img = 3, 96
    vr, prepared, teacher = build_models("fbgemm", "vit_test_tiny", "vit_test_teacher", img)
    images, labels = vr.synthetic_batch(B, seed=3, img=img)
    vr.distill_step(prepared, teacher, images, labels, None, dict(vr.DEFAULT_HPARAMS), clip=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        conv = convert(copy.deepcopy(prepared).eval(), inplace=False)
    ref = ConvertedStudent(conv, B, cuda_dev)(images.to(cuda_dev)).clone()
    with GuardedAlloc() as ga:
        ex = ConvertedStudent(conv, B, cuda_dev)
    got = ex(images.to(cuda_dev)).clone()
    assert ga.check() > 15
    assert torch.equal(got, ref)
