"""Range guard of the teacher's mixed fp16 + fp8 operand format (one fixed 2^7 scale: activations saturate at |x| = 448,
weights at |w| = 3.5 -- csrc/qv_common.cuh).  A fine-tuned in21k ViT-B (ref/src/models/model_registry.py:178-207) has outlier
channels a random-init teacher lacks ("massive activations", large LayerNorm gains): the engine must notice and route the
affected Linear to bf16 hi/lo planes instead of silently clamping the KD target.  Tolerance: north_star's 1e-3 on fp32 logits."""
import copy
import warnings

import pytest
import torch

from parity_utils import rel_max

pytestmark = pytest.mark.gpu


def _teacher(seed=1):
    from oracle import vit_ref as vr
    torch.manual_seed(seed)
    t = vr.create_model("vit_test_teacher", num_classes=10, img_size=64).eval()
    with torch.no_grad():
        t.head.weight.mul_(8.0)
    for p in t.parameters():
        p.requires_grad = False
    return vr, t


def test_ordinary_teacher_stays_mixed_and_raises_no_flag(cuda_dev):
    from qatvit_b200.engine import TeacherEngine
    vr, t = _teacher()
    images, _ = vr.synthetic_batch(4, seed=2, img=64)
    eng = TeacherEngine(copy.deepcopy(t).to(cuda_dev), 4, mixed=True)
    assert eng.mixed and all(all(m.values()) for m in eng.mix)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert eng.calibrate(images.to(cuda_dev)) == []
    assert int(eng.sat.abs().sum()) == 0
    with torch.no_grad():
        ref = t(images)
    assert rel_max(eng.logits, ref) < 1e-3


def test_out_of_range_weight_is_routed_to_three_passes_at_build(cuda_dev):
    from qatvit_b200.engine import TeacherEngine
    vr, t = _teacher()
    with torch.no_grad():
        t.blocks[0].mlp.fc1.weight[3, 5] = 4.0           # e4m3(w * 2^7) would saturate at 3.5
        t.blocks[1].attn.proj.weight[7, 9] = -6.5
    images, _ = vr.synthetic_batch(3, seed=4, img=64)
    eng = TeacherEngine(copy.deepcopy(t).to(cuda_dev), 3, mixed=True)
    assert not eng.mix[0]["fc1"] and not eng.mix[1]["proj"] and eng.mix[0]["qkv"] and eng.mix[1]["fc2"]
    assert any("blocks.0.fc1: weight" in e for e in eng.saturation_events)
    logits = eng.forward(images.to(cuda_dev)).clone()
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = t(images)
    assert rel_max(logits, ref) < 1e-3


def _outlier_teacher():
    vr, t = _teacher(seed=3)
    with torch.no_grad():
        t.blocks[0].norm1.bias[7] = 600.0                 # a "massive activation" channel in the qkv input
        t.blocks[0].attn.qkv.bias[2 * 256 + 5] = 520.0    # ... in V, so the attention output (proj's input) carries it
        t.blocks[1].norm2.weight[11:14] *= 20.0           # large LayerNorm gains on a few channels (fc1 input, |x| up to ~60: in range)
        t.blocks[1].norm2.bias[12] = -700.0               # fc1 input out of range
        t.blocks[1].mlp.fc1.bias[33] = 900.0              # GELU output (fc2's input) out of range
    return vr, t


def test_activation_outliers_are_detected_by_calibrate(cuda_dev):
    from qatvit_b200.engine import TeacherEngine
    vr, t = _outlier_teacher()
    images, _ = vr.synthetic_batch(4, seed=6, img=64)
    with torch.no_grad():
        ref = t(images)
    eng = TeacherEngine(copy.deepcopy(t).to(cuda_dev), 4, mixed=True)
    with pytest.warns(RuntimeWarning, match="outside the mixed"):
        events = eng.calibrate(images.to(cuda_dev))
    assert not eng.mix[0]["qkv"] and not eng.mix[0]["proj"] and not eng.mix[1]["fc1"] and not eng.mix[1]["fc2"]
    assert eng.mix[1]["qkv"] and eng.mix[0]["fc2"]                        # only what left the range moved
    assert len(events) >= 4
    logits = eng.forward(images.to(cuda_dev)).clone()
    torch.cuda.synchronize()
    assert rel_max(logits, ref) < 1e-3, rel_max(logits, ref)
    # the three-pass engine (no mixed format at all) agrees, and a mixed engine WITHOUT the guard would not have
    plain = TeacherEngine(copy.deepcopy(t).to(cuda_dev), 4, mixed=False).forward(images.to(cuda_dev)).clone()
    torch.cuda.synchronize()
    assert rel_max(plain, ref) < 1e-3


def test_activation_outliers_are_caught_asynchronously_inside_the_step(cuda_dev):
    """Without calibrate(): the first forward runs with clamped values and raises the device flags; the flags are read (no sync
    in the step itself) when a later forward starts, the Linears move to bf16 hi/lo planes, and from then on the logits are right."""
    from qatvit_b200.engine import TeacherEngine
    vr, t = _outlier_teacher()
    images, _ = vr.synthetic_batch(4, seed=6, img=64)
    with torch.no_grad():
        ref = t(images)
    eng = TeacherEngine(copy.deepcopy(t).to(cuda_dev), 4, mixed=True)
    x = images.to(cuda_dev)
    first = eng.forward(x).clone()
    torch.cuda.synchronize()
    assert rel_max(first, ref) > 1e-3                                      # the clamp is real: this is what the guard prevents
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for _ in range(4):                                                 # each forward fixes what the previous one flagged
            out = eng.forward(x).clone()
            torch.cuda.synchronize()
    assert len(eng.saturation_events) >= 4
    assert rel_max(out, ref) < 1e-3
