"""Converted int8 student eval at batch 256: compact glue vs the fp32 glue (ms per batch + per-op breakdown of each).
usage (under gpurun): python tools/int8_quick.py"""
import copy
import sys
import warnings

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import qatvit_b200  # noqa: E402,F401
from qatvit_b200 import ops  # noqa: E402
from qatvit_b200.engine import QATDistillStep  # noqa: E402
from qatvit_b200.int8 import ConvertedStudent  # noqa: E402
from torch.ao.quantization import convert  # noqa: E402

dev = torch.device("cuda", 0)
B = 256
student, teacher = bench.build_models(B, dev)
step = QATDistillStep(student, teacher, B, bench.HP)
images = torch.randn(B, 3, 224, 224, device=dev)
labels = torch.randint(0, 10, (B,), device=dev)
for _ in range(2):
    step(images, labels)
torch.cuda.synchronize()
del step, teacher
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    conv = convert(copy.deepcopy(student).cpu().eval(), inplace=False)
torch.cuda.empty_cache()
outs = {}
for mode in (False, ("gelu", "ln"), "attn", True):
    ex = ConvertedStudent(conv, B, dev, compact=mode)
    for _ in range(3):
        ex(images)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        out = ex(images)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ops.profile_begin()
    ex(images)
    prof = ops.profile_end()
    outs[mode] = out.clone()
    print(f"compact={mode!s:5}: {ms:.3f} ms/batch, {B / ms * 1e3:.0f} img/s |",
          {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])})
    del ex
    torch.cuda.empty_cache()
stepq = float(conv.model.head.scale)
print("logits: gelu + ln == fp32 glue:", bool(torch.equal(outs[False], outs[("gelu", "ln")])), "| full vs fp32 glue, max |diff| in head steps:",
      float((outs[True] - outs[False]).abs().max()) / stepq)
