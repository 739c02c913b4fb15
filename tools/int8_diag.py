"""diagnostic: per-launch CUDA-event times of int8_linear inside the compact executor (which launch is the outlier?)"""
import copy, sys, warnings
import torch
sys.path.insert(0, ".")
import bench, qatvit_b200  # noqa
from qatvit_b200 import ops
import qatvit_b200.int8 as I8
from qatvit_b200.engine import QATDistillStep
from torch.ao.quantization import convert
dev = torch.device("cuda", 0)
B = 256
student, teacher = bench.build_models(B, dev)
step = QATDistillStep(student, teacher, B, bench.HP)
images = torch.randn(B, 3, 224, 224, device=dev); labels = torch.randint(0, 10, (B,), device=dev)
step(images, labels); torch.cuda.synchronize()
del step, teacher
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    conv = convert(copy.deepcopy(student).cpu().eval(), inplace=False)
torch.cuda.empty_cache()
orig = ops.int8_linear
for mode in (True, False, True):
    ex = I8.ConvertedStudent(conv, B, dev, compact=mode)
    for _ in range(3): ex(images)
    torch.cuda.synchronize()
    rec = []
    def timed(*a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = orig(*a, **k); e.record()
        rec.append((a[0].shape, a[3].shape, "qy" if k.get("qy") is not None else "y", s, e)); return r
    I8.ops.int8_linear = timed
    for rep in range(2):
        rec.clear()
        ex(images); torch.cuda.synchronize()
        ts = [(x[0], x[1], x[2], x[3].elapsed_time(x[4])) for x in rec]
        tot = sum(t[3] for t in ts)
        worst = sorted(ts, key=lambda t: -t[3])[:3]
        by = {}
        for t in ts: by.setdefault((tuple(t[1]), t[2]), []).append(t[3])
        print(f"mode {mode} rep {rep}: total {tot:.3f} ms; worst {[(tuple(w[1]), w[2], round(w[3], 3)) for w in worst]}")
        print("   per shape (N,K)/out: ", {k: round(sum(v) / len(v) * 1e3, 1) for k, v in by.items()}, "us avg")
    I8.ops.int8_linear = orig
    del ex; torch.cuda.empty_cache()
