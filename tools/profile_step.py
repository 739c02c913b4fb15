"""One profiled QAT-distillation step at the bench workload (B=256) for `ncu --profile-from-start off`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import qatvit_b200  # noqa
from qatvit_b200.engine import QATDistillStep

B = int(os.environ.get("QV_BATCH", "256"))
dev = torch.device("cuda", 0)
student, teacher = bench.build_models(B, dev)
step = QATDistillStep(student, teacher, B, bench.HP)
g = torch.Generator().manual_seed(0)
images = torch.randn(B, 3, 224, 224, generator=g).to(dev)
labels = torch.randint(0, 10, (B,), generator=g).to(dev)
for _ in range(2):
    step(images, labels)
torch.cuda.synchronize()
torch.cuda.profiler.start()
step(images, labels)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", float(step.student_engine.loss3[0]))
