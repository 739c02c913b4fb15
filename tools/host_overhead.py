"""Host-side enqueue time of one step (Python + ctypes launches, no synchronisation) against its GPU time, per batch size.
The step is GPU-bound as long as the first stays well below the second (N = 8 runs 128 images per GPU: ~25 ms of GPU work).
usage (under gpurun): python tools/host_overhead.py [batch ...]"""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import qatvit_b200  # noqa: E402,F401
from qatvit_b200 import ops  # noqa: E402
from qatvit_b200.engine import QATDistillStep  # noqa: E402
from qatvit_b200.optim import FusedClipAdamW  # noqa: E402

dev = torch.device("cuda", 0)
for B in [int(a) for a in sys.argv[1:]] or [128, 256]:
    student, teacher = bench.build_models(B, dev)
    step = QATDistillStep(student, teacher, B, bench.HP)
    opt = FusedClipAdamW(student.parameters(), step.grad_arena, lr=bench.HP["lr"], weight_decay=bench.HP["weight_decay"], max_norm=1.0)
    images = torch.randn(B, 3, 224, 224, device=dev)
    labels = torch.randint(0, 10, (B,), device=dev)
    for _ in range(5):
        step(images, labels)
        opt.step()
    torch.cuda.synchronize()
    host, gpu = [], []
    for _ in range(10):
        torch.cuda.synchronize()
        n0 = ops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        step(images, labels)
        opt.step()
        e1.record()
        host.append((time.perf_counter() - t0) * 1e3)
        torch.cuda.synchronize()
        gpu.append(e0.elapsed_time(e1))
        launches = ops.launch_count() - n0
    host.sort(); gpu.sort()
    print(f"batch {B}: host enqueue {host[len(host) // 2]:.2f} ms (min {host[0]:.2f}), GPU {gpu[len(gpu) // 2]:.2f} ms, {launches} launches, "
          f"{host[len(host) // 2] * 1e3 / launches:.1f} us of host time per launch")
    del step, opt, student, teacher
    torch.cuda.empty_cache()
