import sys, torch, time
sys.path.insert(0, ".")
import bench
import qatvit_b200
from qatvit_b200.engine import QATDistillStep
from qatvit_b200.optim import FusedClipAdamW
dev = torch.device("cuda", 0)
B = 64
student, teacher = bench.build_models(B, dev)
step = QATDistillStep(student, teacher, B, bench.HP)
opt = FusedClipAdamW(student.parameters(), step.grad_arena, lr=bench.HP["lr"] * 0.5, weight_decay=bench.HP["weight_decay"], max_norm=1.0)
g = torch.Generator().manual_seed(0)
data = torch.randn(8, B, 3, 224, 224, generator=g).to(dev)
labels = torch.randint(0, 10, (8, B), generator=g).to(dev)
losses = []
t0 = time.time()
for it in range(400):
    out3 = step(data[it % 8], labels[it % 8])
    opt.step()
    if it % 50 == 0 or it == 399:
        losses.append((it, [round(float(v), 4) for v in out3]))
torch.cuda.synchronize()
print("400 steps in", round(time.time() - t0, 1), "s")
for l in losses: print(l)
sd = student.state_dict()
bad = [k for k, v in sd.items() if v.is_floating_point() and not torch.isfinite(v).all()]
print("non-finite state entries:", bad[:5], "| arena finite:", bool(torch.isfinite(step.grad_arena).all()))
