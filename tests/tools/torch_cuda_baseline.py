"""The reference's own path on the SAME B200 with stock torch kernels (SURVEY.md §2.4 / §8d "the real bar"): restated timm
ViT-B/16 teacher + ViT-S/16 student prepared by stock prepare_qat (fbgemm qconfig), the reference loop body
(oracle.vit_ref.distill_step = ref/src/training/qat_trainer.py:337-361) under eager autograd, ATen CUDA kernels (cuBLAS fp32
SGEMM with TF32 off, SDPA, FusedObsFakeQuant).  Prints one JSON line.  Test/measurement infrastructure only.
Usage (GPU box): python tests/tools/torch_cuda_baseline.py [batch] [dropin]
`dropin`: the SAME unmodified loop body after qatvit_b200.dropin.install() -- the module-level integration: the
teacher stays stock torch CUDA, every fake-quant module / nnqat.Linear / nnqat.Conv2d of the student runs on the sm_100a kernels."""
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from oracle import vit_ref as vr  # noqa: E402

warnings.simplefilter("ignore")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
DROPIN = len(sys.argv) > 2 and sys.argv[2] == "dropin"
if DROPIN:
    import qatvit_b200  # noqa: F401
    from qatvit_b200 import dropin
    dropin.install()
dev = torch.device("cuda", 0)
hp = dict(vr.DEFAULT_HPARAMS)
student = vr.enable_qat(vr.make_student(prefer_reference=False), "fbgemm").to(dev).train()
teacher = vr.make_teacher().to(dev)
opt = vr.make_optimizer(student.parameters(), hp, 0.5)
images, labels = vr.synthetic_batch(B, seed=0)
images, labels = images.to(dev), labels.to(dev)
for _ in range(2):
    vr.distill_step(student, teacher, images, labels, opt, hp)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 3
s.record()
for _ in range(steps):
    vr.distill_step(student, teacher, images, labels, opt, hp)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / steps
print(json.dumps({"impl": ("reference loop body + qatvit_b200.dropin.install() (module-level drop-ins under stock autograd; teacher stock)" if DROPIN else
                           "stock torch CUDA eager (the reference's own GPU path)"), "routes": (dict(dropin.stats) if DROPIN else None), "batch": B, "ms_per_step": ms, "img_per_s": B / ms * 1e3,
                  "allow_tf32_matmul": torch.backends.cuda.matmul.allow_tf32, "allow_tf32_cudnn": torch.backends.cudnn.allow_tf32,
                  "torch": torch.__version__, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))
