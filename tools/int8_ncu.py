"""two forwards of the compact converted executor at batch 256 (for: ncu -k regex:qv_int8_linear_kernel --launch-skip 51 -c 4)"""
import copy, sys, warnings
import torch
sys.path.insert(0, ".")
import bench, qatvit_b200  # noqa
from qatvit_b200.int8 import ConvertedStudent
from torch.ao.quantization import convert
dev = torch.device("cuda", 0)
B = 256
student, teacher = bench.build_models(B, dev)
del teacher
images = torch.randn(B, 3, 224, 224, device=dev)
student.eval()
with torch.no_grad():
    student(images[:8])              # one observer pass so that every activation range is set (stock modules, 8 images)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    conv = convert(copy.deepcopy(student).cpu().eval(), inplace=False)
ex = ConvertedStudent(conv, B, dev)
for _ in range(2):
    out = ex(images)
torch.cuda.synchronize()
print("ok", bool(torch.isfinite(out).all()))
