"""GPU probe for BASELINE.json configs[4]: converted int8 ViT-S/16 student (stock convert() of a QAT-trained student) evaluated on
synthetic 224x224 batches -- our executor (qatvit_b200.int8.ConvertedStudent: tcgen05 kind::i8 linears + fp32 glue) against the
CPU path with stock torch.ops.quantized kernels (oracle/int8_ref.py, the same float glue) on a bounded sample.
Usage: python tests/tools/int8_eval_bench.py [batch]      -> one JSON line"""
import copy
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import qatvit_b200  # noqa: E402,F401
from qatvit_b200.engine import QATDistillStep  # noqa: E402
from qatvit_b200.int8 import ConvertedStudent  # noqa: E402
from oracle import int8_ref  # noqa: E402  (checker / CPU baseline only)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = torch.device("cuda", 0)
    from torch.ao.quantization import convert
    student, teacher = bench.build_models(B, dev)
    step = QATDistillStep(student, teacher, B, bench.HP)
    g = torch.Generator().manual_seed(0)
    images = torch.randn(B, 3, 224, 224, generator=g)
    labels = torch.randint(0, 10, (B,), generator=g)
    for _ in range(2):                                   # observers see data (ref: a trained QAT model)
        step(images.to(dev), labels.to(dev))
    torch.cuda.synchronize()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        conv = convert(copy.deepcopy(student).cpu().eval(), inplace=False)           # ref qat_trainer.py:377-379
    ex = ConvertedStudent(conv, B, dev)
    x = images.to(dev)
    for _ in range(3):
        ex(x)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    n = 10
    for _ in range(n):
        logits = ex(x)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / n
    # CPU reference on a bounded sample (batch 8), all host threads
    torch.set_num_threads(os.cpu_count() or 1)
    xs = images[:8]
    int8_ref.converted_forward(conv, xs)
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < 10.0:
        ref = int8_ref.converted_forward(conv, xs)
        reps += 1
    cpu_dt = (time.perf_counter() - t0) / reps
    ours8 = ConvertedStudent(conv, 8, dev)(xs.to(dev)).cpu()
    stepq = float(conv.model.head.scale)
    print(json.dumps({"workload": f"converted int8 ViT-S/16 student eval, batch {B}, synthetic 224x224 (BASELINE configs[4])",
                      "img_per_s": B / (ms * 1e-3), "ms_per_batch": ms,
                      "cpu_reference": {"img_per_s": 8 / cpu_dt, "ms_per_batch8": cpu_dt * 1e3, "threads": torch.get_num_threads(),
                                        "engine": torch.backends.quantized.engine},
                      "logits_max_abs_diff_in_head_steps": float((ours8 - ref).abs().max()) / stepq,
                      "logits_rel_l2": float((ours8 - ref).norm() / ref.norm()), "finite": bool(torch.isfinite(logits).all())}))


if __name__ == "__main__":
    main()
