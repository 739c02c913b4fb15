"""Isolated timing of the teacher-shape GEMMs: one CTA per tile vs CTA pairs (QV_GEMM_PAIR), CUDA events, 20 launches back to back.
Usage (GPU box): python tools/pair_probe.py > gpurun_out/pair_probe.log"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import qatvit_b200  # noqa: E402,F401
from qatvit_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
M = int(os.environ.get("QV_PROBE_M", 50432))
SHAPES = {"t_qkv": (2304, 768), "t_proj": (768, 768), "t_fc1": (3072, 768), "t_fc2": (768, 3072)}


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    st.record()
    for _ in range(iters):
        fn()
    en.record()
    torch.cuda.synchronize()
    return st.elapsed_time(en) * 1000 / iters


for name, (N, K) in SHAPES.items():
    g = torch.Generator().manual_seed(0)
    a = (torch.randn(M, K, generator=g) * 1.3).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.02).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    am, wm = ops.split_planes_mix(a), ops.split_planes_mix(w, weight=True)
    ab, wb = ops.split_planes(a), ops.split_planes(w)
    out = torch.empty(M, N, device=dev)
    planes = torch.empty(2, M, N, dtype=torch.bfloat16, device=dev)
    cases = {
        "mix fp32-out": lambda: ops.gemm(ops.Op.full(am), ops.Op.full(wm), M, N, K, (2, 2), bias=bias, out=out, mix=True),
        "mix planes-out+gelu": lambda: ops.gemm(ops.Op.full(am), ops.Op.full(wm), M, N, K, (2, 2), bias=bias, out_planes=planes,
                                                gelu=True, mix=True, out_mix=True),
        "bf16x3 fp32-out": lambda: ops.gemm(ops.Op.full(ab), ops.Op.full(wb), M, N, K, (2, 2), bias=bias, out=out),
    }
    for label, fn in cases.items():
        res = []
        for mode in ("0", "31", "63"):
            os.environ["QV_GEMM_PAIR"] = mode
            res.append(timed(fn))
        fl = 2.0 * M * N * K
        print(f"{name} {label}: one-CTA {res[0]:.1f} us ({fl / res[0] * 1e-6:.0f} alg TF/s) | pair {res[1]:.1f} us "
              f"({fl / res[1] * 1e-6:.0f} alg TF/s)  x{res[0] / res[1]:.2f} | pair-256 {res[2]:.1f} us "
              f"({fl / res[2] * 1e-6:.0f} alg TF/s)  x{res[0] / res[2]:.2f}", flush=True)
    del a, w, am, wm, ab, wb, out, planes
