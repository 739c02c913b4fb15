"""GPU probe: teacher-shape GEMM with fp32 output vs bf16 plane output (+GELU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import qatvit_b200  # noqa
from qatvit_b200 import ops
from qatvit_b200.ops import Op, PAIRS_FP32

dev = "cuda"
M = 50432
ONLY = os.environ.get("QV_ONLY")          # e.g. "3072,768,planes+gelu,192": one config, 2 launches (for ncu)
SHAPES = [(2304, 768), (3072, 768), (768, 768), (768, 3072)]
if ONLY:
    SHAPES = [(int(ONLY.split(",")[0]), int(ONLY.split(",")[1]))]
for (N, K) in SHAPES:
    a = ops.split_planes(torch.randn(M, K, device=dev))
    b = ops.split_planes(torch.randn(N, K, device=dev) * 0.05)
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev)
    outp = torch.empty(2, M, N, dtype=torch.bfloat16, device=dev)
    for name, kw in [("fp32", dict(out=out)), ("planes", dict(out_planes=outp)), ("planes+gelu", dict(out_planes=outp, gelu=True))]:
        for tn in (128, 192):
            if ONLY and (name != ONLY.split(",")[2] or tn != int(ONLY.split(",")[3])):
                continue
            for _ in range(1 if ONLY else 3):
                ops.gemm(Op.full(a), Op.full(b), M, N, K, PAIRS_FP32, bias=bias, tile_n=tn, **kw)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(1 if ONLY else 10):
                ops.gemm(Op.full(a), Op.full(b), M, N, K, PAIRS_FP32, bias=bias, tile_n=tn, **kw)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / (1 if ONLY else 10)
            print(f"N={N} K={K} {name:12s} tile_n={tn} {ms*1e3:8.1f} us  bf16 TF/s {2.0*M*N*K*3/ms/1e9:7.1f}", flush=True)
