"""Isolated timing of the STUDENT-shape GEMMs (forward, dgrad, dgrad + gradient-planes epilogue, split-K wgrad) per QV_GEMM_PAIR
mode: CUDA events, 20 launches back to back.  Usage (GPU box): python tools/student_probe.py [modes...] > gpurun_out/student_probe.log"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import qatvit_b200  # noqa: E402,F401
from qatvit_b200 import ops  # noqa: E402
from qatvit_b200.engine import wgrad_splits  # noqa: E402
from qatvit_b200.ops import Op  # noqa: E402

dev = torch.device("cuda", 0)
M = int(os.environ.get("QV_PROBE_M", 50432))
MODES = sys.argv[1:] or ["0", "51", "59", "63"]
LIN = {"qkv": (1152, 384), "proj": (384, 384), "fc1": (1536, 384), "fc2": (384, 1536)}
sms = torch.cuda.get_device_properties(dev).multi_processor_count


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


def report(name, fl, fn, passes):
    res = []
    for mode in MODES:
        os.environ["QV_GEMM_PAIR"] = mode
        res.append(timed(fn))
    cells = " | ".join(f"{m}: {t:7.1f} us {fl / t * 1e-6:5.0f} alg {fl * passes / t * 1e-6:5.0f} mma TF/s" for m, t in zip(MODES, res))
    print(f"{name:28s} {cells}", flush=True)


for name, (N, K) in LIN.items():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(M, K, generator=g).to(dev)
    xp = ops.split_planes(x)
    codes = torch.randint(-127, 128, (1, N, K), generator=g).to(dev).bfloat16()
    codes_t = codes[0].t().contiguous()[None]
    cs = (torch.rand(N, generator=g) * 0.01 + 0.001).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    mm = ops.new_minmax(dev)
    out = torch.empty(M, N, device=dev)
    fl = 2.0 * M * N * K
    report(f"{name} fwd ({N},{K})", fl, lambda: ops.gemm(Op.full(xp), Op.full(codes), M, N, K, (2, 1), out=out, col_scale=cs, bias=bias, minmax=mm), 2)
    # dgrad: gp [M, N] hi/lo x codes_t [K, N] -> [M, K]
    gpl = ops.split_planes(torch.randn(M, N, generator=g).to(dev) * 1e-3)
    gx = torch.empty(M, K, device=dev)
    report(f"{name} dgrad ({K},{N})", fl, lambda: ops.gemm(Op.full(gpl), Op.full(codes_t), M, K, N, (2, 1), out=gx), 2)
    if name == "proj":
        gop = torch.empty(2, M, K, dtype=torch.bfloat16, device=dev)
        report("proj dgrad planes-out", fl, lambda: ops.gemm(Op.full(gpl), Op.full(codes_t), M, K, N, (2, 1), out_planes=gop), 2)
    if name == "fc2":    # fc2 dgrad with fc1's backward prologue in the epilogue (EPI 2): output width K = 1536
        y_raw = torch.randn(M, K, generator=g).to(dev)
        sc, zp = torch.tensor([0.05], device=dev), torch.tensor([64], dtype=torch.int32, device=dev)
        gpo = torch.empty(2, M, K, dtype=torch.bfloat16, device=dev)
        part = torch.empty(-(-M // 32), K, device=dev)
        cs1 = (torch.rand(K, generator=g) * 0.01 + 0.001).to(dev)
        report("fc2 dgrad + gp epilogue", fl, lambda: ops.gemm(Op.full(gpl), Op.full(codes_t), M, K, N, (2, 1), out_planes=gpo, col_scale=cs1,
                                                               grad_of=(y_raw, (sc, zp, 0, 127), True, part)), 2)
    # wgrad: gp^T x, split-K over tokens, both MN-major, (2,2)
    s = wgrad_splits(N, K, M, sms)
    ws = torch.empty(max(s, 1) * N * K, device=dev)
    if s > 1:
        fnw = lambda: ops.gemm(Op.full(gpl, mn_major=True), Op.full(xp, mn_major=True), N, K, M, (2, 2), splits=s, workspace=ws)  # noqa: E731
    else:
        fnw = lambda: ops.gemm(Op.full(gpl, mn_major=True), Op.full(xp, mn_major=True), N, K, M, (2, 2), out=ws[:N * K].view(N, K))  # noqa: E731
    report(f"{name} wgrad splits={s}", fl, fnw, 3)
    del x, xp, codes, codes_t, out, gpl, gx
