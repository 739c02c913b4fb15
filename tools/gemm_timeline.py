"""GPU probe: clock64 timeline of CTA 0 of one tcgen05 GEMM (TMA producer, MMA issuer, one epilogue warp) from the instrumented
build (`make -C qat-vit_b200/csrc debug`, loaded through QV_LIB).  Usage: python tools/gemm_timeline.py M N K planes_a planes_b"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("QV_LIB", os.path.join(ROOT, "qat-vit_b200", "lib", "libqatvit_b200_dbg.so"))
import torch  # noqa: E402
import qatvit_b200  # noqa: E402,F401
from qatvit_b200 import _lib, ops  # noqa: E402
from qatvit_b200.ops import Op  # noqa: E402

TAGS = {1: "tma: wait empty", 2: "tma: empty ok", 30: "epi2: wait raw y", 31: "epi2: raw y ok", 32: "epi2: acc in regs", 33: "epi2: staging free", 34: "epi2: planes stored",
        35: "epi2: colsum done", 10: "mma: wait tmem_empty", 11: "mma: tmem_empty ok", 12: "mma: wait full",
        13: "mma: full ok", 14: "mma: issued", 20: "epi: wait tmem_full", 21: "epi: tmem_full ok", 22: "epi: tmem released"}


def main():
    M, N, K, pa, pb = (int(x) for x in (sys.argv[1:6] if len(sys.argv) >= 6 else (50432, 1152, 384, 2, 1)))
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    a = torch.randn(M, K, device=dev)
    b = torch.randint(-128, 128, (N, K), device=dev).float() if pb == 1 else torch.randn(N, K, device=dev)
    mix = os.environ.get("QV_TL_MIX", "0") != "0"           # teacher format: fp16 + fp8 planes (QV_GEMM_PAIR picks pairs / tile width)
    planes_out = os.environ.get("QV_TL_PLANES", "0") != "0"  # plane-output epilogue (+ GELU, mixed planes) instead of fp32
    if mix:
        b = b * 0.02
        ap, bp = ops.split_planes_mix(a), ops.split_planes_mix(b, weight=True)
    else:
        ap = ops.split_planes(a) if pa == 2 else a.bfloat16()[None].contiguous()
        bp = ops.split_planes(b) if pb == 2 else b.bfloat16()[None].contiguous()
    cs = torch.rand(N, device=dev) + 0.5
    bias = torch.randn(N, device=dev)
    mm = ops.new_minmax(dev)
    out = torch.empty(M, N, device=dev)
    outp = torch.empty(2, M, N, dtype=torch.bfloat16, device=dev) if planes_out else None
    L = _lib.lib()

    grad_epi = os.environ.get("QV_TL_GRADEPI", "0") != "0"    # fc2 dgrad + fc1 gradient-planes epilogue (EPI 2)
    if grad_epi:
        y_raw = torch.randn(M, N, device=dev)
        gsc, gzp = torch.tensor([0.05], device=dev), torch.tensor([64], dtype=torch.int32, device=dev)
        gpo = torch.empty(2, M, N, dtype=torch.bfloat16, device=dev)
        part = torch.empty(-(-M // 32), N, device=dev)

    def run():
        if grad_epi:
            ops.gemm(Op.full(ap), Op.full(bp), M, N, K, (pa, pb), out_planes=gpo, col_scale=cs, grad_of=(y_raw, (gsc, gzp, 0, 127), True, part))
        elif mix and planes_out:
            ops.gemm(Op.full(ap), Op.full(bp), M, N, K, (2, 2), bias=bias, out_planes=outp, gelu=True, mix=True, out_mix=True)
        elif mix:
            ops.gemm(Op.full(ap), Op.full(bp), M, N, K, (2, 2), out=out, bias=bias, mix=True)
        else:
            ops.gemm(Op.full(ap), Op.full(bp), M, N, K, (pa, pb), out=out, col_scale=cs, bias=bias, minmax=mm)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    L.qv_gemm_debug_clear()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    run()
    e.record()
    torch.cuda.synchronize()
    print(f"kernel {s.elapsed_time(e) * 1e3:.1f} us")
    buf = (ctypes.c_ulonglong * (3 * 4096))()
    L.qv_gemm_debug_read.restype = ctypes.c_int
    assert L.qv_gemm_debug_read(buf) == 0
    ev = []
    for who in range(3):
        for i in range(4096):
            v = buf[who * 4096 + i]
            if v == 0:
                break
            ev.append((v & 0xffffffffffff, v >> 48, who))
    ev.sort()
    t0 = ev[0][0]
    # aggregate waiting time per role
    waits = {}
    open_ = {}
    pairs = {1: 2, 10: 11, 12: 13, 20: 21, 30: 31}
    for t, tag, who in ev:
        if tag in pairs:
            open_[tag] = t
        for a_, b_ in pairs.items():
            if tag == b_ and a_ in open_:
                waits[TAGS[a_]] = waits.get(TAGS[a_], 0) + (t - open_.pop(a_))
    total = ev[-1][0] - t0
    print(f"pair launches so far: {ops.gemm_pair_launches()}")
    print(f"CTA 0 timeline: {total / 1.9e3:.1f} us, {sum(1 for x in ev if x[1] == 21)} tiles")
    for k, v in waits.items():
        print(f"  {k:24s} {v / 1.9e3:8.1f} us  ({100.0 * v / total:5.1f} %)")
    lim = int(os.environ.get("QV_LINES", "90"))
    prev = t0
    for t, tag, who in ev[:lim]:
        print(f"{(t - t0) / 1.9:9.0f} ns (+{(t - prev) / 1.9:6.0f})  {TAGS.get(tag, tag)}")
        prev = t


if __name__ == "__main__":
    main()
