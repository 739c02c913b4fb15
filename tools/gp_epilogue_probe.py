"""GPU probe: the fc2 dgrad GEMM with fc1's backward prologue fused in its epilogue (act = 2) at the bench shape
(M = 197*256, N = 1536, K = 384) against the unfused chain.  Usage: python tools/gp_epilogue_probe.py [tile_n ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import qatvit_b200  # noqa: E402,F401
from qatvit_b200 import ops  # noqa: E402
from qatvit_b200.ops import Op, PAIRS_EXACT_B  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


def main():
    dev = torch.device("cuda", 0)
    M, N, K = 197 * int(os.environ.get("QV_B", "256")), int(os.environ.get("QV_N", "1536")), int(os.environ.get("QV_K", "384"))
    gelu = N == 1536
    torch.manual_seed(0)
    a = torch.randn(M, K, device=dev)
    b = torch.randint(-128, 128, (N, K), device=dev).float()
    ap, bp = ops.split_planes(a), b.bfloat16()[None].contiguous()
    y = torch.randn(M, N, device=dev) * 2
    fq = (torch.tensor([4.0 / 127], device=dev), torch.tensor([64], dtype=torch.int32, device=dev), 0, 127)
    wsc = torch.rand(N, device=dev) * 0.02 + 0.001
    gmat = torch.empty(M, N, device=dev)
    planes = torch.empty(2, M, N, dtype=torch.bfloat16, device=dev)
    part = torch.empty(-(-M // 64), N, device=dev)
    slab = torch.empty(-(-M // 32), N, device=dev)
    bias = torch.empty(N, device=dev)
    tiles = [int(t) for t in sys.argv[1:]] or [128, 192]
    t_g = timeit(lambda: ops.gemm(Op.full(ap), Op.full(bp), M, N, K, PAIRS_EXACT_B, out=gmat))
    t_p = timeit(lambda: ops.gp_planes(gmat, y, fq, wsc, True, gelu, M, N, planes, part, 64))
    t_c = timeit(lambda: ops.colsum_reduce(part, part.shape[0], N, bias))
    print(f"unfused: gemm {t_g:.1f} us + gp_planes {t_p:.1f} us + colsum {t_c:.1f} us")
    for tn in tiles:
        t_f = timeit(lambda: ops.gemm(Op.full(ap), Op.full(bp), M, N, K, PAIRS_EXACT_B, out_planes=planes, col_scale=wsc,
                                      grad_of=(y, fq, gelu, slab), tile_n=tn))
        t_n = timeit(lambda: ops.gemm(Op.full(ap), Op.full(bp), M, N, K, PAIRS_EXACT_B, out_planes=planes, col_scale=wsc,
                                      grad_of=(y, fq, gelu, None), tile_n=tn))
        t_c2 = timeit(lambda: ops.colsum_reduce(slab, slab.shape[0], N, bias))
        gb = (M * N * 8 + M * K * 4) / 1e9
        print(f"fused tile_n={tn}: {t_f:.1f} us ({gb / t_f * 1e6:.0f} GB/s algorithmic), without colsum {t_n:.1f} us; colsum_reduce {t_c2:.1f} us")


if __name__ == "__main__":
    main()
