"""GPU probe: student attention forward / backward (integer codes) at the bench shape (B images x 6 heads x 197 tokens),
timed alone.  Usage: python tools/attn_probe.py   (QV_B=256 by default)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import qatvit_b200  # noqa: E402,F401
from qatvit_b200 import ops  # noqa: E402


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
    B, H, T = int(os.environ.get("QV_B", "256")), int(os.environ.get("QV_H", "6")), int(os.environ.get("QV_T", "197"))
    D = H * 64
    torch.manual_seed(0)
    sval = 0.0437
    y_raw = (torch.randint(-60, 68, (B * T, 3 * D), device=dev).float() + 0.3 * torch.randn(B * T, 3 * D, device=dev)) * sval
    s = torch.tensor([sval], device=dev)
    zp = torch.tensor([60], dtype=torch.int32, device=dev)
    fq = (s, zp, 0, 127)
    cp = torch.empty(1, B * T, 3 * D, dtype=torch.bfloat16, device=dev)
    ops.act_planes(y_raw, fq, False, cp, codes_only=True)
    out = torch.empty(2, B * T, D, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B * H * T, device=dev)
    dOp = ops.split_planes(torch.randn(B * T, D, device=dev))
    wsc = torch.rand(3 * D, device=dev) * 0.02 + 0.001
    g_qkv = torch.empty(B * T, 3 * D, device=dev)
    planes = torch.empty(2, B * T, 3 * D, dtype=torch.bfloat16, device=dev)
    slab = torch.empty(B * (-(-T // 128)) * 4, 3 * D, device=dev)
    items = B * H
    t = timeit(lambda: ops.attn_fwd(cp, B, T, H, 0.125, out, qk_scale=s, v_scale=s, lse=lse))
    print(f"attn_fwd (codes): {t:.1f} us, {t * 148 / items:.2f} us per item per SM")
    # teacher form: fp32-grade hi/lo planes, 12 heads, mixed-format planes out (what TeacherEngine launches)
    Ht = 12
    Dt = Ht * 64
    qkvp = ops.split_planes(torch.randn(B * T, 3 * Dt, device=dev) * 1.5)
    out_t = torch.empty(2, B * T, Dt, dtype=torch.bfloat16, device=dev)
    t = timeit(lambda: ops.attn_fwd(qkvp, B, T, Ht, 0.125, out_t))
    print(f"attn_fwd (hi/lo planes, {Ht} heads): {t:.1f} us, {t * 148 / (B * Ht):.2f} us per item per SM")
    del qkvp, out_t
    t = timeit(lambda: ops.attn_bwd(cp, s, out, dOp, lse, B, T, H, 0.125, g_qkv))
    print(f"attn_bwd: {t:.1f} us, {t * 148 / items:.2f} us per item per SM")
    t = timeit(lambda: ops.attn_bwd_gp(cp, s, out, dOp, lse, B, T, H, 0.125, y_raw, fq, wsc, planes, slab))
    print(f"attn_bwd_gp: {t:.1f} us, {t * 148 / items:.2f} us per item per SM")


if __name__ == "__main__":
    main()
