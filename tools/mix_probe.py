"""GPU probe of the mixed (fp16 + fp8 cross terms) GEMM operand format against the bf16 hi/lo three-pass product:
accuracy vs an fp64 matmul of the true fp32 inputs, and CUDA-event time per launch.
Usage: python tools/mix_probe.py [case ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (M, N, K)
    "small_bn64": (333, 64, 128),
    "small_bn128": (197, 128, 64),
    "ragged": (1000, 768, 768),
    "t_qkv": (50432, 2304, 768),
    "t_proj": (50432, 768, 768),
    "t_fc1": (50432, 3072, 768),
    "t_fc2": (50432, 768, 3072),
}


def decode_mix(planes, weight=False):
    import torch
    rows, cols = planes.shape[1], planes.shape[2]
    s16, sh8, sl8 = 128.0, 128.0, 1.0
    h16 = planes[0].view(torch.float16).double()
    b = planes[1].view(torch.uint8).reshape(rows, cols // 64, 2, 64)
    h8 = b[:, :, 0, :].contiguous().view(torch.float8_e4m3fn if weight else torch.float8_e5m2).double().reshape(rows, cols)
    l8 = b[:, :, 1, :].contiguous().view(torch.float8_e5m2).double().reshape(rows, cols)
    return (h16 + l8 / sl8) / s16, h8 / sh8


def run_case(name):
    import torch
    import qatvit_b200  # noqa
    from qatvit_b200 import ops
    M, N, K = CASES[name]
    torch.manual_seed(0)
    dev = "cuda"
    A = torch.randn(M, K, device=dev) * 1.3
    W = torch.randn(N, K, device=dev) * 0.02
    bias = torch.randn(N, device=dev) * 0.1
    ref = A.double() @ W.double().t() + bias.double()[None, :]
    scale = ref.abs().max().item()
    Ab, Wb = ops.split_planes(A), ops.split_planes(W)
    Am, Wm = ops.split_planes_mix(A), ops.split_planes_mix(W, weight=True)
    # the split itself
    a_rec, a_h8 = decode_mix(Am)
    w_rec, w_h8 = decode_mix(Wm, weight=True)
    msg = (f"{name}: split A rel {((a_rec - A.double()).abs().max() / A.abs().max()).item():.2e} hi8 {((a_h8 - A.double()).abs() / A.abs().double().clamp_min(1e-3)).max().item():.2e}"
           f" | W rel {((w_rec - W.double()).abs().max() / W.abs().max()).item():.2e} hi8 {((w_h8 - W.double()).abs() / W.abs().double().clamp_min(1e-4)).max().item():.2e}")
    print(msg, flush=True)
    out_b = ops.gemm(ops.Op.full(Ab), ops.Op.full(Wb), M, N, K, (2, 2), bias=bias)
    out_m = ops.gemm(ops.Op.full(Am), ops.Op.full(Wm), M, N, K, (2, 2), bias=bias, mix=True)
    torch.cuda.synchronize()
    eb = (out_b.double() - ref).abs()
    em = (out_m.double() - ref).abs()
    print(f"{name}: bf16x3 max {eb.max().item() / scale:.2e} rms {eb.pow(2).mean().sqrt().item() / scale:.2e} | "
          f"mix max {em.max().item() / scale:.2e} rms {em.pow(2).mean().sqrt().item() / scale:.2e}", flush=True)
    if N % 64 == 0 and N >= 128:
        # plane outputs: bf16 hi/lo and mixed, with GELU
        gref = torch.nn.functional.gelu(ref)
        pb = ops.gemm(ops.Op.full(Am), ops.Op.full(Wm), M, N, K, (2, 2), bias=bias, mix=True,
                      out_planes=torch.empty(2, M, N, dtype=torch.bfloat16, device=dev), gelu=True)
        pm = ops.gemm(ops.Op.full(Am), ops.Op.full(Wm), M, N, K, (2, 2), bias=bias, mix=True,
                      out_planes=torch.empty(2, M, N, dtype=torch.bfloat16, device=dev), gelu=True, out_mix=True)
        torch.cuda.synchronize()
        e1 = ((pb[0].double() + pb[1].double()) - gref).abs().max().item() / gref.abs().max().item()
        rec, h8 = decode_mix(pm)
        e2 = (rec - gref).abs().max().item() / gref.abs().max().item()
        e3 = ((h8 - gref).abs() / gref.abs().clamp_min(1e-2)).max().item()
        print(f"{name}: plane outputs (GELU) bf16 hi/lo {e1:.2e} | mixed fp16+lo8 {e2:.2e}, hi8 rel {e3:.2e}", flush=True)
    if M * N * K > 1e9:
        for label, fn in (("bf16x3", lambda: ops.gemm(ops.Op.full(Ab), ops.Op.full(Wb), M, N, K, (2, 2), bias=bias, out=out_b)),
                          ("mix", lambda: ops.gemm(ops.Op.full(Am), ops.Op.full(Wm), M, N, K, (2, 2), bias=bias, out=out_m, mix=True))):
            for _ in range(3):
                fn()
            st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            st.record()
            for _ in range(20):
                fn()
            en.record()
            torch.cuda.synchronize()
            us = st.elapsed_time(en) * 1000 / 20
            print(f"{name}: {label} {us:.1f} us  ({2.0 * M * N * K / us * 1e-6:.0f} algorithmic TFLOP/s)", flush=True)


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    if len(names) == 1 and os.environ.get("QV_PROBE_CHILD"):
        run_case(names[0])
    else:
        import subprocess
        for n in names:
            env = dict(os.environ, QV_PROBE_CHILD="1")
            r = subprocess.run([sys.executable, __file__, n], env=env, capture_output=True, text=True, timeout=300)
            sys.stdout.write(r.stdout)
            if r.returncode != 0:
                sys.stdout.write(f"{n}: FAILED rc={r.returncode}\n{r.stderr[-1500:]}\n")
