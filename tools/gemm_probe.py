"""GPU probe: run one tcgen05 GEMM configuration per subprocess (a device trap then only kills that case)
and print error statistics against an fp64 torch matmul.  Usage: python tools/gemm_probe.py [case ...]"""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TILE_N = int(os.environ.get("QV_TILE_N", "0"))

CASES = {
    # name: (M, N, K, a_mn, b_mn, planes_a, planes_b, pairs, splits)
    "kk_small": (256, 128, 64, 0, 0, 1, 1, [(0, 0)], 1),
    "kk_mid": (1024, 384, 384, 0, 0, 1, 1, [(0, 0)], 1),
    "kk_ragged": (1576, 1152, 384, 0, 0, 2, 1, [(0, 0), (1, 0)], 1),
    "kk_3pair": (1000, 768, 768, 0, 0, 2, 2, [(0, 0), (0, 1), (1, 0)], 1),
    "kk_n64": (333, 64, 197, 0, 0, 2, 2, [(0, 0), (0, 1), (1, 0)], 1),
    "kk_n197": (197, 197, 64, 0, 0, 2, 2, [(0, 0), (0, 1), (1, 0)], 1),
    "qkv_fwd": (50432, 1152, 384, 0, 0, 2, 1, [(0, 0), (1, 0)], 1),
    "fc1_fwd": (50432, 1536, 384, 0, 0, 2, 1, [(0, 0), (1, 0)], 1),
    "fc2_fwd": (50432, 384, 1536, 0, 0, 2, 1, [(0, 0), (1, 0)], 1),
    "t_qkv": (50432, 2304, 768, 0, 0, 2, 2, [(0, 0), (0, 1), (1, 0)], 1),
    "t_fc2": (50432, 768, 3072, 0, 0, 2, 2, [(0, 0), (0, 1), (1, 0)], 1),
    "mn_small": (128, 128, 64, 1, 1, 1, 1, [(0, 0)], 1),
    "mn_mid": (384, 1536, 1576, 1, 1, 2, 2, [(0, 0), (0, 1), (1, 0)], 1),
    "wgrad_qkv": (1152, 384, 50432, 1, 1, 2, 2, [(0, 0), (0, 1), (1, 0)], 11),
    "wgrad_fc1": (1536, 384, 50432, 1, 1, 2, 2, [(0, 0), (0, 1), (1, 0)], 8),
    "b_mn_only": (256, 256, 256, 0, 1, 2, 2, [(0, 0), (0, 1), (1, 0)], 1),
}


def run_case(name):
    import torch
    sys.path.insert(0, ROOT)
    import qatvit_b200  # noqa
    from qatvit_b200 import ops
    M, N, K, a_mn, b_mn, pa, pb, pairs, splits = CASES[name]
    torch.manual_seed(0)
    dev = "cuda"
    A32 = torch.randn(M, K, device=dev)
    B32 = torch.randn(N, K, device=dev)

    def planes(x, n):
        if n == 1:
            return x.to(torch.bfloat16)[None].contiguous()
        return ops.split_planes(x.contiguous())
    Ap = planes(A32, pa)   # [p, M, K]
    Bp = planes(B32, pb)
    # effective operand values actually represented by the selected pairs
    ref = torch.zeros(M, N, dtype=torch.float64, device=dev)
    for (i, j) in pairs:
        ref += Ap[i].double() @ Bp[j].double().t()
    a_in = Ap.transpose(1, 2).contiguous() if a_mn else Ap
    b_in = Bp.transpose(1, 2).contiguous() if b_mn else Bp
    bias = torch.randn(N, device=dev)
    cs = torch.rand(N, device=dev) + 0.5
    mm = ops.new_minmax(dev)
    torch.cuda.synchronize()
    if splits == 1:
        out = ops.gemm(ops.Op.full(a_in, bool(a_mn)), ops.Op.full(b_in, bool(b_mn)), M, N, K, (pa, pb), col_scale=cs, bias=bias,
                       minmax=mm, tile_n=TILE_N)
        ref = ref * cs.double()[None, :] + bias.double()[None, :]
    else:
        ws = ops.gemm(ops.Op.full(a_in, bool(a_mn)), ops.Op.full(b_in, bool(b_mn)), M, N, K, (pa, pb), splits=splits, tile_n=TILE_N)
        out = torch.empty(M, N, device=dev)
        ops.splitk_reduce(ws, splits, M, N, out)
    torch.cuda.synchronize()
    err = (out.double() - ref).abs().max().item()
    rel = err / ref.abs().max().item()
    msg = f"{name}: max_abs_err={err:.3e} rel={rel:.3e}"
    if splits == 1:
        import struct
        def dec(u):
            u &= 0xffffffff
            u = (u & 0x7fffffff) if (u & 0x80000000) else (~u & 0xffffffff)
            return struct.unpack("f", struct.pack("I", u))[0]
        mn, mx = dec(int(mm[0, 0])), dec(int(mm[0, 1]))
        msg += f" minmax=({mn:.4f},{mx:.4f}) ref=({out.min().item():.4f},{out.max().item():.4f})"
    # timing
    if M * N * K > 1e9:
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            ops.gemm(ops.Op.full(a_in, bool(a_mn)), ops.Op.full(b_in, bool(b_mn)), M, N, K, (pa, pb), splits=splits,
                     workspace=ws if splits > 1 else None, out=out if splits == 1 else None, tile_n=TILE_N)
        st.record()
        for _ in range(10):
            ops.gemm(ops.Op.full(a_in, bool(a_mn)), ops.Op.full(b_in, bool(b_mn)), M, N, K, (pa, pb), splits=splits,
                     workspace=ws if splits > 1 else None, out=out if splits == 1 else None, tile_n=TILE_N)
        en.record()
        torch.cuda.synchronize()
        ms = st.elapsed_time(en) / 10
        fl = 2.0 * M * N * K * len(pairs)
        msg += f" time={ms*1e3:.1f}us bf16_TFLOPs={fl/ms/1e9:.1f} alg_TFLOPs={2.0*M*N*K/ms/1e9:.1f} tile_n={TILE_N}"
    print(msg, "OK" if rel < 5e-5 else "FAIL", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        run_case(sys.argv[2])
        sys.exit(0)
    names = sys.argv[1:] or list(CASES)
    for n in names:
        try:
            r = subprocess.run([sys.executable, __file__, "--case", n], capture_output=True, text=True, timeout=240)
            tail = (r.stdout + r.stderr).strip().splitlines()[-6:]
            print(f"[{n}] rc={r.returncode}\n  " + "\n  ".join(tail), flush=True)
        except subprocess.TimeoutExpired:
            print(f"[{n}] TIMEOUT", flush=True)
