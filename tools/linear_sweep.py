"""BASELINE.json configs[3]: fake-quant Linear microbench sweep over ViT-S/B shapes (tokens 197 x B, dims 384/768/1536/3072),
forward + STE backward, on one B200.  For each shape it times

  ours   : qv_fq_weight -> qv_gemm_bf16 (hi/lo A planes x exact weight codes, scale+bias epilogue, fused output-observer
           min/max) -> qv_obs_update   |   backward: qv_gp_planes (STE mask, scale fold, bias partials) -> dgrad GEMM ->
           split-K wgrad GEMM -> qv_splitk_reduce (weight STE mask)
  torch  : the reference's own module on the same GPU -- torch.ao.nn.qat.Linear + FusedMovingAvgObsFakeQuantize output hook,
           stock ATen CUDA kernels (cuBLAS fp32 SGEMM, TF32 off), autograd backward

with CUDA events (10 iterations after 3 warm-ups, inputs > L2 at the two larger token counts) and prints one JSON line per
shape.  Usage (GPU box):  python tools/linear_sweep.py > gpurun_out/linear_sweep.jsonl
"""
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import qatvit_b200  # noqa: E402,F401
from qatvit_b200 import ops  # noqa: E402
from qatvit_b200.engine import FQRef, wgrad_splits  # noqa: E402
from qatvit_b200.ops import Op, PAIRS_EXACT_B, PAIRS_FP32  # noqa: E402

warnings.simplefilter("ignore")
dev = torch.device("cuda", 0)
SMS = torch.cuda.get_device_properties(dev).multi_processor_count
PEAK = 1362.7
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:
    pass


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3      # us


def make_module(N, K):
    import torch.ao.nn.qat as nnqat
    from torch.ao.quantization import get_default_qat_qconfig
    m = nnqat.Linear(K, N, bias=True, qconfig=get_default_qat_qconfig("fbgemm"))
    m.activation_post_process = m.qconfig.activation()       # what prepare_qat's output hook holds
    torch.nn.init.trunc_normal_(m.weight, std=0.02)
    return m.to(dev)


def bench_shape(M, N, K):
    mod = make_module(N, K)
    x = torch.randn(M, K, device=dev)
    gy = torch.randn(M, N, device=dev)
    # ---------------- ours ----------------
    wfq, afq = FQRef(mod.weight_fake_quant, channels=N), FQRef(mod.activation_post_process)
    xp = ops.split_planes(x)
    codes = torch.empty(1, N, K, dtype=torch.bfloat16, device=dev)
    codes_t = torch.empty(1, K, N, dtype=torch.bfloat16, device=dev)
    wmask = torch.empty(N, K, dtype=torch.uint8, device=dev)
    y_raw = torch.empty(M, N, device=dev)
    acc = ops.new_minmax(dev)
    w = mod.weight.detach()

    def fwd():
        ops.minmax_reset(acc)
        ops.fq_weight(w, True, wfq.observer_enabled, wfq.fake_quant_enabled, wfq.min_val, wfq.max_val, wfq.scale, wfq.zero_point,
                      wfq.c, wfq.qmin, wfq.qmax, wfq.symmetric, mask=wmask, codes=codes[0], codes_t=codes_t[0])
        ops.gemm(Op.full(xp), Op.full(codes), M, N, K, PAIRS_EXACT_B, out=y_raw, col_scale=wfq.scale, bias=mod.bias.detach(), minmax=acc)
        afq.update_from(acc)

    rpb = 64
    nblk = -(-M // rpb)
    gp = torch.empty(2, M, N, dtype=torch.bfloat16, device=dev)
    part = torch.empty(nblk, N, device=dev)
    gb, gx, gw = torch.empty(N, device=dev), torch.empty(M, K, device=dev), torch.empty(N, K, device=dev)
    s = wgrad_splits(N, K, M, SMS)
    ws = torch.empty(max(s, 1) * N * K, device=dev)

    def bwd():
        ops.gp_planes(gy, y_raw, afq.q, wfq.scale, True, False, M, N, gp, part, rpb)
        ops.colsum_reduce(part, nblk, N, gb)
        ops.gemm(Op.full(gp), Op.full(codes_t), M, K, N, PAIRS_EXACT_B, out=gx)
        if s > 1:
            ops.gemm(Op.full(gp, mn_major=True), Op.full(xp, mn_major=True), N, K, M, PAIRS_FP32, splits=s, workspace=ws)
        else:
            ops.gemm(Op.full(gp, mn_major=True), Op.full(xp, mn_major=True), N, K, M, PAIRS_FP32, out=ws[:N * K].view(N, K))
        ops.splitk_reduce(ws, s, N, K, gw, row_rscale=wfq.scale, mask=wmask)

    t_f, t_b = timed(fwd), timed(bwd)
    t_gemm = timed(lambda: ops.gemm(Op.full(xp), Op.full(codes), M, N, K, PAIRS_EXACT_B, out=y_raw, col_scale=wfq.scale,
                                    bias=mod.bias.detach(), minmax=acc))
    t_gp = timed(lambda: ops.gp_planes(gy, y_raw, afq.q, wfq.scale, True, False, M, N, gp, part, rpb))
    # ---------------- stock torch on the same GPU (the reference's module) ----------------
    ref = make_module(N, K)
    xr = x.clone().requires_grad_(True)

    def ref_fwd_bwd():
        y = ref.activation_post_process(ref(xr))
        y.backward(gy)
        xr.grad = None
        ref.weight.grad = None
        ref.bias.grad = None

    def ref_fwd():
        with torch.no_grad():
            ref.activation_post_process(ref(xr))

    t_ref_all, t_ref_f = timed(ref_fwd_bwd, iters=5, warm=2), timed(ref_fwd, iters=5, warm=2)
    fl = 2.0 * M * N * K
    rec = {"M": M, "N": N, "K": K, "ours_fwd_us": round(t_f, 1), "ours_bwd_us": round(t_b, 1),
           "torch_cuda_fwd_us": round(t_ref_f, 1), "torch_cuda_fwd_bwd_us": round(t_ref_all, 1),
           "speedup_fwd_bwd": round(t_ref_all / (t_f + t_b), 2),
           "fwd_gemm_us": round(t_gemm, 1), "fwd_gemm_alg_tflops": round(fl / t_gemm / 1e6, 1),
           "fwd_gemm_bf16_tflops": round(2 * fl / t_gemm / 1e6, 1), "fwd_gemm_frac_of_measured_sustained": round(2 * fl / t_gemm / 1e6 / PEAK, 3),
           "fwd_bwd_alg_tflops": round(3 * fl / (t_f + t_b) / 1e6, 1),
           "gp_planes_us": round(t_gp, 1), "gp_planes_gbs": round(12.0 * M * N / t_gp / 1e3, 0), "wgrad_splits": s}
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    batches = [int(b) for b in os.environ.get("QV_SWEEP_B", "8,128,256").split(",")]
    shapes = [(1152, 384), (384, 384), (1536, 384), (384, 1536), (2304, 768), (768, 768), (3072, 768), (768, 3072)]
    for B in batches:
        for (N, K) in shapes:
            bench_shape(197 * B, N, K)
