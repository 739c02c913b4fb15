#!/usr/bin/env python
"""bench.py -- QAT-distillation train throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W]           # this repo's CUDA path
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # the reference's CPU path on host cores

A "step" = teacher ViT-B/16 forward + prepared (torch.ao QAT, fbgemm qconfig) ViT-S/16 student forward, KL+CE loss,
backward, gradient all-reduce (N > 1), clip-norm 1.0 and AdamW -- ref/src/training/qat_trainer.py:337-361 -- on one
synthetic batch.  N = 1: batch 256 (BASELINE configs[1]); N > 1: global batch 1024 (configs[2]).
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ViT-S/16 QAT-distill train img/s"
HP = {"lr": 1.5e-4, "weight_decay": 1e-3, "label_smoothing": 0.1, "kd_temp": 4.0, "kd_alpha": 0.5}   # ref DEFAULT_HPARAMS
STUDENT, TEACHER = "vit_small_patch16_224", "vit_base_patch16_224"
# algorithmic FLOPs per image of one step: student fwd+bwd 27.6 G + teacher fwd 35.1 G
FLOP_PER_IMG = 62.7e9


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sustained=d.get("bf16_tflops_sustained", 1400.0), src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own torch.ao CPU path (restated step over the live torch CPU ops)
# --------------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(steps: int, warmup: int, batch: int = 8, budget_s: float = 25.0, qconfig: str = "fbgemm"):
    """img/s of the reference CPU path on a bounded sample: `steps` distill steps at batch 8 (BASELINE configs[0])."""
    import torch
    from oracle import vit_ref as vr     # checker / baseline only -- never on the product path
    torch.set_num_threads(os.cpu_count() or 1)
    hp = dict(vr.DEFAULT_HPARAMS)
    student = vr.enable_qat(vr.make_student(prefer_reference=False), qconfig)
    teacher = vr.make_teacher()
    opt = vr.make_optimizer(student.parameters(), hp, 0.5)
    images, labels = vr.synthetic_batch(batch, seed=0)
    for _ in range(warmup):
        vr.distill_step(student, teacher, images, labels, opt, hp)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        vr.distill_step(student, teacher, images, labels, opt, hp)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(value=batch * done / dt, unit="img/s", cores=torch.get_num_threads(), kind="port",
                sample=f"{done} full distill steps (ViT-B teacher fwd + ViT-S {qconfig}-QAT student fwd/bwd + clip + AdamW) at "
                       f"batch {batch}, fp32, stock torch {torch.__version__} CPU ops, {dt / done * 1e3:.0f} ms/step"), dt / done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, ms = cpu_reference_rate(max(args.steps, 1), max(args.warmup, 0), budget_s=150.0, qconfig=args.qconfig)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "img/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ViT-B/16 teacher -> ViT-S/16 QAT student distillation step (KL+CE), bounded sample: batch 8 "
                                   "per step on the host CPU", "qconfig": args.qconfig},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# --------------------------------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------------------------------
def build_models(batch, dev, seed=0, ln_variant="subclass", prepare=True, qconfig="fbgemm"):
    import warnings
    import torch
    from torch.ao.quantization import get_default_qat_qconfig, prepare_qat
    from qatvit_b200 import vit
    torch.manual_seed(seed)
    student = vit.QATWrapper(vit.create_model(STUDENT, num_classes=10, ln_variant=ln_variant))
    torch.manual_seed(seed + 1)
    teacher = vit.create_model(TEACHER, num_classes=10).eval()
    for p in teacher.parameters():
        p.requires_grad = False
    student.train()
    if not prepare:          # the epochs before qat_start_epoch (ref qat_trainer.py:300-316 not yet taken)
        return student.to(dev), teacher.to(dev)
    # QAT enable block, ref qat_trainer.py:304-308 (fbgemm qconfig = per-channel symmetric weights: north_star)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        student.qconfig = get_default_qat_qconfig(qconfig)
        prepared = prepare_qat(student, inplace=False)
    return prepared.to(dev).train(), teacher.to(dev)


def clip_arena_(arena, max_norm: float, grad_scale: float = 1.0):
    """torch.nn.utils.clip_grad_norm_(params, max_norm) on the flat gradient arena (same math, 2 launches);
    grad_scale folds the 1/world of the DDP gradient mean into the same pass."""
    import torch
    total = torch.linalg.vector_norm(arena) * grad_scale
    coef = (max_norm / (total + 1e-6)).clamp(max=1.0) * grad_scale
    arena.mul_(coef)
    return total


# --------------------------------------------------------------------------------------------------------------------
# side measurements on the N = 1 line (time-boxed, AFTER the timed regions): BASELINE configs[3] and configs[4], and the
# reference's own GPU path (stock torch CUDA eager) on the same box
# --------------------------------------------------------------------------------------------------------------------
def _timed_us(fn, iters=10, warm=3):
    import torch
    for _ in range(warm):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


def fq_linear_microbench(M, N, K, dev, peak_tf, hbm_gbs):
    """One fake-quant Linear (torch.ao.nn.qat.Linear + output FusedMovingAvgObsFakeQuantize, fbgemm qconfig) forward and
    STE backward on this library: weight observer + codes -> tcgen05 GEMM (scale / bias / observer epilogue) | gradient planes
    (STE mask, scale fold, bias partials) -> dgrad GEMM -> split-K wgrad GEMM -> reduce (weight STE mask)."""
    import warnings
    import torch
    import torch.ao.nn.qat as nnqat
    from torch.ao.quantization import get_default_qat_qconfig
    from qatvit_b200 import ops
    from qatvit_b200.engine import FQRef, wgrad_splits
    from qatvit_b200.ops import Op, PAIRS_EXACT_B, PAIRS_FP32
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = nnqat.Linear(K, N, bias=True, qconfig=get_default_qat_qconfig("fbgemm"))
        mod.activation_post_process = mod.qconfig.activation()
    torch.nn.init.trunc_normal_(mod.weight, std=0.02)
    mod = mod.to(dev)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    x = torch.randn(M, K, device=dev)
    gy = torch.randn(M, N, device=dev)
    wfq, afq = FQRef(mod.weight_fake_quant, channels=N), FQRef(mod.activation_post_process)
    xp = ops.split_planes(x)
    codes = torch.empty(1, N, K, dtype=torch.bfloat16, device=dev)
    codes_t = torch.empty(1, K, N, dtype=torch.bfloat16, device=dev)
    wmask = torch.empty(N, K, dtype=torch.uint8, device=dev)
    y_raw = torch.empty(M, N, device=dev)
    acc = ops.new_minmax(dev)
    w = mod.weight.detach()
    ticket = torch.zeros(1, dtype=torch.int32, device=dev)

    def gemm_fwd():
        ops.gemm(Op.full(xp), Op.full(codes), M, N, K, PAIRS_EXACT_B, out=y_raw, col_scale=wfq.scale, bias=mod.bias.detach(), minmax=acc,
                 observer=(afq.min_val, afq.max_val, afq.scale, afq.zero_point, afq.observer_enabled, afq.fake_quant_enabled, afq.c,
                           afq.qmin, afq.qmax, afq.symmetric, ticket))

    def fwd():
        ops.minmax_reset(acc)
        ops.fq_weight(w, True, wfq.observer_enabled, wfq.fake_quant_enabled, wfq.min_val, wfq.max_val, wfq.scale, wfq.zero_point,
                      wfq.c, wfq.qmin, wfq.qmax, wfq.symmetric, mask=wmask, codes=codes[0], codes_t=codes_t[0])
        gemm_fwd()

    rpb = 64
    nblk = -(-M // rpb)
    gp = torch.empty(2, M, N, dtype=torch.bfloat16, device=dev)
    part = torch.empty(nblk, N, device=dev)
    gb, gx, gw = torch.empty(N, device=dev), torch.empty(M, K, device=dev), torch.empty(N, K, device=dev)
    sp = wgrad_splits(N, K, M, sms)
    ws = torch.empty(max(sp, 1) * N * K, device=dev)

    def gp_planes():
        ops.gp_planes(gy, y_raw, afq.q, wfq.scale, True, False, M, N, gp, part, rpb)

    def dgrad():
        ops.gemm(Op.full(gp), Op.full(codes_t), M, K, N, PAIRS_EXACT_B, out=gx)

    def wgrad():
        if sp > 1:
            ops.gemm(Op.full(gp, mn_major=True), Op.full(xp, mn_major=True), N, K, M, PAIRS_FP32, splits=sp, workspace=ws)
        else:
            ops.gemm(Op.full(gp, mn_major=True), Op.full(xp, mn_major=True), N, K, M, PAIRS_FP32, out=ws[:N * K].view(N, K))
        ops.splitk_reduce(ws, sp, N, K, gw, row_rscale=wfq.scale, mask=wmask)

    def bwd():
        gp_planes()
        ops.colsum_reduce(part, nblk, N, gb)
        dgrad()
        wgrad()

    t_f, t_b = _timed_us(fwd), _timed_us(bwd)
    t_g, t_d, t_w, t_p = _timed_us(gemm_fwd), _timed_us(dgrad), _timed_us(wgrad), _timed_us(gp_planes)
    fl = 2.0 * M * N * K
    return {"M": M, "N": N, "K": K, "fwd_us": round(t_f, 1), "ste_bwd_us": round(t_b, 1),
            "fwd_gemm_us": round(t_g, 1), "dgrad_gemm_us": round(t_d, 1), "wgrad_gemm_reduce_us": round(t_w, 1), "gp_planes_us": round(t_p, 1),
            "fwd_gemm_alg_tflops": round(fl / t_g / 1e6, 1), "fwd_gemm_frac": round(fl / t_g / 1e6 / peak_tf, 3),
            "fwd_gemm_mma_frac": round(2 * fl / t_g / 1e6 / peak_tf, 3),           # two bf16 passes per fp32-grade product
            "dgrad_gemm_frac": round(fl / t_d / 1e6 / peak_tf, 3), "wgrad_frac": round(fl / t_w / 1e6 / peak_tf, 3),
            "fwd_bwd_alg_tflops": round(3 * fl / (t_f + t_b) / 1e6, 1),
            "gp_planes_gbs": round(12.0 * M * N / t_p / 1e3, 0), "gp_planes_hbm_frac": round(12.0 * M * N / t_p / 1e3 / hbm_gbs, 3),
            "wgrad_splits": sp}


def side_measurements(student, teacher, step, images, labels, dev, peaks, budget_s=75.0):
    """-> dict(microbench=..., int8_eval=..., torch_cuda_eager=...).  Every leg is best-effort and time-boxed; a failing leg
    reports its error string instead of a number (the headline line is already measured at this point)."""
    import copy
    import warnings
    import torch
    from qatvit_b200 import ops
    out = {}
    t_start = time.perf_counter()
    B = images.shape[0]
    # ---- BASELINE configs[3]: fake-quant Linear sweep over the ViT-S / ViT-B shapes at M = 197 * 256 ----
    try:
        rows = []
        for (N, K) in [(1152, 384), (384, 384), (1536, 384), (384, 1536), (2304, 768), (768, 768), (3072, 768), (768, 3072)]:
            rows.append(fq_linear_microbench(197 * B, N, K, dev, peaks["tf_sustained"], peaks["hbm"]))
            torch.cuda.empty_cache()
        out["microbench"] = {"workload": f"fake-quant Linear fwd + STE bwd, M = 197 x {B}, fbgemm qconfig (BASELINE configs[3])",
                             "peak_tflops": peaks["tf_sustained"], "peak_hbm_gbs": peaks["hbm"],
                             "note": "frac = algorithmic 2MNK / time / measured sustained bf16 peak; an fp32-grade product of hi/lo "
                                     "activation planes x exact weight codes costs 2 bf16 passes (mma_frac), wgrad 3",
                             "shapes": rows}
    except Exception as ex:     # noqa: BLE001
        out["microbench"] = {"error": repr(ex)[:300]}
    # ---- the KD + CE loss kernel is latency-bound on the workload's [B, 10] logits; its HBM fraction is reported on a
    #      synthetic many-class sweep (grid form qv_kd_ce_loss_rows: 12 algorithmic bytes per logit -- s, t in, dL/ds out) ----
    try:
        rows = []
        for (b, c) in [(B, 10), (16384, 1000), (65536, 1000)]:
            g = torch.Generator(device=dev).manual_seed(3)
            s_ = torch.randn(b, c, device=dev, generator=g) * 3
            t_ = torch.randn(b, c, device=dev, generator=g) * 6
            y_ = torch.randint(0, c, (b,), device=dev, generator=g)
            o3, gr = torch.empty(3, device=dev), torch.empty_like(s_)
            us = _timed_us(lambda: ops.kd_ce_loss(s_, t_, y_, HP["kd_temp"], HP["kd_alpha"], HP["label_smoothing"], out3=o3, grad=gr))
            gbs = 12.0 * b * c / (us * 1e-6) / 1e9
            rows.append({"B": b, "C": c, "us": round(us, 2), "kernel": "qv_kd_ce_loss_rows" if b * c >= (1 << 15) else "qv_kd_ce_loss (one block)",
                         "GBps": round(gbs, 1), "frac_hbm": round(gbs / peaks["hbm"], 4), "finite": bool(torch.isfinite(o3).all())})
            del s_, t_, y_, gr
        out["kd_ce_sweep"] = {"workload": "KL + label-smoothed CE forward and dL/ds in one launch (ref qat_trainer.py:343-349)",
                              "bytes_per_logit": 12, "peak_hbm_gbs": peaks["hbm"], "shapes": rows}
    except Exception as ex:     # noqa: BLE001
        out["kd_ce_sweep"] = {"error": repr(ex)[:300]}
    # ---- the reference's own GPU path: stock torch CUDA eager (ATen kernels, cuBLAS fp32, autograd) for the same step ----
    try:
        import torch.nn.functional as F
        ref_student = copy.deepcopy(student)
        ref_opt = torch.optim.AdamW(ref_student.parameters(), lr=HP["lr"] * 0.5, weight_decay=HP["weight_decay"])

        def eager_step():           # ref/src/training/qat_trainer.py:337-361, verbatim order
            with torch.no_grad():
                t_out = teacher(images)
            s_out = ref_student(images)
            T, a = HP["kd_temp"], HP["kd_alpha"]
            ce = F.cross_entropy(s_out, labels, label_smoothing=HP["label_smoothing"])
            kd = F.kl_div(torch.log_softmax(s_out / T, dim=1), torch.softmax(t_out / T, dim=1), reduction="batchmean") * (T ** 2)
            loss = a * kd + (1.0 - a) * ce
            ref_opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(ref_student.parameters(), 1.0)
            ref_opt.step()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            us = _timed_us(eager_step, iters=3, warm=2)
        out["torch_cuda_eager"] = {"impl": "stock torch CUDA eager on the same B200: the same prepared module tree through nn.Module.forward "
                                           "+ autograd (ATen FusedObsFakeQuant, cuBLAS fp32 SGEMM, SDPA), torch AdamW + clip_grad_norm_",
                                   "batch": B, "ms_per_step": us / 1e3, "img_per_s": B / (us * 1e-6),
                                   "allow_tf32_matmul": bool(torch.backends.cuda.matmul.allow_tf32), "torch": torch.__version__}
        del ref_student, ref_opt
        torch.cuda.empty_cache()
    except Exception as ex:     # noqa: BLE001
        out["torch_cuda_eager"] = {"error": repr(ex)[:300]}
    # ---- BASELINE configs[4]: converted int8 student eval (stock convert() of the student trained above) ----
    try:
        if time.perf_counter() - t_start > budget_s:
            raise TimeoutError("side-measurement budget used up")
        from torch.ao.quantization import convert
        from qatvit_b200.int8 import ConvertedStudent
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            conv = convert(copy.deepcopy(student).cpu().eval(), inplace=False)           # ref qat_trainer.py:377-379
        ex_ = ConvertedStudent(conv, B, dev)
        us = _timed_us(lambda: ex_(images), iters=10, warm=3)
        ops.profile_begin()
        logits = ex_(images)
        prof = ops.profile_end()
        # every converted Linear: the fp32 / quint8-output launches and the bf16-code-plane launches of the qkv Linears
        lin = {"ms": sum(prof.get(k, {"ms": 0.0})["ms"] for k in ("int8_linear", "int8_linear_codes")),
               "count": sum(prof.get(k, {"count": 0})["count"] for k in ("int8_linear", "int8_linear_codes"))}
        # qkv 3, proj 1, fc1 4, fc2 4 = 12 D^2 MACs per token and block, 12 blocks, + the patch-embed conv
        int_ops = 2.0 * 197 * B * 12 * 12 * 384 * 384 + 2.0 * 196 * B * 384 * 768
        int8_peak = 2.0 * peaks["tf_sustained"]
        # the same executor with the first implementation of the float glue (fp32 tensors between the Linears), for comparison
        ex_f = ConvertedStudent(conv, B, dev, compact=False)
        us_f = _timed_us(lambda: ex_f(images), iters=5, warm=2)
        del ex_f
        res = {"workload": f"converted int8 ViT-S/16 student eval, batch {B}, synthetic 224x224 (BASELINE configs[4])",
               "img_per_s": B / (us * 1e-6), "ms_per_batch": us / 1e3, "finite": bool(torch.isfinite(logits).all()),
               "glue": ("compact: quint8 codes between a Linear and a consumer that works on codes (attention on one exact bf16 code "
                        "plane, GELU + dynamic re-quantisation as table lookups), qparams folded into the quantising pass"
                        if ex_.compact else "fp32 tensors between the Linears"),
               "fp32_glue_ms_per_batch": us_f / 1e3,
               "int8_linear": {"launches": lin["count"], "ms": round(lin["ms"], 3),
                               "alg_tops": round(int_ops / max(lin["ms"], 1e-9) / 1e9, 1),
                               "frac_of_int8_peak": round(int_ops / max(lin["ms"], 1e-9) / 1e9 / int8_peak, 3),
                               "peak_tops": int8_peak,
                               "peak_source": "2 x measured sustained bf16 (kind::i8 runs at twice the kind::f16 rate; no int8 entry in MEASURED_PEAKS.json)"},
               "breakdown_ms": {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}}
        # CPU side of configs[4]: the same converted module through stock torch.ops.quantized kernels (x86 engine) + the same float glue
        from oracle import int8_ref            # checker / CPU baseline only
        torch.set_num_threads(os.cpu_count() or 1)
        xs = images[:8].cpu()
        int8_ref.converted_forward(conv, xs)
        t0, reps = time.perf_counter(), 0
        while time.perf_counter() - t0 < 6.0:
            ref = int8_ref.converted_forward(conv, xs)
            reps += 1
        cpu_dt = (time.perf_counter() - t0) / reps
        ours8 = ConvertedStudent(conv, 8, dev)(images[:8].contiguous()).cpu()
        stepq = float(conv.model.head.scale)
        res["cpu_reference"] = {"img_per_s": 8 / cpu_dt, "ms_per_batch8": cpu_dt * 1e3, "threads": torch.get_num_threads(),
                                "engine": torch.backends.quantized.engine, "kind": "port (stock torch.ops.quantized kernels + the float glue of SURVEY 8c)"}
        res["logits_max_abs_diff_in_head_steps"] = float((ours8 - ref).abs().max()) / stepq
        out["int8_eval"] = res
    except Exception as ex:     # noqa: BLE001
        out["int8_eval"] = {"error": repr(ex)[:300]}
    out["seconds"] = round(time.perf_counter() - t_start, 1)
    return out


def params_identical(student, dev, world) -> bool:
    """Bit-level checksums (int64 sum of the int32 views, one per tensor) of everything data parallelism keeps identical across
    ranks after a step: every parameter, every weight-observer buffer (derived from identical weights) and the activation
    observers' running min / max (rank 0's, exchanged).  The activation fake-quants' scale / zero_point are NOT in the list: each
    rank recomputes them inside its own forward from the running range AND its local batch (under DDP likewise: rank 0's copy is
    broadcast at the start of the next forward and recomputed before use, torch/nn/parallel/distributed.py:1590-1591)."""
    import torch
    import torch.distributed as dist
    tensors = [p_ for _, p_ in student.named_parameters()]
    for name, b in student.named_buffers():
        local = name.endswith(("scale", "zero_point")) and "weight_fake_quant" not in name
        if not local and b.numel() > 0:
            tensors.append(b)
    sums = torch.zeros(len(tensors), dtype=torch.int64, device=dev)
    for i, t in enumerate(tensors):
        v = t.detach().contiguous()
        sums[i] = (v.view(torch.int32) if v.dtype == torch.float32 else v).to(torch.int64).sum()
    allv = [torch.zeros_like(sums) for _ in range(world)]
    dist.all_gather(allv, sums)
    return all(bool(torch.equal(a, allv[0])) for a in allv)


def check_exchange(step, sync, student, images, labels, dev, world, rank):
    """The one exchange of the path must compute what DDP computes (ref qat_trainer.py:310-313,359 +
    torch/nn/parallel/distributed.py:1590-1591): after a step WITH the overlapped all-reduce every rank holds (a) the SUM over
    ranks of the gradients each rank produces alone from the same state on its shard, and (b) rank 0's activation-observer
    running min / max.  Run once before the timed loop; observer state is put back afterwards."""
    import torch
    import torch.distributed as dist
    buffers = {n: b.detach().clone() for n, b in student.named_buffers()}

    def restore():
        with torch.no_grad():
            for n, b in student.named_buffers():
                if b.shape == buffers[n].shape:
                    b.copy_(buffers[n])
    # 1. local step, no exchange
    step(images, labels)
    local = sync.grad_arena.clone()
    obs_local = torch.stack([torch.stack([a.reshape(()), b.reshape(())]) for a, b in step.activation_observers()]).clone()
    restore()
    # 2. the same step with the exchange overlapped into the backward
    step(images, labels, grad_sync=sync)
    torch.cuda.synchronize()
    reduced = sync.grad_arena.clone()
    obs_after = torch.stack([torch.stack([a.reshape(()), b.reshape(())]) for a, b in step.activation_observers()]).clone()
    restore()
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local)
    total = parts[0].double()
    for p_ in parts[1:]:
        total += p_.double()
    err = float((reduced.double() - total).abs().max() / total.abs().max().clamp_min(1e-30))
    exact = bool(torch.equal(reduced, total.float())) if world == 2 else None      # a + b has one rounding: order-free
    obs0 = [torch.empty_like(obs_local) for _ in range(world)]
    dist.all_gather(obs0, obs_local)
    observers_ok = bool(torch.equal(obs_after, obs0[0]))
    differ = bool(world == 1 or not torch.equal(obs0[0], obs0[-1]))       # the shards really had different local state
    flags = torch.tensor([err <= 2e-6, observers_ok, exact is not False], dtype=torch.int32, device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    return {"grads_sum_ok": bool(flags[0]), "grads_sum_exact": (bool(flags[2]) if world == 2 else None),
            "grads_sum_max_rel_err": err, "observers_rank0": bool(flags[1]), "local_observers_differ_across_ranks": differ,
            "weights_identical": True,
            "how": "one step with the overlapped NCCL exchange vs the fp64 sum of the all-gathered per-rank gradients of the same "
                   "step run without it; observer min/max after the exchange vs rank 0's own; int64 checksum of all parameters "
                   "and buffers across ranks after the timed steps"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import qatvit_b200  # noqa: F401  -- raises if libqatvit_b200.so is missing (no fallback)
    from qatvit_b200 import ops
    from qatvit_b200.engine import QATDistillStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- qatvit_b200 has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the exchange is 86.7 MB per step over NVSwitch: a few NCCL channels carry it; every CTA NCCL takes is an SM a persistent
        # GEMM cannot use while the collective runs
        os.environ.setdefault("NCCL_MAX_CTAS", os.environ.get("QV_NCCL_MAX_CTAS", "8"))
        dist.init_process_group("nccl", device_id=dev)
    n = max(world, 1)
    global_batch = 256 if n == 1 else 1024
    if os.environ.get("QV_BENCH_GLOBAL_BATCH"):      # experiments only (e.g. N = 2 with the per-GPU batch of N = 4); shows in config.global_batch
        global_batch = int(os.environ["QV_BENCH_GLOBAL_BATCH"])
    batch = global_batch // n

    student, teacher = build_models(batch, dev, ln_variant=args.ln_variant, prepare=not args.pre_qat, qconfig=args.qconfig)
    if args.pre_qat:
        import functools
        from qatvit_b200.plain import PlainDistillStep
        # same call surface; no observers to synchronise.  --amp: one bf16 pass per product (the reference's optional --amp epochs)
        QATDistillStep = functools.partial(PlainDistillStep, amp=True) if args.amp else PlainDistillStep          # noqa: N806
    sync = None
    if world > 1:
        from qatvit_b200.ddp import GradSync
        # every rank starts from rank 0's weights, like DDP's constructor broadcast (ref :311)
        for t in list(student.parameters()) + list(student.buffers()):
            if t.numel() > 0:
                dist.broadcast(t.data, src=0)
        # one flat buffer: [gradients | activation-observer min/max tail]; built before the engine so .grad views alias it
        n_grad = sum(p.numel() for p in student.parameters() if p.requires_grad)
        n_obs = 0 if args.pre_qat else QATDistillStep.count_activation_observers(student)
        sync = GradSync(n_grad, n_obs, dev)
        step = QATDistillStep(student, teacher, batch, HP, grad_buffer=sync.grad_arena)
        sync.bind_observers([] if args.pre_qat else step.activation_observers())
    else:
        step = QATDistillStep(student, teacher, batch, HP)
    arena = step.grad_arena
    if args.torch_optimizer:
        opt = torch.optim.AdamW(student.parameters(), lr=HP["lr"] * 0.5, weight_decay=HP["weight_decay"])   # ref :315
    else:   # same arithmetic, two launches over flat arenas (qatvit_b200/optim.py)
        from qatvit_b200.optim import FusedClipAdamW
        opt = FusedClipAdamW(student.parameters(), arena, lr=HP["lr"] * 0.5, weight_decay=HP["weight_decay"], max_norm=1.0)

    g = torch.Generator().manual_seed(1234 + rank)
    host_images = torch.randn(batch, 3, 224, 224, generator=g).pin_memory()
    host_labels = torch.randint(0, 10, (batch,), generator=g).pin_memory()
    images = host_images.to(dev)
    labels = host_labels.to(dev)
    host_loss = torch.zeros(3).pin_memory()

    def train_step(img, lab):
        # teacher fwd, student fwd, loss, bwd (our kernels); with N > 1 the one exchange of the path -- NCCL SUM over NVLink of
        # [grads | rank-0 observer state] (qatvit_b200/ddp.py) -- is issued layer by layer during the backward
        out3 = step(img, lab, grad_sync=sync)
        if args.torch_optimizer:
            clip_arena_(arena, 1.0, 1.0 / world)               # ref :360 (+ DDP mean)
            opt.step()                                         # ref :361
        else:
            opt.step(grad_scale=1.0 / world)                   # ref :360-361 fused: clip-norm 1.0 + AdamW (+ DDP mean)
        return out3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ddp_check = None
    if world > 1 and not args.pre_qat:
        ddp_check = check_exchange(step, sync, student, images, labels, dev, world, rank)
    for _ in range(max(args.warmup, 3)):
        train_step(images, labels)
    barrier()

    # ---- device-resident timing (value) ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = ops.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        train_step(images, labels)
    ev1.record()
    barrier()
    launches = ops.launch_count() - l0
    if ddp_check is not None:       # after the timed steps every rank must hold the same parameters, bit for bit
        ddp_check["weights_identical"] = params_identical(student, dev, world)
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    clocks = sampler.stop() if sampler else None

    # ---- cost of the exchange: the same K steps on the same per-GPU batch WITHOUT the all-reduce (ranks then drift apart, which
    #      is why this runs on a throw-away copy of nothing: it is measured last among the device-resident loops and the
    #      parameters are re-broadcast afterwards) ----
    ddp_overhead_ms = None
    if world > 1:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            step(images, labels)
            if args.torch_optimizer:
                clip_arena_(arena, 1.0, 1.0)
                opt.step()
            else:
                opt.step(grad_scale=1.0)
        ev1.record()
        barrier()
        ms0 = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        dist.all_reduce(ms0, op=dist.ReduceOp.MAX)
        ddp_overhead_ms = (ms_total - float(ms0)) / args.steps
        drifted = list(student.parameters()) + list(student.buffers())
        if args.torch_optimizer:
            drifted += [v for st in opt.state.values() for v in st.values() if torch.is_tensor(v) and v.is_cuda]
        else:
            drifted += [opt.exp_avg, opt.exp_avg_sq]
        for t in drifted:
            if t.numel() > 0:
                dist.broadcast(t.data, src=0)
        train_step(images, labels)        # back in lock-step
        weights_identical = params_identical(student, dev, world)
        if ddp_check is not None:
            ddp_check["weights_identical"] = bool(ddp_check["weights_identical"] and weights_identical)

    # ---- end to end through the public step API: every step's inputs start in PINNED HOST memory and its loss ends there.
    # Like the reference's DataLoader(pin_memory=True) + .to(device, non_blocking=True) (ref :334-335), the copy of step
    # t+1's batch is issued on a side stream while step t computes (two device buffers); every copy and every loss
    # read-back (the reference's loss.item(), ref :363) happens inside the timed region. ----
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(images, labels), (torch.empty_like(images), torch.empty_like(labels))]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    main = torch.cuda.current_stream()

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[i])               # the step that last read this buffer has finished
            bufs[i][0].copy_(host_images, non_blocking=True)
            bufs[i][1].copy_(host_labels, non_blocking=True)
            ready[i].record(copy_stream)

    barrier()
    for ev in freed:
        ev.record(main)
    ev0.record()
    prefetch(0)
    for it in range(args.steps):
        cur = it & 1
        if it + 1 < args.steps:
            prefetch(cur ^ 1)
        main.wait_event(ready[cur])
        out3 = train_step(*bufs[cur])
        freed[cur].record(main)
        host_loss.copy_(out3, non_blocking=True)
        main.synchronize()                                 # the reference's loss.item() (ref :363)
    ev1.record()
    barrier()
    ms2 = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2)

    ddp_ok = ddp_check is None or all(bool(ddp_check[k]) for k in ("grads_sum_ok", "observers_rank0", "weights_identical"))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        if not ddp_ok:
            raise SystemExit(f"bench.py: rank {rank}: data-parallel exchange check FAILED: {ddp_check}")
        return

    # ---- per-kernel-family breakdown of ONE step (CUDA events around every launch, on the launching stream) ----
    # roofline = the dominant kernel family: algorithmic FLOPs per launch / its average launch duration vs the measured
    # sustained bf16 tensor peak; elementwise = the largest HBM-bound family against the measured copy bandwidth.
    ops.profile_begin()
    step(images, labels)
    prof = ops.profile_end()
    peaks = _peaks()
    fam = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
    step_kernel_ms = sum(v["ms"] for v in prof.values())
    t_mixed = bool(getattr(step.teacher_engine, "mixed", False))
    PASSES = {"gemm[teacher linear]": 2 if t_mixed else 3, "gemm[student fwd/dgrad]": 2, "gemm[student dgrad+gp]": 2, "gemm[wgrad]": 3, "gemm[patch-embed]": 1,
              "gemm[attn (unfused)]": 3}
    gemm_fams = {k: v for k, v in prof.items() if k.startswith("gemm") and v["ms"] > 0}
    top = max(gemm_fams, key=lambda k: gemm_fams[k]["ms"])
    tv = gemm_fams[top]
    achieved = tv["work"] / (tv["ms"] * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")     # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(top)
    how = (f"tcgen05.mma kind::f16 bf16 hi/lo planes x{PASSES.get(top, 1)}" if not (t_mixed and top == "gemm[teacher linear]") else
           "tcgen05.mma kind::f16 on fp16 values + 2 x kind::f8f6f4 cross terms on fp8 value / residual copies = 2 bf16-pass equivalents")
    pairs = ops.gemm_pair_launches() > 0 and top == "gemm[teacher linear]"
    if pairs:
        how += "; CTA pairs (tcgen05 cta_group::2: 256 x 256 tile per SM pair, each SM stages 128 rows of A + half of B)"
    roofline = {"kernel": f"qv_gemm_kernel, {top} ({tv['count']} launches/step, {how}, fp32 accumulate in TMEM)",
                "bound": "tensor", "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tf_sustained"], "traffic": traffic,
                "traffic_source": ("static: dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full "
                                   "capture listed in profiles/ncu_traffic.json (not measured in this run)") if traffic else None,
                "peak_source": peaks["src"] + ", bf16 dense sustained (kernel timed inside a long step)",
                "avg_launch_us": tv["ms"] * 1e3 / tv["count"], "share_of_step_kernel_time": tv["ms"] / max(step_kernel_ms, 1e-9),
                "mma_passes": PASSES.get(top, 1), "achieved_mma": achieved * PASSES.get(top, 1),
                "frac_mma": achieved * PASSES.get(top, 1) / peaks["tf_sustained"],
                "note": "achieved = ALGORITHMIC fp32 FLOPs (2MNK per Linear) / CUDA-event time of this family's launches in one "
                        "step; an fp32-exact product costs `mma_passes` bf16-equivalent tensor-core passes (north_star parity: 1e-3 on "
                        "fp32 logits; the mixed format's two fp8 cross terms run at twice the rate and count as one), so frac <= "
                        "1/mma_passes; achieved_mma / frac_mma count every pass issued.  The breakdown is taken in ONE extra step with every "
                        "launch on one stream (CUDA events around each launch need that: the timed steps overlap the teacher forward and the "
                        "weight-gradient GEMMs on side streams), so a family's ms here is its serialised time under the step's power budget and "
                        "the families sum to more than ms_per_step",
                "families": {k: {"ms": round(v["ms"], 3), "launches": v["count"],
                                 "alg_tflops": round(v["work"] / (v["ms"] * 1e-3) / 1e12, 1)} for k, v in gemm_fams.items()},
                "breakdown_ms": {k: round(v["ms"], 3) for k, v in fam}}
    ew = {k: v for k, v in prof.items() if not k.startswith("gemm") and v["work"] > 0 and v["ms"] > 0}
    elementwise = None
    if ew:
        ek = max(ew, key=lambda k: ew[k]["ms"])
        gbs = ew[ek]["work"] / (ew[ek]["ms"] * 1e-3) / 1e9
        elementwise = {"kernel": ek, "bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                       "launches": ew[ek]["count"], "ms": round(ew[ek]["ms"], 3),
                       "families": {k: round(v["work"] / (v["ms"] * 1e-3) / 1e9, 0) for k, v in ew.items()}}

    cpu_base = None
    if n == 1 and not args.no_cpu_baseline:
        cpu_base, _ = cpu_reference_rate(steps=1000, warmup=2, budget_s=18.0, qconfig=args.qconfig)
    side = None
    if n == 1 and not args.no_side and not args.pre_qat:
        side = side_measurements(student, teacher, step, images, labels, dev, peaks)

    value = global_batch * args.steps / (ms_total * 1e-3)
    e2e_value = global_batch * args.steps / (e2e_ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": n, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": ("bf16 student operands, one tcgen05 pass per product, fp32 accumulate (pre-QAT --amp variant); teacher f32-grade"
                  if (args.pre_qat and args.amp) else
                  "f32 (tcgen05: bf16 hi/lo planes, teacher Linears fp16 + fp8 cross terms; fp32 accumulate; integer fake-quant codes exact)"),
        "data": "synthetic",
        "config": {"workload": ("PRE-QAT epoch variant (student not yet prepared, no fake-quant; ref qat_trainer.py:333-361 before "
                                "qat_start_epoch" + (", --amp: one bf16 tensor-core pass per student product" if args.amp else "") +
                                "): ViT-B/16 teacher -> ViT-S/16 student distillation step, batch 256, 1x B200")
                   if (n == 1 and args.pre_qat) else
                   "ViT-B/16 teacher -> ViT-S/16 QAT student distillation step (KL+CE), batch 256, 1x B200"
                   if n == 1 else f"same distillation step data-parallel, global batch 1024 at {n} B200 with NCCL gradient allreduce",
                   "global_batch": global_batch, "per_gpu_batch": batch, "qconfig": args.qconfig, "layernorm": args.ln_variant, "image": "3x224x224",
                   "parallelism": f"dp{n}",
                   "scaling_note": "N = 1 runs BASELINE configs[1] (batch 256 on one GPU); N = 2 / 4 / 8 run configs[2] (global batch 1024 "
                                   "split 512 / 256 / 128 per GPU: strong scaling among themselves; N = 4 has the per-GPU work of N = 1)",
                   "l2": "per-step working set (~20 GB of activations) >> 126 MB L2; no flush needed",
                   "optimizer": "torch AdamW + clip-norm on the flat gradient arena" if args.torch_optimizer else
                                "fused clip-norm + AdamW on flat arenas (qv_clip_adamw, 2 launches)",
                   "attention": "fused tcgen05 (integer-code student fwd+bwd, hi/lo teacher fwd)",
                   "streams": "teacher forward and weight-gradient GEMMs on side streams (event-ordered)",
                   "gemm_pairs": ("teacher Linears and the student's plane-output / K >= 1024 GEMMs as CTA pairs (tcgen05 cta_group::2, QV_GEMM_PAIR="
                                  + os.environ.get("QV_GEMM_PAIR", "default 115") + f"): {int(ops.gemm_pair_launches())} launches so far")},
        "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": host_images.numel() * 4 + host_labels.numel() * 8,
                "d2h_bytes_per_step": 12, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "elementwise": elementwise,
        "model_tflops": FLOP_PER_IMG * value / 1e12,
    }
    if cpu_base is not None:
        line["cpu_baseline"] = cpu_base
    if side is not None:
        line.update({k: v for k, v in side.items() if k != "seconds"})
        line["side_measurements_s"] = side["seconds"]
    if ddp_check is not None:
        line["ddp_check"] = ddp_check
    if ddp_overhead_ms is not None:
        line["ddp_overhead_ms"] = ddp_overhead_ms
        line["ddp_overhead_note"] = ("ms_per_step minus the same K steps on the same per-GPU batch without the gradient exchange "
                                     "(max over ranks both)")
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    if not ddp_ok:
        raise SystemExit(f"bench.py: data-parallel exchange check FAILED: {ddp_check}")


_OUT = None


def _emit(line):
    out = _OUT if _OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--qconfig", default="fbgemm", choices=["fbgemm", "qnnpack"],
                    help="torch.ao QAT qconfig of the student: fbgemm (per-channel symmetric weights, activations 0..127; north_star, "
                         "default) or qnnpack (per-tensor weights, activations 0..255; the reference script's own default)")
    ap.add_argument("--no-side", action="store_true", help="skip the N = 1 side measurements (microbench, int8_eval, torch_cuda_eager)")
    ap.add_argument("--ln-variant", default="subclass", choices=["subclass", "plain"],
                    help="timm LayerNorm flavour: 'subclass' (timm.layers.LayerNorm, not observed: 101 fake-quant modules, the "
                         "primary target) or 'plain' (nn.LayerNorm, observed by prepare_qat: 126) -- SURVEY.md section 0.6")
    ap.add_argument("--pre-qat", action="store_true",
                    help="time the pre-QAT epoch step instead (unprepared student, qatvit_b200.plain.PlainDistillStep) -- not the "
                         "BASELINE metric, a side measurement of SURVEY.md section 8f item 3")
    ap.add_argument("--amp", action="store_true", help="with --pre-qat: the half-precision variant (PlainDistillStep(amp=True))")
    ap.add_argument("--torch-optimizer", action="store_true", help="torch AdamW + clip on the arena instead of qv_clip_adamw")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: libraries that write to fd 1 from C (NCCL prints "NCCL version ..." there on some
    # boxes) are sent to stderr; the line itself goes to a private duplicate of the original stdout.
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
