"""Module-level drop-ins: ``install()`` routes the torch.ao modules the reference's ``prepare_qat`` call creates
(ref/src/training/qat_trainer.py:306-307) to the sm_100a kernels, for CUDA fp32 tensors, WITHOUT changing module types,
attributes, buffers, hooks or state_dict keys -- so ``QATWrapper``, the qconfig plumbing, ``best_qat.pth`` and the stock
``convert()`` flow are untouched.  It replaces

* ``FusedMovingAvgObsFakeQuantize.forward``      (torch/ao/quantization/fake_quantize.py:423-438)
* ``torch.ao.nn.qat.Linear.forward``             (torch/ao/nn/qat/modules/linear.py:50-51)
* the inline loss of the hot loop -> ``distill_loss`` (ref qat_trainer.py:343-349)

with ``torch.autograd.Function``s whose forward/backward are C-ABI calls (observer state is mutated in the modules' own
buffers, enable flags are read on device).  CPU tensors keep the stock path (that path is the oracle).  This is the generic,
un-fused integration: it works under the reference's unmodified loop and autograd.  ``engine.QATDistillStep`` is the fused one.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import ops
from .ops import Op, PAIRS_EXACT_B, PAIRS_FP32

_ORIG = {}
_INF = float("inf")


def _ensure_per_channel_state(mod, channels: int) -> None:
    obs = mod.activation_post_process
    if obs.min_val.numel() != channels:          # what the ATen op does on its first call
        obs.min_val.resize_(channels).fill_(_INF)
        obs.max_val.resize_(channels).fill_(-_INF)
        mod.scale.resize_(channels).fill_(1.0)
        mod.zero_point.resize_(channels).fill_(0)


class _FakeQuantFn(torch.autograd.Function):
    """y = fused observer + fake-quant of x; backward = STE mask."""

    @staticmethod
    def forward(ctx, x, mod):
        obs = mod.activation_post_process
        xc = x.contiguous()
        y = torch.empty_like(xc)
        mask = torch.empty(xc.shape, dtype=torch.uint8, device=xc.device)
        if mod.is_per_channel:
            if mod.ch_axis != 0:
                raise NotImplementedError("qatvit_b200: per-channel fake-quant is implemented for ch_axis 0 (weights)")
            C = xc.shape[0]
            _ensure_per_channel_state(mod, C)
            ops.fq_weight(xc, True, mod.observer_enabled, mod.fake_quant_enabled, obs.min_val, obs.max_val, mod.scale,
                          mod.zero_point, obs.averaging_constant, obs.quant_min, obs.quant_max, mod.is_symmetric_quant,
                          y=y, mask=mask)
        else:
            acc = ops.new_minmax(xc.device)
            ops.minmax_accumulate(xc, acc)
            ops.obs_update(acc, mod.observer_enabled, mod.fake_quant_enabled, obs.min_val, obs.max_val, mod.scale,
                           mod.zero_point, obs.averaging_constant, obs.quant_min, obs.quant_max, mod.is_symmetric_quant)
            ops.fq_apply(xc, mod.scale, mod.zero_point, mod.fake_quant_enabled, obs.quant_min, obs.quant_max, y=y, mask=mask)
        ctx.save_for_backward(mask)
        return y

    @staticmethod
    def backward(ctx, gy):
        (mask,) = ctx.saved_tensors
        return ops.fq_bwd(gy.contiguous(), mask), None


def _fq_forward(self, X):
    if X.is_cuda and X.dtype == torch.float32 and X.numel() > 0:
        return _FakeQuantFn.apply(X, self)
    return _ORIG["fq"](self, X)


class _LearnableFakeQuantFn(torch.autograd.Function):
    """torch._fake_quantize_learnable_per_channel_affine for channel axis 0: y, STE dx and the per-channel scale / zero-point
    gradients (warp-shuffle reduction, qv_fq_learnable_bwd)."""

    @staticmethod
    def forward(ctx, x, scale, zero_point, qmin, qmax, grad_factor):
        xc, sc, zc = x.contiguous(), scale.contiguous(), zero_point.contiguous()
        ctx.save_for_backward(xc, sc, zc)
        ctx.q = (qmin, qmax, grad_factor)
        return ops.fq_learnable_fwd(xc, sc, zc, qmin, qmax)

    @staticmethod
    def backward(ctx, gy):
        xc, sc, zc = ctx.saved_tensors
        qmin, qmax, gf = ctx.q
        dx, ds, dz = ops.fq_learnable_bwd(gy.contiguous(), xc, sc, zc, qmin, qmax, gf)
        return dx, ds, dz, None, None, None


def _learnable_fq_forward(self, X):
    """_LearnableFakeQuantize.forward (torch/ao/quantization/_learnable_fake_quantize.py:158-196) with the per-channel op on our
    kernels; every other branch (static observation, per-tensor, other axes, CPU) stays stock."""
    per_channel = self.qscheme in (torch.per_channel_symmetric, torch.per_channel_affine)
    if not (X.is_cuda and X.dtype == torch.float32 and X.numel() > 0 and per_channel and self.ch_axis == 0
            and int(self.static_enabled[0]) != 1 and int(self.fake_quant_enabled[0]) == 1):
        return _ORIG["learnable_fq"](self, X)
    self.scale.data.clamp_(min=self.eps.item())
    if self.qscheme == torch.per_channel_symmetric:
        self.zero_point.data.zero_()
    grad_factor = 1.0 / (X.numel() * self.quant_max) ** 0.5 if self.use_grad_scaling else 1.0
    return _LearnableFakeQuantFn.apply(X, self.scale, self.zero_point, self.quant_min, self.quant_max, grad_factor)


def _gemm_friendly(M: int, N: int, K: int) -> bool:
    return K % 8 == 0 and N % 32 == 0 and M > 0


class _QATLinearFn(torch.autograd.Function):
    """F.linear(x, fake_quant(W), b): weight observer + codes, tcgen05 GEMM with the scale in the epilogue; backward =
    dgrad / split-K wgrad on the same tensor-core kernel, weight STE mask applied in the reduce."""

    @staticmethod
    def forward(ctx, x, weight, bias, mod):
        wfq = mod.weight_fake_quant
        obs = wfq.activation_post_process
        N, K = weight.shape
        x2 = x.reshape(-1, K).contiguous()
        M = x2.shape[0]
        dev = x.device
        per_channel = bool(wfq.is_per_channel)
        if per_channel:
            _ensure_per_channel_state(wfq, N)
        codes = torch.empty(1, N, K, dtype=torch.bfloat16, device=dev)
        codes_t = torch.empty(1, K, N, dtype=torch.bfloat16, device=dev)
        wmask = torch.empty(N, K, dtype=torch.uint8, device=dev)
        scratch = torch.zeros(2, dtype=torch.int32, device=dev)
        ops.fq_weight(weight.detach().contiguous(), per_channel, wfq.observer_enabled, wfq.fake_quant_enabled, obs.min_val,
                      obs.max_val, wfq.scale, wfq.zero_point, obs.averaging_constant, obs.quant_min, obs.quant_max,
                      wfq.is_symmetric_quant, mask=wmask, codes=codes[0], codes_t=codes_t[0], scratch=scratch)
        scale_vec = wfq.scale if per_channel else wfq.scale.expand(N).contiguous()
        xp = ops.split_planes(x2)
        out = torch.empty(M, N, dtype=torch.float32, device=dev)
        ops.gemm(Op.full(xp), Op.full(codes), M, N, K, PAIRS_EXACT_B, out=out, col_scale=scale_vec,
                 bias=None if bias is None else bias.detach())
        ctx.save_for_backward(xp, codes_t, wmask, scale_vec.clone())
        ctx.shape = (x.shape, M, N, K, bias is not None)
        return out.reshape(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, gy):
        xp, codes_t, wmask, scale_vec = ctx.saved_tensors
        xshape, M, N, K, has_bias = ctx.shape
        dev = gy.device
        g2 = gy.reshape(M, N).contiguous()
        rpb = 64
        nblk = -(-M // rpb)
        gp = torch.empty(2, M, N, dtype=torch.bfloat16, device=dev)
        part = torch.empty(nblk, N, dtype=torch.float32, device=dev)
        ops.gp_planes(g2, None, None, scale_vec, True, False, M, N, gp, part, rpb)
        gb = None
        if has_bias:
            gb = torch.empty(N, dtype=torch.float32, device=dev)
            ops.colsum_reduce(part, nblk, N, gb)
        gx = torch.empty(M, K, dtype=torch.float32, device=dev)
        ops.gemm(Op.full(gp), Op.full(codes_t), M, K, N, PAIRS_EXACT_B, out=gx)
        from .engine import wgrad_splits
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        s = wgrad_splits(N, K, M, sms)
        gw = torch.empty(N, K, dtype=torch.float32, device=dev)
        if s > 1:
            ws = ops.gemm(Op.full(gp, mn_major=True), Op.full(xp, mn_major=True), N, K, M, PAIRS_FP32, splits=s)
        else:
            ws = ops.gemm(Op.full(gp, mn_major=True), Op.full(xp, mn_major=True), N, K, M, PAIRS_FP32)
        ops.splitk_reduce(ws, s, N, K, gw, row_rscale=scale_vec, mask=wmask)
        return gx.reshape(xshape), gw, gb, None


def _qat_linear_forward(self, input):
    if input.is_cuda and input.dtype == torch.float32 and self.weight.dtype == torch.float32:
        N, K = self.weight.shape
        M = input.numel() // max(K, 1)
        if _gemm_friendly(M, N, K) and type(self.weight_fake_quant).__name__ == "FusedMovingAvgObsFakeQuantize":
            return _QATLinearFn.apply(input, self.weight, self.bias, self)
    # shapes the tensor-core kernel does not take (e.g. the 10-class head): fake-quant still runs on our kernels through
    # the patched FusedMovingAvgObsFakeQuantize.forward; the tiny matmul stays in ATen
    return F.linear(input, self.weight_fake_quant(self.weight), self.bias)


class _DistillLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s, t, labels, T, alpha, eps):
        out3, grad = ops.kd_ce_loss(s.contiguous(), t.contiguous(), labels.contiguous(), T, alpha, eps)
        ctx.save_for_backward(grad)
        return out3[0].clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None, None


def distill_loss(student_out: torch.Tensor, teacher_out: torch.Tensor, labels: torch.Tensor, kd_temp: float = 4.0,
                 kd_alpha: float = 0.5, label_smoothing: float = 0.1) -> torch.Tensor:
    """Drop-in for ref qat_trainer.py:343-349: alpha*T^2*KL(softmax(t/T)||softmax(s/T)) + (1-alpha)*CE_ls(s, y)."""
    if not (student_out.is_cuda and student_out.dtype == torch.float32):
        raise RuntimeError("qatvit_b200.distill_loss needs CUDA fp32 logits (there is no CPU fallback)")
    return _DistillLossFn.apply(student_out, teacher_out.detach(), labels, float(kd_temp), float(kd_alpha),
                                float(label_smoothing))


def install(learnable: bool = False) -> None:
    """Patch the torch.ao module classes in place (idempotent).  Module TYPES stay exactly the stock ones.
    learnable=True (default OFF: the reference has no learnable scale) also routes
    ``_LearnableFakeQuantize.forward`` (per-channel, axis 0) to qv_fq_learnable_fwd / _bwd."""
    from torch.ao.quantization.fake_quantize import FusedMovingAvgObsFakeQuantize
    import torch.ao.nn.qat as nnqat
    if learnable and "learnable_fq" not in _ORIG:
        from torch.ao.quantization._learnable_fake_quantize import _LearnableFakeQuantize
        _ORIG["learnable_fq"] = _LearnableFakeQuantize.forward
        _LearnableFakeQuantize.forward = _learnable_fq_forward
    if "fq" in _ORIG:
        return
    _ORIG["fq"] = FusedMovingAvgObsFakeQuantize.forward
    _ORIG["linear"] = nnqat.Linear.forward
    FusedMovingAvgObsFakeQuantize.forward = _fq_forward
    nnqat.Linear.forward = _qat_linear_forward


def uninstall() -> None:
    from torch.ao.quantization.fake_quantize import FusedMovingAvgObsFakeQuantize
    import torch.ao.nn.qat as nnqat
    if "learnable_fq" in _ORIG:
        from torch.ao.quantization._learnable_fake_quantize import _LearnableFakeQuantize
        _LearnableFakeQuantize.forward = _ORIG.pop("learnable_fq")
    if "fq" not in _ORIG:
        return
    FusedMovingAvgObsFakeQuantize.forward = _ORIG.pop("fq")
    nnqat.Linear.forward = _ORIG.pop("linear")
