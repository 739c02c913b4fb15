"""Module-level drop-ins: ``install()`` routes the torch.ao modules the reference's ``prepare_qat`` call creates
(ref/src/training/qat_trainer.py:306-307) to the sm_100a kernels, for CUDA fp32 tensors, WITHOUT changing module types,
attributes, buffers, hooks or state_dict keys -- so ``QATWrapper``, the qconfig plumbing, ``best_qat.pth`` and the stock
``convert()`` flow are untouched.  It replaces

* ``FusedMovingAvgObsFakeQuantize.forward``      (torch/ao/quantization/fake_quantize.py:423-438)
* ``torch.ao.nn.qat.Linear.forward``             (torch/ao/nn/qat/modules/linear.py:50-51; the 10-class head included)
* ``torch.ao.nn.qat.Conv2d.forward``             (torch/ao/nn/qat/modules/conv.py:55-56; the patch embedding, as im2col + GEMM)
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
# which route each patched module call took since install() (tests assert that the reference's prepared student leaves nothing
# on the stock F.linear / cuDNN route)
stats = {"linear_gemm": 0, "linear_small": 0, "linear_stock": 0, "conv_gemm": 0, "conv_stock": 0}


def _ensure_per_channel_state(mod, channels: int) -> None:
    obs = mod.activation_post_process
    if obs.min_val.numel() != channels:          # what the ATen op does on its first call
        obs.min_val.resize_(channels).fill_(_INF)
        obs.max_val.resize_(channels).fill_(-_INF)
        mod.scale.resize_(channels).fill_(1.0)
        mod.zero_point.resize_(channels).fill_(0)


def _dense_view(x: torch.Tensor):
    """(contiguous tensor holding x's elements, restore) -- per-tensor fake-quant is elementwise, so a channels_last 4-D tensor
    (what the patch-embedding drop-in returns) is processed in its own memory order instead of being copied to NCHW."""
    if x.is_contiguous():
        return x, None
    if x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last):
        return x.permute(0, 2, 3, 1), (0, 3, 1, 2)
    return x.contiguous(), None


class _FakeQuantFn(torch.autograd.Function):
    """y = fused observer + fake-quant of x; backward = STE mask."""

    @staticmethod
    def forward(ctx, x, mod):
        obs = mod.activation_post_process
        if mod.is_per_channel:
            if mod.ch_axis != 0:
                raise NotImplementedError("qatvit_b200: per-channel fake-quant is implemented for ch_axis 0 (weights)")
            xc, perm = x.contiguous(), None
            y = torch.empty_like(xc)
            mask = torch.empty(xc.shape, dtype=torch.uint8, device=xc.device)
            C = xc.shape[0]
            _ensure_per_channel_state(mod, C)
            ops.fq_weight(xc, True, mod.observer_enabled, mod.fake_quant_enabled, obs.min_val, obs.max_val, mod.scale,
                          mod.zero_point, obs.averaging_constant, obs.quant_min, obs.quant_max, mod.is_symmetric_quant,
                          y=y, mask=mask)
        else:
            xc, perm = _dense_view(x)
            y = torch.empty_like(xc)
            mask = torch.empty(xc.shape, dtype=torch.uint8, device=xc.device)
            acc = mod.__dict__.get("_qv_acc")
            if acc is None or acc.device != xc.device:
                acc = mod.__dict__["_qv_acc"] = ops.new_minmax(xc.device)
            else:
                ops.minmax_reset(acc)
            ops.minmax_accumulate(xc, acc)
            ops.obs_update(acc, mod.observer_enabled, mod.fake_quant_enabled, obs.min_val, obs.max_val, mod.scale,
                           mod.zero_point, obs.averaging_constant, obs.quant_min, obs.quant_max, mod.is_symmetric_quant)
            ops.fq_apply(xc, mod.scale, mod.zero_point, mod.fake_quant_enabled, obs.quant_min, obs.quant_max, y=y, mask=mask)
        ctx.save_for_backward(mask)
        ctx.perm = perm
        return y if perm is None else y.permute(*perm)

    @staticmethod
    def backward(ctx, gy):
        (mask,) = ctx.saved_tensors
        if ctx.perm is None:
            return ops.fq_bwd(gy.contiguous(), mask), None
        g = ops.fq_bwd(gy.permute(0, 2, 3, 1).contiguous(), mask)       # mask is in channels_last order
        return g.permute(*ctx.perm), None


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


def _flag(t: torch.Tensor, owner) -> int:
    """int(t[0]) of an enable flag that lives on the device, WITHOUT a sync per call: the value is cached on the owning module
    and re-read only when the buffer was modified in place (``_version`` moves on ``t[0] = 0``, ``copy_``, load_state_dict) or
    replaced (``.to()``)."""
    key = (t.data_ptr(), t._version)
    c = owner.__dict__.get("_qv_flag")
    if c is None or c[0] != key:
        c = (key, int(t.reshape(-1)[0].item()))
        owner.__dict__["_qv_flag"] = c
    return c[1]


class _WeightCache:
    """Per-module scratch for everything that is a function of the WEIGHT only (codes, transposed codes, STE mask, the
    fake-quantised fp32 copy of a small weight, split-K workspace): allocated once, overwritten by every forward -- the stock
    module re-runs weight_fake_quant per call too, and so do we (the observer's EMA needs it), but without a cudaMalloc-class
    call or a memset launch per step.  Buffers a pending backward still needs are never overwritten: a second forward before
    the first one's backward (weight sharing, two passes per step) gets fresh tensors instead (``pending``)."""

    def __init__(self):
        self.key, self.bufs, self.pending, self.ws = None, None, 0, None

    def __deepcopy__(self, memo):
        return _WeightCache()

    def get(self, N: int, K: int, dev, small: bool, fresh: bool):
        key = (N, K, dev, small)
        if fresh or self.key != key or self.bufs is None:
            b = dict(wmask=torch.empty(N, K, dtype=torch.uint8, device=dev), scratch=torch.zeros(2, dtype=torch.int32, device=dev),
                     scale_vec=torch.empty(N, dtype=torch.float32, device=dev))
            if small:
                b["wq"] = torch.empty(N, K, dtype=torch.float32, device=dev)
            else:
                b["codes"] = torch.empty(1, N, K, dtype=torch.bfloat16, device=dev)
                b["codes_t"] = torch.empty(1, K, N, dtype=torch.bfloat16, device=dev)
            if fresh:
                return b
            self.key, self.bufs = key, b
        return self.bufs

    def workspace(self, n: int, dev) -> torch.Tensor:
        if self.ws is None or self.ws.numel() < n or self.ws.device != dev:
            self.ws = torch.empty(n, dtype=torch.float32, device=dev)
        return self.ws


def _cache(mod) -> _WeightCache:
    c = mod.__dict__.get("_qv_cache")
    if c is None:
        c = mod.__dict__["_qv_cache"] = _WeightCache()
    return c


def _quantize_weight(mod, w2: torch.Tensor, small: bool, track: bool):
    """weight_fake_quant(weight) on our kernel: observer EMA + qparams in the module's own buffers, STE mask, and either the
    integer codes (both orientations, the tensor-core operands) or the fake-quantised fp32 weight (small Linears)."""
    wfq = mod.weight_fake_quant
    obs = wfq.activation_post_process
    N, K = w2.shape
    per_channel = bool(wfq.is_per_channel)
    if per_channel:
        _ensure_per_channel_state(wfq, N)
    cache = _cache(mod)
    b = cache.get(N, K, w2.device, small, fresh=track and cache.pending > 0)
    if small:
        ops.fq_weight(w2, per_channel, wfq.observer_enabled, wfq.fake_quant_enabled, obs.min_val, obs.max_val, wfq.scale,
                      wfq.zero_point, obs.averaging_constant, obs.quant_min, obs.quant_max, wfq.is_symmetric_quant,
                      y=b["wq"], mask=b["wmask"], scratch=b["scratch"])
    else:
        ops.fq_weight(w2, per_channel, wfq.observer_enabled, wfq.fake_quant_enabled, obs.min_val, obs.max_val, wfq.scale,
                      wfq.zero_point, obs.averaging_constant, obs.quant_min, obs.quant_max, wfq.is_symmetric_quant,
                      mask=b["wmask"], codes=b["codes"][0], codes_t=b["codes_t"][0], scratch=b["scratch"], scale_vec=b["scale_vec"])
    if track:
        cache.pending += 1
    return b, cache


def _codes_linear_fwd(xp: torch.Tensor, M: int, N: int, K: int, b: dict, bias) -> torch.Tensor:
    out = torch.empty(M, N, dtype=torch.float32, device=xp.device)
    ops.gemm(Op.full(xp), Op.full(b["codes"]), M, N, K, PAIRS_EXACT_B, out=out, col_scale=b["scale_vec"],
             bias=None if bias is None else bias.detach())
    return out


def _codes_linear_bwd(g2: torch.Tensor, xp: torch.Tensor, b: dict, cache: _WeightCache, M: int, N: int, K: int, has_bias: bool,
                      need_gx: bool):
    """dgrad / split-K wgrad of y = x (codes * scale)^T + bias; the weight STE mask and 1/scale are applied in the reduce."""
    dev = g2.device
    rpb = 64
    nblk = -(-M // rpb)
    gp = torch.empty(2, M, N, dtype=torch.bfloat16, device=dev)
    part = torch.empty(nblk, N, dtype=torch.float32, device=dev)
    ops.gp_planes(g2, None, None, b["scale_vec"], True, False, M, N, gp, part, rpb)
    gb = None
    if has_bias:
        gb = torch.empty(N, dtype=torch.float32, device=dev)
        ops.colsum_reduce(part, nblk, N, gb)
    gx = None
    if need_gx:
        gx = torch.empty(M, K, dtype=torch.float32, device=dev)
        ops.gemm(Op.full(gp), Op.full(b["codes_t"]), M, K, N, PAIRS_EXACT_B, out=gx)
    from .engine import wgrad_splits
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    s = wgrad_splits(N, K, M, sms)
    ws = cache.workspace(s * N * K, dev)
    gw = torch.empty(N, K, dtype=torch.float32, device=dev)
    if s > 1:
        ops.gemm(Op.full(gp, mn_major=True), Op.full(xp, mn_major=True), N, K, M, PAIRS_FP32, splits=s, workspace=ws)
    else:
        ops.gemm(Op.full(gp, mn_major=True), Op.full(xp, mn_major=True), N, K, M, PAIRS_FP32, out=ws[:N * K].view(N, K))
    ops.splitk_reduce(ws, s, N, K, gw, row_rscale=b["scale_vec"], mask=b["wmask"])
    cache.pending = max(0, cache.pending - 1)
    return gx, gw, gb


class _QATLinearFn(torch.autograd.Function):
    """F.linear(x, fake_quant(W), b): weight observer + codes, tcgen05 GEMM with the scale in the epilogue; backward =
    dgrad / split-K wgrad on the same tensor-core kernel, weight STE mask applied in the reduce."""

    @staticmethod
    def forward(ctx, x, weight, bias, mod):
        N, K = weight.shape
        x2 = x.reshape(-1, K).contiguous()
        M = x2.shape[0]
        track = torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad)
        b, cache = _quantize_weight(mod, weight.detach().contiguous(), small=False, track=track)
        xp = ops.split_planes(x2)
        out = _codes_linear_fwd(xp, M, N, K, b, bias)
        ctx.saved = (xp, b, cache)
        ctx.shape = (x.shape, M, N, K, bias is not None)
        return out.reshape(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, gy):
        xp, b, cache = ctx.saved
        xshape, M, N, K, has_bias = ctx.shape
        gx, gw, gb = _codes_linear_bwd(gy.reshape(M, N).contiguous(), xp, b, cache, M, N, K, has_bias, ctx.needs_input_grad[0])
        return (None if gx is None else gx.reshape(xshape)), gw, gb, None


class _SmallLinearFn(torch.autograd.Function):
    """Linears the tensor-core kernel does not take (the 10-class head: N = 10): one warp per output in fp32 FMA order, on the
    fake-quantised fp32 weight; backward = the same small kernel the fused engine uses (qv_head_fwd / qv_head_bwd)."""

    @staticmethod
    def forward(ctx, x, weight, bias, mod):
        N, K = weight.shape
        x2 = x.reshape(-1, K).contiguous()
        M = x2.shape[0]
        track = torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad)
        b, cache = _quantize_weight(mod, weight.detach().contiguous(), small=True, track=track)
        out = torch.empty(M, N, dtype=torch.float32, device=x.device)
        ops.head_fwd(x2, b["wq"], None if bias is None else bias.detach(), M, K, N, out)
        ctx.saved = (x2, b, cache)
        ctx.shape = (x.shape, M, N, K, bias is not None)
        return out.reshape(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, gy):
        x2, b, cache = ctx.saved
        xshape, M, N, K, has_bias = ctx.shape
        dev = gy.device
        gx = torch.empty(M, K, dtype=torch.float32, device=dev)
        gw = torch.empty(N, K, dtype=torch.float32, device=dev)
        gb = torch.empty(N, dtype=torch.float32, device=dev)
        ops.head_bwd(gy.reshape(M, N).contiguous(), x2, b["wq"], b["wmask"], M, K, N, gx, gw, gb)
        cache.pending = max(0, cache.pending - 1)
        return gx.reshape(xshape), gw, (gb if has_bias else None), None


_SMALL_MAX_ROWS = 4096       # qv_head_bwd walks the batch serially per weight element: keep it to head-sized problems


def _is_fused_fq(m) -> bool:
    return type(m).__name__ == "FusedMovingAvgObsFakeQuantize"


def _qat_linear_forward(self, input):
    if input.is_cuda and input.dtype == torch.float32 and self.weight.dtype == torch.float32 and _is_fused_fq(self.weight_fake_quant) \
            and input.numel() > 0 and _flag(self.weight_fake_quant.fake_quant_enabled, self.weight_fake_quant) == 1:
        N, K = self.weight.shape
        M = input.numel() // max(K, 1)
        if _gemm_friendly(M, N, K):
            stats["linear_gemm"] += 1
            return _QATLinearFn.apply(input, self.weight, self.bias, self)
        if M <= _SMALL_MAX_ROWS and N <= 1024:
            stats["linear_small"] += 1
            return _SmallLinearFn.apply(input, self.weight, self.bias, self)
    stats["linear_stock"] += int(input.is_cuda)      # CPU tensors are the oracle path, not a missed route
    # fake-quant switched off (weights are then not integer codes), exotic fake-quant classes, huge odd shapes: the stock
    # composition -- the fake-quant itself still runs on our kernels through the patched FusedMovingAvgObsFakeQuantize.forward
    return F.linear(input, self.weight_fake_quant(self.weight), self.bias)


class _QATConv2dFn(torch.autograd.Function):
    """Patch-embedding convolution (kernel == stride, no padding / dilation / groups: timm PatchEmbed.proj, the only conv of the
    reference path) as im2col + the fake-quant Linear GEMM above.  Output is the NCHW view of the [B * patches, N] GEMM result
    (channels_last memory), so timm's flatten(2).transpose(1, 2) that follows is a view again."""

    @staticmethod
    def forward(ctx, x, weight, bias, mod):
        B, C, H, W = x.shape
        N = weight.shape[0]
        ps = weight.shape[2]
        K = C * ps * ps
        P = (H // ps) * (W // ps)
        M = B * P
        track = torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad)
        b, cache = _quantize_weight(mod, weight.detach().reshape(N, K).contiguous(), small=False, track=track)
        xp = torch.empty(2, M, K, dtype=torch.bfloat16, device=x.device)
        ops.im2col_fq(x.contiguous(), None, B, C, H, ps, xp)
        out = _codes_linear_fwd(xp, M, N, K, b, bias)
        ctx.saved = (xp, b, cache)
        ctx.shape = (x.shape, weight.shape, M, N, K, ps, bias is not None)
        return out.view(B, H // ps, W // ps, N).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        xp, b, cache = ctx.saved
        xshape, wshape, M, N, K, ps, has_bias = ctx.shape
        B, C, H, W = xshape
        g2 = gy.permute(0, 2, 3, 1).reshape(M, N).contiguous()       # a view when gy arrives in the layout we produced
        gx, gw, gb = _codes_linear_bwd(g2, xp, b, cache, M, N, K, has_bias, ctx.needs_input_grad[0])
        if gx is not None:   # col2im of non-overlapping patches is a permutation (the input of the reference's patch embedding
            # is the fake-quantised image and needs no gradient: this branch only serves other callers)
            gx = gx.view(B, H // ps, W // ps, C, ps, ps).permute(0, 3, 1, 4, 2, 5).reshape(B, C, H, W)
        return gx, gw.view(wshape), gb, None


def _qat_conv2d_forward(self, input):
    w = self.weight
    if (input.is_cuda and input.dtype == torch.float32 and w.dtype == torch.float32 and input.dim() == 4 and input.numel() > 0
            and _is_fused_fq(self.weight_fake_quant) and self.groups == 1 and tuple(self.dilation) == (1, 1)
            and tuple(self.padding) == (0, 0) and self.padding_mode == "zeros" and w.shape[2] == w.shape[3]
            and tuple(self.stride) == (w.shape[2], w.shape[3]) and input.shape[2] == input.shape[3]
            and input.shape[2] % w.shape[2] == 0 and _gemm_friendly(1, w.shape[0], w.shape[1] * w.shape[2] * w.shape[3])
            and _flag(self.weight_fake_quant.fake_quant_enabled, self.weight_fake_quant) == 1):
        stats["conv_gemm"] += 1
        return _QATConv2dFn.apply(input, w, self.bias, self)
    stats["conv_stock"] += int(input.is_cuda)
    return _ORIG["conv2d"](self, input)          # any other convolution: stock (cuDNN) path, fake-quant on our kernels


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
    _ORIG["conv2d"] = nnqat.Conv2d.forward
    FusedMovingAvgObsFakeQuantize.forward = _fq_forward
    nnqat.Linear.forward = _qat_linear_forward
    nnqat.Conv2d.forward = _qat_conv2d_forward


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
    nnqat.Conv2d.forward = _ORIG.pop("conv2d")
