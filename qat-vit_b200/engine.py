"""Fused QAT-distillation step executor (the hot path of bdina9/qat-vit on B200).

Replaces, for one training iteration, everything between ``images.to(device)`` and ``optimizer.step()`` in
ref/src/training/qat_trainer.py:337-361:

    teacher(images)  (no_grad)                      -> TeacherEngine.forward
    ddp_model(images)  [QATWrapper -> prepared ViT] -> StudentEngine.forward
    KL + CE loss                                    -> qv_kd_ce_loss (inside StudentEngine.forward)
    loss.backward()                                 -> StudentEngine.backward  (hand-written backward, no autograd)

The module tree is NOT modified: parameters stay the torch Parameters of the prepared ``QATWrapper`` (their
``.grad`` become views into one flat gradient arena), observer state stays in the buffers of the
``FusedMovingAvgObsFakeQuantize`` modules ``prepare_qat`` created, so ``state_dict()`` (best_qat.pth),
``convert()`` (best_converted.pth), DDP-style gradient all-reduce and the torch optimizer keep working unchanged.

All arithmetic runs in the sm_100a kernels behind the C-ABI (``ops``); this file only sequences launches on
the current CUDA stream (it is CUDA-graph capturable: no host sync, no allocation after construction).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from .ops import Op, Out, PAIRS_EXACT_B, PAIRS_FP32, PAIRS_SINGLE

_INF = float("inf")


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class FQRef:
    """The buffers of one FusedMovingAvgObsFakeQuantize module (SURVEY.md §8b 'state ownership')."""

    def __init__(self, mod: nn.Module, channels: Optional[int] = None):
        from torch.ao.quantization.fake_quantize import FusedMovingAvgObsFakeQuantize
        if not isinstance(mod, FusedMovingAvgObsFakeQuantize):
            raise TypeError(f"expected FusedMovingAvgObsFakeQuantize, got {type(mod)}")
        obs = mod.activation_post_process
        self.mod = mod
        self.qmin, self.qmax = int(obs.quant_min), int(obs.quant_max)
        self.symmetric = bool(mod.is_symmetric_quant)
        self.per_channel = bool(mod.is_per_channel)
        self.c = float(obs.averaging_constant)
        if self.per_channel:
            if channels is None:
                raise ValueError("per-channel fake-quant needs the channel count")
            if obs.min_val.numel() != channels:     # what the ATen op does on its first call
                obs.min_val.resize_(channels).fill_(_INF)
                obs.max_val.resize_(channels).fill_(-_INF)
                mod.scale.resize_(channels).fill_(1.0)
                mod.zero_point.resize_(channels).fill_(0)
        self.min_val, self.max_val = obs.min_val, obs.max_val
        self.scale, self.zero_point = mod.scale, mod.zero_point
        self.observer_enabled, self.fake_quant_enabled = mod.observer_enabled, mod.fake_quant_enabled
        for t in (self.min_val, self.max_val, self.scale, self.zero_point, self.observer_enabled, self.fake_quant_enabled):
            if not t.is_cuda:
                raise RuntimeError("qatvit_b200: move the prepared model to the GPU before building the engine")

    @property
    def q(self):
        """(scale, zero_point, qmin, qmax) for kernels that fake-quantise on load."""
        return (self.scale, self.zero_point, self.qmin, self.qmax)

    def update_from(self, acc: torch.Tensor) -> None:
        ops.obs_update(acc, self.observer_enabled, self.fake_quant_enabled, self.min_val, self.max_val, self.scale,
                       self.zero_point, self.c, self.qmin, self.qmax, self.symmetric)


class _QLinear:
    """One fake-quant Linear / Conv2d-as-GEMM of the student: parameters, observer state, per-step derived operands."""

    def __init__(self, mod: nn.Module, dev, acc: torch.Tensor, small: bool = False):
        self.mod = mod
        self.weight, self.bias = mod.weight, mod.bias
        self.N = mod.weight.shape[0]
        self.K = mod.weight.numel() // self.N
        self.wfq = FQRef(mod.weight_fake_quant, channels=self.N)
        self.afq = FQRef(mod.activation_post_process)
        if self.afq.per_channel:
            raise NotImplementedError("per-channel activation fake-quant is not part of the reference path")
        self.acc = acc                                    # uint32[2] slot of the output observer
        self.small = small
        self.wmask = torch.empty(self.N, self.K, dtype=torch.uint8, device=dev)
        if small:
            self.wq = torch.empty(self.N, self.K, dtype=torch.float32, device=dev)
        else:
            self.codes = torch.empty(1, self.N, self.K, dtype=torch.bfloat16, device=dev)
            self.codes_t = torch.empty(1, self.K, self.N, dtype=torch.bfloat16, device=dev)
        self.scratch = torch.zeros(2, dtype=torch.int32, device=dev)
        # per-output-channel scale vector for the GEMM epilogues (aliases the module buffer when per-channel)
        self.wscale_vec = self.wfq.scale if self.wfq.per_channel else torch.ones(self.N, device=dev)

    def quantize_weight(self) -> None:
        f = self.wfq
        w = self.weight.detach()
        if self.small:
            ops.fq_weight(w, f.per_channel, f.observer_enabled, f.fake_quant_enabled, f.min_val, f.max_val, f.scale,
                          f.zero_point, f.c, f.qmin, f.qmax, f.symmetric, y=self.wq, mask=self.wmask, scratch=self.scratch)
        else:
            ops.fq_weight(w, f.per_channel, f.observer_enabled, f.fake_quant_enabled, f.min_val, f.max_val, f.scale,
                          f.zero_point, f.c, f.qmin, f.qmax, f.symmetric, mask=self.wmask, codes=self.codes[0],
                          codes_t=self.codes_t[0], scratch=self.scratch)
        if not f.per_channel:
            self.wscale_vec.copy_(f.scale.expand(self.N))


def gemm_tile_n(N: int) -> int:
    """N-tile width qv_gemm_bf16 picks for an output width N (mirrors pick_bn in csrc/gemm_sm100.cu)."""
    if N <= 64:
        return 64
    if N % 192 == 0:
        return 192
    if N <= 128 or N % 128 == 0:
        return 128
    return 192 if N > 1024 else 128


def _splits_for(tiles: int, kblocks: int, sms: int, max_splits: int = 64) -> int:
    """Split-K factor for a persistent GEMM with `tiles` output tiles and `kblocks` 64-deep k-blocks on `sms` SMs.
    Work items are dealt round-robin to one CTA per SM, so the kernel lasts  ceil(tiles*s / sms) * ceil(kblocks / s)
    k-block times: pick the smallest s within 5 % of the best such cost (a 2-wave choice that leaves the second wave a
    third full wastes ~30 %), never leaving a split empty."""
    tiles = max(tiles, 1)
    best_cost, cands = None, []
    for s in range(1, min(kblocks, max_splits) + 1):
        per = -(-kblocks // s)
        if -(-kblocks // per) != s:          # would leave an empty split
            continue
        cost = -(-(tiles * s) // sms) * per
        cands.append((s, cost))
        best_cost = cost if best_cost is None else min(best_cost, cost)
    for s, cost in cands:
        if cost <= 1.05 * best_cost:
            return s
    return 1


def wgrad_splits(n_out: int, k_in: int, tokens: int, sms: int) -> int:
    """Split-K factor for weight.grad[n_out, k_in] = gy^T x over `tokens` rows (128 x gemm_tile_n(k_in) output tiles)."""
    tiles = (-(-n_out // 128)) * (-(-k_in // gemm_tile_n(k_in)))
    return _splits_for(tiles, -(-tokens // 64), sms)


class _ViTDims:
    def __init__(self, vit: nn.Module, batch: int):
        pe = vit.patch_embed
        self.B = batch
        self.D = vit.embed_dim
        self.L = len(vit.blocks)
        self.H = vit.blocks[0].attn.num_heads
        self.hd = self.D // self.H
        if self.hd != 64:
            raise NotImplementedError("attention kernels are specialised for head_dim 64 (timm ViT-S/B)")
        self.F = vit.blocks[0].mlp.fc1.weight.shape[0]
        self.C = vit.head.weight.shape[0]
        self.ps = pe.proj.kernel_size[0]
        self.in_ch = pe.proj.in_channels
        self.HW = pe.img_size[0] if hasattr(pe, "img_size") else None
        self.P = pe.num_patches
        self.T = self.P + 1
        self.M = batch * self.T
        self.Kc = self.in_ch * self.ps * self.ps
        self.ldS = _round_up(self.T, 4)
        self.ldP = _round_up(self.T, 8)
        if self.T > 224:
            raise NotImplementedError("the fused attention kernel holds all keys in one tile: at most 224 tokens")
        self.eps = float(vit.blocks[0].norm1.eps)
        self.attn_scale = float(vit.blocks[0].attn.scale)


def _attention_forward(d: _ViTDims, qkvp: torch.Tensor, S: torch.Tensor, Pp: torch.Tensor, o: torch.Tensor) -> None:
    """softmax(Q K^T / 8) V per (image, head) as two batched tcgen05 GEMMs + one row-softmax kernel
    (replaces F.scaled_dot_product_attention in timm Attention.forward)."""
    B, H, T, D = d.B, d.H, d.T, d.D
    BH = B * H
    q_op = Op.tokens(qkvp, B, T, 0, 64)
    k_op = Op.tokens(qkvp, B, T, D, 64)
    ops.gemm(q_op, k_op, T, T, 64, PAIRS_FP32, out=Out.per_head(S, BH, H, T, T), nbatch=BH, batch_inner=H)
    ops.softmax_planes(S, d.ldS, BH * T, T, d.attn_scale, Pp)
    p_op = Op.per_head(Pp, BH, H, T, T)
    v_op = Op.tokens(qkvp, B, T, 2 * D, 64, mn_major=True)
    ops.gemm(p_op, v_op, T, 64, T, PAIRS_FP32, out=Out.tokens(o, B, T, 0, 64), nbatch=BH, batch_inner=H)


class TeacherEngine:
    """Frozen fp32 ViT forward (ref qat_trainer.py:337-338) on the tcgen05 GEMMs, fp32 accumulation.  The block Linears
    take their operands in the mixed format (fp16 value + fp8 copies of the value and of the fp16 rounding residual): one
    kind::f16 product plus two kind::f8f6f4 cross terms at twice the rate -- fp32-grade (~2^-16 per product) for the cost of
    two bf16 passes instead of three.  ``mixed=False`` (or QV_TEACHER_MIX=0) keeps bf16 hi/lo planes and three MMAs per product
    (hi*hi + hi*lo + lo*hi) everywhere; the patch embedding always does."""

    def __init__(self, vit: nn.Module, batch: int, mixed: Optional[bool] = None):
        import os
        self.vit = vit
        dev = next(vit.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("qatvit_b200: the teacher must live on a CUDA device (there is no CPU fallback)")
        self.dev = dev
        d = self.d = _ViTDims(vit, batch)
        M, D, F = d.M, d.D, d.F
        bf, f32 = torch.bfloat16, torch.float32

        if mixed is None:
            mixed = os.environ.get("QV_TEACHER_MIX", "1") != "0"
        self.mixed = mx = bool(mixed) and D % 64 == 0 and F % 64 == 0

        def planes_of(w: torch.Tensor, mix: bool = False) -> torch.Tensor:
            w2 = w.detach().reshape(w.shape[0], -1).contiguous()
            return ops.split_planes_mix(w2, weight=True) if mix else ops.split_planes(w2)

        self.w_conv = planes_of(vit.patch_embed.proj.weight)
        self.blocks = []
        for blk in vit.blocks:
            self.blocks.append(dict(
                n1=(blk.norm1.weight.detach(), blk.norm1.bias.detach()), n2=(blk.norm2.weight.detach(), blk.norm2.bias.detach()),
                qkv=(planes_of(blk.attn.qkv.weight, mx), blk.attn.qkv.bias.detach()),
                proj=(planes_of(blk.attn.proj.weight, mx), blk.attn.proj.bias.detach()),
                fc1=(planes_of(blk.mlp.fc1.weight, mx), blk.mlp.fc1.bias.detach()),
                fc2=(planes_of(blk.mlp.fc2.weight, mx), blk.mlp.fc2.bias.detach())))
        e = lambda *s, dt=f32: torch.empty(*s, dtype=dt, device=dev)  # noqa: E731
        self.img_planes = e(2, d.B * d.P, d.Kc, dt=bf)
        self.p_raw = e(d.B * d.P, D)
        self.x = [e(M, D), e(M, D)]
        self.hp = e(2, M, D, dt=bf)
        self.qkvp = e(2, M, 3 * D, dt=bf)
        self.op = e(2, M, D, dt=bf)
        self.y = e(M, D)
        self.fp = e(2, M, F, dt=bf)
        self.xn = e(d.B, D)
        self.logits = e(d.B, d.C)

    @torch.no_grad()
    def forward(self, images: torch.Tensor) -> torch.Tensor:
        d, v = self.d, self.vit
        B, T, D, F, M = d.B, d.T, d.D, d.F, d.M
        mx = self.mixed
        if tuple(images.shape) != (B, d.in_ch, d.HW, d.HW):
            raise RuntimeError(f"teacher engine built for batch {B}, got {tuple(images.shape)}")
        ops.im2col_fq(images, None, B, d.in_ch, d.HW, d.ps, self.img_planes)
        ops.gemm(Op.full(self.img_planes), Op.full(self.w_conv), B * d.P, D, d.Kc, PAIRS_FP32, out=self.p_raw,
                 bias=v.patch_embed.proj.bias.detach())
        ops.embed_fwd(self.p_raw, None, v.cls_token.detach().reshape(-1), v.pos_embed.detach().reshape(T, D), B, d.P, D,
                      self.x[0])
        cur = 0
        x_in, y_prev = self.x[0], None
        for li, blk in enumerate(self.blocks):
            g, b = blk["n1"]
            if li == 0:
                ops.resid_ln_fwd(x_in, None, None, g, b, d.eps, M, D, h_planes=self.hp, planes_mix=mx)
            else:
                ops.resid_ln_fwd(x_in, y_prev, None, g, b, d.eps, M, D, x_out=self.x[cur ^ 1], h_planes=self.hp, planes_mix=mx)
                cur ^= 1
                x_in = self.x[cur]
            w, bias = blk["qkv"]
            # no observer sits between the teacher's Linears: epilogues emit the next operand's bf16 planes directly
            ops.gemm(Op.full(self.hp), Op.full(w), M, 3 * D, D, PAIRS_FP32, bias=bias, out_planes=self.qkvp, mix=mx)
            # fused softmax attention (bf16 hi/lo q, k, v): scores / probabilities stay in tensor memory, output lands as proj's
            # A operand
            ops.attn_fwd(self.qkvp, B, T, d.H, d.attn_scale, self.op, out_mix=mx)
            w, bias = blk["proj"]
            ops.gemm(Op.full(self.op), Op.full(w), M, D, D, PAIRS_FP32, out=self.y, bias=bias, mix=mx)
            g, b = blk["n2"]
            ops.resid_ln_fwd(x_in, self.y, None, g, b, d.eps, M, D, x_out=self.x[cur ^ 1], h_planes=self.hp, planes_mix=mx)
            cur ^= 1
            x_in = self.x[cur]
            w, bias = blk["fc1"]
            ops.gemm(Op.full(self.hp), Op.full(w), M, F, D, PAIRS_FP32, bias=bias, out_planes=self.fp, gelu=True, mix=mx, out_mix=mx)
            w, bias = blk["fc2"]
            ops.gemm(Op.full(self.fp), Op.full(w), M, D, F, PAIRS_FP32, out=self.y, bias=bias, mix=mx)
            y_prev = self.y
        ops.resid_ln_fwd(x_in, y_prev, None, v.norm.weight.detach(), v.norm.bias.detach(), d.eps, B, D, in_row_stride=T,
                         h_f32=self.xn)
        ops.head_fwd(self.xn, v.head.weight.detach(), v.head.bias.detach(), B, D, d.C, self.logits)
        return self.logits


class StudentEngine:
    """Forward + hand-written backward of the prepared (torch.ao eager-mode QAT) ``QATWrapper`` student."""

    def __init__(self, student: nn.Module, batch: int, hparams: Dict, grad_buffer: Optional[torch.Tensor] = None,
                 fused_attention: Optional[bool] = None, fused_gp: bool = True):
        self.student = student
        vit = self.vit = student.model
        dev = next(student.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("qatvit_b200: the student must live on a CUDA device (there is no CPU fallback)")
        self.dev = dev
        self.hp_ = dict(hparams)
        d = self.d = _ViTDims(vit, batch)
        # with plain nn.LayerNorm blocks (older timm) prepare_qat also observes every LayerNorm output (126
        # fake-quant modules instead of 101).  The LN outputs are then exact integer codes, so qkv / fc1 become single-pass
        # integer GEMMs (codes x codes) and their wgrads two-pass.
        self.ln_obs = hasattr(vit.blocks[0].norm1, "activation_post_process")
        if self.ln_obs != hasattr(vit.norm, "activation_post_process"):
            raise NotImplementedError("mixed observed / unobserved LayerNorm modules")
        self.sms = torch.cuda.get_device_properties(dev).multi_processor_count
        L, M, D, F, B, T = d.L, d.M, d.D, d.F, d.B, d.T
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda *s, dt=f32: torch.empty(*s, dtype=dt, device=dev)  # noqa: E731

        # ---- observer slots: input, conv out, 4 per block, head ----
        n_act = 2 + 4 * L + 1
        n_ln = (2 * L + 1) if self.ln_obs else 0
        self.acc = torch.empty(n_act + n_ln, 2, dtype=torch.int32, device=dev)
        self.obs_ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        import os
        self._fused_obs = os.environ.get("QV_FUSED_OBS", "1") != "0"
        self.fq_in = FQRef(student.quant.activation_post_process)
        self.conv = _QLinear(vit.patch_embed.proj, dev, self.acc[1])
        self.lin: List[Dict[str, _QLinear]] = []
        for i, blk in enumerate(vit.blocks):
            base = 2 + 4 * i
            self.lin.append(dict(qkv=_QLinear(blk.attn.qkv, dev, self.acc[base]), proj=_QLinear(blk.attn.proj, dev, self.acc[base + 1]),
                                 fc1=_QLinear(blk.mlp.fc1, dev, self.acc[base + 2]), fc2=_QLinear(blk.mlp.fc2, dev, self.acc[base + 3])))
        self.head = _QLinear(vit.head, dev, self.acc[n_act - 1], small=True)
        self.all_linears = [self.conv] + [q for blk in self.lin for q in blk.values()] + [self.head]
        # every per-channel weight fake-quant with the same qparams in ONE launch; the rest (per-tensor qnnpack weights, the
        # 10-row head that keeps an fp32 fake-quantised copy) one by one
        self._wgroup, self._wsingle = [], []
        for ql in self.all_linears:
            f = ql.wfq
            same = not self._wgroup or (f.c, f.qmin, f.qmax, f.symmetric) == self._wgroup_key
            if f.per_channel and not ql.small and ql.N % 8 == 0 and ql.K % 4 == 0 and same:
                self._wgroup_key = (f.c, f.qmin, f.qmax, f.symmetric)
                self._wgroup.append(ql)
            else:
                self._wsingle.append(ql)
        self._wtable = None
        self._wptrs = None
        self.ln_fq: List[FQRef] = []          # [norm1_0, norm2_0, norm1_1, ..., final norm] (observed-LN variant only)
        self.ln_acc = [self.acc[n_act + i] for i in range(n_ln)]
        if self.ln_obs:
            for blk in vit.blocks:
                self.ln_fq += [FQRef(blk.norm1.activation_post_process), FQRef(blk.norm2.activation_post_process)]
            self.ln_fq.append(FQRef(vit.norm.activation_post_process))
            if not all(int(f.fake_quant_enabled.item()) != 0 for f in self.ln_fq):
                raise NotImplementedError("observed LayerNorm with fake-quant disabled is not on the fused path")

        # ---- flat gradient arena (the buffer a DDP-style all-reduce runs over) ----
        self.params = [p for p in student.parameters() if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        if grad_buffer is not None:      # e.g. the head of a ddp.GradSync buffer (gradients + observer tail)
            if grad_buffer.numel() != total or grad_buffer.dtype != f32 or grad_buffer.device != dev:
                raise ValueError("grad_buffer must be a flat fp32 tensor with one element per trainable parameter")
            self.grad_arena = grad_buffer
        else:
            self.grad_arena = torch.zeros(total, dtype=f32, device=dev)
        self._grad_views = []
        off = 0
        self._goff = {}
        for p in self.params:
            self._grad_views.append(self.grad_arena[off:off + p.numel()].view_as(p))
            self._goff[id(p)] = off
            off += p.numel()
        self.attach_grads()
        # arena offset below which nothing belongs to block l or later (parameters are laid out in module order)
        self._block_lo = [min(self._goff[id(p)] for p in blk.parameters()) for blk in vit.blocks]
        order_ok = all(self._block_lo[i] < self._block_lo[i + 1] for i in range(L - 1)) and \
            all(self._goff[id(p)] > self._block_lo[-1] for p in list(vit.norm.parameters()) + list(vit.head.parameters()))
        if not order_ok:
            raise RuntimeError("unexpected parameter order: blocks / norm / head must follow each other in the gradient arena")

        # Fused attention works on the integer codes of the fake-quantised q, k, v (FQ(x) = code * scale): one exact bf16
        # plane instead of hi/lo planes, scores never leave tensor memory, backward recomputes P from the saved logsumexp.
        # It needs fake-quant to be ON for every qkv output observer (the reference never turns it off); the flags are read
        # once here -- if any is off, fall back to the unfused hi/lo-plane kernels.
        if fused_attention is None:
            fused_attention = all(int(ql["qkv"].afq.fake_quant_enabled.item()) != 0 for ql in self.lin)
        self.fused_attn = bool(fused_attention)
        # The backward prologue of every block Linear (gradient x STE mask of its output fake-quant [x gelu'] x weight scale ->
        # bf16 planes, bias-grad partial sums; qv_gp_planes) runs inside the kernel that PRODUCES the gradient: the fc2 dgrad
        # GEMM epilogue (for fc1), the attention backward's output stage (for qkv) and the LayerNorm backward (for proj and the
        # previous block's fc2).  fused_gp=False keeps the standalone kernel (parity reference for the fused forms).
        self.fused_gp = bool(fused_gp)
        import os
        self._wstream = torch.cuda.Stream(device=dev) if os.environ.get("QV_OVERLAP_WGRAD", "1") != "0" else None
        self._w_ready = {k: torch.cuda.Event() for k in ("fc2", "fc1", "proj", "qkv", "conv")}
        self._w_done = {k: torch.cuda.Event() for k in ("fc2", "fc1", "proj", "qkv", "conv")}
        self._w_pending = set()

        # ---- forward activations (saved for backward) ----
        self.img_codes = e(1, B * d.P, d.Kc, dt=bf)
        self.p_raw = e(B * d.P, D)
        self.x_in = [e(M, D) for _ in range(L)]
        self.x_mid = [e(M, D) for _ in range(L)]
        # A operand of qkv / fc1: LayerNorm output as hi/lo planes, or (observed LN) one plane of codes + the raw output
        npl = 1 if self.ln_obs else 2
        self.h1p = [e(npl, M, D, dt=bf) for _ in range(L)]
        self.h2p = [e(npl, M, D, dt=bf) for _ in range(L)]
        if self.ln_obs:
            self.h1_raw = [e(M, D) for _ in range(L)]
            self.h2_raw = [e(M, D) for _ in range(L)]
            self.hN_raw = e(M, D)
            self.xn_raw = e(B, D)
            self.xn_mask = e(B, D, dt=torch.uint8)
        self.qkv_raw = [e(M, 3 * D) for _ in range(L)]
        if self.fused_attn:
            self.qkvc = [e(1, M, 3 * D, dt=bf) for _ in range(L)]
            self.lse = [e(B * d.H * T) for _ in range(L)]
        else:
            self.qkvp = [e(2, M, 3 * D, dt=bf) for _ in range(L)]
            self.Pp = [torch.zeros(2, B * d.H * T, d.ldP, dtype=bf, device=dev) for _ in range(L)]
        self.op = [e(2, M, D, dt=bf) for _ in range(L)]
        self.a_raw = [e(M, D) for _ in range(L)]
        self.f_raw = [e(M, F) for _ in range(L)]
        self.gelp = [e(2, M, F, dt=bf) for _ in range(L)]
        self.m_raw = [e(M, D) for _ in range(L)]
        self.stats1 = [(e(M), e(M)) for _ in range(L)]
        self.stats2 = [(e(M), e(M)) for _ in range(L)]
        if not self.fused_attn:
            self.S = e(B * d.H * T, d.ldS)
            self.o = e(M, D)
        self.xcls = e(B, D)
        self.xn = e(B, D)
        self.statsF = (e(B), e(B))
        self.logits_raw = e(B, d.C)
        self.loss3 = e(3)
        self.g_logits = e(B, d.C)

        # ---- backward scratch ----
        self.gx = [e(M, D), e(M, D)]
        self.g_xn = e(B, D)
        self.gpD = e(2, M, D, dt=bf)         # fc2's gradient planes
        self.gpDp = e(2, M, D, dt=bf)        # proj's (separate, so a weight-gradient GEMM on the side stream can still read one
        self.gpF = e(2, M, F, dt=bf)         #         while the main chain already writes the other)
        self.gp3 = e(2, M, 3 * D, dt=bf)
        self.gpP = e(2, B * d.P, D, dt=bf)
        if not self.fused_gp:
            self.g_big = e(M, F)
        self.g_h = e(M, D)
        self.g_op = e(2, M, D, dt=bf)
        if not (self.fused_gp and self.fused_attn):
            self.g_qkv = e(M, 3 * D)
        if not self.fused_attn:
            self.g_o = e(M, D)
            self.dP = e(B * d.H * T, d.ldS)
            self.dSp = torch.zeros(2, B * d.H * T, d.ldP, dtype=bf, device=dev)
        import os
        self.rpb_gp = 64
        self.rpb_ln = int(os.environ.get("QV_RPB_LN", "64"))
        # bias-grad partial sums: standalone gp_planes [M/64][N]; GEMM epilogue [M/32][F]; attention backward [B*mt*4][3D]
        self.slabs_attn = B * (-(-T // 128)) * 4
        self.bias_part = e(max(-(-M // self.rpb_gp) * max(F, 3 * D), -(-M // 32) * F, self.slabs_attn * 3 * D))
        self.ln_part = e(-(-M // self.rpb_ln), 2, D)
        max_ws = 0
        self._splits = {}
        for (n, k, kdim) in [(3 * D, D, M), (D, D, M), (F, D, M), (D, F, M), (D, d.Kc, B * d.P)]:
            s = wgrad_splits(n, k, kdim, self.sms)
            self._splits[(n, k)] = s
            max_ws = max(max_ws, s * n * k)
        self.ws = e(max_ws)

    # ------------------------------------------------------------------------------------------
    def attach_grads(self) -> None:
        """(Re)point every parameter's .grad at its slice of the arena (optimizer.zero_grad(set_to_none) drops them)."""
        for p, g in zip(self.params, self._grad_views):
            p.grad = g

    def _grad(self, p: torch.Tensor) -> torch.Tensor:
        off = self._goff[id(p)]
        return self.grad_arena[off:off + p.numel()]

    def _ln_param_grads(self, norm: nn.Module, nblk: int) -> None:
        D = self.d.D
        gw, gb = self._goff[id(norm.weight)], self._goff[id(norm.bias)]
        if gb == gw + D:
            ops.colsum_reduce(self.ln_part, nblk, 2 * D, self.grad_arena[gw:gw + 2 * D])
        else:  # not adjacent in the arena: reduce into scratch, then copy
            tmp = torch.empty(2 * D, device=self.dev)
            ops.colsum_reduce(self.ln_part, nblk, 2 * D, tmp)
            self._grad(norm.weight).copy_(tmp[:D])
            self._grad(norm.bias).copy_(tmp[D:])

    # ------------------------------------------------------------------------------------------
    def _linear_fwd(self, ql: _QLinear, a_planes: torch.Tensor, M: int, out: torch.Tensor, pairs=PAIRS_EXACT_B, alpha=None):
        """y_raw = x @ (codes*scale)^T + b with the output observer's min/max fused in the epilogue, then EMA + qparams."""
        f = ql.afq      # the observer's EMA + qparams run in the GEMM's tail (last epilogue warp of the grid): no extra launch
        if not self._fused_obs:
            ops.gemm(Op.full(a_planes), Op.full(ql.codes), M, ql.N, ql.K, pairs, out=out, col_scale=ql.wscale_vec, alpha=alpha,
                     bias=ql.bias.detach(), minmax=ql.acc)
            return ql.afq.update_from(ql.acc)
        ops.gemm(Op.full(a_planes), Op.full(ql.codes), M, ql.N, ql.K, pairs, out=out, col_scale=ql.wscale_vec, alpha=alpha,
                 bias=ql.bias.detach(), minmax=ql.acc,
                 observer=(f.min_val, f.max_val, f.scale, f.zero_point, f.observer_enabled, f.fake_quant_enabled, f.c, f.qmin,
                           f.qmax, f.symmetric, self.obs_ticket))

    def _ln_fwd(self, slot: int, norm: nn.Module, x_in, y_raw, fq, x_out, h_out, h_raw, stats) -> None:
        """x_out = x_in + FQ(y_raw); LayerNorm -> the A operand of the next Linear: bf16 hi/lo planes, or (observed LN) raw
        output + fused observer + one plane of integer codes."""
        d = self.d
        g, b = norm.weight.detach(), norm.bias.detach()
        if not self.ln_obs:
            ops.resid_ln_fwd(x_in, y_raw, fq, g, b, d.eps, d.M, d.D, x_out=x_out, h_planes=h_out, mean=stats[0], rstd=stats[1])
            return
        f, acc = self.ln_fq[slot], self.ln_acc[slot]
        ops.resid_ln_fwd(x_in, y_raw, fq, g, b, d.eps, d.M, d.D, x_out=x_out, h_f32=h_raw, mean=stats[0], rstd=stats[1], minmax=acc)
        f.update_from(acc)
        ops.act_planes(h_raw, f.q, False, h_out, codes_only=True)

    def _lin_after_ln(self, ql: _QLinear, slot: int, a, M: int, out) -> None:
        if self.ln_obs:     # codes x codes, the activation scale rides in alpha
            self._linear_fwd(ql, a, M, out, pairs=PAIRS_SINGLE, alpha=self.ln_fq[slot].scale)
        else:
            self._linear_fwd(ql, a, M, out)

    def quantize_weights(self) -> None:
        """Observer + fake-quant of every weight (ref: weight_fake_quant inside each nnqat module's forward)."""
        if self._wgroup:
            # the descriptor table holds raw pointers: rebuild it whenever a weight's storage has moved (an optimizer that
            # re-homes parameters into a flat arena, FusedClipAdamW, does that once after the engine is built)
            ptrs = [ql.weight.data_ptr() for ql in self._wgroup]
            if ptrs != self._wptrs:
                self._wtable, self._wblocks, self._wmaxk = ops.fq_weight_group_table(
                    [dict(w=ql.weight.detach().reshape(ql.N, ql.K), min_val=ql.wfq.min_val, max_val=ql.wfq.max_val,
                          scale=ql.wfq.scale, zero_point=ql.wfq.zero_point, observer_enabled=ql.wfq.observer_enabled,
                          fake_quant_enabled=ql.wfq.fake_quant_enabled, mask=ql.wmask, codes=ql.codes[0], codes_t=ql.codes_t[0])
                     for ql in self._wgroup], self.dev)
                self._wptrs = ptrs
            c, qmin, qmax, sym = self._wgroup_key
            ops.fq_weight_grouped(self._wtable, len(self._wgroup), self._wblocks, self._wmaxk, c, qmin, qmax, sym)
        for ql in self._wsingle:
            ql.quantize_weight()

    def predict(self, images: torch.Tensor) -> torch.Tensor:
        """Forward only: the prepared student's output (fake-quantised logits), as ``ddp_model(images)`` returns it -- the call
        inside ref qat_trainer.py:49-61 (evaluate_fp32).  Like the stock modules, the observers keep following the data unless
        they were disabled (torch.ao.quantization.disable_observer): FusedMovingAvgObsFakeQuantize ignores train / eval mode."""
        self.forward(images, None, None)
        hd = self.head
        y, _ = ops.fq_apply(self.logits_raw, hd.afq.scale, hd.afq.zero_point, hd.afq.fake_quant_enabled, hd.afq.qmin, hd.afq.qmax,
                            want_mask=False)
        return y

    def forward(self, images: torch.Tensor, labels: Optional[torch.Tensor], teacher_logits: Optional[torch.Tensor],
                teacher_ready=None) -> Optional[torch.Tensor]:
        """teacher_ready: optional CUDA event; the current stream waits for it right before the loss (the teacher forward may
        then run concurrently on another stream).  labels = None: stop after the head (predict)."""
        d, v = self.d, self.vit
        B, T, D, F, M, L = d.B, d.T, d.D, d.F, d.M, d.L
        if tuple(images.shape) != (B, d.in_ch, d.HW, d.HW):
            raise RuntimeError(f"student engine built for batch {B}, got {tuple(images.shape)}")
        ops.minmax_reset(self.acc)
        self.quantize_weights()
        # input fake-quant (QuantStub hook) fused into im2col; patch-embed conv as an exact-integer GEMM
        ops.minmax_accumulate(images, self.acc[0])
        self.fq_in.update_from(self.acc[0])
        ops.im2col_fq(images, self.fq_in.q, B, d.in_ch, d.HW, d.ps, self.img_codes)
        self._linear_fwd(self.conv, self.img_codes, B * d.P, self.p_raw, pairs=PAIRS_SINGLE, alpha=self.fq_in.scale)
        ops.embed_fwd(self.p_raw, self.conv.afq.q, v.cls_token.detach().reshape(-1), v.pos_embed.detach().reshape(T, D), B,
                      d.P, D, self.x_in[0])
        for l, blk in enumerate(v.blocks):
            ql = self.lin[l]
            if l == 0:
                self._ln_fwd(0, blk.norm1, self.x_in[0], None, None, None, self.h1p[0], self.h1_raw[0] if self.ln_obs else None,
                             self.stats1[0])
            self._lin_after_ln(ql["qkv"], 2 * l, self.h1p[l], M, self.qkv_raw[l])
            if self.fused_attn:
                qs = ql["qkv"].afq.scale
                ops.act_planes(self.qkv_raw[l], ql["qkv"].afq.q, False, self.qkvc[l], codes_only=True)
                ops.attn_fwd(self.qkvc[l], B, T, d.H, d.attn_scale, self.op[l], qk_scale=qs, v_scale=qs, lse=self.lse[l])
            else:
                ops.act_planes(self.qkv_raw[l], ql["qkv"].afq.q, False, self.qkvp[l])
                _attention_forward(d, self.qkvp[l], self.S, self.Pp[l], self.o)
                ops.split_planes(self.o, self.op[l])
            self._linear_fwd(ql["proj"], self.op[l], M, self.a_raw[l])
            self._ln_fwd(2 * l + 1, blk.norm2, self.x_in[l], self.a_raw[l], ql["proj"].afq.q, self.x_mid[l], self.h2p[l],
                         self.h2_raw[l] if self.ln_obs else None, self.stats2[l])
            self._lin_after_ln(ql["fc1"], 2 * l + 1, self.h2p[l], M, self.f_raw[l])
            ops.act_planes(self.f_raw[l], ql["fc1"].afq.q, True, self.gelp[l])
            self._linear_fwd(ql["fc2"], self.gelp[l], M, self.m_raw[l])
            if l + 1 < L:
                nb = v.blocks[l + 1]
                self._ln_fwd(2 * l + 2, nb.norm1, self.x_mid[l], self.m_raw[l], ql["fc2"].afq.q, self.x_in[l + 1], self.h1p[l + 1],
                             self.h1_raw[l + 1] if self.ln_obs else None, self.stats1[l + 1])
            else:   # final norm: only the cls rows reach the head
                gN, bN = v.norm.weight.detach(), v.norm.bias.detach()
                if self.ln_obs:
                    # the reference normalises (and OBSERVES) all tokens before taking x[:, 0]: min / max over every row
                    fN, accN = self.ln_fq[2 * L], self.ln_acc[2 * L]
                    ops.resid_ln_fwd(self.x_mid[l], self.m_raw[l], ql["fc2"].afq.q, gN, bN, d.eps, M, D, h_f32=self.hN_raw, minmax=accN)
                    fN.update_from(accN)
                    ops.resid_ln_fwd(self.x_mid[l], self.m_raw[l], ql["fc2"].afq.q, gN, bN, d.eps, B, D, in_row_stride=T,
                                     x_out=self.xcls, h_f32=self.xn_raw, mean=self.statsF[0], rstd=self.statsF[1])
                    ops.fq_apply(self.xn_raw, fN.scale, fN.zero_point, fN.fake_quant_enabled, fN.qmin, fN.qmax, y=self.xn,
                                 mask=self.xn_mask)
                else:
                    ops.resid_ln_fwd(self.x_mid[l], self.m_raw[l], ql["fc2"].afq.q, gN, bN, d.eps, B, D, in_row_stride=T,
                                     x_out=self.xcls, h_f32=self.xn, mean=self.statsF[0], rstd=self.statsF[1])
        hd = self.head
        ops.head_fwd(self.xn, hd.wq, hd.bias.detach(), B, D, d.C, self.logits_raw, minmax=hd.acc)
        hd.afq.update_from(hd.acc)
        if labels is None:
            return None
        hp = self.hp_
        if teacher_ready is not None:
            torch.cuda.current_stream().wait_event(teacher_ready)
        ops.kd_ce_loss(self.logits_raw, teacher_logits, labels, hp["kd_temp"], hp["kd_alpha"], hp["label_smoothing"],
                       s_scale=hd.afq.scale, s_zp=hd.afq.zero_point, qmin=hd.afq.qmin, qmax=hd.afq.qmax, out3=self.loss3,
                       grad=self.g_logits)
        return self.loss3

    # ------------------------------------------------------------------------------------------
    def _wgrad(self, ql: _QLinear, gp: torch.Tensor, x_planes: torch.Tensor, kdim: int, pairs, alpha=None, key=None) -> None:
        """weight.grad[N,K] = mask * (gp'^T @ x) / scale[n]  (split-K over the token dimension, deterministic reduce).
        pairs: (2,2) -> gp hi/lo x x hi/lo ; (2,1) -> gp hi/lo x exact codes.
        Nothing downstream in the backward chain reads a weight gradient, so the GEMM + reduce go to a side stream (ordered
        after the kernel that wrote `gp`); `key` names the gp buffer so its next writer can wait for this read (_w_wait)."""
        if self._wstream is None or key is None or ops.profiling():
            return self._wgrad_now(ql, gp, x_planes, kdim, pairs, alpha)
        main = torch.cuda.current_stream()
        self._w_ready[key].record(main)
        with torch.cuda.stream(self._wstream):
            self._wstream.wait_event(self._w_ready[key])
            self._wgrad_now(ql, gp, x_planes, kdim, pairs, alpha)
            self._w_done[key].record(self._wstream)
        self._w_pending.add(key)

    def _w_wait(self, key: str) -> None:
        """The current stream is about to overwrite gp buffer `key`: wait for the side-stream weight-gradient GEMM reading it."""
        if key in self._w_pending:
            torch.cuda.current_stream().wait_event(self._w_done[key])
            self._w_pending.discard(key)

    def _w_join(self) -> None:
        """Every weight gradient issued so far is complete (before gradients are declared final / the optimizer runs)."""
        if self._w_pending:
            torch.cuda.current_stream().wait_stream(self._wstream)
            self._w_pending.clear()

    def _wgrad_now(self, ql: _QLinear, gp: torch.Tensor, x_planes: torch.Tensor, kdim: int, pairs, alpha=None) -> None:
        s = self._splits[(ql.N, ql.K)]
        if s > 1:
            ops.gemm(Op.full(gp, mn_major=True), Op.full(x_planes, mn_major=True), ql.N, ql.K, kdim, pairs, splits=s,
                     workspace=self.ws)
        else:   # a single split writes its (raw) sums straight into slice 0 of the workspace
            ops.gemm(Op.full(gp, mn_major=True), Op.full(x_planes, mn_major=True), ql.N, ql.K, kdim, pairs,
                     out=self.ws[:ql.N * ql.K].view(ql.N, ql.K))
        ops.splitk_reduce(self.ws, s, ql.N, ql.K, self._grad(ql.weight), row_rscale=ql.wscale_vec, alpha=alpha, mask=ql.wmask)

    def _dgrad(self, ql: _QLinear, gp: torch.Tensor, M: int, out: torch.Tensor) -> None:
        ops.gemm(Op.full(gp), Op.full(ql.codes_t), M, ql.K, ql.N, PAIRS_EXACT_B, out=out)

    def _part(self, nblk: int, N: int) -> torch.Tensor:
        return self.bias_part[:nblk * N].view(nblk, N)

    def _gp(self, g, y_raw, ql: _QLinear, gelu: bool, R: int, out_planes, remap=(0, 0)) -> None:
        nblk = -(-R // self.rpb_gp)
        part = self._part(nblk, ql.N)
        ops.gp_planes(g, y_raw, ql.afq.q, ql.wscale_vec, True, gelu, R, ql.N, out_planes, part, self.rpb_gp, remap[0], remap[1])
        ops.colsum_reduce(part, nblk, ql.N, self._grad(ql.bias))

    def backward(self, grads_final_from=None) -> None:
        """grads_final_from(lo): optional callback, called as soon as every gradient with arena offset >= lo is final (after
        the head, after each block, after the embeddings) -- ddp.GradSync uses it to overlap the all-reduce with backward."""
        d, v = self.d, self.vit
        B, T, D, F, M, L, H = d.B, d.T, d.D, d.F, d.M, d.L, d.H
        BH = B * H
        hd = self.head
        ops.head_bwd(self.g_logits, self.xn, hd.wq, hd.wmask, B, D, d.C, self.g_xn, self._grad(hd.weight), self._grad(hd.bias))
        if self.ln_obs:     # STE mask of the final norm's output fake-quant
            ops.fq_bwd(self.g_xn, self.xn_mask, gx=self.g_xn)
        gx, gx2 = self.gx
        gx.zero_()
        nblk_ln = -(-B // self.rpb_ln)
        ops.ln_bwd(self.g_xn, self.xcls, self.statsF[0], self.statsF[1], v.norm.weight.detach(), None, B, D, gx, self.ln_part,
                   self.rpb_ln, out_row_stride=T)
        self._ln_param_grads(v.norm, nblk_ln)
        if grads_final_from is not None:
            grads_final_from(min(self._goff[id(p)] for p in list(v.norm.parameters()) + list(v.head.parameters())))
        nblk_ln = -(-M // self.rpb_ln)
        for l in range(L - 1, -1, -1):
            blk, ql = v.blocks[l], self.lin[l]
            # ---- MLP ----
            if not (self.fused_gp and l < L - 1):     # else: emitted by the LayerNorm backward of block l + 1
                self._w_wait("fc2")
                self._gp(gx, self.m_raw[l], ql["fc2"], False, M, self.gpD)
            if self.fused_gp:
                self._w_wait("fc1")
                # fc2 dgrad with fc1's backward prologue in its epilogue: gelu'(FQ(f_raw)) * STE mask * fc1 weight scale
                nslab = -(-M // 32)
                part = self._part(nslab, F)
                ops.gemm(Op.full(self.gpD), Op.full(ql["fc2"].codes_t), M, F, D, PAIRS_EXACT_B, out_planes=self.gpF,
                         col_scale=ql["fc1"].wscale_vec, grad_of=(self.f_raw[l], ql["fc1"].afq.q, True, part))
                ops.colsum_reduce(part, nslab, F, self._grad(ql["fc1"].bias))
                self._wgrad(ql["fc2"], self.gpD, self.gelp[l], M, PAIRS_FP32, key="fc2")
            else:
                self._dgrad(ql["fc2"], self.gpD, M, self.g_big)
                self._wgrad(ql["fc2"], self.gpD, self.gelp[l], M, PAIRS_FP32, key="fc2")
                self._w_wait("fc1")
                self._gp(self.g_big, self.f_raw[l], ql["fc1"], True, M, self.gpF)
            self._dgrad(ql["fc1"], self.gpF, M, self.g_h)
            # norm2 backward (+ residual grad) also emits proj's gradient planes
            nblk_gp = -(-M // self.rpb_ln)
            part = self._part(nblk_gp, D)
            self._w_wait("proj")
            gp_proj = (self.a_raw[l], ql["proj"].afq.q, ql["proj"].wscale_vec, self.gpDp, part) if self.fused_gp else None
            if self.ln_obs:
                f2 = self.ln_fq[2 * l + 1]
                self._wgrad(ql["fc1"], self.gpF, self.h2p[l], M, PAIRS_EXACT_B, alpha=f2.scale, key="fc1")
                ops.ln_bwd(self.g_h, self.x_mid[l], self.stats2[l][0], self.stats2[l][1], blk.norm2.weight.detach(), gx, M, D, gx2,
                           self.ln_part, self.rpb_ln, h_raw=self.h2_raw[l], h_fq=f2.q, gp=gp_proj)
            else:
                self._wgrad(ql["fc1"], self.gpF, self.h2p[l], M, PAIRS_FP32, key="fc1")
                ops.ln_bwd(self.g_h, self.x_mid[l], self.stats2[l][0], self.stats2[l][1], blk.norm2.weight.detach(), gx, M, D, gx2,
                           self.ln_part, self.rpb_ln, gp=gp_proj)
            self._ln_param_grads(blk.norm2, nblk_ln)
            # ---- attention ----
            if self.fused_gp:
                ops.colsum_reduce(part, nblk_gp, D, self._grad(ql["proj"].bias))
            else:
                self._gp(gx2, self.a_raw[l], ql["proj"], False, M, self.gpDp)
            if self.fused_attn:
                # proj dgrad emits dL/dO directly as bf16 hi/lo planes; one fused kernel recomputes P and writes dQ | dK | dV
                ops.gemm(Op.full(self.gpDp), Op.full(ql["proj"].codes_t), M, D, D, PAIRS_EXACT_B, out_planes=self.g_op)
                self._wgrad(ql["proj"], self.gpDp, self.op[l], M, PAIRS_FP32, key="proj")
                if self.fused_gp:   # ... with qkv's backward prologue applied on the way out
                    self._w_wait("qkv")
                    part = self._part(self.slabs_attn, 3 * D)
                    ops.attn_bwd_gp(self.qkvc[l], ql["qkv"].afq.scale, self.op[l], self.g_op, self.lse[l], B, T, H, d.attn_scale,
                                    self.qkv_raw[l], ql["qkv"].afq.q, ql["qkv"].wscale_vec, self.gp3, part)
                    ops.colsum_reduce(part, self.slabs_attn, 3 * D, self._grad(ql["qkv"].bias))
                else:
                    ops.attn_bwd(self.qkvc[l], ql["qkv"].afq.scale, self.op[l], self.g_op, self.lse[l], B, T, H, d.attn_scale,
                                 self.g_qkv)
            else:
                self._dgrad(ql["proj"], self.gpDp, M, self.g_o)
                self._wgrad(ql["proj"], self.gpDp, self.op[l], M, PAIRS_FP32, key="proj")
                ops.split_planes(self.g_o, self.g_op)
                qkvp, Pp = self.qkvp[l], self.Pp[l]
                # dP = dO V^T
                ops.gemm(Op.tokens(self.g_op, B, T, 0, 64), Op.tokens(qkvp, B, T, 2 * D, 64), T, T, 64, PAIRS_FP32,
                         out=Out.per_head(self.dP, BH, H, T, T), nbatch=BH, batch_inner=H)
                ops.attn_ds(Pp, self.dP, d.ldS, BH * T, T, d.attn_scale, self.dSp)
                # dQ = dS K ; dK = dS^T Q ; dV = P^T dO   -> column blocks of g_qkv
                ops.gemm(Op.per_head(self.dSp, BH, H, T, T), Op.tokens(qkvp, B, T, D, 64, mn_major=True), T, 64, T, PAIRS_FP32,
                         out=Out.tokens(self.g_qkv, B, T, 0, 64), nbatch=BH, batch_inner=H)
                ops.gemm(Op.per_head(self.dSp, BH, H, T, T, mn_major=True), Op.tokens(qkvp, B, T, 0, 64, mn_major=True), T, 64, T,
                         PAIRS_FP32, out=Out.tokens(self.g_qkv, B, T, D, 64), nbatch=BH, batch_inner=H)
                ops.gemm(Op.per_head(Pp, BH, H, T, T, mn_major=True), Op.tokens(self.g_op, B, T, 0, 64, mn_major=True), T, 64, T,
                         PAIRS_FP32, out=Out.tokens(self.g_qkv, B, T, 2 * D, 64), nbatch=BH, batch_inner=H)
            if not (self.fused_gp and self.fused_attn):
                self._w_wait("qkv")
                self._gp(self.g_qkv, self.qkv_raw[l], ql["qkv"], False, M, self.gp3)
            self._dgrad(ql["qkv"], self.gp3, M, self.g_h)
            # norm1 backward also emits the gradient planes of the PREVIOUS block's fc2 (its output joined this residual stream)
            gp_fc2 = None
            if self.fused_gp and l > 0:
                self._w_wait("fc2")
                pf = self.lin[l - 1]["fc2"]
                part = self._part(nblk_gp, D)
                gp_fc2 = (self.m_raw[l - 1], pf.afq.q, pf.wscale_vec, self.gpD, part)
            if self.ln_obs:
                f1 = self.ln_fq[2 * l]
                self._wgrad(ql["qkv"], self.gp3, self.h1p[l], M, PAIRS_EXACT_B, alpha=f1.scale, key="qkv")
                ops.ln_bwd(self.g_h, self.x_in[l], self.stats1[l][0], self.stats1[l][1], blk.norm1.weight.detach(), gx2, M, D, gx,
                           self.ln_part, self.rpb_ln, h_raw=self.h1_raw[l], h_fq=f1.q, gp=gp_fc2)
            else:
                self._wgrad(ql["qkv"], self.gp3, self.h1p[l], M, PAIRS_FP32, key="qkv")
                ops.ln_bwd(self.g_h, self.x_in[l], self.stats1[l][0], self.stats1[l][1], blk.norm1.weight.detach(), gx2, M, D, gx,
                           self.ln_part, self.rpb_ln, gp=gp_fc2)
            self._ln_param_grads(blk.norm1, nblk_ln)
            if gp_fc2 is not None:
                ops.colsum_reduce(part, nblk_gp, D, self._grad(self.lin[l - 1]["fc2"].bias))
            if grads_final_from is not None:
                self._w_join()
                grads_final_from(self._block_lo[l])
        # ---- embeddings: pos_embed, cls_token, patch-embed conv ----
        ops.colsum_rows(gx, B, T * D, T * D, self._grad(v.pos_embed))
        ops.colsum_rows(gx, B, D, T * D, self._grad(v.cls_token))
        self._gp(gx, self.p_raw, self.conv, False, B * d.P, self.gpP, remap=(d.P, T))
        self._wgrad(self.conv, self.gpP, self.img_codes, B * d.P, PAIRS_EXACT_B, alpha=self.fq_in.scale, key="conv")
        self._w_join()
        if grads_final_from is not None:
            grads_final_from(0)


class QATDistillStep:
    """One object per (student, teacher, batch size): ``loss = step(images, labels)`` runs teacher forward, student
    forward, loss and backward on the current stream and leaves gradients in ``student`` parameters' ``.grad``."""

    def __init__(self, student: nn.Module, teacher: nn.Module, batch: int, hparams: Dict,
                 grad_buffer: Optional[torch.Tensor] = None, fused_attention: Optional[bool] = None, fused_gp: bool = True,
                 overlap_teacher: Optional[bool] = None, teacher_mixed: Optional[bool] = None):
        self.student_engine = StudentEngine(student, batch, hparams, grad_buffer=grad_buffer, fused_attention=fused_attention,
                                            fused_gp=fused_gp)
        self.teacher_engine = TeacherEngine(teacher, batch, mixed=teacher_mixed)
        self.grad_arena = self.student_engine.grad_arena
        # The frozen teacher's forward (ref qat_trainer.py:337-338) is independent of the student's until the loss: it runs on a
        # second stream, so each stream's kernel-boundary bubbles (tails of persistent kernels, small reduces / observer updates)
        # are filled by the other's work.
        import os
        if overlap_teacher is None:
            overlap_teacher = os.environ.get("QV_OVERLAP_TEACHER", "1") != "0"
        self.overlap_teacher = bool(overlap_teacher)
        self._tstream = torch.cuda.Stream(device=self.student_engine.dev) if self.overlap_teacher else None
        self._tdone = torch.cuda.Event() if self.overlap_teacher else None

    def __call__(self, images: torch.Tensor, labels: torch.Tensor, grad_sync=None) -> torch.Tensor:
        """grad_sync: a ddp.GradSync whose buffer holds the gradient arena -- its all-reduce then overlaps the backward."""
        if self.overlap_teacher and not ops.profiling():
            main = torch.cuda.current_stream()
            self._tstream.wait_stream(main)                      # images are ready; the previous step's loss has read the logits
            with torch.cuda.stream(self._tstream):
                t_logits = self.teacher_engine.forward(images)
                self._tdone.record(self._tstream)
            out3 = self.student_engine.forward(images, labels, t_logits, teacher_ready=self._tdone)
        else:
            t_logits = self.teacher_engine.forward(images)
            out3 = self.student_engine.forward(images, labels, t_logits)
        if grad_sync is None:
            self.student_engine.backward()
        else:
            grad_sync.begin_step()
            self.student_engine.backward(grads_final_from=grad_sync.grads_final_from)
            grad_sync.end_step()
        return out3

    def predict(self, images: torch.Tensor) -> torch.Tensor:
        """student(images) without loss / backward (validation loop of ref qat_trainer.py:49-61)."""
        return self.student_engine.predict(images)

    @property
    def student_logits_raw(self) -> torch.Tensor:
        return self.student_engine.logits_raw

    def activation_observers(self):
        """[(min_val, max_val)] of every activation fake-quant, in forward order (for ddp.GradSync)."""
        se = self.student_engine
        fqs = [se.fq_in, se.conv.afq] + [ql[k].afq for ql in se.lin for k in ("qkv", "proj", "fc1", "fc2")] + [se.head.afq]
        fqs += se.ln_fq
        return [(f.min_val, f.max_val) for f in fqs]

    @staticmethod
    def count_activation_observers(student: nn.Module) -> int:
        """activation fake-quant modules of the prepared student (51, or 76 with observed LayerNorms) -- the length of
        activation_observers(), needed to size ddp.GradSync before the engine exists."""
        from torch.ao.quantization.fake_quantize import FusedMovingAvgObsFakeQuantize
        return sum(1 for n, m in student.named_modules()
                   if isinstance(m, FusedMovingAvgObsFakeQuantize) and n.endswith("activation_post_process"))

    @staticmethod
    def count_trainable(student: nn.Module) -> int:
        return sum(p.numel() for p in student.parameters() if p.requires_grad)
