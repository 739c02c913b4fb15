"""Pre-QAT epochs on the same kernels (ref/src/training/qat_trainer.py:286,320,333-361 with
``qat_enabled == False``): before ``qat_start_epoch`` the reference trains the UNPREPARED ``QATWrapper`` student (identity
Quant/DeQuant stubs, plain nn.Linear / nn.Conv2d, no fake-quant) against the same frozen teacher with the same KL + CE loss.

``PlainDistillStep`` is that step on the tcgen05 GEMM family: every fp32 tensor that feeds a GEMM travels as bf16 hi/lo planes,
three MMAs per product, fp32 accumulation (fp32-grade results, ~2^-16 relative) -- the reference's non-AMP arithmetic.  Its
optional ``--amp`` variant (fp16 autocast + GradScaler around the student, ref :286,340,353-357) has a half-precision counterpart
here: ``PlainDistillStep(..., amp=True)`` runs ONE bf16 pass per product (a third of the tensor work; the teacher stays fp32-grade
as in the reference, where it runs outside the autocast region).  Weights are re-split into planes
every step (they change every step); dgrad reads the same [N, K] planes as an MN-major operand, so no transposed copy exists.
Attention uses the unfused kernels (scores / probabilities as planes in HBM, saved for backward): the fused integer-code
kernels need fake-quantised q, k, v.  Gradients land in one flat arena exactly like the QAT engine's, so the optimizer and the
gradient all-reduce are shared (qatvit_b200.optim.FusedClipAdamW, qatvit_b200.ddp.GradSync).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from .engine import TeacherEngine, _BatchBuffers, _ViTDims, _attention_forward, wgrad_splits
from .ops import Op, Out, PAIRS_FP32, PAIRS_SINGLE


class _Lin:
    """One plain Linear / Conv2d-as-GEMM: parameters + the per-step hi/lo planes of its weight."""

    def __init__(self, mod: nn.Module, dev):
        self.mod = mod
        self.weight, self.bias = mod.weight, mod.bias
        self.N = mod.weight.shape[0]
        self.K = mod.weight.numel() // self.N
        self.planes = torch.empty(2, self.N, self.K, dtype=torch.bfloat16, device=dev)

    def split(self) -> None:
        ops.split_planes(self.weight.detach().reshape(self.N, self.K), self.planes)


class PlainStudentEngine(_BatchBuffers):
    """Forward + hand-written backward of the unprepared ``QATWrapper`` student (no fake-quant anywhere)."""

    def __init__(self, student: nn.Module, batch: int, hparams: Dict, grad_buffer: Optional[torch.Tensor] = None,
                 amp: bool = False):
        # amp: the half-precision variant of the pre-QAT step (the reference's optional --amp: fp16 autocast + GradScaler around
        # the student forward / loss, ref qat_trainer.py:286,340,353-357).  Here: ONE tensor-core pass per product on the bf16
        # hi planes of both operands (the lo planes are simply not read), fp32 accumulation and fp32 outputs, LayerNorm /
        # softmax / loss in fp32 as under autocast.  bf16 keeps the fp32 exponent range, so no loss scaling is needed; operand
        # rounding is 2^-9 (fp16 autocast: 2^-11) -- reduced-precision by design, checked against the fp32 reference to 3e-2.
        self.amp = bool(amp)
        self.pairs = PAIRS_SINGLE if self.amp else PAIRS_FP32
        for m in student.modules():
            if type(m).__name__ == "FusedMovingAvgObsFakeQuantize":
                raise RuntimeError("qatvit_b200: PlainStudentEngine takes the student BEFORE prepare_qat (use QATDistillStep after)")
        self.student = student
        vit = self.vit = student.model
        dev = next(student.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("qatvit_b200: the student must live on a CUDA device (there is no CPU fallback)")
        self.dev = dev
        self.hp_ = dict(hparams)
        d = self.d = _ViTDims(vit, batch)
        self.sms = torch.cuda.get_device_properties(dev).multi_processor_count
        L, M, D, F, B, T = d.L, d.M, d.D, d.F, d.B, d.T
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda *s, dt=f32: torch.empty(*s, dtype=dt, device=dev)  # noqa: E731
        self.conv = _Lin(vit.patch_embed.proj, dev)
        self.lin: List[Dict[str, _Lin]] = [dict(qkv=_Lin(b.attn.qkv, dev), proj=_Lin(b.attn.proj, dev), fc1=_Lin(b.mlp.fc1, dev),
                                                fc2=_Lin(b.mlp.fc2, dev)) for b in vit.blocks]
        self.all_linears = [self.conv] + [q for blk in self.lin for q in blk.values()]
        # ---- flat gradient arena ----
        self.params = [p for p in student.parameters() if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        if grad_buffer is not None:
            if grad_buffer.numel() != total or grad_buffer.dtype != f32 or grad_buffer.device != dev:
                raise ValueError("grad_buffer must be a flat fp32 tensor with one element per trainable parameter")
            self.grad_arena = grad_buffer
        else:
            self.grad_arena = torch.zeros(total, dtype=f32, device=dev)
        self._goff, self._grad_views, off = {}, [], 0
        for p in self.params:
            self._grad_views.append(self.grad_arena[off:off + p.numel()].view_as(p))
            self._goff[id(p)] = off
            off += p.numel()
        self.attach_grads()
        # ---- forward activations (saved for backward) and backward scratch (re-viewed for a smaller batch: engine._BatchBuffers) ----
        self._bb_init(dev, d)
        add = self._bb_add
        add("img_planes", lambda d: (2, d.B * d.P, d.Kc), bf)
        add("p_raw", lambda d: (d.B * d.P, d.D))
        add("x_in", lambda d: (d.M, d.D), count=L)
        add("x_mid", lambda d: (d.M, d.D), count=L)
        add("h1p", lambda d: (2, d.M, d.D), bf, count=L)
        add("h2p", lambda d: (2, d.M, d.D), bf, count=L)
        add("qkvp", lambda d: (2, d.M, 3 * d.D), bf, count=L)
        add("Pp", lambda d: (2, d.B * d.H * d.T, d.ldP), bf, count=L, zero=True)
        add("op", lambda d: (2, d.M, d.D), bf, count=L)
        add("a_raw", lambda d: (d.M, d.D))
        add("f_raw", lambda d: (d.M, d.F), count=L)
        add("gelp", lambda d: (2, d.M, d.F), bf, count=L)
        add("m_raw", lambda d: (d.M, d.D))
        for nm in ("st1m", "st1r", "st2m", "st2r"):
            add(nm, lambda d: (d.M,), count=L)
        add("S", lambda d: (d.B * d.H * d.T, d.ldS))
        add("o", lambda d: (d.M, d.D))
        add("xcls", lambda d: (d.B, d.D))
        add("xn", lambda d: (d.B, d.D))
        add("stFm", lambda d: (d.B,))
        add("stFr", lambda d: (d.B,))
        add("logits", lambda d: (d.B, d.C))
        add("g_logits", lambda d: (d.B, d.C))
        self.loss3 = e(3)
        # ---- backward scratch ----
        add("gx", lambda d: (d.M, d.D), count=2)
        add("g_xn", lambda d: (d.B, d.D))
        add("gpD", lambda d: (2, d.M, d.D), bf)
        add("gpF", lambda d: (2, d.M, d.F), bf)
        add("gp3", lambda d: (2, d.M, 3 * d.D), bf)
        add("gpP", lambda d: (2, d.B * d.P, d.D), bf)
        add("g_big", lambda d: (d.M, d.F))
        add("g_h", lambda d: (d.M, d.D))
        add("g_o", lambda d: (d.M, d.D))
        add("g_op", lambda d: (2, d.M, d.D), bf)
        add("g_qkv", lambda d: (d.M, 3 * d.D))
        add("dP", lambda d: (d.B * d.H * d.T, d.ldS))
        add("dSp", lambda d: (2, d.B * d.H * d.T, d.ldP), bf, zero=True)
        self.rpb = 64
        self.bias_part = e(-(-M // self.rpb) * max(F, 3 * D))
        self.ln_part = e(-(-M // self.rpb), 2, D)
        self._ln_tmp = e(2 * D)
        self._splits_by_b, max_ws = {}, 0
        for bb in range(1, B + 1):
            tab = {}
            for (n, k, kdim) in [(3 * D, D, bb * T), (D, D, bb * T), (F, D, bb * T), (D, F, bb * T), (D, d.Kc, bb * d.P)]:
                sp = wgrad_splits(n, k, kdim, self.sms)
                tab[(n, k)] = sp
                max_ws = max(max_ws, sp * n * k)
            self._splits_by_b[bb] = tab
        self.ws = e(max_ws)
        self._bb_on_bind(B)

    def _bb_on_bind(self, b: int) -> None:
        self._splits = self._splits_by_b[b]
        self.stats1 = list(zip(self.st1m, self.st1r))
        self.stats2 = list(zip(self.st2m, self.st2r))
        self.statsF = (self.stFm, self.stFr)

    # ------------------------------------------------------------------------------------------
    def attach_grads(self) -> None:
        for p, g in zip(self.params, self._grad_views):
            p.grad = g

    def _grad(self, p: torch.Tensor) -> torch.Tensor:
        off = self._goff[id(p)]
        return self.grad_arena[off:off + p.numel()]

    def _ln_param_grads(self, norm: nn.Module, nblk: int) -> None:
        D = self.d.D
        gw, gb = self._goff[id(norm.weight)], self._goff[id(norm.bias)]
        if gb == gw + D:        # weight and bias gradients are neighbours in the arena: one reduce writes both
            ops.colsum_reduce(self.ln_part, nblk, 2 * D, self.grad_arena[gw:gw + 2 * D])
            return
        tmp = self._ln_tmp
        ops.colsum_reduce(self.ln_part, nblk, 2 * D, tmp)
        self._grad(norm.weight).copy_(tmp[:D])
        self._grad(norm.bias).copy_(tmp[D:])

    def _fwd(self, ql: _Lin, a_planes, M, out=None, out_planes=None) -> None:
        ops.gemm(Op.full(a_planes), Op.full(ql.planes), M, ql.N, ql.K, self.pairs, out=out, out_planes=out_planes,
                 bias=ql.bias.detach())

    def forward(self, images: torch.Tensor, labels: Optional[torch.Tensor], teacher_logits: Optional[torch.Tensor],
                teacher_ready=None) -> Optional[torch.Tensor]:
        self._bb_bind(self._bb_check(images, "student"))
        d, v = self.d, self.vit
        B, T, D, F, M, L = d.B, d.T, d.D, d.F, d.M, d.L
        for ql in self.all_linears:
            ql.split()
        ops.im2col_fq(images, None, B, d.in_ch, d.HW, d.ps, self.img_planes)
        self._fwd(self.conv, self.img_planes, B * d.P, out=self.p_raw)
        ops.embed_fwd(self.p_raw, None, v.cls_token.detach().reshape(-1), v.pos_embed.detach().reshape(T, D), B, d.P, D, self.x_in[0])
        for l, blk in enumerate(v.blocks):
            ql = self.lin[l]
            if l == 0:
                ops.resid_ln_fwd(self.x_in[0], None, None, blk.norm1.weight.detach(), blk.norm1.bias.detach(), d.eps, M, D,
                                 h_planes=self.h1p[0], mean=self.stats1[0][0], rstd=self.stats1[0][1])
            self._fwd(ql["qkv"], self.h1p[l], M, out_planes=self.qkvp[l])
            _attention_forward(d, self.qkvp[l], self.S, self.Pp[l], self.o, pairs=self.pairs)
            ops.split_planes(self.o, self.op[l])
            self._fwd(ql["proj"], self.op[l], M, out=self.a_raw)
            ops.resid_ln_fwd(self.x_in[l], self.a_raw, None, blk.norm2.weight.detach(), blk.norm2.bias.detach(), d.eps, M, D,
                             x_out=self.x_mid[l], h_planes=self.h2p[l], mean=self.stats2[l][0], rstd=self.stats2[l][1])
            self._fwd(ql["fc1"], self.h2p[l], M, out=self.f_raw[l])
            ops.act_planes(self.f_raw[l], None, True, self.gelp[l])
            self._fwd(ql["fc2"], self.gelp[l], M, out=self.m_raw)
            if l + 1 < L:
                nb = v.blocks[l + 1]
                ops.resid_ln_fwd(self.x_mid[l], self.m_raw, None, nb.norm1.weight.detach(), nb.norm1.bias.detach(), d.eps, M, D,
                                 x_out=self.x_in[l + 1], h_planes=self.h1p[l + 1], mean=self.stats1[l + 1][0],
                                 rstd=self.stats1[l + 1][1])
            else:
                ops.resid_ln_fwd(self.x_mid[l], self.m_raw, None, v.norm.weight.detach(), v.norm.bias.detach(), d.eps, B, D,
                                 in_row_stride=T, x_out=self.xcls, h_f32=self.xn, mean=self.statsF[0], rstd=self.statsF[1])
        ops.head_fwd(self.xn, v.head.weight.detach(), v.head.bias.detach(), B, D, d.C, self.logits)
        if labels is None:
            return None
        hp = self.hp_
        if teacher_ready is not None:
            torch.cuda.current_stream().wait_event(teacher_ready)
        ops.kd_ce_loss(self.logits, teacher_logits, labels, hp["kd_temp"], hp["kd_alpha"], hp["label_smoothing"], out3=self.loss3,
                       grad=self.g_logits)
        return self.loss3

    def predict(self, images: torch.Tensor) -> torch.Tensor:
        self.forward(images, None, None)
        return self.logits

    # ------------------------------------------------------------------------------------------
    def _gp(self, g, y_raw, ql: _Lin, gelu: bool, R: int, out_planes, remap=(0, 0)) -> None:
        """gradient planes of a plain Linear: g [* gelu'(y_raw)] -> hi/lo planes; bias grad = column sums."""
        nblk = -(-R // self.rpb)
        part = self.bias_part[:nblk * ql.N].view(nblk, ql.N)
        ops.gp_planes(g, y_raw, None, None, False, gelu, R, ql.N, out_planes, part, self.rpb, remap[0], remap[1])
        ops.colsum_reduce(part, nblk, ql.N, self._grad(ql.bias))

    def _dgrad(self, ql: _Lin, gp, M: int, out) -> None:
        # g[M, N] @ W[N, K]: the weight planes as they lie ([contraction N][K]) are the MN-major B operand
        ops.gemm(Op.full(gp), Op.full(ql.planes, mn_major=True), M, ql.K, ql.N, self.pairs, out=out)

    def _wgrad(self, ql: _Lin, gp, x_planes, kdim: int) -> None:
        s = self._splits[(ql.N, ql.K)]
        if s > 1:
            ops.gemm(Op.full(gp, mn_major=True), Op.full(x_planes, mn_major=True), ql.N, ql.K, kdim, self.pairs, splits=s,
                     workspace=self.ws)
        else:
            ops.gemm(Op.full(gp, mn_major=True), Op.full(x_planes, mn_major=True), ql.N, ql.K, kdim, self.pairs,
                     out=self.ws[:ql.N * ql.K].view(ql.N, ql.K))
        ops.splitk_reduce(self.ws, s, ql.N, ql.K, self._grad(ql.weight))

    def backward(self, grads_final_from=None) -> None:
        d, v = self.d, self.vit
        B, T, D, F, M, L, H = d.B, d.T, d.D, d.F, d.M, d.L, d.H
        BH = B * H
        ops.head_bwd(self.g_logits, self.xn, v.head.weight.detach(), None, B, D, d.C, self.g_xn, self._grad(v.head.weight),
                     self._grad(v.head.bias))
        gx, gx2 = self.gx
        ops.zero_(gx)
        ops.ln_bwd(self.g_xn, self.xcls, self.statsF[0], self.statsF[1], v.norm.weight.detach(), None, B, D, gx, self.ln_part,
                   self.rpb, out_row_stride=T)
        self._ln_param_grads(v.norm, -(-B // self.rpb))
        nblk_ln = -(-M // self.rpb)
        for l in range(L - 1, -1, -1):
            blk, ql = v.blocks[l], self.lin[l]
            # ---- MLP ----
            self._gp(gx, None, ql["fc2"], False, M, self.gpD)
            self._dgrad(ql["fc2"], self.gpD, M, self.g_big)
            self._wgrad(ql["fc2"], self.gpD, self.gelp[l], M)
            self._gp(self.g_big, self.f_raw[l], ql["fc1"], True, M, self.gpF)
            self._dgrad(ql["fc1"], self.gpF, M, self.g_h)
            self._wgrad(ql["fc1"], self.gpF, self.h2p[l], M)
            ops.ln_bwd(self.g_h, self.x_mid[l], self.stats2[l][0], self.stats2[l][1], blk.norm2.weight.detach(), gx, M, D, gx2,
                       self.ln_part, self.rpb)
            self._ln_param_grads(blk.norm2, nblk_ln)
            # ---- attention (unfused: probabilities saved as planes) ----
            self._gp(gx2, None, ql["proj"], False, M, self.gpD)
            self._dgrad(ql["proj"], self.gpD, M, self.g_o)
            self._wgrad(ql["proj"], self.gpD, self.op[l], M)
            ops.split_planes(self.g_o, self.g_op)
            qkvp, Pp = self.qkvp[l], self.Pp[l]
            ops.gemm(Op.tokens(self.g_op, B, T, 0, 64), Op.tokens(qkvp, B, T, 2 * D, 64), T, T, 64, self.pairs,
                     out=Out.per_head(self.dP, BH, H, T, T), nbatch=BH, batch_inner=H)                      # dP = dO V^T
            ops.attn_ds(Pp, self.dP, d.ldS, BH * T, T, d.attn_scale, self.dSp)
            ops.gemm(Op.per_head(self.dSp, BH, H, T, T), Op.tokens(qkvp, B, T, D, 64, mn_major=True), T, 64, T, self.pairs,
                     out=Out.tokens(self.g_qkv, B, T, 0, 64), nbatch=BH, batch_inner=H)                      # dQ = dS K
            ops.gemm(Op.per_head(self.dSp, BH, H, T, T, mn_major=True), Op.tokens(qkvp, B, T, 0, 64, mn_major=True), T, 64, T,
                     self.pairs, out=Out.tokens(self.g_qkv, B, T, D, 64), nbatch=BH, batch_inner=H)          # dK = dS^T Q
            ops.gemm(Op.per_head(Pp, BH, H, T, T, mn_major=True), Op.tokens(self.g_op, B, T, 0, 64, mn_major=True), T, 64, T,
                     self.pairs, out=Out.tokens(self.g_qkv, B, T, 2 * D, 64), nbatch=BH, batch_inner=H)      # dV = P^T dO
            self._gp(self.g_qkv, None, ql["qkv"], False, M, self.gp3)
            self._dgrad(ql["qkv"], self.gp3, M, self.g_h)
            self._wgrad(ql["qkv"], self.gp3, self.h1p[l], M)
            ops.ln_bwd(self.g_h, self.x_in[l], self.stats1[l][0], self.stats1[l][1], blk.norm1.weight.detach(), gx2, M, D, gx,
                       self.ln_part, self.rpb)
            self._ln_param_grads(blk.norm1, nblk_ln)
            if grads_final_from is not None:
                grads_final_from(min(self._goff[id(p)] for p in blk.parameters()))
        # ---- embeddings ----
        ops.colsum_rows(gx, B, T * D, T * D, self._grad(v.pos_embed))
        ops.colsum_rows(gx, B, D, T * D, self._grad(v.cls_token))
        self._gp(gx, None, self.conv, False, B * d.P, self.gpP, remap=(d.P, T))
        self._wgrad(self.conv, self.gpP, self.img_planes, B * d.P)
        if grads_final_from is not None:
            grads_final_from(0)


class PlainDistillStep:
    """``loss3 = step(images, labels)``: teacher forward (side stream), plain student forward, KL + CE, backward -- the
    reference's training iteration before QAT is enabled (ref qat_trainer.py:333-361 with qat_enabled == False, no AMP)."""

    def __init__(self, student: nn.Module, teacher: nn.Module, batch: int, hparams: Dict,
                 grad_buffer: Optional[torch.Tensor] = None, teacher_mixed: Optional[bool] = None, amp: bool = False):
        self.student_engine = PlainStudentEngine(student, batch, hparams, grad_buffer=grad_buffer, amp=amp)
        self.teacher_engine = TeacherEngine(teacher, batch, mixed=teacher_mixed)
        self.grad_arena = self.student_engine.grad_arena
        self._tstream = torch.cuda.Stream(device=self.student_engine.dev)
        self._tdone = torch.cuda.Event()

    def __call__(self, images: torch.Tensor, labels: torch.Tensor, grad_sync=None) -> torch.Tensor:
        main = torch.cuda.current_stream()
        self.student_engine.attach_grads()       # optimizer.zero_grad(set_to_none=True) drops the arena views (ref :351)
        if ops.profiling():
            t_logits = self.teacher_engine.forward(images)
            out3 = self.student_engine.forward(images, labels, t_logits)
        else:
            self._tstream.wait_stream(main)
            with torch.cuda.stream(self._tstream):
                t_logits = self.teacher_engine.forward(images)
                self._tdone.record(self._tstream)
            out3 = self.student_engine.forward(images, labels, t_logits, teacher_ready=self._tdone)
        if grad_sync is None:
            self.student_engine.backward()
        else:
            grad_sync.begin_step()
            self.student_engine.backward(grads_final_from=grad_sync.grads_final_from)
            grad_sync.end_step()
        return out3

    def predict(self, images: torch.Tensor) -> torch.Tensor:
        return self.student_engine.predict(images)
