"""Converted int8 student executor (BASELINE.json configs[4]; SURVEY.md §8 a12, §8c "config 5 oracle").

Input: the module stock ``convert()`` returns for the trained student (ref/src/training/qat_trainer.py:377-388), or its
``state_dict()`` (= best_converted.pth, read directly by ``ConvertedStudent.from_state_dict``).  Every converted
``nnq.Linear`` / ``nnq.Conv2d`` runs as ``qv_int8_linear`` (tcgen05 kind::i8, requantising epilogue) and is bit-identical
to ``torch.ops.quantized.linear`` on identical quint8 inputs (tests/test_int8_gpu.py).

The reference's converted model cannot run end to end (timm's ``cat`` / ``+ pos_embed`` receive quantized tensors,
SURVEY.md §0.9), so the glue between the quantized modules is OURS, defined in SURVEY.md §8c and mirrored on CPU with
stock ops by oracle/int8_ref.py: a quantized module whose producer is quantized (patch-embed conv <- QuantStub) consumes
its codes directly; every other quantized module quantises its fp32 input per tensor with DYNAMIC affine quint8 qparams
(min / max of the batch, Python-observer formula); cls / pos-embed / LayerNorm / attention / GELU / residuals are fp32.

Two realisations of that glue, same arithmetic (``ConvertedStudent(..., compact=...)``):
  * compact (default): a converted Linear's output is (q - z_y) * s_y -- an integer code times one scale -- so where the consumer
    can work on codes the GEMM writes codes instead of the dequantised fp32 tensor: the qkv Linear emits the centred codes
    q - z_y as ONE exact bf16 plane (qv_int8_linear_codes; the QAT student's fused attention kernel, s_y applied inside), fc1 emits
    quint8 codes (1 byte / element), and GELU + the
    dynamic re-quantisation between fc1 and fc2 are 256-entry table lookups on the codes (bit-identical codes: every table entry
    is computed with the elementwise expressions); the qparams kernel is folded into the quantising pass, and the LayerNorm
    outputs are quantised from the saved row statistics instead of an fp32 copy (qv_ln_quantize_u8_dyn).
  * ``compact=False``: every Linear writes fp32, the glue runs elementwise on fp32 tensors (the first implementation; kept as the
    parity reference of the compact path, tests/test_int8_gpu.py).  ``compact={"gelu", "ln"}`` selects single features: those
    two reproduce the fp32 glue bit for bit, "attn" (exact integer products) does not.
"""
from __future__ import annotations

import math
import re
from typing import Dict

import torch
import torch.nn as nn

from . import ops


class _QLin:
    """Device-side operands of one converted nnq.Linear / nnq.Conv2d (weights symmetric: zero point 0)."""

    def __init__(self, qweight: torch.Tensor, bias, scale: float, zero_point: int, dev):
        if qweight.dtype != torch.qint8:
            raise TypeError("converted weights must be qint8")
        w = qweight.int_repr().reshape(qweight.shape[0], -1).contiguous()
        if qweight.qscheme() in (torch.per_channel_affine, torch.per_channel_symmetric):
            if int(qweight.q_per_channel_zero_points().abs().max()) != 0:
                raise NotImplementedError("weight zero points must be 0 (symmetric qint8)")
            sw = qweight.q_per_channel_scales().to(torch.float32)
        else:
            if int(qweight.q_zero_point()) != 0:
                raise NotImplementedError("weight zero point must be 0 (symmetric qint8)")
            sw = torch.tensor([qweight.q_scale()], dtype=torch.float32)
        self.N, self.K = w.shape
        self.qw = w.to(dev)
        self.sw = sw.contiguous().to(dev)
        self.wsum = w.to(torch.int32).sum(1).to(torch.int32).contiguous().to(dev)      # one-time weight prep
        self.bias = None if bias is None else bias.detach().to(torch.float32).contiguous().to(dev)
        self.sy, self.zy = float(scale), int(zero_point)
        self.sy_dev = torch.tensor([self.sy], dtype=torch.float32, device=dev)          # for kernels that take the scale by pointer


def _unpack_linear(mod) -> tuple:
    w, b = mod._packed_params._weight_bias() if hasattr(mod._packed_params, "_weight_bias") else mod._weight_bias()
    return w, b


def _spec_from_module(converted: nn.Module) -> dict:
    """Operands of the module stock ``convert()`` returns (ref qat_trainer.py:377-379)."""
    vit = converted.model
    if type(vit.blocks[0].norm1).__module__.startswith("torch.ao.nn.quantized"):
        raise NotImplementedError("quantized LayerNorm (observed nn.LayerNorm variant) is not supported by the int8 executor")
    pe = vit.patch_embed
    conv = pe.proj

    def lin(mod):
        w, b = _unpack_linear(mod)
        return (w, b, mod.scale, mod.zero_point)

    blocks = []
    for blk in vit.blocks:
        blocks.append({"qkv": lin(blk.attn.qkv), "proj": lin(blk.attn.proj), "fc1": lin(blk.mlp.fc1), "fc2": lin(blk.mlp.fc2),
                       "n1": (blk.norm1.weight, blk.norm1.bias), "n2": (blk.norm2.weight, blk.norm2.bias)})
    return dict(in_scale=converted.quant.scale, in_zp=converted.quant.zero_point,
                conv=(conv.weight(), conv.bias(), conv.scale, conv.zero_point), ps=conv.kernel_size[0], HW=pe.img_size[0],
                blocks=blocks, head=lin(vit.head), norm=(vit.norm.weight, vit.norm.bias), cls=vit.cls_token, pos=vit.pos_embed,
                H=vit.blocks[0].attn.num_heads, eps=float(vit.blocks[0].norm1.eps), attn_scale=float(vit.blocks[0].attn.scale))


def _spec_from_state_dict(sd: dict, num_heads=None, eps: float = 1e-6) -> dict:
    """Operands straight from ``converted.state_dict()`` -- the file the reference writes as best_converted.pth
    (ref qat_trainer.py:386-388; key layout SURVEY.md §3.4): ``quant.{scale,zero_point}``,
    ``model.patch_embed.proj.{weight (qint8 tensor),bias,scale,zero_point}``, per Linear ``<name>.{scale,zero_point}`` and
    ``<name>._packed_params._packed_params`` = ``(qint8 weight, fp32 bias)``, plus the float LayerNorm / cls / pos tensors.
    The head count and LayerNorm eps are module attributes, not tensors: timm's ViTs use 64-wide heads and eps 1e-6
    (SURVEY.md App. B); pass ``num_heads`` / ``eps`` for anything else."""
    missing = [k for k in ("quant.scale", "quant.zero_point", "model.cls_token", "model.pos_embed",
                           "model.patch_embed.proj.weight", "model.head._packed_params._packed_params") if k not in sd]
    if missing:
        raise KeyError(f"not a converted QATWrapper state_dict (best_converted.pth): missing {missing}")
    if "model.blocks.0.norm1.scale" in sd:
        raise NotImplementedError("quantized LayerNorm (observed nn.LayerNorm variant) is not supported by the int8 executor")

    def lin(prefix):
        w, b = sd[prefix + "._packed_params._packed_params"]
        return (w, b, float(sd[prefix + ".scale"]), int(sd[prefix + ".zero_point"]))

    depth = 1 + max(int(m.group(1)) for m in (re.match(r"model\.blocks\.(\d+)\.", k) for k in sd) if m)
    blocks = []
    for i in range(depth):
        p = f"model.blocks.{i}."
        blocks.append({"qkv": lin(p + "attn.qkv"), "proj": lin(p + "attn.proj"), "fc1": lin(p + "mlp.fc1"), "fc2": lin(p + "mlp.fc2"),
                       "n1": (sd[p + "norm1.weight"], sd[p + "norm1.bias"]), "n2": (sd[p + "norm2.weight"], sd[p + "norm2.bias"])})
    cw = sd["model.patch_embed.proj.weight"]
    D, ps = int(cw.shape[0]), int(cw.shape[-1])
    T = int(sd["model.pos_embed"].shape[1])
    side = int(round((T - 1) ** 0.5))
    if side * side != T - 1:
        raise ValueError(f"pos_embed holds {T} tokens: not a square patch grid + cls token")
    H = int(num_heads) if num_heads is not None else max(D // 64, 1)
    if D % H:
        raise ValueError(f"embed dim {D} is not divisible by {H} heads")
    return dict(in_scale=sd["quant.scale"], in_zp=sd["quant.zero_point"],
                conv=(cw, sd.get("model.patch_embed.proj.bias"), float(sd["model.patch_embed.proj.scale"]),
                      int(sd["model.patch_embed.proj.zero_point"])), ps=ps, HW=side * ps,
                blocks=blocks, head=lin("model.head"), norm=(sd["model.norm.weight"], sd["model.norm.bias"]),
                cls=sd["model.cls_token"], pos=sd["model.pos_embed"], H=H, eps=float(eps), attn_scale=float(D // H) ** -0.5)


class ConvertedStudent:
    """Runs the converted QATWrapper student on one GPU.  ``forward(images) -> fp32 logits [B, classes]``.

    ``ConvertedStudent(converted_module, batch, device)`` takes the module stock ``convert()`` returns;
    ``ConvertedStudent.from_state_dict(path_or_dict, batch, device)`` reads best_converted.pth directly (no module tree)."""

    @classmethod
    def from_state_dict(cls, state_dict, batch: int, device, num_heads=None, eps: float = 1e-6, compact: bool = True):
        if not isinstance(state_dict, dict):
            state_dict = torch.load(state_dict, map_location="cpu", weights_only=False)   # packed params are (qtensor, bias) tuples
        return cls(_spec_from_state_dict(state_dict, num_heads, eps), batch, device, compact=compact)

    def __init__(self, converted, batch: int, device, compact: bool = True):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("qatvit_b200: the converted executor needs a CUDA device (there is no CPU fallback)")
        spec = converted if isinstance(converted, dict) else _spec_from_module(converted)
        self.dev = dev
        self.B = batch
        self.D = int(spec["pos"].shape[-1])
        self.H = int(spec["H"])
        self.ps = int(spec["ps"])
        self.HW = int(spec["HW"])
        self.P = (self.HW // self.ps) ** 2
        self.T = self.P + 1
        self.L = len(spec["blocks"])
        self.eps = float(spec["eps"])
        self.attn_scale = float(spec["attn_scale"])
        B, T, D, P = batch, self.T, self.D, self.P
        M = B * T
        f32 = dict(dtype=torch.float32, device=dev)
        # nnq.Quantize (from QuantStub): static input qparams
        self.in_scale = spec["in_scale"].detach().reshape(1).to(**f32)
        self.in_zp = spec["in_zp"].detach().reshape(1).to(torch.int32).to(dev)
        self.conv = _QLin(*spec["conv"], dev)
        self.blocks = []
        for blk in spec["blocks"]:
            d: Dict[str, object] = {}
            for name in ("qkv", "proj", "fc1", "fc2"):
                d[name] = _QLin(*blk[name], dev)
            d["n1"] = tuple(t.detach().to(**f32) for t in blk["n1"])
            d["n2"] = tuple(t.detach().to(**f32) for t in blk["n2"])
            self.blocks.append(d)
        self.head = _QLin(*spec["head"], dev)
        self.norm = tuple(t.detach().to(**f32) for t in spec["norm"])
        self.cls = spec["cls"].detach().reshape(-1).to(**f32)
        self.pos = spec["pos"].detach().reshape(T, D).to(**f32)
        self.F = self.blocks[0]["fc1"].N
        self.C = self.head.N
        F = self.F
        e = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, device=dev)  # noqa: E731
        self.q_img = e(B * P, self.conv.K, dt=torch.uint8)
        self.p = e(B * P, D)
        self.x = [e(M, D), e(M, D)]
        self.h = e(M, D)
        self.qh = e(M, D, dt=torch.uint8)
        # compact: True (every feature) / False (none), one feature name, or a collection of names out of
        #   "attn": qkv leaves its Linear as one bf16 code plane for the integer-code attention kernel (exact products: not
        #           bit-identical to the fp32 glue, whose hi/lo products are ~2^-16 relative)
        #   "gelu": GELU + re-quantisation between fc1 and fc2 as table lookups, qparams folded into the quantising passes
        #   "ln"  : the LayerNorm outputs are quantised from the saved row statistics instead of an fp32 copy
        # "gelu" and "ln" reproduce the fp32 glue bit for bit (the parity reference of the compact path).
        feats = {True: {"attn", "gelu", "ln"}, False: set()}.get(compact) if isinstance(compact, bool) else \
            ({compact} if isinstance(compact, str) else set(compact))
        if not feats <= {"attn", "gelu", "ln"}:
            raise ValueError("compact must be True, False, or a selection of 'attn', 'gelu', 'ln'")
        self.c_attn = "attn" in feats and D % 8 == 0
        self.c_gelu = "gelu" in feats and F % 16 == 0           # the table kernels move 16 codes per thread
        self.c_ln = "ln" in feats and D % 128 == 0
        self.compact = self.c_attn and self.c_gelu and self.c_ln
        if self.c_ln:
            self.ln_mean = e(M)
            self.ln_rstd = e(M)
        if self.c_attn:
            self.qkv_codes = e(1, M, 3 * D, dt=torch.bfloat16)
        else:
            self.qkv = e(M, 3 * D)
            self.qkvp = e(2, M, 3 * D, dt=torch.bfloat16)
        if self.c_gelu:
            self.q_f = e(M, F, dt=torch.uint8)
        else:
            self.f = e(M, F)
            self.g = e(M, F)
        self.o = e(M, D)
        self.y = e(M, D)
        self.qg = e(M, F, dt=torch.uint8)
        self.xn = e(B, D)
        self.qxn = e(B, D, dt=torch.uint8)
        self.logits = e(B, self.C)
        # Ragged last batch (the reference's eval DataLoader has no drop_last: 10 000 % 256 = 16 images, ref qat_trainer.py:237-254):
        # every activation buffer is re-viewed as the CONTIGUOUS tensor an executor built for b images would own (same storage, first
        # numel(b) elements); kernels take their row counts at launch, so b images on an executor built for B give bit for bit what
        # an executor built for b gives.  Views are cached per b.
        self._full = {k: v for k, v in vars(self).items()
                      if torch.is_tensor(v) and k in ("q_img", "p", "h", "qh", "qkv_codes", "qkv", "qkvp", "q_f", "f", "g", "o", "y", "qg",
                                                      "xn", "qxn", "logits", "ln_mean", "ln_rstd")}
        self._full["x0"], self._full["x1"] = self.x
        self._views = {}
        self._cur = batch
        n_dyn = 4 * self.L + 1
        self.acc = torch.empty(n_dyn, 2, dtype=torch.int32, device=dev)
        self.dyn_scale = e(n_dyn)
        self.dyn_zp = torch.empty(n_dyn, dtype=torch.int32, device=dev)

    def _bind(self, b: int) -> None:
        """Point every activation buffer attribute at its batch-b view."""
        if b == self._cur:
            return
        views = self._views.get(b)
        if views is None:
            views = {}
            for name, full in self._full.items():
                shape = list(full.shape)
                ax = 1 if full.dim() == 3 else 0                 # [planes, rows, cols] or [rows, cols]: rows scale with the batch
                shape[ax] = shape[ax] * b // self.B
                views[name] = full.view(-1)[:math.prod(shape)].view(shape)
            self._views[b] = views
        for name, v in views.items():
            if name not in ("x0", "x1"):
                setattr(self, name, v)
        self.x = [views["x0"], views["x1"]]
        self._cur = b

    def _dyn_quant(self, slot: int, x: torch.Tensor, out: torch.Tensor, have_minmax: bool = False):
        """dynamic per-tensor affine quint8 (0..255) of an fp32 tensor: min/max -> qparams -> codes, all on device."""
        acc = self.acc[slot]
        if not have_minmax:
            ops.minmax_accumulate(x, acc)
        s, z = self.dyn_scale[slot:slot + 1], self.dyn_zp[slot:slot + 1]
        if self.c_gelu:
            ops.quantize_u8_dyn(x, acc, s, z, out)             # qparams from the finished accumulator + codes, one launch
        else:
            ops.qparams_from_minmax(acc, 0, 255, s, z)
            ops.quantize_u8(x, s, z, out)
        return s, z

    def _ln_quant(self, slot: int, cur: int, y_prev, gamma, beta, M: int):
        """x <- x + y_prev (when given); qh <- dynamically quantised LayerNorm(x).  Returns (index of the current residual
        buffer, scale, zero point)."""
        D = self.D
        fused = self.c_ln
        out = dict(mean=self.ln_mean, rstd=self.ln_rstd) if fused else dict(h_f32=self.h)
        if y_prev is None:
            ops.resid_ln_fwd(self.x[cur], None, None, gamma, beta, self.eps, M, D, minmax=self.acc[slot], **out)
        else:
            ops.resid_ln_fwd(self.x[cur], y_prev, None, gamma, beta, self.eps, M, D, x_out=self.x[cur ^ 1], minmax=self.acc[slot], **out)
            cur ^= 1
        if fused:      # the LayerNorm output is recomputed from (x, mean, rstd) inside the quantising pass: no fp32 copy of it
            s, z = self.dyn_scale[slot:slot + 1], self.dyn_zp[slot:slot + 1]
            ops.ln_quantize_u8_dyn(self.x[cur], self.ln_mean, self.ln_rstd, gamma, beta, M, D, self.acc[slot], s, z, self.qh)
        else:
            s, z = self._dyn_quant(slot, self.h, self.qh, have_minmax=True)
        return cur, s, z

    def _lin(self, ql: _QLin, qx, sx, zx, y=None, qy=None):
        ops.int8_linear(qx, sx, zx, ql.qw, ql.sw, ql.wsum, ql.bias, ql.sy, ql.zy, y=y, qy=qy)

    @torch.no_grad()
    def forward(self, images: torch.Tensor, trace=None) -> torch.Tensor:
        """trace (optional dict): name -> clone of the uint8 input codes of each quantized module (parity diagnostics)."""
        T, D, P, H, F = self.T, self.D, self.P, self.H, self.F
        B = int(images.shape[0]) if images.dim() == 4 else -1
        if images.dim() != 4 or tuple(images.shape[1:]) != (3, self.HW, self.HW) or not (1 <= B <= self.B) or not images.is_cuda:
            raise RuntimeError(f"converted executor built for CUDA batches of 1..{self.B} 3x{self.HW}x{self.HW} images, got "
                               f"{tuple(images.shape)} on {images.device}")
        self._bind(B)
        M = B * T
        ops.minmax_reset(self.acc)
        ops.im2col_u8(images, self.in_scale, self.in_zp, B, 3, self.HW, self.ps, self.q_img)
        self._lin(self.conv, self.q_img, self.in_scale, self.in_zp, self.p)
        ops.embed_fwd(self.p, None, self.cls, self.pos, B, P, D, self.x[0])
        cur = 0
        slot = 0
        y_prev = None
        for li, blk in enumerate(self.blocks):
            g, b = blk["n1"]
            cur, s, z = self._ln_quant(slot, cur, y_prev, g, b, M); slot += 1
            if trace is not None:
                trace[f"blocks.{li}.attn.qkv"] = (self.qh.clone(), s.clone(), z.clone())
            if self.c_attn:
                # q, k, v = (codes - z_y) * s_y: the centred codes are one exact bf16 plane, s_y is applied by the attention kernel
                ql = blk["qkv"]
                ops.int8_linear_codes(self.qh, s, z, ql.qw, ql.sw, ql.wsum, ql.bias, ql.sy, ql.zy, self.qkv_codes)
                ops.attn_fwd(self.qkv_codes, B, T, H, self.attn_scale, None, qk_scale=ql.sy_dev, v_scale=ql.sy_dev, out_f32=self.o)
            else:
                self._lin(blk["qkv"], self.qh, s, z, self.qkv)
                ops.split_planes(self.qkv, self.qkvp)
                ops.attn_fwd(self.qkvp, B, T, H, self.attn_scale, None, out_f32=self.o)
            s, z = self._dyn_quant(slot, self.o, self.qh); slot += 1
            self._lin(blk["proj"], self.qh, s, z, self.y)
            g, b = blk["n2"]
            cur, s, z = self._ln_quant(slot, cur, self.y, g, b, M); slot += 1
            if self.c_gelu:
                # GELU((q - z_y) s_y) takes <= 256 values: min / max and the re-quantised codes are table lookups on fc1's codes
                ql = blk["fc1"]
                self._lin(ql, self.qh, s, z, qy=self.q_f)
                ops.gelu_u8_minmax(self.q_f, ql.sy, ql.zy, self.acc[slot])
                s, z = self.dyn_scale[slot:slot + 1], self.dyn_zp[slot:slot + 1]
                ops.gelu_u8_requant(self.q_f, ql.sy, ql.zy, self.acc[slot], s, z, self.qg)
                slot += 1
            else:
                self._lin(blk["fc1"], self.qh, s, z, self.f)
                ops.gelu_minmax(self.f, self.g, self.acc[slot])
                s, z = self._dyn_quant(slot, self.g, self.qg, have_minmax=True); slot += 1
            if trace is not None:
                trace[f"blocks.{li}.mlp.fc2"] = (self.qg.clone(), s.clone(), z.clone())
            self._lin(blk["fc2"], self.qg, s, z, self.y)
            y_prev = self.y
        g, b = self.norm
        ops.resid_ln_fwd(self.x[cur], y_prev, None, g, b, self.eps, B, D, in_row_stride=T, h_f32=self.xn)
        s, z = self._dyn_quant(slot, self.xn, self.qxn)
        self._lin(self.head, self.qxn, s, z, self.logits)        # DeQuantStub: (q - z) * s, fused in the epilogue
        return self.logits

    def __call__(self, images, trace=None):
        return self.forward(images, trace)
