"""Converted int8 student executor (BASELINE.json configs[4]; SURVEY.md §8 a12, §8c "config 5 oracle").

Input: the module stock ``convert()`` returns for the trained student (ref/src/training/qat_trainer.py:377-388), or its
``state_dict()`` (= best_converted.pth) together with the float student it was converted from.  Every converted
``nnq.Linear`` / ``nnq.Conv2d`` runs as ``qv_int8_linear`` (tcgen05 kind::i8, requantising epilogue) and is bit-identical
to ``torch.ops.quantized.linear`` on identical quint8 inputs (tests/test_int8_gpu.py).

The reference's converted model cannot run end to end (timm's ``cat`` / ``+ pos_embed`` receive quantized tensors,
SURVEY.md §0.9), so the glue between the quantized modules is OURS, defined in SURVEY.md §8c and mirrored on CPU with
stock ops by oracle/int8_ref.py: a quantized module whose producer is quantized (patch-embed conv <- QuantStub) consumes
its codes directly; every other quantized module quantises its fp32 input per tensor with DYNAMIC affine quint8 qparams
(min / max of the batch, Python-observer formula); cls / pos-embed / LayerNorm / attention / GELU / residuals are fp32.
"""
from __future__ import annotations

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


def _unpack_linear(mod) -> tuple:
    w, b = mod._packed_params._weight_bias() if hasattr(mod._packed_params, "_weight_bias") else mod._weight_bias()
    return w, b


class ConvertedStudent:
    """Runs the converted QATWrapper student on one GPU.  ``forward(images) -> fp32 logits [B, classes]``."""

    def __init__(self, converted: nn.Module, batch: int, device):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("qatvit_b200: the converted executor needs a CUDA device (there is no CPU fallback)")
        vit = converted.model
        if type(vit.blocks[0].norm1).__module__.startswith("torch.ao.nn.quantized"):
            raise NotImplementedError("quantized LayerNorm (observed nn.LayerNorm variant) is not supported by the int8 executor")
        self.dev = dev
        self.B = batch
        pe = vit.patch_embed
        self.D = vit.embed_dim
        self.H = vit.blocks[0].attn.num_heads
        self.ps = pe.proj.kernel_size[0]
        self.HW = pe.img_size[0]
        self.P = pe.num_patches
        self.T = self.P + 1
        self.L = len(vit.blocks)
        self.eps = float(vit.blocks[0].norm1.eps)
        self.attn_scale = float(vit.blocks[0].attn.scale)
        B, T, D, P = batch, self.T, self.D, self.P
        M = B * T
        f32 = dict(dtype=torch.float32, device=dev)
        # nnq.Quantize (from QuantStub): static input qparams
        self.in_scale = converted.quant.scale.detach().reshape(1).to(**f32)
        self.in_zp = converted.quant.zero_point.detach().reshape(1).to(torch.int32).to(dev)
        conv = pe.proj
        self.conv = _QLin(conv.weight(), conv.bias(), conv.scale, conv.zero_point, dev)
        self.blocks = []
        for blk in vit.blocks:
            d: Dict[str, object] = {}
            for name, mod in (("qkv", blk.attn.qkv), ("proj", blk.attn.proj), ("fc1", blk.mlp.fc1), ("fc2", blk.mlp.fc2)):
                w, b = _unpack_linear(mod)
                d[name] = _QLin(w, b, mod.scale, mod.zero_point, dev)
            d["n1"] = (blk.norm1.weight.detach().to(**f32), blk.norm1.bias.detach().to(**f32))
            d["n2"] = (blk.norm2.weight.detach().to(**f32), blk.norm2.bias.detach().to(**f32))
            self.blocks.append(d)
        w, b = _unpack_linear(vit.head)
        self.head = _QLin(w, b, vit.head.scale, vit.head.zero_point, dev)
        self.norm = (vit.norm.weight.detach().to(**f32), vit.norm.bias.detach().to(**f32))
        self.cls = vit.cls_token.detach().reshape(-1).to(**f32)
        self.pos = vit.pos_embed.detach().reshape(T, D).to(**f32)
        self.F = self.blocks[0]["fc1"].N
        self.C = self.head.N
        F = self.F
        e = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, device=dev)  # noqa: E731
        self.q_img = e(B * P, self.conv.K, dt=torch.uint8)
        self.p = e(B * P, D)
        self.x = [e(M, D), e(M, D)]
        self.h = e(M, D)
        self.qh = e(M, D, dt=torch.uint8)
        self.qkv = e(M, 3 * D)
        self.qkvp = e(2, M, 3 * D, dt=torch.bfloat16)
        self.o = e(M, D)
        self.y = e(M, D)
        self.f = e(M, F)
        self.g = e(M, F)
        self.qg = e(M, F, dt=torch.uint8)
        self.xn = e(B, D)
        self.qxn = e(B, D, dt=torch.uint8)
        self.logits = e(B, self.C)
        n_dyn = 4 * self.L + 1
        self.acc = torch.empty(n_dyn, 2, dtype=torch.int32, device=dev)
        self.dyn_scale = e(n_dyn)
        self.dyn_zp = torch.empty(n_dyn, dtype=torch.int32, device=dev)

    def _dyn_quant(self, slot: int, x: torch.Tensor, out: torch.Tensor, have_minmax: bool = False):
        """dynamic per-tensor affine quint8 (0..255) of an fp32 tensor: min/max -> qparams -> codes, all on device."""
        acc = self.acc[slot]
        if not have_minmax:
            ops.minmax_accumulate(x, acc)
        s, z = self.dyn_scale[slot:slot + 1], self.dyn_zp[slot:slot + 1]
        ops.qparams_from_minmax(acc, 0, 255, s, z)
        ops.quantize_u8(x, s, z, out)
        return s, z

    def _lin(self, ql: _QLin, qx, sx, zx, y):
        ops.int8_linear(qx, sx, zx, ql.qw, ql.sw, ql.wsum, ql.bias, ql.sy, ql.zy, y=y)

    @torch.no_grad()
    def forward(self, images: torch.Tensor, trace=None) -> torch.Tensor:
        """trace (optional dict): name -> clone of the uint8 input codes of each quantized module (parity diagnostics)."""
        B, T, D, P, H, F = self.B, self.T, self.D, self.P, self.H, self.F
        M = B * T
        if tuple(images.shape) != (B, 3, self.HW, self.HW) or not images.is_cuda:
            raise RuntimeError(f"converted executor built for CUDA batch {B}, got {tuple(images.shape)} on {images.device}")
        ops.minmax_reset(self.acc)
        ops.im2col_u8(images, self.in_scale, self.in_zp, B, 3, self.HW, self.ps, self.q_img)
        self._lin(self.conv, self.q_img, self.in_scale, self.in_zp, self.p)
        ops.embed_fwd(self.p, None, self.cls, self.pos, B, P, D, self.x[0])
        cur = 0
        slot = 0
        y_prev = None
        for li, blk in enumerate(self.blocks):
            g, b = blk["n1"]
            if y_prev is None:
                ops.resid_ln_fwd(self.x[cur], None, None, g, b, self.eps, M, D, h_f32=self.h, minmax=self.acc[slot])
            else:
                ops.resid_ln_fwd(self.x[cur], y_prev, None, g, b, self.eps, M, D, x_out=self.x[cur ^ 1], h_f32=self.h,
                                 minmax=self.acc[slot])
                cur ^= 1
            s, z = self._dyn_quant(slot, self.h, self.qh, have_minmax=True); slot += 1
            if trace is not None:
                trace[f"blocks.{li}.attn.qkv"] = (self.qh.clone(), s.clone(), z.clone())
            self._lin(blk["qkv"], self.qh, s, z, self.qkv)
            ops.split_planes(self.qkv, self.qkvp)
            ops.attn_fwd(self.qkvp, B, T, H, self.attn_scale, None, out_f32=self.o)
            s, z = self._dyn_quant(slot, self.o, self.qh); slot += 1
            self._lin(blk["proj"], self.qh, s, z, self.y)
            g, b = blk["n2"]
            ops.resid_ln_fwd(self.x[cur], self.y, None, g, b, self.eps, M, D, x_out=self.x[cur ^ 1], h_f32=self.h,
                             minmax=self.acc[slot])
            cur ^= 1
            s, z = self._dyn_quant(slot, self.h, self.qh, have_minmax=True); slot += 1
            self._lin(blk["fc1"], self.qh, s, z, self.f)
            ops.gelu_minmax(self.f, self.g, self.acc[slot])
            s, z = self._dyn_quant(slot, self.g, self.qg, have_minmax=True); slot += 1
            self._lin(blk["fc2"], self.qg, s, z, self.y)
            y_prev = self.y
        g, b = self.norm
        ops.resid_ln_fwd(self.x[cur], y_prev, None, g, b, self.eps, B, D, in_row_stride=T, h_f32=self.xn)
        s, z = self._dyn_quant(slot, self.xn, self.qxn)
        self._lin(self.head, self.qxn, s, z, self.logits)        # DeQuantStub: (q - z) * s, fused in the epilogue
        return self.logits

    def __call__(self, images, trace=None):
        return self.forward(images, trace)
