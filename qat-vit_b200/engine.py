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
        # per-tensor weights (qnnpack qconfig): the kernel also writes the scale broadcast over the output channels
        sv = None if f.per_channel else self.wscale_vec
        if self.small:
            ops.fq_weight(w, f.per_channel, f.observer_enabled, f.fake_quant_enabled, f.min_val, f.max_val, f.scale,
                          f.zero_point, f.c, f.qmin, f.qmax, f.symmetric, y=self.wq, mask=self.wmask, scratch=self.scratch, scale_vec=sv)
        else:
            ops.fq_weight(w, f.per_channel, f.observer_enabled, f.fake_quant_enabled, f.min_val, f.max_val, f.scale,
                          f.zero_point, f.c, f.qmin, f.qmax, f.symmetric, mask=self.wmask, codes=self.codes[0],
                          codes_t=self.codes_t[0], scratch=self.scratch, scale_vec=sv)


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


def _is_identity(m) -> bool:
    return m is None or isinstance(m, nn.Identity) or (isinstance(m, nn.Dropout) and m.p == 0.0)


def _check_supported_vit(vit: nn.Module) -> None:
    """The engines execute ONE function: the timm VisionTransformer of SURVEY.md App. B (what ``timm.create_model`` returns for
    vit_small/base_patch16_224 with default arguments, ref model_registry.py:167-172,228-233).  ``create_student`` /
    ``create_teacher`` forward **kwargs to timm, so other variants are reachable: refuse them instead of silently computing a
    different function than the module tree."""
    bad = []
    for name in ("norm_pre", "patch_drop", "fc_norm", "pos_drop", "head_drop"):
        if hasattr(vit, name) and not _is_identity(getattr(vit, name)):
            bad.append(name)
    if getattr(vit, "global_pool", "token") != "token":
        bad.append(f"global_pool={vit.global_pool!r}")
    if getattr(vit, "num_prefix_tokens", 1) != 1 or getattr(vit, "reg_token", None) is not None:
        bad.append("prefix / register tokens")
    if getattr(vit, "no_embed_class", False):
        bad.append("no_embed_class")
    if hasattr(vit.patch_embed, "norm") and not _is_identity(vit.patch_embed.norm):
        bad.append("patch_embed.norm")
    for i, blk in enumerate(vit.blocks):
        for name in ("ls1", "ls2", "drop_path1", "drop_path2"):
            if hasattr(blk, name) and not _is_identity(getattr(blk, name)):
                bad.append(f"blocks.{i}.{name}")
        for name in ("q_norm", "k_norm", "norm", "attn_drop", "proj_drop"):
            if hasattr(blk.attn, name) and not _is_identity(getattr(blk.attn, name)):
                bad.append(f"blocks.{i}.attn.{name}")
        for name in ("drop1", "drop2", "norm"):
            if hasattr(blk.mlp, name) and not _is_identity(getattr(blk.mlp, name)):
                bad.append(f"blocks.{i}.mlp.{name}")
        act = getattr(blk.mlp, "act", None)
        if act is not None and not (isinstance(act, nn.GELU) and getattr(act, "approximate", "none") == "none"):
            bad.append(f"blocks.{i}.mlp.act={type(act).__name__}")
        if blk.attn.qkv.bias is None or blk.attn.proj.bias is None or blk.mlp.fc1.bias is None or blk.mlp.fc2.bias is None:
            bad.append(f"blocks.{i}: Linear without bias")
    if bad:
        raise NotImplementedError("qatvit_b200: this VisionTransformer variant is not on the fused path (the engine would compute a "
                                  "different function than the module tree): " + ", ".join(bad[:8]))


class _ViTDims:
    def __init__(self, vit: nn.Module, batch: int):
        _check_supported_vit(vit)
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

    def with_batch(self, b: int) -> "_ViTDims":
        import copy
        d = copy.copy(self)
        d.B, d.M = b, b * self.T
        return d


class _BatchBuffers:
    """Activation / scratch buffers of an engine: allocated ONCE for the construction batch B and, for a smaller batch b (the
    ragged last batch of an epoch -- the reference's DataLoaders have no drop_last, ref qat_trainer.py:227-254: 50 000 % 256 = 80
    train, 10 000 % 256 = 16 eval images), re-viewed as the CONTIGUOUS tensors an engine built for b would own (same storage, first
    numel(b) elements).  Kernels take their row counts at launch and TMA clips ragged tiles, so an engine built for B and fed b
    images computes bit for bit what an engine built for b computes (tests/test_partial_batch_gpu.py).  Views are cached per b:
    nothing is allocated and nothing synchronises after the first step of a given size."""

    def _bb_init(self, dev, dims: _ViTDims) -> None:
        self._bb_dev, self._bb_dims = dev, dims
        self._bb_specs, self._bb_full, self._bb_views = {}, {}, {}
        self._bb_cur = dims.B

    def _bb_add(self, name: str, shape_fn, dtype=torch.float32, count: Optional[int] = None, zero: bool = False) -> None:
        shape = tuple(int(x) for x in shape_fn(self._bb_dims))
        alloc = torch.zeros if zero else torch.empty
        full = [alloc(shape, dtype=dtype, device=self._bb_dev) for _ in range(count or 1)]
        self._bb_specs[name] = (shape_fn, count, zero)
        self._bb_full[name] = full
        setattr(self, name, full if count else full[0])

    def _bb_bind(self, b: int) -> None:
        """Point every registered buffer attribute (and self.d) at its batch-b view."""
        if b == self._bb_cur:
            return
        views = self._bb_views.get(b)
        if views is None:
            d = self._bb_dims.with_batch(b)
            views = {"d": d}
            for name, (shape_fn, count, _) in self._bb_specs.items():
                shape = tuple(int(x) for x in shape_fn(d))
                n = math.prod(shape)
                vs = [f.view(-1)[:n].view(shape) for f in self._bb_full[name]]
                views[name] = vs if count else vs[0]
            self._bb_views[b] = views
        for name, v in views.items():
            setattr(self, name, v)
        for name, (_, _, zero) in self._bb_specs.items():
            if zero:        # padded buffers whose padding must read as zero: the layout changed, so clear them (fallback paths only)
                for t in self._bb_full[name]:
                    t.zero_()
        self._bb_cur = b
        self._bb_on_bind(b)

    def _bb_on_bind(self, b: int) -> None:
        pass

    def _bb_check(self, images: torch.Tensor, what: str) -> int:
        d0 = self._bb_dims
        b = int(images.shape[0]) if images.dim() == 4 else -1
        if images.dim() != 4 or tuple(images.shape[1:]) != (d0.in_ch, d0.HW, d0.HW) or not (1 <= b <= d0.B):
            raise RuntimeError(f"{what} engine built for batch {d0.B} of {d0.in_ch}x{d0.HW}x{d0.HW} images (any batch 1..{d0.B} "
                               f"is accepted), got {tuple(images.shape)}")
        return b


def _attention_forward(d: _ViTDims, qkvp: torch.Tensor, S: torch.Tensor, Pp: torch.Tensor, o: torch.Tensor, pairs=PAIRS_FP32) -> None:
    """softmax(Q K^T / 8) V per (image, head) as two batched tcgen05 GEMMs + one row-softmax kernel
    (replaces F.scaled_dot_product_attention in timm Attention.forward)."""
    B, H, T, D = d.B, d.H, d.T, d.D
    BH = B * H
    q_op = Op.tokens(qkvp, B, T, 0, 64)
    k_op = Op.tokens(qkvp, B, T, D, 64)
    ops.gemm(q_op, k_op, T, T, 64, pairs, out=Out.per_head(S, BH, H, T, T), nbatch=BH, batch_inner=H)
    ops.softmax_planes(S, d.ldS, BH * T, T, d.attn_scale, Pp)
    p_op = Op.per_head(Pp, BH, H, T, T)
    v_op = Op.tokens(qkvp, B, T, 2 * D, 64, mn_major=True)
    ops.gemm(p_op, v_op, T, 64, T, pairs, out=Out.tokens(o, B, T, 0, 64), nbatch=BH, batch_inner=H)


class TeacherEngine(_BatchBuffers):
    """Frozen fp32 ViT forward (ref qat_trainer.py:337-338) on the tcgen05 GEMMs, fp32 accumulation.  The block Linears
    take their operands in the mixed format (fp16 value + fp8 copies of the value and of the fp16 rounding residual): one
    kind::f16 product plus two kind::f8f6f4 cross terms at twice the rate -- fp32-grade (~2^-16 per product) for the cost of
    two bf16 passes instead of three.  ``mixed=False`` (or QV_TEACHER_MIX=0) keeps bf16 hi/lo planes and three MMAs per product
    (hi*hi + hi*lo + lo*hi) everywhere; the patch embedding always does.

    Range guard of the mixed format (one fixed 2^7 scale: activations saturate at |x| = 448, weights at |w| = 3.5 --
    csrc/qv_common.cuh).  Per Linear and per block: (a) a weight with max |w| >= 3.5 is routed to the three-pass path when the
    engine is built; (b) every producer of a mixed ACTIVATION tensor (LayerNorm, attention output, fc1 + GELU epilogue) raises a
    device flag bit when a value leaves the range; the flags are copied to pinned host memory after every forward (no sync) and
    looked at when the NEXT forward starts: that Linear then moves to bf16 hi/lo planes for good (``saturation_events``), with a
    RuntimeWarning.  ``calibrate(images)`` does the same synchronously -- call it once on a representative batch before training
    so that no step ever runs with a clamped activation (a fine-tuned in21k ViT-B has outlier channels a random-init one lacks)."""

    SITES = ("qkv", "proj", "fc1", "fc2")

    def __init__(self, vit: nn.Module, batch: int, mixed: Optional[bool] = None):
        import os
        self.vit = vit
        dev = next(vit.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("qatvit_b200: the teacher must live on a CUDA device (there is no CPU fallback)")
        self.dev = dev
        d = self.d = _ViTDims(vit, batch)
        D, F = d.D, d.F
        bf, f32 = torch.bfloat16, torch.float32

        if mixed is None:
            mixed = os.environ.get("QV_TEACHER_MIX", "1") != "0"
        self.mixed = bool(mixed) and D % 64 == 0 and F % 64 == 0
        self.saturation_events: List[str] = []        # "blocks.3.fc2: activation |x| > 448" ... (what left the mixed format, and why)

        self.w_conv = ops.split_planes(vit.patch_embed.proj.weight.detach().reshape(D, -1).contiguous())
        self.blocks = []
        self.mix: List[Dict[str, bool]] = []          # per block and Linear: operands in the mixed format?
        for li, blk in enumerate(vit.blocks):
            lin = dict(qkv=blk.attn.qkv, proj=blk.attn.proj, fc1=blk.mlp.fc1, fc2=blk.mlp.fc2)
            entry = dict(n1=(blk.norm1.weight.detach(), blk.norm1.bias.detach()), n2=(blk.norm2.weight.detach(), blk.norm2.bias.detach()))
            mix = {}
            for name, m in lin.items():
                w2 = m.weight.detach().reshape(m.weight.shape[0], -1).contiguous()
                ok = self.mixed
                if ok and not bool(w2.abs().max() < 3.5):           # build time, once: e4m3(w * 2^7) would saturate (or w is not finite)
                    ok = False
                    self.saturation_events.append(f"blocks.{li}.{name}: weight max |w| >= 3.5")
                mix[name] = ok
                entry[name] = [ops.split_planes_mix(w2, weight=True) if ok else ops.split_planes(w2), m.bias.detach(), m]
            self.blocks.append(entry)
            self.mix.append(mix)
        # range guard of the activation producers: one int32 word per site, bit l = block l
        if d.L > 31:
            raise NotImplementedError("the mixed-format range guard keeps one flag bit per block: depth <= 31")
        self.sat = torch.zeros(len(self.SITES), dtype=torch.int32, device=dev)
        self._sat_of = [self.sat[i:i + 1] for i in range(len(self.SITES))]
        self._sat_host = torch.zeros(len(self.SITES), dtype=torch.int32).pin_memory()
        self._sat_event = torch.cuda.Event()
        self._sat_pending = False
        self._bb_init(dev, d)
        add = self._bb_add
        add("img_planes", lambda d: (2, d.B * d.P, d.Kc), bf)
        add("p_raw", lambda d: (d.B * d.P, d.D))
        add("x", lambda d: (d.M, d.D), count=2)
        add("hp", lambda d: (2, d.M, d.D), bf)
        add("qkvp", lambda d: (2, d.M, 3 * d.D), bf)
        add("op", lambda d: (2, d.M, d.D), bf)
        add("y", lambda d: (d.M, d.D))
        add("fp", lambda d: (2, d.M, d.F), bf)
        add("xn", lambda d: (d.B, d.D))
        add("logits", lambda d: (d.B, d.C))

    # ---- range guard ------------------------------------------------------------------------------
    def _demote(self, li: int, site: str, why: str) -> None:
        """Linear `site` of block li leaves the mixed format: weight re-split as bf16 hi/lo planes, producers follow self.mix."""
        if not self.mix[li][site]:
            return
        import warnings
        ent = self.blocks[li][site]
        m = ent[2]
        ent[0] = ops.split_planes(m.weight.detach().reshape(m.weight.shape[0], -1).contiguous())
        self.mix[li][site] = False
        msg = f"blocks.{li}.{site}: {why}"
        self.saturation_events.append(msg)
        warnings.warn("qatvit_b200 teacher: " + msg + " -- this Linear now runs on bf16 hi/lo planes (three MMA passes)", RuntimeWarning)

    def _apply_flags(self, words) -> bool:
        hit = False
        for si, site in enumerate(self.SITES):
            w = int(words[si])
            for li in range(self.d.L):
                if w >> li & 1 and self.mix[li][site]:
                    self._demote(li, site, "activation |x| > 448 (outside the mixed fp16 + fp8 format)")
                    hit = True
        return hit

    def poll_saturation(self, wait: bool = False) -> bool:
        """Look at the flags of the last forward whose copy has arrived (wait=True: synchronise on it).  True if a Linear was
        demoted.  Called at the start of every forward; never blocks unless asked to."""
        if not self._sat_pending:
            return False
        if wait:
            self._sat_event.synchronize()
        elif not self._sat_event.query():
            return False
        self._sat_pending = False
        words = self._sat_host.tolist()
        if not any(words):
            return False
        hit = self._apply_flags(words)
        ops.zero_(self.sat)
        return hit

    @torch.no_grad()
    def calibrate(self, images: torch.Tensor, max_rounds: int = 4) -> List[str]:
        """Run the forward on `images` until no activation leaves the mixed format (each round demotes the Linears whose input
        saturated).  Synchronises; call once before training.  Returns saturation_events."""
        for _ in range(max_rounds):
            self.forward(images)
            if not self.poll_saturation(wait=True):
                break
        return self.saturation_events

    @torch.no_grad()
    def forward(self, images: torch.Tensor) -> torch.Tensor:
        v = self.vit
        b = self._bb_check(images, "teacher")
        self._bb_bind(b)
        self.poll_saturation()
        d = self.d
        B, T, D, F, M = d.B, d.T, d.D, d.F, d.M
        any_mix = False
        ops.im2col_fq(images, None, B, d.in_ch, d.HW, d.ps, self.img_planes)
        ops.gemm(Op.full(self.img_planes), Op.full(self.w_conv), B * d.P, D, d.Kc, PAIRS_FP32, out=self.p_raw,
                 bias=v.patch_embed.proj.bias.detach())
        ops.embed_fwd(self.p_raw, None, v.cls_token.detach().reshape(-1), v.pos_embed.detach().reshape(T, D), B, d.P, D,
                      self.x[0])
        cur = 0
        x_in, y_prev = self.x[0], None
        for li, blk in enumerate(self.blocks):
            mx, bit = self.mix[li], 1 << li
            any_mix = any_mix or any(mx.values())
            g, b_ = blk["n1"]
            if li == 0:
                ops.resid_ln_fwd(x_in, None, None, g, b_, d.eps, M, D, h_planes=self.hp, planes_mix=mx["qkv"], sat=(self._sat_of[0], bit))
            else:
                ops.resid_ln_fwd(x_in, y_prev, None, g, b_, d.eps, M, D, x_out=self.x[cur ^ 1], h_planes=self.hp,
                                 planes_mix=mx["qkv"], sat=(self._sat_of[0], bit))
                cur ^= 1
                x_in = self.x[cur]
            w, bias, _ = blk["qkv"]
            # no observer sits between the teacher's Linears: epilogues emit the next operand's bf16 planes directly
            ops.gemm(Op.full(self.hp), Op.full(w), M, 3 * D, D, PAIRS_FP32, bias=bias, out_planes=self.qkvp, mix=mx["qkv"])
            # fused softmax attention (bf16 hi/lo q, k, v): scores / probabilities stay in tensor memory, output lands as proj's
            # A operand
            ops.attn_fwd(self.qkvp, B, T, d.H, d.attn_scale, self.op, out_mix=mx["proj"], sat=(self._sat_of[1], bit))
            w, bias, _ = blk["proj"]
            ops.gemm(Op.full(self.op), Op.full(w), M, D, D, PAIRS_FP32, out=self.y, bias=bias, mix=mx["proj"])
            g, b_ = blk["n2"]
            ops.resid_ln_fwd(x_in, self.y, None, g, b_, d.eps, M, D, x_out=self.x[cur ^ 1], h_planes=self.hp, planes_mix=mx["fc1"],
                             sat=(self._sat_of[2], bit))
            cur ^= 1
            x_in = self.x[cur]
            w, bias, _ = blk["fc1"]
            ops.gemm(Op.full(self.hp), Op.full(w), M, F, D, PAIRS_FP32, bias=bias, out_planes=self.fp, gelu=True, mix=mx["fc1"],
                     out_mix=mx["fc2"], sat=(self._sat_of[3], bit))
            w, bias, _ = blk["fc2"]
            ops.gemm(Op.full(self.fp), Op.full(w), M, D, F, PAIRS_FP32, out=self.y, bias=bias, mix=mx["fc2"])
            y_prev = self.y
        ops.resid_ln_fwd(x_in, y_prev, None, v.norm.weight.detach(), v.norm.bias.detach(), d.eps, B, D, in_row_stride=T,
                         h_f32=self.xn)
        ops.head_fwd(self.xn, v.head.weight.detach(), v.head.bias.detach(), B, D, d.C, self.logits)
        if any_mix and not self._sat_pending:
            self._sat_host.copy_(self.sat, non_blocking=True)     # 16 bytes D2H on this stream; read when the next forward starts
            self._sat_event.record()
            self._sat_pending = True
        return self.logits


class StudentEngine(_BatchBuffers):
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
        # fused_attention=False keeps the unfused batched-GEMM kernels (scores / probabilities as planes in HBM): the parity
        # reference for the fused ones.
        # Every consumer applies the producer's fake-quant ON LOAD and every GEMM consumes weight CODES: that is the arithmetic of
        # fake_quant_enabled == 1, the only state the reference ever runs in (it never calls disable_fake_quant).  The flags are
        # read once here; a model with fake-quant switched off on any module is refused instead of silently computing
        # something else (the module-level install() path honours the flags per call).
        fq_off = [n for n, m in student.named_modules() if hasattr(m, "fake_quant_enabled") and hasattr(m, "observer_enabled")
                  and int(m.fake_quant_enabled.item()) == 0]
        if fq_off:
            raise NotImplementedError("qatvit_b200: fake-quant is disabled on " + ", ".join(fq_off[:4]) + (" ..." if len(fq_off) > 4 else "")
                                      + ": the fused engine implements fake_quant_enabled == 1 only (use qatvit_b200.dropin.install() "
                                      "for models with fake-quant switched off)")
        if fused_attention is None:
            fused_attention = True
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
        self._pub_event = torch.cuda.Event()

        # ---- forward activations (saved for backward) and backward scratch: registered with _BatchBuffers so that a smaller
        #      batch (the ragged tail of an epoch) re-views the same storage ----
        self._bb_init(dev, d)
        add = self._bb_add
        lo, fa, fgp = self.ln_obs, self.fused_attn, self.fused_gp
        add("img_codes", lambda d: (1, d.B * d.P, d.Kc), bf)
        add("p_raw", lambda d: (d.B * d.P, d.D))
        add("x_in", lambda d: (d.M, d.D), count=L)
        add("x_mid", lambda d: (d.M, d.D), count=L)
        # A operand of qkv / fc1: LayerNorm output as hi/lo planes, or (observed LN) one plane of codes + the raw output
        npl = 1 if lo else 2
        add("h1p", lambda d: (npl, d.M, d.D), bf, count=L)
        add("h2p", lambda d: (npl, d.M, d.D), bf, count=L)
        if lo:
            add("h1_raw", lambda d: (d.M, d.D), count=L)
            add("h2_raw", lambda d: (d.M, d.D), count=L)
            add("hN_raw", lambda d: (d.M, d.D))
            add("xn_raw", lambda d: (d.B, d.D))
            add("xn_mask", lambda d: (d.B, d.D), torch.uint8)
        add("qkv_raw", lambda d: (d.M, 3 * d.D), count=L)
        if fa:
            add("qkvc", lambda d: (1, d.M, 3 * d.D), bf, count=L)
            add("lse", lambda d: (d.B * d.H * d.T,), count=L)
        else:
            add("qkvp", lambda d: (2, d.M, 3 * d.D), bf, count=L)
            add("Pp", lambda d: (2, d.B * d.H * d.T, d.ldP), bf, count=L, zero=True)
        add("op", lambda d: (2, d.M, d.D), bf, count=L)
        add("a_raw", lambda d: (d.M, d.D), count=L)
        add("f_raw", lambda d: (d.M, d.F), count=L)
        add("gelp", lambda d: (2, d.M, d.F), bf, count=L)
        add("m_raw", lambda d: (d.M, d.D), count=L)
        for nm in ("st1m", "st1r", "st2m", "st2r"):       # LayerNorm mean / rstd per row (norm1, norm2)
            add(nm, lambda d: (d.M,), count=L)
        if not fa:
            add("S", lambda d: (d.B * d.H * d.T, d.ldS))
            add("o", lambda d: (d.M, d.D))
        add("xcls", lambda d: (d.B, d.D))
        add("xn", lambda d: (d.B, d.D))
        add("stFm", lambda d: (d.B,))
        add("stFr", lambda d: (d.B,))
        add("logits_raw", lambda d: (d.B, d.C))
        add("g_logits", lambda d: (d.B, d.C))
        self.loss3 = e(3)

        # ---- backward scratch ----
        add("gx", lambda d: (d.M, d.D), count=2)
        add("g_xn", lambda d: (d.B, d.D))
        add("gpD", lambda d: (2, d.M, d.D), bf)          # fc2's gradient planes
        add("gpDp", lambda d: (2, d.M, d.D), bf)         # proj's (separate, so a weight-gradient GEMM on the side stream can still
        add("gpF", lambda d: (2, d.M, d.F), bf)          #         read one while the main chain already writes the other)
        add("gp3", lambda d: (2, d.M, 3 * d.D), bf)
        add("gpP", lambda d: (2, d.B * d.P, d.D), bf)
        if not fgp:
            add("g_big", lambda d: (d.M, d.F))
        add("g_h", lambda d: (d.M, d.D))
        add("g_op", lambda d: (2, d.M, d.D), bf)
        if not (fgp and fa):
            add("g_qkv", lambda d: (d.M, 3 * d.D))
        if not fa:
            add("g_o", lambda d: (d.M, d.D))
            add("dP", lambda d: (d.B * d.H * d.T, d.ldS))
            add("dSp", lambda d: (2, d.B * d.H * d.T, d.ldP), bf, zero=True)
        import os
        self.rpb_gp = 64
        # token rows per block of the LayerNorm backward: by default ONE full wave of 2 blocks per SM (788 blocks of 64 rows ran
        # 2.66 waves at batch 256: 4 700 -> 5 080 GB/s with 171 rows per block); a function of the CURRENT batch so that b images on
        # an engine built for B reduce in the same order as on an engine built for b.  QV_RPB_LN=<n> pins it (A/B).
        self._rpb_ln_env = int(os.environ.get("QV_RPB_LN", "0"))
        self.rpb_ln = self._rpb_ln_for(M)
        # bias-grad partial sums: standalone gp_planes [M/64][N]; GEMM epilogue [M/32][F]; attention backward [B*mt*4][3D]
        # (sized for the construction batch; a smaller batch uses a prefix)
        slabs_full = B * (-(-T // 128)) * 4
        self.bias_part = e(max(-(-M // self.rpb_gp) * max(F, 3 * D), -(-M // 32) * F, slabs_full * 3 * D, -(-M // 8) * D))
        self.ln_part = e(max(-(-(bb * T) // self._rpb_ln_for(bb * T)) for bb in range(1, B + 1)), 2, D)
        # split-K factors of the weight-gradient GEMMs depend on the token count: one table per batch size 1..B, and a
        # workspace that holds the largest of them (so that b images on an engine built for B split exactly like an engine
        # built for b -- bit-identical sums)
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

    def _rpb_ln_for(self, rows: int) -> int:
        if self._rpb_ln_env > 0:
            return self._rpb_ln_env
        return max(8, -(-rows // (2 * self.sms)))

    def _bb_on_bind(self, b: int) -> None:
        d = self.d
        self.rpb_ln = self._rpb_ln_for(b * d.T)
        self._splits = self._splits_by_b[b]
        self.slabs_attn = b * (-(-d.T // 128)) * 4
        self.stats1 = list(zip(self.st1m, self.st1r))
        self.stats2 = list(zip(self.st2m, self.st2r))
        self.statsF = (self.stFm, self.stFr)

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
        self._bb_bind(self._bb_check(images, "student"))          # any batch 1..B: the ragged tail of an epoch re-views the buffers
        d, v = self.d, self.vit
        B, T, D, F, M, L = d.B, d.T, d.D, d.F, d.M, d.L
        if labels is not None and (labels.dim() != 1 or labels.shape[0] != B):
            raise RuntimeError(f"labels must be a [batch] int64 tensor matching the {B} images, got {tuple(labels.shape)}")
        if teacher_logits is not None and tuple(teacher_logits.shape) != (B, d.C):
            raise RuntimeError(f"teacher logits must be [{B}, {d.C}], got {tuple(teacher_logits.shape)}")
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

    def _publish(self, grads_final_from, lo: int) -> None:
        """Every gradient with arena offset >= lo has been ENQUEUED: hand the suffix to the exchange (ddp.GradSync) without making
        the main chain wait for the weight-gradient side stream.  The collective is issued from the side stream's context after
        that stream has caught up with the main one (bias / LayerNorm gradients are written there), so it is ordered after both,
        while the main stream runs on into the next block."""
        ws = self._wstream
        if ws is None or not self._w_pending:
            return grads_final_from(lo)
        main = torch.cuda.current_stream()
        self._pub_event.record(main)
        with torch.cuda.stream(ws):
            ws.wait_event(self._pub_event)
            grads_final_from(lo)

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
        ops.zero_(gx)                  # only the cls rows carry a gradient out of the final norm (x[:, 0]); the rest is zero
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
                self._publish(grads_final_from, self._block_lo[l])
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
        # optimizer.zero_grad(set_to_none=True) -- every iteration of the reference loop, ref qat_trainer.py:351 -- drops the
        # .grad views into the arena: put them back (152 pointer assignments, no launch, no sync)
        self.student_engine.attach_grads()
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
