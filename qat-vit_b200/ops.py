"""Thin torch-tensor wrappers over the C-ABI (include/qatvit_b200.h).

PyTorch is plumbing here: device memory (caching allocator), streams, dtype checks.  Every function
enqueues hand-written sm_100a kernels on torch's current CUDA stream and returns immediately.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import GemmArgs, check

_NULL = ctypes.c_void_p(0)


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor], dtype=None, name: str = "tensor") -> ctypes.c_void_p:
    if t is None:
        return _NULL
    if not t.is_cuda:
        raise RuntimeError(f"qatvit_b200: {name} must be a CUDA tensor (there is no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"qatvit_b200: {name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"qatvit_b200: {name} must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


def new_minmax(device, count: int = 1) -> torch.Tensor:
    """uint32[count][2] ordered min/max accumulators (stored as int32), reset to the identity."""
    acc = torch.empty(count, 2, dtype=torch.int32, device=device)
    minmax_reset(acc)
    return acc


def minmax_reset(acc: torch.Tensor) -> None:
    check(_lib.lib().qv_minmax_reset(_p(acc, torch.int32, "acc"), acc.numel() // 2, _stream()), "minmax_reset")


def minmax_accumulate(x: torch.Tensor, acc: torch.Tensor) -> None:
    check(_lib.lib().qv_minmax_accumulate(_p(x, torch.float32, "x"), x.numel(), _p(acc, torch.int32, "acc"), _stream()),
          "minmax_accumulate")


def obs_update(acc, observer_enabled, fake_quant_enabled, min_val, max_val, scale, zero_point, averaging_const,
               qmin, qmax, symmetric) -> None:
    check(_lib.lib().qv_obs_update(_p(acc, torch.int32, "acc"), _p(observer_enabled, torch.int64, "observer_enabled"),
                                   _p(fake_quant_enabled, torch.int64, "fake_quant_enabled"),
                                   _p(min_val, torch.float32, "min_val"), _p(max_val, torch.float32, "max_val"),
                                   _p(scale, torch.float32, "scale"), _p(zero_point, torch.int32, "zero_point"),
                                   float(averaging_const), int(qmin), int(qmax), int(bool(symmetric)), _stream()),
          "obs_update")


def fq_apply(x, scale, zero_point, fake_quant_enabled, qmin, qmax, y=None, mask=None, want_mask=True):
    if y is None:
        y = torch.empty_like(x)
    if mask is None and want_mask:
        mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    check(_lib.lib().qv_fq_apply(_p(x, torch.float32, "x"), x.numel(), _p(scale, torch.float32, "scale"),
                                 _p(zero_point, torch.int32, "zero_point"),
                                 _p(fake_quant_enabled, torch.int64, "fake_quant_enabled"), int(qmin), int(qmax),
                                 _p(y, torch.float32, "y"), _p(mask, torch.uint8, "mask"), _stream()), "fq_apply")
    return y, mask


def fq_weight(w, per_channel, observer_enabled, fake_quant_enabled, min_val, max_val, scale, zero_point,
              averaging_const, qmin, qmax, symmetric, y=None, mask=None, codes=None, codes_t=None, scratch=None, scale_vec=None):
    rows = w.shape[0]
    cols = w.numel() // rows
    check(_lib.lib().qv_fq_weight(_p(w, torch.float32, "w"), rows, cols, int(bool(per_channel)),
                                  _p(observer_enabled, torch.int64), _p(fake_quant_enabled, torch.int64),
                                  _p(min_val, torch.float32, "min_val"), _p(max_val, torch.float32, "max_val"),
                                  _p(scale, torch.float32, "scale"), _p(zero_point, torch.int32, "zero_point"),
                                  float(averaging_const), int(qmin), int(qmax), int(bool(symmetric)),
                                  _p(y, torch.float32, "y"), _p(mask, torch.uint8, "mask"),
                                  _p(codes, torch.bfloat16, "codes"), _p(codes_t, torch.bfloat16, "codes_t"),
                                  _p(scratch, torch.int32, "scratch"), _p(scale_vec, torch.float32, "scale_vec"), _stream()),
          "fq_weight")


FQW_ROWS = 16      # output channels per block of qv_fq_weight_grouped (csrc/fakequant.cu)


def fq_weight_group_table(entries, device) -> Tuple[torch.Tensor, int, int]:
    """Device descriptor table for fq_weight_grouped.  entries: dicts with w [N, K] fp32, min_val / max_val / scale fp32 [N],
    zero_point int32 [N], observer_enabled / fake_quant_enabled int64 [1], mask uint8 [N, K], codes bf16 [N, K], codes_t bf16 [K, N].
    -> (uint8 tensor holding the qv_fqw_desc array, total_blocks, max_cols).  The tensors must stay alive and in place."""
    arr = (_lib.FqwDesc * len(entries))()
    start, max_cols = 0, 0
    for d, e in zip(arr, entries):
        w = e["w"]
        rows, cols = w.shape[0], w.numel() // w.shape[0]
        if cols % 4 or rows % 8:
            raise RuntimeError("qatvit_b200: grouped weight fake-quant needs cols % 4 == 0 and rows % 8 == 0")
        d.w = _p(w, torch.float32, "weight").value
        d.min_val, d.max_val = _p(e["min_val"], torch.float32).value, _p(e["max_val"], torch.float32).value
        d.scale, d.zero_point = _p(e["scale"], torch.float32).value, _p(e["zero_point"], torch.int32).value
        d.observer_enabled = _p(e["observer_enabled"], torch.int64).value
        d.fake_quant_enabled = _p(e["fake_quant_enabled"], torch.int64).value
        d.mask, d.codes, d.codes_t = _p(e["mask"], torch.uint8).value, _p(e["codes"], torch.bfloat16).value, \
            _p(e["codes_t"], torch.bfloat16).value
        for t in (e["min_val"], e["max_val"], e["scale"], e["zero_point"]):
            if t.numel() != rows:
                raise RuntimeError("qatvit_b200: per-channel observer state must have one entry per output channel")
        d.rows, d.cols, d.block_start = rows, cols, start
        start += -(-rows // FQW_ROWS)
        max_cols = max(max_cols, cols)
    table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
    return table, start, max_cols


def fq_weight_grouped(table, n_desc, total_blocks, max_cols, averaging_const, qmin, qmax, symmetric):
    """All per-channel fake-quantised weights in one launch (include/qatvit_b200.h: qv_fq_weight_grouped)."""
    check(_lib.lib().qv_fq_weight_grouped(_p(table, torch.uint8, "descriptor table"), n_desc, total_blocks, max_cols,
                                          float(averaging_const), int(qmin), int(qmax), int(bool(symmetric)), _stream()),
          "fq_weight_grouped")


def fq_bwd(gy, mask, gx=None):
    if gx is None:
        gx = torch.empty_like(gy)
    check(_lib.lib().qv_fq_bwd(_p(gy, torch.float32, "gy"), _p(mask, torch.uint8, "mask"), gy.numel(),
                               _p(gx, torch.float32, "gx"), _stream()), "fq_bwd")
    return gx


def fq_learnable_fwd(x, scale, zero_point, qmin, qmax, y=None):
    """Learnable per-channel fake-quant forward, channel axis 0 (include/qatvit_b200.h: qv_fq_learnable_fwd)."""
    rows = x.shape[0]
    if y is None:
        y = torch.empty_like(x)
    check(_lib.lib().qv_fq_learnable_fwd(_p(x, torch.float32, "x"), rows, x.numel() // max(rows, 1), _p(scale, torch.float32, "scale"),
                                         _p(zero_point, torch.float32, "zero_point"), int(qmin), int(qmax),
                                         _p(y, torch.float32, "y"), _stream()), "fq_learnable_fwd")
    return y


def fq_learnable_bwd(gy, x, scale, zero_point, qmin, qmax, grad_factor, want_dx=True):
    """-> (dx, dscale[rows], dzero_point[rows]): STE gradient and the per-channel scale / zero-point gradients (one warp per
    channel, shuffle reduction; qv_fq_learnable_bwd)."""
    rows = x.shape[0]
    dx = torch.empty_like(x) if want_dx else None
    ds = torch.empty(rows, dtype=torch.float32, device=x.device)
    dz = torch.empty(rows, dtype=torch.float32, device=x.device)
    check(_lib.lib().qv_fq_learnable_bwd(_p(gy, torch.float32, "gy"), _p(x, torch.float32, "x"), rows, x.numel() // max(rows, 1),
                                         _p(scale, torch.float32, "scale"), _p(zero_point, torch.float32, "zero_point"),
                                         int(qmin), int(qmax), float(grad_factor), _p(dx, torch.float32, "dx"),
                                         _p(ds, torch.float32), _p(dz, torch.float32), _stream()), "fq_learnable_bwd")
    return dx, ds, dz


def split_planes(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [..] -> bf16 [2, ..] (hi, lo)."""
    if out is None:
        out = torch.empty((2,) + tuple(x.shape), dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().qv_split_planes(_p(x, torch.float32, "x"), x.numel(), _p(out[0], torch.bfloat16),
                                     _p(out[1], torch.bfloat16), _stream()), "split_planes")
    return out


def split_planes_mix(x: torch.Tensor, weight: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [rows, cols] -> the mixed GEMM operand format as a 2-byte-element stack [2, rows, cols]: region 0 = fp16(x * 2^s),
    region 1 = per 64-column block 64 fp8 of the value + 64 e5m2 of the fp16 rounding residual (qv_split_planes_mix)."""
    rows, cols = x.shape
    if out is None:
        out = torch.empty((2, rows, cols), dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().qv_split_planes_mix(_p(x, torch.float32, "x"), rows, cols, int(bool(weight)), _p(out[0], torch.bfloat16),
                                         _p(out[1], torch.bfloat16), _stream()), "split_planes_mix")
    return out


_KD_ROWS_MIN_ELEMS = 1 << 15      # above this many logits the one-block kernel stops being latency-bound: use the grid form
_kd_ws = {}                        # device -> zeroed workspace of the grid form (ticket word + per-block partial sums)


def kd_ce_loss(s_raw, t, labels, T, alpha, eps, s_scale=None, s_zp=None, qmin=0, qmax=255, want_grad=True, out3=None,
               grad=None, rows=None):
    """KL + label-smoothed CE and dL/ds in one launch (ref qat_trainer.py:343-349).  rows: force (True) / forbid (False) the
    grid form ``qv_kd_ce_loss_rows`` (default: by size -- the reference's [B, 10] logits take the one-block kernel)."""
    B, C = s_raw.shape
    if out3 is None:
        out3 = torch.empty(3, dtype=torch.float32, device=s_raw.device)
    if grad is None and want_grad:
        grad = torch.empty_like(s_raw)
    if rows is None:
        rows = B * C >= _KD_ROWS_MIN_ELEMS
    if rows:
        L = _lib.lib()
        need = int(L.qv_kd_ce_rows_workspace_floats(B))
        key = (s_raw.device, torch.cuda.current_stream(s_raw.device).cuda_stream)     # one workspace per stream: launches that may overlap never share a ticket
        ws = _kd_ws.get(key)
        if ws is None or ws.numel() < need:
            ws = _kd_ws[key] = torch.zeros(max(need, 4096), dtype=torch.float32, device=s_raw.device)
        check(L.qv_kd_ce_loss_rows(_p(s_raw, torch.float32, "s"), _p(t, torch.float32, "t"), _p(labels, torch.int64, "labels"),
                                   B, C, float(T), float(alpha), float(eps), _p(s_scale, torch.float32), _p(s_zp, torch.int32),
                                   int(qmin), int(qmax), _p(out3), _p(grad), _p(ws, torch.float32), _stream()), "kd_ce_loss_rows")
        return out3, grad
    check(_lib.lib().qv_kd_ce_loss(_p(s_raw, torch.float32, "s"), _p(t, torch.float32, "t"),
                                   _p(labels, torch.int64, "labels"), B, C, float(T), float(alpha), float(eps),
                                   _p(s_scale, torch.float32), _p(s_zp, torch.int32), int(qmin), int(qmax),
                                   _p(out3), _p(grad), _stream()), "kd_ce_loss")
    return out3, grad


class Op:
    """One GEMM operand: a bf16 plane stack plus how batch items map onto it (struct qv_operand)."""
    __slots__ = ("t", "ptr", "ld", "plane_stride", "mn_major", "rows", "cols", "nb", "batch_stride", "c2_outer",
                 "c2_inner", "col0", "col_inner")

    def __init__(self, t: torch.Tensor, rows: int, cols: int, ld: int, plane_stride: int, mn_major: bool = False,
                 nb: int = 1, batch_stride: int = 0, c2_outer: int = 0, c2_inner: int = 0, col0: int = 0,
                 col_inner: int = 0):
        if t.dtype != torch.bfloat16 or not t.is_cuda:
            raise RuntimeError("qatvit_b200: GEMM operands must be CUDA bf16 plane stacks (no CPU fallback)")
        self.t = t
        self.ptr = t.data_ptr()
        self.rows, self.cols, self.ld, self.plane_stride = rows, cols, ld, plane_stride
        self.mn_major = mn_major
        self.nb, self.batch_stride = nb, batch_stride
        self.c2_outer, self.c2_inner, self.col0, self.col_inner = c2_outer, c2_inner, col0, col_inner

    @staticmethod
    def full(t: torch.Tensor, mn_major: bool = False) -> "Op":
        """t: [planes, rows, cols] -- one unbatched matrix per plane."""
        assert t.dim() == 3 and t.stride(2) == 1
        return Op(t, t.shape[1], t.shape[2], t.stride(1), t.stride(0), mn_major)

    @staticmethod
    def tokens(t: torch.Tensor, B: int, T: int, col0: int, col_inner: int, mn_major: bool = False) -> "Op":
        """t: [planes, B*T, W] token tensor; batch item (b, h) = rows of image b, columns col0 + h*col_inner."""
        assert t.dim() == 3 and t.stride(2) == 1 and t.shape[1] == B * T
        W = t.shape[2]
        return Op(t, T, W, t.stride(1), t.stride(0), mn_major, nb=B, batch_stride=T * t.stride(1), c2_outer=1,
                  c2_inner=0, col0=col0, col_inner=col_inner)

    @staticmethod
    def per_head(t: torch.Tensor, BH: int, H: int, T: int, cols: int, mn_major: bool = False) -> "Op":
        """t: [planes, B*H*T, ld] -- one [T, cols] matrix per (image, head)."""
        assert t.dim() == 3 and t.stride(2) == 1 and t.shape[1] == BH * T
        return Op(t, T, cols, t.stride(1), t.stride(0), mn_major, nb=BH, batch_stride=T * t.stride(1), c2_outer=H,
                  c2_inner=1)

    def fill(self, o: "_lib.Operand") -> None:
        o.ptr = self.ptr
        o.ld, o.plane_stride, o.mn_major = self.ld, self.plane_stride, int(self.mn_major)
        o.rows, o.cols, o.nb, o.batch_stride = self.rows, self.cols, self.nb, self.batch_stride
        o.c2_outer, o.c2_inner, o.col0, o.col_inner = self.c2_outer, self.c2_inner, self.col0, self.col_inner


class Out:
    """Where a GEMM writes: an fp32 tensor plus the batch-item mapping (struct qv_out)."""
    __slots__ = ("t", "ptr", "ld", "rows", "cols", "nb", "batch_stride", "c2_outer", "c2_inner", "col0", "col_inner")

    def __init__(self, t, rows, cols, ld, nb=1, batch_stride=0, c2_outer=0, c2_inner=0, col0=0, col_inner=0):
        if t.dtype != torch.float32 or not t.is_cuda:
            raise RuntimeError("qatvit_b200: GEMM output must be a CUDA fp32 tensor (no CPU fallback)")
        self.t, self.ptr = t, t.data_ptr()
        self.rows, self.cols, self.ld, self.nb, self.batch_stride = rows, cols, ld, nb, batch_stride
        self.c2_outer, self.c2_inner, self.col0, self.col_inner = c2_outer, c2_inner, col0, col_inner

    @staticmethod
    def full(t: torch.Tensor) -> "Out":
        assert t.dim() == 2 and t.stride(1) == 1
        return Out(t, t.shape[0], t.shape[1], t.stride(0))

    @staticmethod
    def tokens(t: torch.Tensor, B: int, T: int, col0: int, col_inner: int) -> "Out":
        """t: [B*T, W]; item (b, h) -> rows of image b, columns col0 + h*col_inner."""
        assert t.dim() == 2 and t.stride(1) == 1 and t.shape[0] == B * T
        return Out(t, T, t.shape[1], t.stride(0), nb=B, batch_stride=T * t.stride(0), c2_outer=1, c2_inner=0, col0=col0,
                   col_inner=col_inner)

    @staticmethod
    def per_head(t: torch.Tensor, BH: int, H: int, T: int, cols: int) -> "Out":
        """t: [B*H*T, ld]; one [T, cols] matrix per (image, head)."""
        assert t.dim() == 2 and t.stride(1) == 1 and t.shape[0] == BH * T
        return Out(t, T, cols, t.stride(0), nb=BH, batch_stride=T * t.stride(0), c2_outer=H, c2_inner=1)

    def fill(self, o) -> None:
        o.ptr, o.ld, o.rows, o.cols, o.nb, o.batch_stride = self.ptr, self.ld, self.rows, self.cols, self.nb, self.batch_stride
        o.c2_outer, o.c2_inner, o.col0, o.col_inner = self.c2_outer, self.c2_inner, self.col0, self.col_inner


# (a_planes, b_planes): which hi/lo plane products the tensor cores accumulate
PAIRS_SINGLE = (1, 1)        # exact integer codes on both sides (patch-embed conv forward)
PAIRS_EXACT_B = (2, 1)       # A = fp32 as hi/lo, B = exact integer codes: A0*B + A1*B
PAIRS_FP32 = (2, 2)          # both fp32 as hi/lo: A0*B0 + A0*B1 + A1*B0 (lo*lo dropped, ~2^-16 relative)


def gemm(a: Op, b: Op, M: int, N: int, K: int, planes: Tuple[int, int], *, out=None, col_scale=None, col_rscale=None,
         alpha=None, bias=None, minmax=None, splits: int = 1, workspace: Optional[torch.Tensor] = None, nbatch: int = 1,
         batch_inner: int = 1, tile_n: int = 0, out_planes: Optional[torch.Tensor] = None, gelu: bool = False,
         grad_of=None, observer=None, mix: bool = False, out_mix: bool = False, sat=None):
    """D[M,N] = sum_pairs A[pa] @ B[pb]^T on tcgen05 (include/qatvit_b200.h: qv_gemm_bf16).
    mix: both operands are in the mixed fp16 + fp8 format (split_planes_mix / out_mix producers); out_mix: out_planes is
    written in the mixed activation format instead of bf16 hi/lo.
    out: an ``Out`` descriptor, a 2-D fp32 tensor, or None (allocated).
    out_planes: bf16 [2, M, N] -- write [GELU](D) as hi/lo planes straight from the epilogue instead of fp32.
    grad_of = (y_raw [M,N], (scale, zp, qmin, qmax), gelu: bool, colsum [ceil(M/32), N] | None): with out_planes, the
    gradient-planes epilogue (act = 2): D * [gelu'(FQ(y))] * STEmask(y) * col_scale -> planes, bias-grad slab sums.
    observer = (min_val, max_val, scale, zero_point, observer_enabled, fake_quant_enabled, c, qmin, qmax, symmetric, ticket):
    with minmax, the output observer's update (obs_update) runs in the GEMM's tail instead of a separate launch."""
    args = GemmArgs()
    a.fill(args.a)
    b.fill(args.b)
    args.a_planes, args.b_planes = planes
    args.M, args.N, args.K = M, N, K
    ret = None
    if out_planes is not None:
        if out_planes.dtype != torch.bfloat16 or not out_planes.is_cuda or out_planes.dim() != 3 or out_planes.stride(2) != 1:
            raise RuntimeError("qatvit_b200: out_planes must be a CUDA bf16 [2, M, N] plane stack (no CPU fallback)")
        o = args.out
        o.ptr, o.ld, o.rows, o.cols, o.nb, o.batch_stride = out_planes.data_ptr(), out_planes.stride(1), M, N, 1, 0
        args.out_kind, args.act, args.out_plane_stride = (2 if out_mix else 1), int(bool(gelu)), out_planes.stride(0)
        if grad_of is not None:
            y_raw, fq, g_gelu, colsum = grad_of
            if y_raw.dtype != torch.float32 or not y_raw.is_cuda or y_raw.dim() != 2 or y_raw.stride(1) != 1:
                raise RuntimeError("qatvit_b200: grad_of[0] must be a CUDA fp32 [M, N] tensor (no CPU fallback)")
            args.act = 2
            args.ep_raw, args.ep_raw_ld = y_raw.data_ptr(), y_raw.stride(0)
            args.ep_scale, args.ep_zp = _p(fq[0], torch.float32, "scale"), _p(fq[1], torch.int32, "zero_point")
            args.ep_qmin, args.ep_qmax, args.ep_gelu = int(fq[2]), int(fq[3]), int(bool(g_gelu))
            if colsum is not None:
                if colsum.dtype != torch.float32 or colsum.numel() < -(-M // 32) * N or not colsum.is_contiguous():
                    raise RuntimeError("qatvit_b200: colsum must be a contiguous fp32 buffer of at least ceil(M/32) * N elements")
                args.ep_colsum = colsum.data_ptr()
        ret = out_planes
    elif splits <= 1:
        if out is None:
            out = torch.empty(M, N, dtype=torch.float32, device=a.t.device)
        ret = out.t if isinstance(out, Out) else out
        (out if isinstance(out, Out) else Out.full(out)).fill(args.out)
    else:
        if workspace is None:
            workspace = torch.empty(splits, M, N, dtype=torch.float32, device=a.t.device)
        args.workspace = workspace.data_ptr()
        ret = workspace
    args.col_scale = None if col_scale is None else col_scale.data_ptr()
    args.col_rscale = None if col_rscale is None else col_rscale.data_ptr()
    args.alpha = None if alpha is None else alpha.data_ptr()
    args.bias = None if bias is None else bias.data_ptr()
    args.minmax = None if minmax is None else minmax.data_ptr()
    if observer is not None:
        mn, mx, sc, zp, on, fq_on, c, qmin, qmax, sym, ticket = observer
        args.obs_min_val, args.obs_max_val = _p(mn, torch.float32, "min_val"), _p(mx, torch.float32, "max_val")
        args.obs_scale, args.obs_zero_point = _p(sc, torch.float32, "scale"), _p(zp, torch.int32, "zero_point")
        args.obs_enabled, args.obs_fq_enabled = _p(on, torch.int64, "observer_enabled"), _p(fq_on, torch.int64, "fake_quant_enabled")
        args.obs_c, args.obs_qmin, args.obs_qmax, args.obs_symmetric = float(c), int(qmin), int(qmax), int(bool(sym))
        args.obs_ticket = _p(ticket, torch.int32, "observer ticket")
    args.splits = splits
    args.nbatch, args.batch_inner = nbatch, batch_inner
    args.tile_n = tile_n
    args.mix = int(bool(mix))
    if sat is not None and out_mix:
        sat_p, args.sat_bit = _sat(sat)
        args.sat_flag = sat_p.value
    check(_lib.lib().qv_gemm_bf16(ctypes.byref(args), _stream()), "gemm_bf16")
    return ret


def splitk_reduce(workspace, splits, M, N, out, row_rscale=None, alpha=None, mask=None, accumulate=False):
    check(_lib.lib().qv_splitk_reduce(_p(workspace, torch.float32), splits, M, N, _p(row_rscale, torch.float32),
                                      _p(alpha, torch.float32), _p(mask, torch.uint8), _p(out, torch.float32),
                                      int(bool(accumulate)), _stream()), "splitk_reduce")
    return out


def launch_count() -> int:
    return int(_lib.lib().qv_launch_count())


def gemm_pair_launches() -> int:
    """GEMM launches that ran as CTA pairs (cta_group::2) so far in this process."""
    return int(_lib.lib().qv_gemm_pair_launches())


def _sat(sat):
    """(flag tensor int32[1], bit) -> ctypes pair for the mixed-format range guard; None -> (NULL, 0)."""
    if sat is None:
        return _NULL, 0
    flag, bit = sat
    if flag.dtype != torch.int32 or not flag.is_cuda or flag.numel() != 1:
        raise RuntimeError("qatvit_b200: the saturation flag must be a CUDA int32 scalar")
    return ctypes.c_void_p(flag.data_ptr()), int(bit)


def zero_(t: torch.Tensor) -> torch.Tensor:
    """Stream-ordered clear of a contiguous CUDA tensor (qv_zero)."""
    check(_lib.lib().qv_zero(_p(t, None, "tensor"), t.numel() * t.element_size(), _stream()), "zero")
    return t


def resid_ln_fwd(x_in, y_raw, fq, gamma, beta, eps, R, D, *, in_row_stride=1, x_out=None, h_planes=None, h_f32=None,
                 mean=None, rstd=None, minmax=None, planes_mix=False, sat=None):
    """x_out = x_in + FQ(y_raw); h = LN(x_out).  fq = (scale, zero_point, qmin, qmax) or None.
    planes_mix: h_planes in the mixed fp16 + fp8 operand format instead of bf16 hi/lo; sat = (flag, bit): its range guard."""
    sc, zp, qmin, qmax = fq if fq is not None else (None, None, 0, 0)
    sat_p, sat_b = _sat(sat if planes_mix else None)
    check(_lib.lib().qv_resid_ln_fwd(_p(x_in, torch.float32), _p(y_raw, torch.float32), _p(sc, torch.float32),
                                     _p(zp, torch.int32), qmin, qmax, _p(gamma, torch.float32), _p(beta, torch.float32),
                                     float(eps), R, D, in_row_stride, _p(x_out, torch.float32),
                                     _p(h_planes, torch.bfloat16), 0 if h_planes is None else h_planes.stride(0),
                                     _p(h_f32, torch.float32), _p(mean, torch.float32), _p(rstd, torch.float32),
                                     _p(minmax, torch.int32), int(bool(planes_mix)), sat_p, sat_b, _stream()), "resid_ln_fwd")


def ln_bwd(g_h, x, mean, rstd, gamma, g_res, R, D, g_x, partials, rows_per_block, out_row_stride=1, h_raw=None, h_fq=None,
           gp=None):
    """h_raw + h_fq = (scale, zero_point, qmin, qmax): g_h first passes the STE mask of an observed LayerNorm's fake-quant.
    gp = (y_raw [R, D], (scale, zp, qmin, qmax), w_scale [D], out_planes bf16 [2, R, D], bias_partials | None): also emit the
    gradient planes g_x * STEmask(y_raw) * w_scale of the Linear whose output y_raw fed this residual stream (qv_ln_bwd_gp)."""
    sc, zp, qmin, qmax = h_fq if (h_raw is not None and h_fq is not None) else (None, None, 0, 0)
    if gp is not None:
        if out_row_stride != 1:
            raise RuntimeError("qatvit_b200: ln_bwd with fused gradient planes needs out_row_stride == 1")
        y_raw, fq, w_scale, out_planes, bias_part = gp
        check(_lib.lib().qv_ln_bwd_gp(_p(g_h, torch.float32), _p(x, torch.float32), _p(mean, torch.float32),
                                      _p(rstd, torch.float32), _p(gamma, torch.float32), _p(g_res, torch.float32), R, D,
                                      _p(g_x, torch.float32), _p(partials, torch.float32), rows_per_block,
                                      _p(h_raw if sc is not None else None, torch.float32), _p(sc, torch.float32),
                                      _p(zp, torch.int32), qmin, qmax, _p(y_raw, torch.float32, "gp y_raw"),
                                      _p(fq[0], torch.float32), _p(fq[1], torch.int32), int(fq[2]), int(fq[3]),
                                      _p(w_scale, torch.float32), _p(out_planes, torch.bfloat16, "gp out_planes"),
                                      out_planes.stride(0), _p(bias_part, torch.float32), _stream()), "ln_bwd_gp")
        return
    check(_lib.lib().qv_ln_bwd(_p(g_h, torch.float32), _p(x, torch.float32), _p(mean, torch.float32),
                               _p(rstd, torch.float32), _p(gamma, torch.float32), _p(g_res, torch.float32), R, D,
                               out_row_stride, _p(g_x, torch.float32), _p(partials, torch.float32), rows_per_block,
                               _p(h_raw if sc is not None else None, torch.float32), _p(sc, torch.float32), _p(zp, torch.int32),
                               qmin, qmax, _stream()), "ln_bwd")


def colsum_reduce(partials, nblk, ncols, out, accumulate=False):
    check(_lib.lib().qv_colsum_reduce(_p(partials, torch.float32), nblk, ncols, _p(out, torch.float32),
                                      int(bool(accumulate)), _stream()), "colsum_reduce")


def colsum_rows(x, R, N, ld, out, accumulate=False):
    check(_lib.lib().qv_colsum_rows(_p(x, torch.float32), R, N, ld, _p(out, torch.float32), int(bool(accumulate)),
                                    _stream()), "colsum_rows")


def gp_planes(g, y_raw, fq, w_scale, per_channel, gelu, R, N, out_planes, bias_partials, rows_per_block, remap_P=0,
              remap_T=0):
    sc, zp, qmin, qmax = fq if fq is not None else (None, None, 0, 0)
    check(_lib.lib().qv_gp_planes(_p(g, torch.float32), _p(y_raw, torch.float32), _p(sc, torch.float32),
                                  _p(zp, torch.int32), qmin, qmax, _p(w_scale, torch.float32), int(bool(per_channel)),
                                  int(bool(gelu)), R, N, remap_P, remap_T, _p(out_planes, torch.bfloat16),
                                  out_planes.stride(0), _p(bias_partials, torch.float32), rows_per_block, _stream()),
          "gp_planes")


def act_planes(y_raw, fq, gelu, out_planes, codes_only=False):
    """out_planes [2, ...] <- hi/lo planes of [GELU](FQ(y_raw)); codes_only: out_planes [1, ...] <- centred integer codes."""
    sc, zp, qmin, qmax = fq if fq is not None else (None, None, 0, 0)
    check(_lib.lib().qv_act_planes(_p(y_raw, torch.float32), _p(sc, torch.float32), _p(zp, torch.int32), qmin, qmax,
                                   int(bool(gelu)), int(bool(codes_only)), y_raw.numel(), _p(out_planes, torch.bfloat16),
                                   out_planes.stride(0), _stream()), "act_planes")


def embed_fwd(p_raw, fq, cls, pos, B, P, D, x0):
    sc, zp, qmin, qmax = fq if fq is not None else (None, None, 0, 0)
    check(_lib.lib().qv_embed_fwd(_p(p_raw, torch.float32), _p(sc, torch.float32), _p(zp, torch.int32), qmin, qmax,
                                  _p(cls, torch.float32), _p(pos, torch.float32), B, P, D, _p(x0, torch.float32),
                                  _stream()), "embed_fwd")


def im2col_fq(img, fq, B, C, HW, patch, out_planes):
    """fq given: out_planes [1, B*P, K] integer codes; fq None: out_planes [2, B*P, K] hi/lo of the raw pixels."""
    sc, zp, qmin, qmax = fq if fq is not None else (None, None, 0, 0)
    lo = None if fq is not None else out_planes[1]
    check(_lib.lib().qv_im2col_fq(_p(img, torch.float32), _p(sc, torch.float32), _p(zp, torch.int32), qmin, qmax, B, C, HW,
                                  patch, _p(out_planes[0], torch.bfloat16), _p(lo, torch.bfloat16), _stream()), "im2col_fq")


def softmax_planes(S, ldS, rows, T, scale, P_planes):
    check(_lib.lib().qv_softmax_planes(_p(S, torch.float32), ldS, rows, T, float(scale), _p(P_planes, torch.bfloat16),
                                       P_planes.stride(1), P_planes.stride(0), _stream()), "softmax_planes")


def attn_ds(P_planes, dP, lddP, rows, T, scale, dS_planes):
    check(_lib.lib().qv_attn_ds(_p(P_planes, torch.bfloat16), P_planes.stride(1), P_planes.stride(0),
                                _p(dP, torch.float32), lddP, rows, T, float(scale), _p(dS_planes, torch.bfloat16),
                                dS_planes.stride(1), dS_planes.stride(0), _stream()), "attn_ds")


def attn_fwd(qkv_planes, B, T, H, scale, out_planes, qk_scale=None, v_scale=None, lse=None, out_f32=None, out_mix=False,
             sat=None):
    """Fused softmax(Q K^T * scale) V per (image, head): qkv_planes bf16 [1 or 2, B*T, 3*H*64] -> out_planes bf16
    [2, B*T, H*64] (include/qatvit_b200.h: qv_attn_fwd)."""
    if qkv_planes.dim() != 3 or qkv_planes.stride(2) != 1 or (out_planes is not None and (out_planes.dim() != 3 or out_planes.stride(2) != 1)):
        raise RuntimeError("qatvit_b200: attn_fwd takes [planes, tokens, cols] plane stacks")
    ops_, opl = (out_planes.stride(0), out_planes.stride(1)) if out_planes is not None else (0, 0)
    sat_p, sat_b = _sat(sat if out_mix else None)
    check(_lib.lib().qv_attn_fwd(_p(qkv_planes, torch.bfloat16, "qkv_planes"), qkv_planes.shape[0], qkv_planes.stride(0),
                                 qkv_planes.stride(1), B, T, H, float(scale), _p(qk_scale, torch.float32),
                                 _p(v_scale, torch.float32), _p(out_planes, torch.bfloat16, "out_planes"), ops_, opl,
                                 _p(out_f32, torch.float32, "out_f32"), _p(lse, torch.float32), int(bool(out_mix)), sat_p, sat_b,
                                 _stream()), "attn_fwd")
    return out_planes if out_planes is not None else out_f32


def _check_attn_bwd_planes(qkv_codes, o_planes, do_planes):
    if qkv_codes.dim() != 3 or qkv_codes.shape[0] != 1 or do_planes.dim() != 3 or do_planes.shape[0] != 2 or \
            o_planes.dim() != 3 or o_planes.shape[0] != 2:
        raise RuntimeError("qatvit_b200: attn_bwd takes a [1, tokens, 3D] code plane and [2, tokens, D] output / gradient planes")


def attn_bwd(qkv_codes, qscale, o_planes, do_planes, lse, B, T, H, scale, g_qkv):
    """Fused attention backward on integer codes: g_qkv fp32 [B*T, 3*H*64] <- dQ | dK | dV (include/qatvit_b200.h: qv_attn_bwd).
    o_planes: the forward's output planes (attn_fwd out_planes)."""
    _check_attn_bwd_planes(qkv_codes, o_planes, do_planes)
    check(_lib.lib().qv_attn_bwd(_p(qkv_codes, torch.bfloat16, "qkv_codes"), qkv_codes.stride(1), _p(qscale, torch.float32),
                                 _p(o_planes, torch.bfloat16, "o_planes"), o_planes.stride(0), o_planes.stride(1),
                                 _p(do_planes, torch.bfloat16, "do_planes"), do_planes.stride(0), do_planes.stride(1),
                                 _p(lse, torch.float32, "lse"), B, T, H, float(scale), _p(g_qkv, torch.float32, "g_qkv"),
                                 _stream()), "attn_bwd")
    return g_qkv


def attn_bwd_gp(qkv_codes, qscale, o_planes, do_planes, lse, B, T, H, scale, y_raw, fq, w_scale, gp_planes, colsum):
    """attn_bwd with the qkv Linear's gradient prologue fused: gp_planes bf16 [2, B*T, 3*H*64] <- (dQ|dK|dV) * STEmask(y_raw) *
    w_scale; colsum fp32 [B * ceil(T/128) * 4, 3*H*64] <- per-slab bias-grad partials (include/qatvit_b200.h: qv_attn_bwd_gp)."""
    _check_attn_bwd_planes(qkv_codes, o_planes, do_planes)
    nslab = B * (-(-T // 128)) * 4
    if colsum.numel() < nslab * 3 * H * 64 or gp_planes.dim() != 3 or gp_planes.stride(1) != 3 * H * 64:
        raise RuntimeError("qatvit_b200: attn_bwd_gp needs dense [2, B*T, 3D] planes and a [B*ceil(T/128)*4, 3D] colsum buffer")
    check(_lib.lib().qv_attn_bwd_gp(_p(qkv_codes, torch.bfloat16, "qkv_codes"), qkv_codes.stride(1), _p(qscale, torch.float32),
                                    _p(o_planes, torch.bfloat16, "o_planes"), o_planes.stride(0), o_planes.stride(1),
                                    _p(do_planes, torch.bfloat16, "do_planes"), do_planes.stride(0), do_planes.stride(1),
                                    _p(lse, torch.float32, "lse"), B, T, H, float(scale), _p(y_raw, torch.float32, "y_raw"),
                                    _p(fq[0], torch.float32), _p(fq[1], torch.int32), int(fq[2]), int(fq[3]),
                                    _p(w_scale, torch.float32), _p(gp_planes, torch.bfloat16, "gp_planes"), gp_planes.stride(0),
                                    _p(colsum, torch.float32), _stream()), "attn_bwd_gp")
    return nslab


def int8_linear(qx, sx, zx, qw, sw, wsum, bias, sy, zy, qy=None, y=None, engine="x86"):
    """quantized::linear on the integer tensor cores (include/qatvit_b200.h: qv_int8_linear).  qx uint8 [M,K], qw int8 [N,K]."""
    M, K = qx.shape
    N = qw.shape[0]
    if qy is None and y is None:
        qy = torch.empty(M, N, dtype=torch.uint8, device=qx.device)
    check(_lib.lib().qv_int8_linear(_p(qx, torch.uint8, "qx"), M, K, _p(sx, torch.float32, "sx"), _p(zx, torch.int32, "zx"),
                                    _p(qw, torch.int8, "qw"), N, _p(sw, torch.float32, "sw"), int(sw.numel() > 1),
                                    _p(wsum, torch.int32, "wsum"), _p(bias, torch.float32, "bias"), float(sy), int(zy),
                                    int(engine == "qnnpack"), _p(qy, torch.uint8, "qy"), _p(y, torch.float32, "y"), _stream()), "int8_linear")
    return qy if qy is not None else y


def int8_linear_codes(qx, sx, zx, qw, sw, wsum, bias, sy, zy, codes, engine="x86"):
    """quantized::linear whose output is the centred code q_y - zy as ONE bf16 plane (the integer operand of attn_fwd):
    codes bf16 [M, N] or [1, M, N] (include/qatvit_b200.h: qv_int8_linear_codes)."""
    M, K = qx.shape
    N = qw.shape[0]
    if codes.numel() != M * N:
        raise RuntimeError("qatvit_b200: codes must hold M * N bf16 values")
    check(_lib.lib().qv_int8_linear_codes(_p(qx, torch.uint8, "qx"), M, K, _p(sx, torch.float32, "sx"), _p(zx, torch.int32, "zx"),
                                          _p(qw, torch.int8, "qw"), N, _p(sw, torch.float32, "sw"), int(sw.numel() > 1),
                                          _p(wsum, torch.int32, "wsum"), _p(bias, torch.float32, "bias"), float(sy), int(zy),
                                          int(engine == "qnnpack"), _p(codes, torch.bfloat16, "codes"), _stream()), "int8_linear_codes")
    return codes


def quantize_u8(x, scale, zero_point, out=None):
    if out is None:
        out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    check(_lib.lib().qv_quantize_u8(_p(x, torch.float32, "x"), x.numel(), _p(scale, torch.float32), _p(zero_point, torch.int32),
                                    _p(out, torch.uint8, "out"), _stream()), "quantize_u8")
    return out


def qparams_from_minmax(acc, qmin, qmax, scale, zero_point):
    check(_lib.lib().qv_qparams_from_minmax(_p(acc, torch.int32, "acc"), int(qmin), int(qmax), _p(scale, torch.float32),
                                            _p(zero_point, torch.int32), _stream()), "qparams_from_minmax")


def im2col_u8(img, scale, zero_point, B, C, HW, patch, out):
    check(_lib.lib().qv_im2col_u8(_p(img, torch.float32, "img"), _p(scale, torch.float32), _p(zero_point, torch.int32), B, C, HW,
                                  patch, _p(out, torch.uint8, "out"), _stream()), "im2col_u8")


def gelu_minmax(x, y, acc=None):
    check(_lib.lib().qv_gelu_minmax(_p(x, torch.float32, "x"), x.numel(), _p(y, torch.float32, "y"), _p(acc, torch.int32),
                                    _stream()), "gelu_minmax")


def quantize_u8_dyn(x, acc, scale_out, zero_point_out, out):
    """qparams_from_minmax(acc, 0, 255) + quantize_u8 in one launch; the qparams land in scale_out / zero_point_out."""
    check(_lib.lib().qv_quantize_u8_dyn(_p(x, torch.float32, "x"), x.numel(), _p(acc, torch.int32, "acc"),
                                        _p(scale_out, torch.float32, "scale_out"), _p(zero_point_out, torch.int32, "zero_point_out"),
                                        _p(out, torch.uint8, "out"), _stream()), "quantize_u8_dyn")
    return out


def ln_quantize_u8_dyn(x, mean, rstd, gamma, beta, R, D, acc, scale_out, zero_point_out, out):
    """out = quantize_u8(LayerNorm(x)) with the dynamic qparams of acc, from the row statistics resid_ln_fwd saved
    (include/qatvit_b200.h: qv_ln_quantize_u8_dyn); the qparams land in scale_out / zero_point_out."""
    check(_lib.lib().qv_ln_quantize_u8_dyn(_p(x, torch.float32, "x"), _p(mean, torch.float32, "mean"), _p(rstd, torch.float32, "rstd"),
                                           _p(gamma, torch.float32, "gamma"), _p(beta, torch.float32, "beta"), R, D,
                                           _p(acc, torch.int32, "acc"), _p(scale_out, torch.float32, "scale_out"),
                                           _p(zero_point_out, torch.int32, "zero_point_out"), _p(out, torch.uint8, "out"), _stream()),
          "ln_quantize_u8_dyn")
    return out


def codes_from_u8(q, zero_point, codes):
    """codes = bf16(q - zero_point): a converted Linear's quint8 output as the one-plane integer operand of attn_fwd."""
    check(_lib.lib().qv_codes_from_u8(_p(q, torch.uint8, "q"), q.numel(), int(zero_point), _p(codes, torch.bfloat16, "codes"),
                                      _stream()), "codes_from_u8")
    return codes


def gelu_u8_minmax(q, sy, zy, acc):
    check(_lib.lib().qv_gelu_u8_minmax(_p(q, torch.uint8, "q"), q.numel(), float(sy), int(zy), _p(acc, torch.int32, "acc"),
                                       _stream()), "gelu_u8_minmax")


def gelu_u8_requant(q, sy, zy, acc, scale_out, zero_point_out, out):
    check(_lib.lib().qv_gelu_u8_requant(_p(q, torch.uint8, "q"), q.numel(), float(sy), int(zy), _p(acc, torch.int32, "acc"),
                                        _p(scale_out, torch.float32, "scale_out"), _p(zero_point_out, torch.int32, "zero_point_out"),
                                        _p(out, torch.uint8, "out"), _stream()), "gelu_u8_requant")
    return out


def head_fwd(x, wq, bias, B, K, N, out, minmax=None):
    check(_lib.lib().qv_head_fwd(_p(x, torch.float32), _p(wq, torch.float32), _p(bias, torch.float32), B, K, N,
                                 _p(out, torch.float32), _p(minmax, torch.int32), _stream()), "head_fwd")


def head_bwd(g, x, wq, wmask, B, K, N, gx, gw, gb, accumulate=False):
    check(_lib.lib().qv_head_bwd(_p(g, torch.float32), _p(x, torch.float32), _p(wq, torch.float32),
                                 _p(wmask, torch.uint8), B, K, N, _p(gx, torch.float32), _p(gw, torch.float32),
                                 _p(gb, torch.float32), int(bool(accumulate)), _stream()), "head_bwd")


# ------------------------------------------------------------------------------------------------
# optional per-op CUDA-event profiling (bench.py's roofline section; off by default, zero overhead when off)
# ------------------------------------------------------------------------------------------------
_prof = None
_prof_gc = True


def _gemm_tag(args, kw):
    """profile tag + algorithmic FLOPs (2 M N K per product of the ORIGINAL fp32 / integer operands; the hi/lo bf16 passes
    that realise an fp32 product on the tensor cores are not counted)."""
    a, b, M, N, K, planes = args[0], args[1], args[2], args[3], args[4], args[5]
    nb = kw.get("nbatch", 1)
    if nb > 1:
        kind = "attn (unfused)"
    elif a.mn_major:
        kind = "wgrad"
    elif tuple(planes) == (2, 2):
        kind = "teacher linear"
    elif kw.get("grad_of") is not None:
        kind = "student dgrad+gp"
    elif tuple(planes) == (2, 1):
        kind = "student fwd/dgrad"
    else:
        kind = "patch-embed"
    return f"gemm[{kind}]", 2.0 * M * N * K * nb


def _n(t):
    return 0 if t is None else t.numel()


def _bytes_act_planes(a, k):
    n = a[0].numel()
    return "act_planes", 4.0 * n + (2.0 * n if k.get("codes_only", False) else 4.0 * n)


def _bytes_gp_planes(a, k):
    R, N = a[6], a[7]
    return "gp_planes", (12.0 if a[1] is not None else 8.0) * R * N


def _bytes_resid_ln(a, k):
    R, D = a[6], a[7]
    n = sum(1 for t in (a[0], a[1], k.get("x_out"), k.get("h_planes"), k.get("h_f32")) if t is not None)
    return "resid_ln_fwd", 4.0 * n * R * D


def _bytes_ln_bwd(a, k):
    R, D = a[6], a[7]
    return "ln_bwd", ((16.0 if a[5] is not None else 12.0) + (8.0 if k.get("gp") is not None else 0.0)) * R * D


def _wrap(name, fn, tag_fn=None):
    def inner(*a, **k):
        if _prof is None:
            return fn(*a, **k)
        tag, work = tag_fn(a, k) if tag_fn else (name, 0.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = fn(*a, **k)
        e.record()
        _prof.append((tag, work, s, e))
        return r
    inner.__name__ = name
    inner.__doc__ = fn.__doc__
    return inner


for _nm in ("zero_", "minmax_reset", "minmax_accumulate", "obs_update", "fq_apply", "fq_weight", "fq_weight_grouped", "fq_bwd", "fq_learnable_fwd", "fq_learnable_bwd", "split_planes",
           "kd_ce_loss", "splitk_reduce", "colsum_reduce", "colsum_rows",
           "embed_fwd", "im2col_fq", "softmax_planes", "attn_ds", "head_fwd", "head_bwd", "attn_fwd", "attn_bwd", "attn_bwd_gp",
           "int8_linear", "quantize_u8", "qparams_from_minmax", "im2col_u8", "gelu_minmax", "quantize_u8_dyn", "codes_from_u8",
           "gelu_u8_minmax", "gelu_u8_requant", "int8_linear_codes", "ln_quantize_u8_dyn"):
    globals()[_nm] = _wrap(_nm, globals()[_nm])
gemm = _wrap("gemm", gemm, _gemm_tag)
act_planes = _wrap("act_planes", act_planes, _bytes_act_planes)
gp_planes = _wrap("gp_planes", gp_planes, _bytes_gp_planes)
resid_ln_fwd = _wrap("resid_ln_fwd", resid_ln_fwd, _bytes_resid_ln)
ln_bwd = _wrap("ln_bwd", ln_bwd, _bytes_ln_bwd)


def profiling() -> bool:
    return _prof is not None


def profile_begin() -> None:
    """Start recording a CUDA-event pair around every op.  The cyclic garbage collector is paused until profile_end(): a
    generation-2 collection landing between an op's two events (tens of ms with a model's worth of tensors alive) would be
    charged to that op -- seen as a 35-45 ms `int8_linear` / `resid_ln_fwd` in tools/int8_quick.py."""
    global _prof, _prof_gc
    import gc
    _prof_gc = gc.isenabled()
    gc.collect()
    gc.disable()
    _prof = []


def profile_end():
    """-> {tag: dict(count, ms, work)} with CUDA-event durations of every op since profile_begin()."""
    global _prof
    torch.cuda.synchronize()
    if _prof_gc:
        import gc
        gc.enable()
    out = {}
    for tag, work, s, e in _prof:
        d = out.setdefault(tag, dict(count=0, ms=0.0, work=0.0))
        d["count"] += 1
        d["ms"] += s.elapsed_time(e)
        d["work"] += work
    _prof = None
    return out
