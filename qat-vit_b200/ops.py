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
              averaging_const, qmin, qmax, symmetric, y=None, mask=None, codes=None, codes_t=None, scratch=None):
    rows = w.shape[0]
    cols = w.numel() // rows
    check(_lib.lib().qv_fq_weight(_p(w, torch.float32, "w"), rows, cols, int(bool(per_channel)),
                                  _p(observer_enabled, torch.int64), _p(fake_quant_enabled, torch.int64),
                                  _p(min_val, torch.float32, "min_val"), _p(max_val, torch.float32, "max_val"),
                                  _p(scale, torch.float32, "scale"), _p(zero_point, torch.int32, "zero_point"),
                                  float(averaging_const), int(qmin), int(qmax), int(bool(symmetric)),
                                  _p(y, torch.float32, "y"), _p(mask, torch.uint8, "mask"),
                                  _p(codes, torch.bfloat16, "codes"), _p(codes_t, torch.bfloat16, "codes_t"),
                                  _p(scratch, torch.int32, "scratch"), _stream()), "fq_weight")


def fq_bwd(gy, mask, gx=None):
    if gx is None:
        gx = torch.empty_like(gy)
    check(_lib.lib().qv_fq_bwd(_p(gy, torch.float32, "gy"), _p(mask, torch.uint8, "mask"), gy.numel(),
                               _p(gx, torch.float32, "gx"), _stream()), "fq_bwd")
    return gx


def split_planes(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [..] -> bf16 [2, ..] (hi, lo)."""
    if out is None:
        out = torch.empty((2,) + tuple(x.shape), dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().qv_split_planes(_p(x, torch.float32, "x"), x.numel(), _p(out[0], torch.bfloat16),
                                     _p(out[1], torch.bfloat16), _stream()), "split_planes")
    return out


def kd_ce_loss(s_raw, t, labels, T, alpha, eps, s_scale=None, s_zp=None, qmin=0, qmax=255, want_grad=True):
    B, C = s_raw.shape
    out3 = torch.empty(3, dtype=torch.float32, device=s_raw.device)
    grad = torch.empty_like(s_raw) if want_grad else None
    check(_lib.lib().qv_kd_ce_loss(_p(s_raw, torch.float32, "s"), _p(t, torch.float32, "t"),
                                   _p(labels, torch.int64, "labels"), B, C, float(T), float(alpha), float(eps),
                                   _p(s_scale, torch.float32), _p(s_zp, torch.int32), int(qmin), int(qmax),
                                   _p(out3), _p(grad), _stream()), "kd_ce_loss")
    return out3, grad


def gemm(a: torch.Tensor, b: torch.Tensor, M: int, N: int, K: int, pairs: Sequence[Tuple[int, int]], *,
         a_mn_major: bool = False, b_mn_major: bool = False, out: Optional[torch.Tensor] = None,
         col_scale=None, col_rscale=None, alpha=None, bias=None, minmax=None, splits: int = 1,
         workspace: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """D[M,N] = sum_pairs A[pa] @ B[pb]^T on tcgen05.  a, b: bf16 plane stacks [planes, rows, ld]."""
    if a.dim() != 3 or b.dim() != 3:
        raise RuntimeError("gemm operands must be [planes, rows, cols] bf16 plane stacks")
    args = GemmArgs()
    args.a = a.data_ptr(); args.lda = a.stride(1); args.a_plane_stride = a.stride(0); args.a_mn_major = int(a_mn_major)
    args.b = b.data_ptr(); args.ldb = b.stride(1); args.b_plane_stride = b.stride(0); args.b_mn_major = int(b_mn_major)
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16 or not a.is_cuda or not b.is_cuda:
        raise RuntimeError("gemm operands must be CUDA bf16")
    if a.stride(2) != 1 or b.stride(2) != 1:
        raise RuntimeError("gemm operand inner stride must be 1")
    args.npairs = len(pairs)
    for i, (pa, pb) in enumerate(pairs):
        args.pair_a[i] = pa
        args.pair_b[i] = pb
    args.M, args.N, args.K = M, N, K
    if splits <= 1:
        if out is None:
            out = torch.empty(M, N, dtype=torch.float32, device=a.device)
        args.d = out.data_ptr(); args.ldd = out.stride(0)
    else:
        if workspace is None:
            workspace = torch.empty(splits, M, N, dtype=torch.float32, device=a.device)
        args.workspace = workspace.data_ptr()
    args.col_scale = None if col_scale is None else col_scale.data_ptr()
    args.col_rscale = None if col_rscale is None else col_rscale.data_ptr()
    args.alpha = None if alpha is None else alpha.data_ptr()
    args.bias = None if bias is None else bias.data_ptr()
    args.minmax = None if minmax is None else minmax.data_ptr()
    args.splits = splits
    check(_lib.lib().qv_gemm_bf16(ctypes.byref(args), _stream()), "gemm_bf16")
    return out if splits <= 1 else workspace


def splitk_reduce(workspace, splits, M, N, out, row_rscale=None, alpha=None, mask=None, accumulate=False):
    check(_lib.lib().qv_splitk_reduce(_p(workspace, torch.float32), splits, M, N, _p(row_rscale, torch.float32),
                                      _p(alpha, torch.float32), _p(mask, torch.uint8), _p(out, torch.float32),
                                      int(bool(accumulate)), _stream()), "splitk_reduce")
    return out


def launch_count() -> int:
    return int(_lib.lib().qv_launch_count())
