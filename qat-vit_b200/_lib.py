"""ctypes loader for libqatvit_b200.so (the C-ABI in include/qatvit_b200.h).

There is NO fallback: if the shared library is missing, importing this module raises, and every compute
entry point fails without a CUDA device.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QV_LIB") or os.path.join(_HERE, "lib", "libqatvit_b200.so")    # QV_LIB: an instrumented build
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "qatvit_b200.h")


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    res = subprocess.run(cmd, capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libqatvit_b200.so failed:\n" + (res.stdout or "") + (res.stderr or ""))
    return LIB_PATH


class Operand(ctypes.Structure):
    """struct qv_operand (include/qatvit_b200.h)."""
    _fields_ = [
        ("ptr", c_void_p), ("ld", c_int64), ("plane_stride", c_int64), ("mn_major", c_int32),
        ("rows", c_int64), ("cols", c_int64), ("nb", c_int64), ("batch_stride", c_int64),
        ("c2_outer", c_int32), ("c2_inner", c_int32), ("col0", c_int32), ("col_inner", c_int32),
    ]


class Out(ctypes.Structure):
    """struct qv_out (include/qatvit_b200.h)."""
    _fields_ = [
        ("ptr", c_void_p), ("ld", c_int64), ("rows", c_int64), ("cols", c_int64), ("nb", c_int64),
        ("batch_stride", c_int64), ("c2_outer", c_int32), ("c2_inner", c_int32), ("col0", c_int32), ("col_inner", c_int32),
    ]


class FqwDesc(ctypes.Structure):
    """struct qv_fqw_desc (include/qatvit_b200.h)."""
    _fields_ = [
        ("w", c_void_p), ("min_val", c_void_p), ("max_val", c_void_p), ("scale", c_void_p), ("zero_point", c_void_p),
        ("observer_enabled", c_void_p), ("fake_quant_enabled", c_void_p),
        ("mask", c_void_p), ("codes", c_void_p), ("codes_t", c_void_p),
        ("rows", c_int32), ("cols", c_int32), ("block_start", c_int32), ("reserved", c_int32),
    ]


class GemmArgs(ctypes.Structure):
    """struct qv_gemm_args (include/qatvit_b200.h)."""
    _fields_ = [
        ("a", Operand), ("b", Operand),
        ("a_planes", c_int32), ("b_planes", c_int32),
        ("M", c_int64), ("N", c_int64), ("K", c_int64),
        ("out", Out),
        ("col_scale", c_void_p), ("col_rscale", c_void_p), ("alpha", c_void_p), ("bias", c_void_p),
        ("minmax", c_void_p),
        ("splits", c_int32), ("workspace", c_void_p),
        ("nbatch", c_int32), ("batch_inner", c_int32), ("tile_n", c_int32),
        ("out_kind", c_int32), ("act", c_int32), ("out_plane_stride", c_int64),
        ("ep_raw", c_void_p), ("ep_raw_ld", c_int64), ("ep_scale", c_void_p), ("ep_zp", c_void_p),
        ("ep_qmin", c_int32), ("ep_qmax", c_int32), ("ep_gelu", c_int32), ("ep_colsum", c_void_p),
        ("obs_min_val", c_void_p), ("obs_max_val", c_void_p), ("obs_scale", c_void_p), ("obs_zero_point", c_void_p),
        ("obs_enabled", c_void_p), ("obs_fq_enabled", c_void_p),
        ("obs_c", c_float), ("obs_qmin", c_int32), ("obs_qmax", c_int32), ("obs_symmetric", c_int32),
        ("obs_ticket", c_void_p),
        ("mix", c_int32),
        ("sat_flag", c_void_p), ("sat_bit", c_int32),
    ]


_P = c_void_p
_SIGNATURES = {
    "qv_version": (c_int, []),
    "qv_last_error": (c_char_p, []),
    "qv_device_sm_count": (c_int, []),
    "qv_launch_count": (c_int64, []),
    "qv_gemm_pair_launches": (c_int64, []),
    "qv_zero": (c_int, [_P, c_int64, _P]),
    "qv_minmax_reset": (c_int, [_P, c_int, _P]),
    "qv_minmax_accumulate": (c_int, [_P, c_int64, _P, _P]),
    "qv_obs_update": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_float, c_int32, c_int32, c_int32, _P]),
    "qv_fq_apply": (c_int, [_P, c_int64, _P, _P, _P, c_int32, c_int32, _P, _P, _P]),
    "qv_fq_weight": (c_int, [_P, c_int64, c_int64, c_int32, _P, _P, _P, _P, _P, _P, c_float, c_int32, c_int32,
                             c_int32, _P, _P, _P, _P, _P, _P, _P]),
    "qv_fq_weight_grouped": (c_int, [_P, c_int32, c_int32, c_int32, c_float, c_int32, c_int32, c_int32, _P]),
    "qv_fq_bwd": (c_int, [_P, _P, c_int64, _P, _P]),
    "qv_fq_learnable_fwd": (c_int, [_P, c_int64, c_int64, _P, _P, c_int32, c_int32, _P, _P]),
    "qv_fq_learnable_bwd": (c_int, [_P, _P, c_int64, c_int64, _P, _P, c_int32, c_int32, c_float, _P, _P, _P, _P]),
    "qv_split_planes": (c_int, [_P, c_int64, _P, _P, _P]),
    "qv_split_planes_mix": (c_int, [_P, c_int64, c_int64, c_int32, _P, _P, _P]),
    "qv_kd_ce_loss": (c_int, [_P, _P, _P, c_int32, c_int32, c_float, c_float, c_float, _P, _P, c_int32, c_int32,
                              _P, _P, _P]),
    "qv_kd_ce_rows_workspace_floats": (c_int64, [c_int32]),
    "qv_kd_ce_loss_rows": (c_int, [_P, _P, _P, c_int32, c_int32, c_float, c_float, c_float, _P, _P, c_int32, c_int32,
                                   _P, _P, _P, _P]),
    "qv_gemm_bf16": (c_int, [POINTER(GemmArgs), _P]),
    "qv_splitk_reduce": (c_int, [_P, c_int32, c_int64, c_int64, _P, _P, _P, _P, c_int32, _P]),
    "qv_resid_ln_fwd": (c_int, [_P, _P, _P, _P, c_int32, c_int32, _P, _P, c_float, c_int64, c_int32, c_int64, _P, _P,
                                c_int64, _P, _P, _P, _P, c_int32, _P, c_int32, _P]),
    "qv_ln_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int32, c_int64, _P, _P, c_int32, _P, _P, _P, c_int32, c_int32, _P]),
    "qv_ln_bwd_gp": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int32, _P, _P, c_int32, _P, _P, _P, c_int32, c_int32,
                             _P, _P, _P, c_int32, c_int32, _P, _P, c_int64, _P, _P]),
    "qv_colsum_reduce": (c_int, [_P, c_int32, c_int64, _P, c_int32, _P]),
    "qv_colsum_rows": (c_int, [_P, c_int64, c_int64, c_int64, _P, c_int32, _P]),
    "qv_gp_planes": (c_int, [_P, _P, _P, _P, c_int32, c_int32, _P, c_int32, c_int32, c_int64, c_int64, c_int32, c_int32,
                             _P, c_int64, _P, c_int32, _P]),
    "qv_act_planes": (c_int, [_P, _P, _P, c_int32, c_int32, c_int32, c_int32, c_int64, _P, c_int64, _P]),
    "qv_embed_fwd": (c_int, [_P, _P, _P, c_int32, c_int32, _P, _P, c_int64, c_int32, c_int32, _P, _P]),
    "qv_im2col_fq": (c_int, [_P, _P, _P, c_int32, c_int32, c_int64, c_int32, c_int32, c_int32, _P, _P, _P]),
    "qv_softmax_planes": (c_int, [_P, c_int64, c_int64, c_int32, c_float, _P, c_int64, c_int64, _P]),
    "qv_attn_ds": (c_int, [_P, c_int64, c_int64, _P, c_int64, c_int64, c_int32, c_float, _P, c_int64, c_int64, _P]),
    "qv_attn_fwd": (c_int, [_P, c_int32, c_int64, c_int64, c_int32, c_int32, c_int32, c_float, _P, _P, _P, c_int64, c_int64,
                            _P, _P, c_int32, _P, c_int32, _P]),
    "qv_attn_bwd": (c_int, [_P, c_int64, _P, _P, c_int64, c_int64, _P, c_int64, c_int64, _P, c_int32, c_int32, c_int32, c_float,
                            _P, _P]),
    "qv_attn_bwd_gp": (c_int, [_P, c_int64, _P, _P, c_int64, c_int64, _P, c_int64, c_int64, _P, c_int32, c_int32, c_int32,
                               c_float, _P, _P, _P, c_int32, c_int32, _P, _P, c_int64, _P, _P]),
    "qv_clip_adamw": (c_int, [_P, _P, _P, _P, c_int64, _P, c_int32, c_float, c_float, c_float, c_float, c_float, c_float, c_float,
                              c_int64, _P, c_int32, _P]),
    "qv_int8_linear": (c_int, [_P, c_int64, c_int64, _P, _P, _P, c_int64, _P, c_int32, _P, _P, c_float, c_int32, c_int32, _P, _P,
                               _P]),
    "qv_int8_linear_codes": (c_int, [_P, c_int64, c_int64, _P, _P, _P, c_int64, _P, c_int32, _P, _P, c_float, c_int32, c_int32, _P, _P]),
    "qv_quantize_u8": (c_int, [_P, c_int64, _P, _P, _P, _P]),
    "qv_qparams_from_minmax": (c_int, [_P, c_int32, c_int32, _P, _P, _P]),
    "qv_im2col_u8": (c_int, [_P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P, _P]),
    "qv_resize_normalize_u8": (c_int, [_P, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, _P, _P, _P, _P, c_int32, _P, _P,
                                       _P, _P]),
    "qv_gelu_minmax": (c_int, [_P, c_int64, _P, _P, _P]),
    "qv_quantize_u8_dyn": (c_int, [_P, c_int64, _P, _P, _P, _P, _P]),
    "qv_codes_from_u8": (c_int, [_P, c_int64, c_int32, _P, _P]),
    "qv_ln_quantize_u8_dyn": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int32, _P, _P, _P, _P, _P]),
    "qv_gelu_u8_minmax": (c_int, [_P, c_int64, c_float, c_int32, _P, _P]),
    "qv_gelu_u8_requant": (c_int, [_P, c_int64, c_float, c_int32, _P, _P, _P, _P, _P]),
    "qv_head_fwd": (c_int, [_P, _P, _P, c_int32, c_int32, c_int32, _P, _P, _P]),
    "qv_head_bwd": (c_int, [_P, _P, _P, _P, c_int32, c_int32, c_int32, _P, _P, _P, c_int32, _P]),
}

_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(qatvit_b200 has no CPU / PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().qv_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"qatvit_b200 {what} failed (code {rc}): {msg}")
