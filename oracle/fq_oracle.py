"""oracle/fq_oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/fq_oracle.c (the plain-C restatement of torch's CPU
``fused_moving_avg_obs_fake_quant`` arithmetic) with numpy in/out.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfq_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile fq_oracle.c with gcc (no FMA contraction)."""
    src = os.path.join(_HERE, "fq_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_build/libfq_oracle.so"])
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.qo_choose_qparams.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_int32, ctypes.c_int32,
                                           ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _lib.qo_choose_qparams.restype = None
        _lib.qo_calculate_qparams.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_int32, ctypes.c_int32,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _lib.qo_calculate_qparams.restype = None
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def choose_qparams(mn: float, mx: float, qmin: int, qmax: int, symmetric: bool):
    s = np.zeros(1, np.float32)
    z = np.zeros(1, np.int32)
    lib().qo_choose_qparams(ctypes.c_float(mn), ctypes.c_float(mx), qmin, qmax, int(symmetric), _p(s), _p(z))
    return float(s[0]), int(z[0])


def calculate_qparams(mn: float, mx: float, qmin: int, qmax: int, symmetric: bool, unsigned: bool = False):
    s = np.zeros(1, np.float32)
    z = np.zeros(1, np.int32)
    lib().qo_calculate_qparams(ctypes.c_float(mn), ctypes.c_float(mx), qmin, qmax, int(symmetric), int(unsigned),
                               _p(s), _p(z))
    return float(s[0]), int(z[0])


@dataclass
class FQState:
    """Mirror of the buffers a FusedMovingAvgObsFakeQuantize module owns."""
    qmin: int
    qmax: int
    symmetric: bool
    channels: int = 0                      # 0 => per-tensor
    averaging_constant: float = 0.01
    observer_enabled: int = 1
    fake_quant_enabled: int = 1
    min_val: np.ndarray = field(default=None)
    max_val: np.ndarray = field(default=None)
    scale: np.ndarray = field(default=None)
    zero_point: np.ndarray = field(default=None)

    def __post_init__(self):
        n = max(self.channels, 1)
        if self.min_val is None:
            self.min_val = np.full(n, np.inf, np.float32)
            self.max_val = np.full(n, -np.inf, np.float32)
            self.scale = np.ones(n, np.float32)
            self.zero_point = np.zeros(n, np.int32)


def fused_obs_fq(x: np.ndarray, st: FQState, want_codes: bool = True):
    """Run one forward of the fused observer + fake-quant; mutates ``st``; returns (y, mask, codes)."""
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    mask = np.empty(x.shape, np.uint8)
    codes = np.empty(x.shape, np.int32) if want_codes else None
    L = lib()
    if st.channels:
        C = st.channels
        assert x.shape[0] == C
        inner = x.size // C
        L.qo_fused_obs_fq_per_channel(_p(x), ctypes.c_int64(C), ctypes.c_int64(inner),
                                      ctypes.c_int64(st.observer_enabled), ctypes.c_int64(st.fake_quant_enabled),
                                      _p(st.min_val), _p(st.max_val), _p(st.scale), _p(st.zero_point),
                                      ctypes.c_float(st.averaging_constant), st.qmin, st.qmax, int(st.symmetric),
                                      _p(y), _p(mask), _p(codes))
    else:
        L.qo_fused_obs_fq_per_tensor(_p(x), ctypes.c_int64(x.size), ctypes.c_int64(st.observer_enabled),
                                     ctypes.c_int64(st.fake_quant_enabled), _p(st.min_val), _p(st.max_val),
                                     _p(st.scale), _p(st.zero_point), ctypes.c_float(st.averaging_constant),
                                     st.qmin, st.qmax, int(st.symmetric), _p(y), _p(mask), _p(codes))
    if not st.fake_quant_enabled and codes is not None:
        codes[...] = 0
    return y, mask, codes


def fq_bwd(gy: np.ndarray, mask: np.ndarray) -> np.ndarray:
    gy = np.ascontiguousarray(gy, np.float32)
    mask = np.ascontiguousarray(mask, np.uint8)
    gx = np.empty_like(gy)
    lib().qo_fq_bwd(_p(gy), _p(mask), ctypes.c_int64(gy.size), _p(gx))
    return gx


def fq_learnable_fwd(x: np.ndarray, scale: np.ndarray, zero_point: np.ndarray, qmin: int, qmax: int) -> np.ndarray:
    """torch._fake_quantize_learnable_per_channel_affine forward, channel axis 0 (qo_fq_learnable_fwd)."""
    x = np.ascontiguousarray(x, np.float32)
    scale = np.ascontiguousarray(scale, np.float32)
    zero_point = np.ascontiguousarray(zero_point, np.float32)
    C = x.shape[0]
    y = np.empty_like(x)
    lib().qo_fq_learnable_fwd(_p(x), ctypes.c_int64(C), ctypes.c_int64(x.size // max(C, 1)), _p(scale), _p(zero_point),
                              int(qmin), int(qmax), _p(y))
    return y


def fq_learnable_bwd(gy: np.ndarray, x: np.ndarray, scale: np.ndarray, zero_point: np.ndarray, qmin: int, qmax: int,
                     grad_factor: float):
    """-> (dx fp32, dscale float64 [C], dzero_point float64 [C], magnitude float64 [C]) of the learnable per-channel
    fake-quant (qo_fq_learnable_bwd).  magnitude = sum |g| (|xq - zr| + 1) grad_factor: what fp32 rounding acts on (see the C file)."""
    gy = np.ascontiguousarray(gy, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    scale = np.ascontiguousarray(scale, np.float32)
    zero_point = np.ascontiguousarray(zero_point, np.float32)
    C = x.shape[0]
    dx = np.empty_like(x)
    ds, dz, da = np.zeros(C, np.float64), np.zeros(C, np.float64), np.zeros(C, np.float64)
    lib().qo_fq_learnable_bwd(_p(gy), _p(x), ctypes.c_int64(C), ctypes.c_int64(x.size // max(C, 1)), _p(scale), _p(zero_point),
                              int(qmin), int(qmax), ctypes.c_float(grad_factor), _p(dx), _p(ds), _p(dz), _p(da))
    return dx, ds, dz, da


def distill_loss(s: np.ndarray, t: np.ndarray, labels: np.ndarray, T: float, alpha: float, eps: float):
    """Returns (loss, loss_kd*T^2, loss_ce, dL/ds) -- ref qat_trainer.py:343-349."""
    s = np.ascontiguousarray(s, np.float32)
    t = np.ascontiguousarray(t, np.float32)
    labels = np.ascontiguousarray(labels, np.int64)
    B, C = s.shape
    out = np.zeros(3, np.float32)
    g = np.empty_like(s)
    lib().qo_distill_loss(_p(s), _p(t), _p(labels), ctypes.c_int64(B), ctypes.c_int64(C), ctypes.c_float(T),
                          ctypes.c_float(alpha), ctypes.c_float(eps), _p(out[0:1]), _p(out[1:2]), _p(out[2:3]), _p(g))
    return float(out[0]), float(out[1]), float(out[2]), g


def int8_linear(qx, sx, zx, qw, sw, bias, sy, zy, engine="x86"):
    """quantized::linear restatement; qx uint8 [M,K], qw int8 [N,K].  engine: "x86"/"fbgemm" (float bias,
    the default CPU engine) or "qnnpack" (int32-quantised bias)."""
    qx = np.ascontiguousarray(qx, np.uint8)
    qw = np.ascontiguousarray(qw, np.int8)
    sw = np.ascontiguousarray(np.atleast_1d(sw), np.float32)
    bias = None if bias is None else np.ascontiguousarray(bias, np.float32)
    M, K = qx.shape
    N = qw.shape[0]
    qy = np.empty((M, N), np.uint8)
    lib().qo_int8_linear(_p(qx), ctypes.c_int64(M), ctypes.c_int64(K), ctypes.c_float(sx), ctypes.c_int32(zx),
                         _p(qw), ctypes.c_int64(N), _p(sw), int(sw.size > 1), _p(bias), ctypes.c_float(sy),
                         ctypes.c_int32(zy), int(engine == "qnnpack"), _p(qy))
    return qy
