// qv_common.cuh -- shared device/host helpers for the qatvit_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#include "../../include/qatvit_b200.h"

// ------------------------------------------------------------------------------------------------
// host-side error plumbing (thread-local message, negative return codes)
// ------------------------------------------------------------------------------------------------
int qv_set_error(int code, const char* fmt, ...);
int qv_check_launch(const char* what);

#define QV_REQUIRE(cond, code, ...)                   \
  do {                                                \
    if (!(cond)) return qv_set_error((code), __VA_ARGS__); \
  } while (0)

static inline bool qv_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ------------------------------------------------------------------------------------------------
// order-preserving float <-> uint32 encoding for atomicMin/atomicMax on floats
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t qv_f2ord(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float qv_ord2f(uint32_t u) {
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
// Identity elements of the encoded min / max reductions.
#define QV_ORD_MIN_INIT 0xffffffffu   // encoded +NaN-ish top: any real value is smaller
#define QV_ORD_MAX_INIT 0x00000000u

// ------------------------------------------------------------------------------------------------
// fake-quant element math: q = rint(x * inv) + zp ; clamp ; (q - zp) * scale
// All fp32, no FMA contraction (__fmul_rn / __fadd_rn keep ptxas from fusing).
// ------------------------------------------------------------------------------------------------
struct QvQParams {
  float scale;
  float inv;     // 1.0f / scale, fp32 division
  float zp;      // zero point as float
  float qmin;
  float qmax;
};

__device__ __forceinline__ QvQParams qv_load_qparams(const float* scale, const int32_t* zp, int qmin, int qmax) {
  QvQParams q;
  q.scale = __ldg(scale);
  q.inv = __fdiv_rn(1.0f, q.scale);
  q.zp = (float)__ldg(zp);
  q.qmin = (float)qmin;
  q.qmax = (float)qmax;
  return q;
}

// returns the fake-quantised value; *in_range = STE mask; centered code (q_clamped - zp) in *ccode
__device__ __forceinline__ float qv_fq(float x, const QvQParams& q, bool* in_range, float* ccode) {
  float r = __fadd_rn(rintf(__fmul_rn(x, q.inv)), q.zp);
  float c = fminf(fmaxf(r, q.qmin), q.qmax);
  if (in_range) *in_range = (q.qmin <= r) && (r <= q.qmax);
  float cc = __fsub_rn(c, q.zp);
  if (ccode) *ccode = cc;
  return __fmul_rn(cc, q.scale);
}

// exact-erf GELU (timm Mlp.act = nn.GELU()) and its derivative; shared so that every kernel evaluates the same expression
__device__ __forceinline__ float qv_gelu_fwd(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float qv_gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Column sums across the 32 lanes of a warp for 32 columns at once (lane r holds row r's 32 values in v): a transposing
// butterfly, 31 shuffles in total; returns the sum of column `lane` over all rows.  Fixed order -> deterministic.  Destroys v.
__device__ __forceinline__ float qv_warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int off = 16 >> s, n = 16 >> s;
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = up ? v[i] : v[i + n];
      const float keep = up ? v[i + n] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// split an fp32 value into bf16 hi + bf16 lo (x ~= hi + lo, relative error <= 2^-17)
__device__ __forceinline__ void qv_split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__device__ __forceinline__ float qv_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float qv_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float qv_warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

int qv_num_sms();
