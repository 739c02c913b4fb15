// qv_common.cuh -- shared device/host helpers for the qatvit_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
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

// ------------------------------------------------------------------------------------------------
// "mixed" operand format of the frozen teacher's Linears: an fp32 value x is carried as  s = x * 2^7  in three pieces
//   region 0: h16 = fp16(s)                                             (11 significant bits)
//   region 1: per 64-column block one 128-byte row = 64 x hi8 = fp8(s), then 64 x lo8 = e5m2(s - h16), the fp16 rounding residual
// so that  A.W = [h16(A).h16(W) + hi8(A).lo8(W) + lo8(A).hi8(W)] * 2^-14  to ~2^-16 per product: the two cross terms are 2^-12 of
// the main one and only need fp8 precision, which the tensor cores run at twice the fp16 rate (three bf16 hi/lo passes -> the
// cost of two).  One fixed power-of-two scale for everything (no per-tensor statistics, no multiplies besides x * 128): e5m2
// spans the fp16 range.  Activations: hi8 = e5m2 (|x| <= 448 before saturation, fp16 part |x| <= 511);  weights: hi8 = e4m3
// (one more bit; saturates at |w| = 3.5, far above any ViT weight).  Every product term carries 2^14.
// ------------------------------------------------------------------------------------------------
#define QV_MIX_ACT 0
#define QV_MIX_WGT 1
#define QV_MIX_SCALE 128.0f
#define QV_MIX_ACC_SCALE 6.103515625e-05f     /* 2^-14 */
// Range of the format: the e5m2 copy of an activation saturates at 57344 / 2^7 = 448 (its fp16 part at 511.75), the e4m3 copy
// of a weight at 448 / 2^7 = 3.5.  Producers of mixed ACTIVATION planes take an optional (flag, bit) pair and OR the bit into
// the flag when a value leaves the range (one atomic per warp, only then); the host routes that tensor to bf16 hi/lo planes.
#define QV_MIX_ACT_MAX 448.0f
#define QV_MIX_WGT_MAX 3.5f
// returns the packed fp16 pair (a0 low half, a1 high half); ph / pl = packed hi8 / lo8 pairs (a0 low byte)
template <int KIND>
__device__ __forceinline__ uint32_t qv_mix_split2(float a0, float a1, uint32_t& ph, uint32_t& pl) {
  const float s0 = a0 * QV_MIX_SCALE, s1 = a1 * QV_MIX_SCALE;
  uint32_t h;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(s1), "f"(s0));
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h));
  ph = __nv_cvt_float2_to_fp8x2(make_float2(s0, s1), __NV_SATFINITE, KIND == QV_MIX_ACT ? __NV_E5M2 : __NV_E4M3);
  pl = __nv_cvt_float2_to_fp8x2(make_float2(s0 - hf.x, s1 - hf.y), __NV_SATFINITE, __NV_E5M2);
  return h;
}

// exact-erf GELU to ~3e-7 absolute without erff's branches (Abramowitz-Stegun 7.1.26, |erf error| <= 1.5e-7): the frozen
// teacher's fc1 epilogue is instruction-bound (50 instructions per element with erff + the operand split).  erfc(z) = poly(t) e^{-z^2},
// t = 1 / (1 + p z);  x >= 0: gelu = x (1 - erfc / 2),  x < 0: gelu = x erfc / 2 (no cancellation in the tail).
__device__ __forceinline__ float qv_gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t, e;                                               // raw MUFU ops: 1 + p z >= 1 and e^{-z^2} may flush to zero
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  float pl = fmaf(1.061405429f * 0.5f, t, -1.453152027f * 0.5f);
  pl = fmaf(pl, t, 1.421413741f * 0.5f);
  pl = fmaf(pl, t, -0.284496736f * 0.5f);
  pl = fmaf(pl, t, 0.254829592f * 0.5f);
  const float half_erfc = pl * t * e;
  return x * (x >= 0.f ? 1.0f - half_erfc : half_erfc);
}

// ---- the same GELU + mixed-format split on PAIRS, for the teacher's fc1 epilogue (gemm_sm100.cu, EPI 1, out_fmt 1) ----
// The epilogue warps of that kernel are bound by the FP32 pipe's issue rate (~25 instructions per element against a 12-k-block
// tile of MMAs), so the pair form uses packed fma / mul (one issue slot for two elements), folds every constant factor (the
// format's x 128, the 1/2 of erfc / 2, |x| = |z| / c) into the polynomial's coefficients, takes gelu = relu(x) - |x| erfc / 2
// (no compare / select / 1 - ...), and forms the fp16 rounding residual with ONE mixed-precision fma per element.
__device__ __forceinline__ uint64_t qv2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void qv2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t qv2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t qv2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// (x0, x1) -> 2^7 * gelu(x0), 2^7 * gelu(x1)  (same approximation as qv_gelu_fast: Abramowitz-Stegun 7.1.26, MUFU rcp / ex2)
__device__ __forceinline__ uint64_t qv_gelu128_pair(uint64_t x2) {
  constexpr float C = 0.8493218002880191f;            // sqrt(log2 e) / sqrt 2:  (C x)^2 = x^2 log2(e) / 2
  constexpr float P = 0.3275911f / 1.2011224087864498f;
  constexpr float K = -64.0f / C;                     // - 2^7 / 2 / C: the polynomial then yields -2^7 |x| erfc(|x| / sqrt 2) / 2 per |z|
  float x0, x1, z0, z1, e0, e1, d0, d1, t0, t1;
  qv2_unpack(x2, x0, x1);
  const uint64_t z2 = qv2_mul(x2, qv2_pack(C, C));
  const uint64_t zz2 = qv2_mul(z2, z2);
  qv2_unpack(z2, z0, z1);
  qv2_unpack(zz2, e0, e1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-e0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-e1));
  const uint64_t az2 = qv2_pack(fabsf(z0), fabsf(z1));
  qv2_unpack(qv2_fma(az2, qv2_pack(P, P), qv2_pack(1.0f, 1.0f)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  const uint64_t t2 = qv2_pack(t0, t1);
  uint64_t pl = qv2_fma(qv2_pack(K * 1.061405429f, K * 1.061405429f), t2, qv2_pack(K * -1.453152027f, K * -1.453152027f));
  pl = qv2_fma(pl, t2, qv2_pack(K * 1.421413741f, K * 1.421413741f));
  pl = qv2_fma(pl, t2, qv2_pack(K * -0.284496736f, K * -0.284496736f));
  pl = qv2_fma(pl, t2, qv2_pack(K * 0.254829592f, K * 0.254829592f));
  const uint64_t u2 = qv2_mul(qv2_mul(qv2_mul(pl, t2), qv2_pack(e0, e1)), az2);      // - 2^7 |x| erfc / 2
  return qv2_fma(qv2_pack(fmaxf(x0, 0.f), fmaxf(x1, 0.f)), qv2_pack(QV_MIX_SCALE, QV_MIX_SCALE), u2);
}
// split of an ALREADY SCALED pair (s = x * 2^7): packed fp16 pair returned, ph / pl = packed hi8 / lo8 pairs (element 0 low byte)
template <int KIND>
__device__ __forceinline__ uint32_t qv_mix_split2_scaled(uint64_t s2, uint32_t& ph, uint32_t& pl) {
  float s0, s1, r0, r1;
  qv2_unpack(s2, s0, s1);
  uint32_t h;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(s1), "f"(s0));
  const unsigned short h0 = static_cast<unsigned short>(h & 0xffffu), h1 = static_cast<unsigned short>(h >> 16), m1 = 0xBC00;   // -1.0
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r0) : "h"(h0), "h"(m1), "f"(s0));      // s - fp16(s), exact
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r1) : "h"(h1), "h"(m1), "f"(s1));
  ph = __nv_cvt_float2_to_fp8x2(make_float2(s0, s1), __NV_SATFINITE, KIND == QV_MIX_ACT ? __NV_E5M2 : __NV_E4M3);
  pl = __nv_cvt_float2_to_fp8x2(make_float2(r0, r1), __NV_SATFINITE, __NV_E5M2);
  return h;
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
