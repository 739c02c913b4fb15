// fakequant.cu -- observer + fake-quant kernels (HBM-bound, vectorised, SM-count sized grids).
//
// Replaces the single ATen op behind FusedMovingAvgObsFakeQuantize.forward
// (torch/ao/quantization/fake_quantize.py:423-438 -> fused_moving_avg_obs_fake_quant), i.e. torch's
// aminmax + MovingAverageMinMax + ChooseQuantizationParamsKernelImpl + FakeQuantizeCore launches
// (SURVEY.md §2.4 K1-K6).  Arithmetic contract: the CPU path:
//   EMA fp32 mul-then-add, fbgemm ChooseQuantizationParams (float scale, double zero-point math),
//   q = rint(x * (1/s)) + zp, clamp, (q - zp) * s.
#include <stdio.h>

#include <algorithm>
#include <mutex>

#include "qv_common.cuh"
#include "qv_observer.cuh"

namespace {

// ---------------- min/max accumulation ----------------
__global__ void qv_minmax_reset_kernel(uint32_t* acc, int count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    acc[2 * i] = QV_ORD_MIN_INIT;
    acc[2 * i + 1] = QV_ORD_MAX_INIT;
  }
}

__device__ __forceinline__ void qv_block_minmax_commit(float mn, float mx, uint32_t* acc) {
  __shared__ float smn[32], smx[32];
  mn = qv_warp_min(mn);
  mx = qv_warp_max(mx);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) { smn[warp] = mn; smx[warp] = mx; }
  __syncthreads();
  if (warp == 0) {
    mn = lane < nw ? smn[lane] : INFINITY;
    mx = lane < nw ? smx[lane] : -INFINITY;
    mn = qv_warp_min(mn);
    mx = qv_warp_max(mx);
    if (lane == 0 && mn <= mx) {
      atomicMin(acc, qv_f2ord(mn));
      atomicMax(acc + 1, qv_f2ord(mx));
    }
  }
}

__global__ void __launch_bounds__(256) qv_minmax_kernel(const float* __restrict__ x, int64_t n, uint32_t* acc) {
  float mn = INFINITY, mx = -INFINITY;
  const int64_t n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  // 4 independent 16-byte loads in flight per thread
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 a = __ldg(x4 + i), b = __ldg(x4 + i + stride), c = __ldg(x4 + i + 2 * stride), d = __ldg(x4 + i + 3 * stride);
    mn = fminf(mn, fminf(fminf(fminf(a.x, a.y), fminf(a.z, a.w)), fminf(fminf(b.x, b.y), fminf(b.z, b.w))));
    mn = fminf(mn, fminf(fminf(fminf(c.x, c.y), fminf(c.z, c.w)), fminf(fminf(d.x, d.y), fminf(d.z, d.w))));
    mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w))));
    mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(c.x, c.y), fmaxf(c.z, c.w)), fmaxf(fmaxf(d.x, d.y), fmaxf(d.z, d.w))));
  }
  for (; i < n4; i += stride) {
    float4 a = __ldg(x4 + i);
    mn = fminf(mn, fminf(fminf(a.x, a.y), fminf(a.z, a.w)));
    mx = fmaxf(mx, fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    float v = x[(n4 << 2) + threadIdx.x];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  qv_block_minmax_commit(mn, mx, acc);
}

// ---------------- per-tensor observer update (1 thread) ----------------
__global__ void qv_obs_update_kernel(const uint32_t* acc, const int64_t* obs_on, const int64_t* fq_on, float* min_val,
                                     float* max_val, float* scale, int32_t* zp, float c, int qmin, int qmax,
                                     int symmetric) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  qv_observer_step(acc[0], acc[1], obs_on, fq_on, min_val, max_val, scale, zp, c, qmin, qmax, symmetric);
}

// ---------------- elementwise fake-quant ----------------
__global__ void __launch_bounds__(256) qv_fq_apply_kernel(const float* __restrict__ x, int64_t n,
                                                          const float* __restrict__ scale,
                                                          const int32_t* __restrict__ zp,
                                                          const int64_t* __restrict__ fq_on, int qmin, int qmax,
                                                          float* __restrict__ y, uint8_t* __restrict__ mask) {
  const bool on = (*fq_on != 0);
  const QvQParams q = qv_load_qparams(scale, zp, qmin, qmax);
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    float4 o;
    uchar4 m;
    if (on) {
      bool b0, b1, b2, b3;
      o.x = qv_fq(v.x, q, &b0, nullptr);
      o.y = qv_fq(v.y, q, &b1, nullptr);
      o.z = qv_fq(v.z, q, &b2, nullptr);
      o.w = qv_fq(v.w, q, &b3, nullptr);
      m = make_uchar4(b0, b1, b2, b3);
    } else {
      o = v;
      m = make_uchar4(1, 1, 1, 1);
    }
    reinterpret_cast<float4*>(y)[i] = o;
    if (mask) reinterpret_cast<uchar4*>(mask)[i] = m;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    bool b = true;
    y[i] = on ? qv_fq(x[i], q, &b, nullptr) : x[i];
    if (mask) mask[i] = b;
  }
}

__global__ void __launch_bounds__(256) qv_fq_bwd_kernel(const float* __restrict__ gy, const uint8_t* __restrict__ mask,
                                                        int64_t n, float* __restrict__ gx) {
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gy) + i);
    const uchar4 m = __ldg(reinterpret_cast<const uchar4*>(mask) + i);
    reinterpret_cast<float4*>(gx)[i] = make_float4(m.x ? g.x : 0.f, m.y ? g.y : 0.f, m.z ? g.z : 0.f, m.w ? g.w : 0.f);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    gx[i] = mask[i] ? gy[i] : 0.f;
  }
}

// ---------------- weight flavour: one block per row ----------------
// PER_CHANNEL: row min/max -> EMA -> qparams inside the block; else qparams were produced by
// qv_minmax_kernel + qv_obs_update_kernel and are read from element 0.
template <bool PER_CHANNEL>
__global__ void __launch_bounds__(128) qv_fq_weight_kernel(const float* __restrict__ w, int64_t rows, int64_t cols,
                                                           const int64_t* obs_on, const int64_t* fq_on, float* min_val,
                                                           float* max_val, float* scale, int32_t* zp, float c, int qmin,
                                                           int qmax, int symmetric, float* __restrict__ y,
                                                           uint8_t* __restrict__ mask, __nv_bfloat16* __restrict__ codes,
                                                           __nv_bfloat16* __restrict__ codes_t, float* __restrict__ scale_vec) {
  const int64_t r = blockIdx.x;
  const float* wr = w + r * cols;
  __shared__ float s_scale;
  __shared__ int32_t s_zp;
  __shared__ float smn[4], smx[4];
  const bool fq = (*fq_on != 0);
  if (PER_CHANNEL) {
    if (*obs_on != 0) {
      float mn = INFINITY, mx = -INFINITY;
      for (int64_t k = threadIdx.x; k < cols; k += blockDim.x) {
        const float v = __ldg(wr + k);
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
      }
      mn = qv_warp_min(mn);
      mx = qv_warp_max(mx);
      if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
      __syncthreads();
      if (threadIdx.x == 0) {
        mn = fminf(fminf(smn[0], smn[1]), fminf(smn[2], smn[3]));
        mx = fmaxf(fmaxf(smx[0], smx[1]), fmaxf(smx[2], smx[3]));
        const float rmin = qv_ema(min_val[r], mn, c), rmax = qv_ema(max_val[r], mx, c);
        min_val[r] = rmin;
        max_val[r] = rmax;
        if (fq) {
          float s;
          int32_t z;
          qv_choose_qparams(rmin, rmax, qmin, qmax, symmetric != 0, &s, &z);
          scale[r] = s;
          zp[r] = z;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) { s_scale = scale[r]; s_zp = zp[r]; }
  } else {
    if (threadIdx.x == 0) { s_scale = scale[0]; s_zp = zp[0]; }
  }
  __syncthreads();
  if (scale_vec && threadIdx.x == 0) scale_vec[r] = s_scale;      // per-row copy of the scale in use (GEMM epilogue vector)
  QvQParams q;
  q.scale = s_scale;
  q.inv = __fdiv_rn(1.0f, q.scale);
  q.zp = (float)s_zp;
  q.qmin = (float)qmin;
  q.qmax = (float)qmax;
  for (int64_t k = threadIdx.x; k < cols; k += blockDim.x) {
    const float v = __ldg(wr + k);
    bool in = true;
    float cc = 0.f;
    const float o = fq ? qv_fq(v, q, &in, &cc) : v;
    const int64_t idx = r * cols + k;
    if (y) y[idx] = o;
    if (mask) mask[idx] = in;
    if (codes) codes[idx] = __float2bfloat16_rn(cc);
    if (codes_t) codes_t[k * rows + r] = __float2bfloat16_rn(cc);
  }
}

// ---------------- fp32 -> bf16 hi/lo planes ----------------
__global__ void __launch_bounds__(256) qv_split_planes_kernel(const float* __restrict__ x, int64_t n,
                                                              __nv_bfloat16* __restrict__ hi,
                                                              __nv_bfloat16* __restrict__ lo) {
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    __nv_bfloat16 h[4], l[4];
    qv_split_bf16(v.x, h[0], l[0]);
    qv_split_bf16(v.y, h[1], l[1]);
    qv_split_bf16(v.z, h[2], l[2]);
    qv_split_bf16(v.w, h[3], l[3]);
    reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
    reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    qv_split_bf16(x[i], hi[i], lo[i]);
  }
}

// ---------------- fp32 -> mixed planes (fp16 | fp8 hi / residual blocks; see qv_common.cuh) ----------------
// one thread = 8 consecutive columns of one row
template <int KIND>
__global__ void __launch_bounds__(256) qv_split_planes_mix_kernel(const float* __restrict__ x, int64_t rows, int64_t cols,
                                                                  uint8_t* __restrict__ r0, uint8_t* __restrict__ r1) {
  const int64_t groups_per_row = cols >> 3;
  const int64_t total = rows * groups_per_row;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / groups_per_row, c = (i - r * groups_per_row) << 3;
    const float4 va = __ldg(reinterpret_cast<const float4*>(x + r * cols + c));
    const float4 vb = __ldg(reinterpret_cast<const float4*>(x + r * cols + c + 4));
    uint32_t ph[4], pl[4];
    uint4 h16;
    h16.x = qv_mix_split2<KIND>(va.x, va.y, ph[0], pl[0]);
    h16.y = qv_mix_split2<KIND>(va.z, va.w, ph[1], pl[1]);
    h16.z = qv_mix_split2<KIND>(vb.x, vb.y, ph[2], pl[2]);
    h16.w = qv_mix_split2<KIND>(vb.z, vb.w, ph[3], pl[3]);
    *reinterpret_cast<uint4*>(r0 + (r * cols + c) * 2) = h16;
    uint8_t* row1 = r1 + r * cols * 2 + (c >> 6) * 128 + (c & 63);
    *reinterpret_cast<uint2*>(row1) = make_uint2(ph[0] | (ph[1] << 16), ph[2] | (ph[3] << 16));
    *reinterpret_cast<uint2*>(row1 + 64) = make_uint2(pl[0] | (pl[1] << 16), pl[2] | (pl[3] << 16));
  }
}

inline int ew_blocks(int64_t n4) {
  const int sms = qv_num_sms();
  int64_t b = (n4 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sms > 0 ? sms : 1) * 8;   // 8 resident 256-thread CTAs per SM
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

extern "C" int qv_minmax_reset(uint32_t* acc, int count, void* stream) {
  QV_REQUIRE(acc && count > 0, QV_ERR_INVALID, "bad minmax_reset arguments");
  qv_minmax_reset_kernel<<<(count + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(acc, count);
  return qv_check_launch("qv_minmax_reset");
}

extern "C" int qv_minmax_accumulate(const float* x, int64_t n, uint32_t* acc, void* stream) {
  QV_REQUIRE(acc != nullptr && n >= 0, QV_ERR_INVALID, "bad minmax arguments");
  if (n == 0) return QV_OK;
  QV_REQUIRE(x && qv_aligned16(x), QV_ERR_INVALID, "x must be a 16-byte aligned device pointer");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  qv_minmax_kernel<<<ew_blocks((n >> 2) / 4 + 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, acc);
  return qv_check_launch("qv_minmax_accumulate");
}

extern "C" int qv_obs_update(const uint32_t* acc, const int64_t* observer_enabled, const int64_t* fake_quant_enabled,
                             float* min_val, float* max_val, float* scale, int32_t* zero_point, float averaging_const,
                             int32_t qmin, int32_t qmax, int32_t symmetric, void* stream) {
  QV_REQUIRE(acc && observer_enabled && fake_quant_enabled && min_val && max_val && scale && zero_point,
             QV_ERR_INVALID, "null observer state pointer");
  QV_REQUIRE(qmin < qmax, QV_ERR_INVALID, "qmin must be < qmax");
  qv_obs_update_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(acc, observer_enabled, fake_quant_enabled,
                                                                       min_val, max_val, scale, zero_point,
                                                                       averaging_const, qmin, qmax, symmetric);
  return qv_check_launch("qv_obs_update");
}

extern "C" int qv_fq_apply(const float* x, int64_t n, const float* scale, const int32_t* zero_point,
                           const int64_t* fake_quant_enabled, int32_t qmin, int32_t qmax, float* y, uint8_t* mask,
                           void* stream) {
  QV_REQUIRE(n >= 0 && scale && zero_point && fake_quant_enabled, QV_ERR_INVALID, "bad fq_apply arguments");
  if (n == 0) return QV_OK;
  QV_REQUIRE(x && y && qv_aligned16(x) && qv_aligned16(y), QV_ERR_INVALID, "x / y must be 16-byte aligned");
  QV_REQUIRE(mask == nullptr || (reinterpret_cast<uintptr_t>(mask) & 3) == 0, QV_ERR_INVALID, "mask must be 4-byte aligned");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  qv_fq_apply_kernel<<<ew_blocks(n >> 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, scale, zero_point,
                                                                                      fake_quant_enabled, qmin, qmax, y,
                                                                                      mask);
  return qv_check_launch("qv_fq_apply");
}

extern "C" int qv_fq_bwd(const float* gy, const uint8_t* mask, int64_t n, float* gx, void* stream) {
  QV_REQUIRE(n >= 0, QV_ERR_INVALID, "bad fq_bwd arguments");
  if (n == 0) return QV_OK;
  QV_REQUIRE(gy && gx && mask && qv_aligned16(gy) && qv_aligned16(gx), QV_ERR_INVALID, "gy / gx must be 16-byte aligned");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  qv_fq_bwd_kernel<<<ew_blocks(n >> 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(gy, mask, n, gx);
  return qv_check_launch("qv_fq_bwd");
}

// ------------------------------------------------------------------------------------------------
// Grouped per-channel weight fake-quant: ALL fake-quantised weights of the model in one launch (the reference runs the
// observer + fake-quant of each of its 50 weights as ~4 launches apiece).  A block takes FQW_ROWS consecutive output
// channels of one weight (descriptor table in device memory, built once by the host side): a warp owns two rows at a
// time -- row min/max (float4 reads), EMA + qparams by one lane, quantise on the second (L1-resident) read -> codes
// (row-major), STE mask, and the codes staged in shared memory so that the TRANSPOSED copy (the dgrad operand) leaves as
// 32-byte runs instead of one scattered 2-byte store per element.
// ------------------------------------------------------------------------------------------------
constexpr int FQW_ROWS = 16;

__global__ void __launch_bounds__(256) qv_fq_weight_grouped_kernel(const qv_fqw_desc* __restrict__ descs, int n_desc, float c,
                                                                   int qmin, int qmax, int symmetric) {
  extern __shared__ __nv_bfloat16 s_codes[];       // [FQW_ROWS][cols + 8] (row pitch keeps the column reads conflict-light)
  __shared__ int s_d;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_desc - 1;
    while (lo < hi) {                               // last descriptor whose first block is <= blockIdx.x
      const int mid = (lo + hi + 1) >> 1;
      if (descs[mid].block_start <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid - 1;
    }
    s_d = lo;
  }
  __syncthreads();
  const qv_fqw_desc d = descs[s_d];
  const int r0 = (static_cast<int>(blockIdx.x) - d.block_start) * FQW_ROWS;
  const int nrows = min(FQW_ROWS, d.rows - r0);
  const int cols = d.cols, pitch = cols + 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool obs = (*d.observer_enabled != 0), fq = (*d.fake_quant_enabled != 0);
  for (int rr = warp; rr < nrows; rr += 8) {
    const int r = r0 + rr;
    const float4* wr = reinterpret_cast<const float4*>(d.w + static_cast<int64_t>(r) * cols);
    float sc, zpf;
    if (obs) {
      float mn = INFINITY, mx = -INFINITY;
      for (int k = lane; k < cols / 4; k += 32) {
        const float4 v = __ldg(wr + k);
        mn = fminf(mn, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
        mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
      }
      mn = qv_warp_min(mn);
      mx = qv_warp_max(mx);
      if (lane == 0) {
        const float rmin = qv_ema(d.min_val[r], mn, c), rmax = qv_ema(d.max_val[r], mx, c);
        d.min_val[r] = rmin;
        d.max_val[r] = rmax;
        if (fq) {
          float s_;
          int32_t z_;
          qv_choose_qparams(rmin, rmax, qmin, qmax, symmetric != 0, &s_, &z_);
          d.scale[r] = s_;
          d.zero_point[r] = z_;
        }
      }
      __syncwarp();
    }
    if (lane == 0) { sc = d.scale[r]; zpf = static_cast<float>(d.zero_point[r]); }
    sc = __shfl_sync(0xffffffffu, sc, 0);
    zpf = __shfl_sync(0xffffffffu, zpf, 0);
    QvQParams q;
    q.scale = sc;
    q.inv = __fdiv_rn(1.0f, sc);
    q.zp = zpf;
    q.qmin = static_cast<float>(qmin);
    q.qmax = static_cast<float>(qmax);
    for (int k = lane; k < cols / 4; k += 32) {
      const float4 v = __ldg(wr + k);
      const float a[4] = {v.x, v.y, v.z, v.w};
      __nv_bfloat16 cb[4];
      uint8_t mb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bool in = true;
        float cc = a[j];                            // fake-quant disabled: the "code" plane carries the raw value (not used then)
        if (fq) qv_fq(a[j], q, &in, &cc);
        cb[j] = __float2bfloat16_rn(cc);
        mb[j] = in ? 1 : 0;
      }
      const int64_t idx = static_cast<int64_t>(r) * cols + 4 * k;
      *reinterpret_cast<uint2*>(d.codes + idx) = *reinterpret_cast<uint2*>(cb);
      *reinterpret_cast<uchar4*>(d.mask + idx) = make_uchar4(mb[0], mb[1], mb[2], mb[3]);
      *reinterpret_cast<uint2*>(s_codes + rr * pitch + 4 * k) = *reinterpret_cast<uint2*>(cb);
    }
  }
  __syncthreads();
  // transposed copy: codes_t[k][r0 .. r0 + nrows) -- 32 contiguous bytes per k when the block is full
  if (nrows == FQW_ROWS) {
    for (int k = threadIdx.x; k < cols; k += blockDim.x) {
      __nv_bfloat16 v[FQW_ROWS];
#pragma unroll
      for (int rr = 0; rr < FQW_ROWS; ++rr) v[rr] = s_codes[rr * pitch + k];
      uint4* dst = reinterpret_cast<uint4*>(d.codes_t + static_cast<int64_t>(k) * d.rows + r0);
      dst[0] = *reinterpret_cast<uint4*>(v);
      dst[1] = *reinterpret_cast<uint4*>(v + 8);
    }
  } else {
    for (int k = threadIdx.x; k < cols; k += blockDim.x)
      for (int rr = 0; rr < nrows; ++rr)
        d.codes_t[static_cast<int64_t>(k) * d.rows + r0 + rr] = __bfloat16_as_ushort(s_codes[rr * pitch + k]);
  }
}

extern "C" int qv_fq_weight_grouped(const qv_fqw_desc* descs_device, int32_t n_desc, int32_t total_blocks, int32_t max_cols,
                                    float averaging_const, int32_t qmin, int32_t qmax, int32_t symmetric, void* stream) {
  QV_REQUIRE(descs_device && n_desc > 0 && total_blocks > 0 && max_cols > 0 && max_cols % 4 == 0, QV_ERR_INVALID,
             "bad fq_weight_grouped arguments (cols must be multiples of 4)");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  const size_t smem = static_cast<size_t>(FQW_ROWS) * (max_cols + 8) * sizeof(__nv_bfloat16);
  QV_REQUIRE(smem <= 200 * 1024, QV_ERR_UNSUPPORTED, "fq_weight_grouped: rows of %d elements do not fit the staging tile", max_cols);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(qv_fq_weight_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  QV_REQUIRE(attr_err == cudaSuccess, QV_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  qv_fq_weight_grouped_kernel<<<static_cast<unsigned>(total_blocks), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      descs_device, n_desc, averaging_const, qmin, qmax, symmetric);
  return qv_check_launch("qv_fq_weight_grouped");
}

extern "C" int qv_fq_weight(const float* w, int64_t rows, int64_t cols, int32_t per_channel,
                            const int64_t* observer_enabled, const int64_t* fake_quant_enabled, float* min_val,
                            float* max_val, float* scale, int32_t* zero_point, float averaging_const, int32_t qmin,
                            int32_t qmax, int32_t symmetric, float* y, uint8_t* mask, uint16_t* codes, uint16_t* codes_t,
                            uint32_t* scratch, float* scale_vec, void* stream) {
  QV_REQUIRE(w && rows > 0 && cols > 0, QV_ERR_INVALID, "bad weight shape");
  QV_REQUIRE(observer_enabled && fake_quant_enabled && min_val && max_val && scale && zero_point, QV_ERR_INVALID,
             "null observer state pointer");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(codes);
  __nv_bfloat16* ct = reinterpret_cast<__nv_bfloat16*>(codes_t);
  if (per_channel) {
    qv_fq_weight_kernel<true><<<static_cast<unsigned>(rows), 128, 0, st>>>(w, rows, cols, observer_enabled,
                                                                          fake_quant_enabled, min_val, max_val, scale,
                                                                          zero_point, averaging_const, qmin, qmax,
                                                                          symmetric, y, mask, c, ct, scale_vec);
    return qv_check_launch("qv_fq_weight");
  }
  QV_REQUIRE(scratch != nullptr, QV_ERR_INVALID, "per-tensor weight fake-quant needs a uint32[2] scratch");
  int rc = qv_minmax_reset(scratch, 1, stream);
  if (rc) return rc;
  rc = qv_minmax_accumulate(w, rows * cols, scratch, stream);
  if (rc) return rc;
  rc = qv_obs_update(scratch, observer_enabled, fake_quant_enabled, min_val, max_val, scale, zero_point, averaging_const,
                     qmin, qmax, symmetric, stream);
  if (rc) return rc;
  qv_fq_weight_kernel<false><<<static_cast<unsigned>(rows), 128, 0, st>>>(w, rows, cols, observer_enabled,
                                                                         fake_quant_enabled, min_val, max_val, scale,
                                                                         zero_point, averaging_const, qmin, qmax,
                                                                         symmetric, y, mask, c, ct, scale_vec);
  return qv_check_launch("qv_fq_weight");
}

// ------------------------------------------------------------------------------------------------
// Learnable per-channel fake-quant (torch._fake_quantize_learnable_per_channel_affine, channel axis 0 of x[rows][cols]): the
// op behind torch/ao/quantization/_learnable_fake_quantize.py:158-196.  Not on the reference's path (its scales are buffers,
// SURVEY.md 0.10) -- the opt-in counterpart of north_star's "per-channel scale gradient done as a warp-shuffle reduction".
// Arithmetic: oracle/fq_oracle.c qo_fq_learnable_fwd / _bwd (note the backward rounds AFTER adding the zero point, the forward
// before: they differ on ties, as in ATen).  One WARP per channel: lanes stride the row with 16-byte loads, per-lane partial sums
// in a fixed order, then an xor-shuffle butterfly -- no atomics, no second pass, bit-reproducible run to run.
// ------------------------------------------------------------------------------------------------
struct QvLearnQ { float s, inv, zr, qmin, qmax; };
__device__ __forceinline__ QvLearnQ qv_learn_q(const float* scale, const float* zp, int64_t r, int qmin, int qmax) {
  QvLearnQ q;
  q.s = __ldg(scale + r);
  q.inv = __fdiv_rn(1.0f, q.s);
  q.qmin = (float)qmin;
  q.qmax = (float)qmax;
  q.zr = fminf(fmaxf(rintf(__ldg(zp + r)), q.qmin), q.qmax);
  return q;
}
__device__ __forceinline__ float qv_learn_fwd1(float x, const QvLearnQ& q) {
  const float v = fminf(fmaxf(__fadd_rn(q.zr, rintf(__fmul_rn(x, q.inv))), q.qmin), q.qmax);
  return __fmul_rn(__fsub_rn(v, q.zr), q.s);
}
// one element of the backward: returns dx; adds this element's terms to ds / dz
__device__ __forceinline__ float qv_learn_bwd1(float g, float x, const QvLearnQ& q, float gf, float& ds, float& dz) {
  const float xq = rintf(__fadd_rn(q.zr, __fmul_rn(x, q.inv)));
  const bool in = (xq >= q.qmin) && (xq <= q.qmax);
  if (in) {
    const float xfq = __fmul_rn(__fsub_rn(xq, q.zr), q.s);
    ds += __fmul_rn(__fmul_rn(__fmul_rn(g, __fsub_rn(xfq, x)), q.inv), gf);
  } else {
    const float edge = __fsub_rn(xq < q.qmin ? q.qmin : q.qmax, q.zr);
    ds += __fmul_rn(__fmul_rn(g, edge), gf);
    dz += __fmul_rn(__fmul_rn(__fmul_rn(g, -1.0f), q.s), gf);
  }
  return in ? g : 0.0f;
}

__global__ void __launch_bounds__(256) qv_fq_learnable_fwd_kernel(const float* __restrict__ x, int64_t rows, int64_t cols,
                                                                  const float* __restrict__ scale, const float* __restrict__ zp,
                                                                  int qmin, int qmax, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const bool vec = (cols & 3) == 0;
  for (int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const QvLearnQ q = qv_learn_q(scale, zp, r, qmin, qmax);
    const float* xr = x + r * cols;
    float* yr = y + r * cols;
    if (vec) {
      for (int64_t i = lane; i < (cols >> 2); i += 32) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(xr) + i);
        reinterpret_cast<float4*>(yr)[i] = make_float4(qv_learn_fwd1(v.x, q), qv_learn_fwd1(v.y, q), qv_learn_fwd1(v.z, q),
                                                       qv_learn_fwd1(v.w, q));
      }
    } else {
      for (int64_t i = lane; i < cols; i += 32) yr[i] = qv_learn_fwd1(__ldg(xr + i), q);
    }
  }
}

__global__ void __launch_bounds__(128) qv_fq_learnable_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                                                  int64_t rows, int64_t cols, const float* __restrict__ scale,
                                                                  const float* __restrict__ zp, int qmin, int qmax, float gf,
                                                                  float* __restrict__ dx, float* __restrict__ dscale,
                                                                  float* __restrict__ dzp) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const bool vec = (cols & 3) == 0;
  for (int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const QvLearnQ q = qv_learn_q(scale, zp, r, qmin, qmax);
    const float* xr = x + r * cols;
    const float* gr = gy + r * cols;
    float* dr = dx ? dx + r * cols : nullptr;
    float ds = 0.f, dz = 0.f;
    if (vec) {
      const int64_t n4 = cols >> 2;
      for (int64_t i = lane; i < n4; i += 64) {          // two independent 16-byte load pairs in flight per lane
        const bool two = i + 32 < n4;
        const float4 xa = __ldg(reinterpret_cast<const float4*>(xr) + i);
        const float4 ga = __ldg(reinterpret_cast<const float4*>(gr) + i);
        float4 xb = make_float4(0.f, 0.f, 0.f, 0.f), gb = xb;
        if (two) {
          xb = __ldg(reinterpret_cast<const float4*>(xr) + i + 32);
          gb = __ldg(reinterpret_cast<const float4*>(gr) + i + 32);
        }
        float4 o;
        o.x = qv_learn_bwd1(ga.x, xa.x, q, gf, ds, dz);
        o.y = qv_learn_bwd1(ga.y, xa.y, q, gf, ds, dz);
        o.z = qv_learn_bwd1(ga.z, xa.z, q, gf, ds, dz);
        o.w = qv_learn_bwd1(ga.w, xa.w, q, gf, ds, dz);
        if (dr) reinterpret_cast<float4*>(dr)[i] = o;
        if (two) {
          o.x = qv_learn_bwd1(gb.x, xb.x, q, gf, ds, dz);
          o.y = qv_learn_bwd1(gb.y, xb.y, q, gf, ds, dz);
          o.z = qv_learn_bwd1(gb.z, xb.z, q, gf, ds, dz);
          o.w = qv_learn_bwd1(gb.w, xb.w, q, gf, ds, dz);
          if (dr) reinterpret_cast<float4*>(dr)[i + 32] = o;
        }
      }
    } else {
      for (int64_t i = lane; i < cols; i += 32) {
        const float o = qv_learn_bwd1(__ldg(gr + i), __ldg(xr + i), q, gf, ds, dz);
        if (dr) dr[i] = o;
      }
    }
    ds = qv_warp_sum(ds);
    dz = qv_warp_sum(dz);
    if (lane == 0) {
      if (dscale) dscale[r] = ds;
      if (dzp) dzp[r] = dz;
    }
  }
}

extern "C" int qv_fq_learnable_fwd(const float* x, int64_t rows, int64_t cols, const float* scale, const float* zero_point,
                                   int32_t qmin, int32_t qmax, float* y, void* stream) {
  QV_REQUIRE(rows >= 0 && cols >= 0 && qmin < qmax, QV_ERR_INVALID, "bad fq_learnable_fwd arguments");
  if (rows == 0 || cols == 0) return QV_OK;
  QV_REQUIRE(x && y && scale && zero_point, QV_ERR_INVALID, "null pointer");
  QV_REQUIRE((cols & 3) != 0 || (qv_aligned16(x) && qv_aligned16(y)), QV_ERR_INVALID, "x / y must be 16-byte aligned");
  const int sms = qv_num_sms();
  QV_REQUIRE(sms > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  const int64_t blocks = std::min<int64_t>((rows + 7) / 8, static_cast<int64_t>(sms) * 8);
  qv_fq_learnable_fwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, rows, cols, scale, zero_point, qmin, qmax, y);
  return qv_check_launch("qv_fq_learnable_fwd");
}

extern "C" int qv_fq_learnable_bwd(const float* gy, const float* x, int64_t rows, int64_t cols, const float* scale,
                                   const float* zero_point, int32_t qmin, int32_t qmax, float grad_factor, float* dx,
                                   float* dscale, float* dzero_point, void* stream) {
  QV_REQUIRE(rows >= 0 && cols >= 0 && qmin < qmax, QV_ERR_INVALID, "bad fq_learnable_bwd arguments");
  if (rows == 0) return QV_OK;
  QV_REQUIRE(scale && zero_point && (dx || dscale || dzero_point), QV_ERR_INVALID, "null pointer");
  QV_REQUIRE(cols == 0 || (gy && x), QV_ERR_INVALID, "null pointer");
  QV_REQUIRE((cols & 3) != 0 || (qv_aligned16(gy) && qv_aligned16(x) && qv_aligned16(dx)), QV_ERR_INVALID,
             "gy / x / dx must be 16-byte aligned");
  const int sms = qv_num_sms();
  QV_REQUIRE(sms > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  const int64_t blocks = std::min<int64_t>((rows + 3) / 4, static_cast<int64_t>(sms) * 16);
  qv_fq_learnable_bwd_kernel<<<static_cast<unsigned>(blocks), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      gy, x, rows, cols, scale, zero_point, qmin, qmax, grad_factor, dx, dscale, dzero_point);
  return qv_check_launch("qv_fq_learnable_bwd");
}

extern "C" int qv_split_planes(const float* x, int64_t n, uint16_t* hi, uint16_t* lo, void* stream) {
  QV_REQUIRE(n >= 0, QV_ERR_INVALID, "bad split_planes arguments");
  if (n == 0) return QV_OK;
  QV_REQUIRE(x && hi && lo && qv_aligned16(x) && (reinterpret_cast<uintptr_t>(hi) & 7) == 0 &&
                 (reinterpret_cast<uintptr_t>(lo) & 7) == 0,
             QV_ERR_INVALID, "split_planes pointers must be aligned (x:16, planes:8)");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  qv_split_planes_kernel<<<ew_blocks(n >> 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, n, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo));
  return qv_check_launch("qv_split_planes");
}

extern "C" int qv_split_planes_mix(const float* x, int64_t rows, int64_t cols, int32_t kind, uint16_t* region0, uint16_t* region1,
                                   void* stream) {
  QV_REQUIRE(rows >= 0 && cols >= 0 && (kind == QV_MIX_ACT || kind == QV_MIX_WGT), QV_ERR_INVALID, "bad split_planes_mix arguments");
  if (rows == 0 || cols == 0) return QV_OK;
  QV_REQUIRE(cols % 64 == 0, QV_ERR_UNSUPPORTED, "mixed planes need a multiple of 64 columns (got %lld)", (long long)cols);
  QV_REQUIRE(x && region0 && region1 && qv_aligned16(x) && qv_aligned16(region0) && qv_aligned16(region1), QV_ERR_INVALID,
             "split_planes_mix pointers must be 16-byte aligned");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  const int blocks = ew_blocks(rows * (cols >> 3));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (kind == QV_MIX_ACT)
    qv_split_planes_mix_kernel<QV_MIX_ACT><<<blocks, 256, 0, st>>>(x, rows, cols, reinterpret_cast<uint8_t*>(region0),
                                                                   reinterpret_cast<uint8_t*>(region1));
  else
    qv_split_planes_mix_kernel<QV_MIX_WGT><<<blocks, 256, 0, st>>>(x, rows, cols, reinterpret_cast<uint8_t*>(region0),
                                                                   reinterpret_cast<uint8_t*>(region1));
  return qv_check_launch("qv_split_planes_mix");
}
