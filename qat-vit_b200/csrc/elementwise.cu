// elementwise.cu -- HBM-bound row / elementwise kernels around the fake-quant GEMMs.
//
// The reference keeps every one of these as a separate ATen launch with fp32 tensors in between
// (SURVEY.md §2.4 K4/K6/K9/K11/K12).  Here the activation fake-quant is applied ON LOAD by the consumer
// (residual add, LayerNorm, GELU, attention pre-pass, loss) from the raw GEMM output and the step's
// (scale, zero_point); the STE mask is recomputed in backward from the same raw tensor, so no mask or
// fake-quantised copy ever goes to HBM.  Tensors that feed a tensor-core GEMM are written directly as
// bf16 hi/lo planes.
#include "qv_common.cuh"

namespace {

struct OptQ {          // optional per-tensor fake-quant applied on load
  bool on;
  QvQParams q;
};

__device__ __forceinline__ OptQ load_optq(const float* scale, const int32_t* zp, int qmin, int qmax) {
  OptQ o;
  o.on = (scale != nullptr);
  if (o.on) o.q = qv_load_qparams(scale, zp, qmin, qmax);
  return o;
}

__device__ __forceinline__ float gelu_fwd(float x) { return qv_gelu_fwd(x); }
__device__ __forceinline__ float gelu_grad(float x) { return qv_gelu_grad(x); }

__device__ __forceinline__ void store_planes4(__nv_bfloat16* hi, __nv_bfloat16* lo, int64_t idx, float a, float b,
                                              float c, float d) {
  __nv_bfloat16 h[4], l[4];
  qv_split_bf16(a, h[0], l[0]);
  qv_split_bf16(b, h[1], l[1]);
  qv_split_bf16(c, h[2], l[2]);
  qv_split_bf16(d, h[3], l[3]);
  *reinterpret_cast<uint2*>(hi + idx) = *reinterpret_cast<uint2*>(h);
  *reinterpret_cast<uint2*>(lo + idx) = *reinterpret_cast<uint2*>(l);
}

// ------------------------------------------------------------------------------------------------
// x_out = x_in + FQ(y_raw);  h = LayerNorm(x_out) -> bf16 hi/lo planes (and/or fp32);  one warp per row.
// VPL = float4 vectors per lane (D = 128 * VPL).  Persistent: the grid is a fixed number of blocks per SM, a warp
// walks rows warp_id, warp_id + n_warps, ... and issues the NEXT row's loads before it reduces the current one, so
// every warp always has a full row (2 x VPL x 512 B) in flight while it computes.
// ------------------------------------------------------------------------------------------------
template <int VPL>
__device__ __forceinline__ void resid_ln_load(const float* __restrict__ x_in, const float* __restrict__ y_raw, int64_t src, int lane,
                                              float4 (&xa)[VPL], float4 (&ya)[VPL]) {
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    xa[i] = x_in ? __ldg(reinterpret_cast<const float4*>(x_in + src + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (y_raw) ya[i] = __ldg(reinterpret_cast<const float4*>(y_raw + src + c));
  }
}

template <int VPL>
__global__ void __launch_bounds__(256, (VPL <= 4 ? 3 : 2)) resid_ln_fwd_kernel(
    const float* __restrict__ x_in, const float* __restrict__ y_raw, const float* y_scale, const int32_t* y_zp, int qmin, int qmax,
    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int64_t R, int64_t in_row_stride,
    float* __restrict__ x_out, __nv_bfloat16* __restrict__ h_planes, int64_t plane_stride, float* __restrict__ h_f32,
    float* __restrict__ mean_out, float* __restrict__ rstd_out, uint32_t* minmax, int plane_fmt, int32_t* sat_flag, int32_t sat_bit) {
  constexpr int D = 128 * VPL;
  const int lane = threadIdx.x & 31;
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  __shared__ float s_mn[8], s_mx[8];
  float omn = INFINITY, omx = -INFINITY;
  const OptQ oq = load_optq(y_scale, y_zp, qmin, qmax);
  float4 xa[VPL], ya[VPL];
  if (r < R) resid_ln_load<VPL>(x_in, y_raw, r * in_row_stride * D, lane, xa, ya);
  while (r < R) {
    const int64_t rn = r + n_warps;
    float4 v[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float4 a = xa[i];
      if (y_raw) {
        float4 y = ya[i];
        if (oq.on) {
          y.x = qv_fq(y.x, oq.q, nullptr, nullptr);
          y.y = qv_fq(y.y, oq.q, nullptr, nullptr);
          y.z = qv_fq(y.z, oq.q, nullptr, nullptr);
          y.w = qv_fq(y.w, oq.q, nullptr, nullptr);
        }
        a.x += y.x; a.y += y.y; a.z += y.z; a.w += y.w;
      }
      v[i] = a;
    }
    if (rn < R) resid_ln_load<VPL>(x_in, y_raw, rn * in_row_stride * D, lane, xa, ya);   // next row in flight
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = qv_warp_sum(s) * (1.0f / D);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      ss += (a * a + b * b) + (c * c + d * d);
    }
    const float var = qv_warp_sum(ss) * (1.0f / D);
    const float rstd = rsqrtf(var + eps);
    if (lane == 0) {
      if (mean_out) mean_out[r] = mean;
      if (rstd_out) rstd_out[r] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (x_out) *reinterpret_cast<float4*>(x_out + r * D + c) = v[i];
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      const float o0 = (v[i].x - mean) * rstd * g.x + b.x;
      const float o1 = (v[i].y - mean) * rstd * g.y + b.y;
      const float o2 = (v[i].z - mean) * rstd * g.z + b.z;
      const float o3 = (v[i].w - mean) * rstd * g.w + b.w;
      if (h_planes) {
        if (plane_fmt == 1) {           // mixed operand format: fp16 | per 64-column block 64 hi8 + 64 lo8 (qv_common.cuh)
          uint32_t ph0, pl0, ph1, pl1;
          const uint32_t h0 = qv_mix_split2<QV_MIX_ACT>(o0, o1, ph0, pl0), h1 = qv_mix_split2<QV_MIX_ACT>(o2, o3, ph1, pl1);
          *reinterpret_cast<uint2*>(h_planes + r * D + c) = make_uint2(h0, h1);
          uint8_t* row1 = reinterpret_cast<uint8_t*>(h_planes + plane_stride + r * D) + (c >> 6) * 128 + (c & 63);
          *reinterpret_cast<uint32_t*>(row1) = ph0 | (ph1 << 16);
          *reinterpret_cast<uint32_t*>(row1 + 64) = pl0 | (pl1 << 16);
        } else {
          store_planes4(h_planes, h_planes + plane_stride, r * D + c, o0, o1, o2, o3);
        }
      }
      if (h_f32) *reinterpret_cast<float4*>(h_f32 + r * D + c) = make_float4(o0, o1, o2, o3);
      omn = fminf(omn, fminf(fminf(o0, o1), fminf(o2, o3)));
      omx = fmaxf(omx, fmaxf(fmaxf(o0, o1), fmaxf(o2, o3)));
    }
    r = rn;
  }
  if (sat_flag && plane_fmt == 1) {   // mixed format range guard: |h| beyond what fp8 e5m2(h * 2^7) holds -> raise the caller's flag
    const bool over = fmaxf(-omn, omx) > QV_MIX_ACT_MAX;
    if (__any_sync(0xffffffffu, over) && lane == 0) atomicOr(sat_flag, sat_bit);
  }
  if (minmax) {                       // observed-LayerNorm variant: min / max of the LN output, one atomic pair per block
    omn = qv_warp_min(omn);
    omx = qv_warp_max(omx);
    if (lane == 0) { s_mn[threadIdx.x >> 5] = omn; s_mx[threadIdx.x >> 5] = omx; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) { omn = fminf(omn, s_mn[w]); omx = fmaxf(omx, s_mx[w]); }
      if (omn <= omx) {
        atomicMin(minmax, qv_f2ord(omn));
        atomicMax(minmax + 1, qv_f2ord(omx));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward fused with the residual-gradient add.  One warp per row, a block walks
// `rows_per_block` rows and emits partial dgamma / dbeta column sums (deterministic two-stage reduce).
//   g_x[r*out_row_stride] = (g_res ? g_res[r] : 0) + rstd * (gy - mean(gy) - xhat * mean(gy * xhat)),  gy = g_h * gamma
// ------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(256, 2) ln_bwd_kernel(const float* __restrict__ g_h, const float* __restrict__ x,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, const float* __restrict__ g_res,
                                                     int64_t R, int64_t out_row_stride, float* __restrict__ g_x,
                                                     float* __restrict__ partials, int rows_per_block,
                                                     const float* __restrict__ h_raw, const float* h_scale,
                                                     const int32_t* h_zp, int qmin, int qmax,
                                                     const float* __restrict__ gp_y, const float* gp_scale, const int32_t* gp_zp,
                                                     int gp_qmin, int gp_qmax, const float* __restrict__ gp_wscale,
                                                     __nv_bfloat16* __restrict__ gp_out, int64_t gp_plane_stride,
                                                     float* __restrict__ gp_partials) {
  constexpr int D = 128 * VPL;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const OptQ hq = load_optq(h_raw ? h_scale : nullptr, h_zp, qmin, qmax);   // observed LN output: g_h passes its STE mask
  // fused gradient planes of the Linear whose (fake-quantised) output y was added into this residual stream:
  // gp = g_x * STEmask(y) * w_scale -> hi/lo planes, per-block column sums of g_x * mask (its bias grad)
  const OptQ gq_ = load_optq(gp_y ? gp_scale : nullptr, gp_zp, gp_qmin, gp_qmax);
  float4 dg[VPL], db[VPL], dbias[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) dg[i] = db[i] = dbias[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  for (int rr = warp; rr < rows_per_block; rr += nwarps) {
    const int64_t r = r0 + rr;
    if (r >= R) break;
    // every global read of the row is issued before the first use (gamma / w_scale are L1-resident re-reads, which keeps
    // them out of the persistent register state): 4-5 independent 16-byte loads per lane and vector in flight
    float4 gh[VPL], xh[VPL], gr[VPL], yv[VPL], hv[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      gh[i] = __ldg(reinterpret_cast<const float4*>(g_h + r * D + c));
      xh[i] = __ldg(reinterpret_cast<const float4*>(x + r * D + c));
      if (g_res) gr[i] = __ldg(reinterpret_cast<const float4*>(g_res + r * D + c));
      if (gq_.on) yv[i] = __ldg(reinterpret_cast<const float4*>(gp_y + r * D + c));
      if (hq.on) hv[i] = __ldg(reinterpret_cast<const float4*>(h_raw + r * D + c));
    }
    const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (hq.on) {
        bool in0, in1, in2, in3;
        qv_fq(hv[i].x, hq.q, &in0, nullptr); qv_fq(hv[i].y, hq.q, &in1, nullptr);
        qv_fq(hv[i].z, hq.q, &in2, nullptr); qv_fq(hv[i].w, hq.q, &in3, nullptr);
        gh[i].x = in0 ? gh[i].x : 0.f; gh[i].y = in1 ? gh[i].y : 0.f; gh[i].z = in2 ? gh[i].z : 0.f; gh[i].w = in3 ? gh[i].w : 0.f;
      }
      const float4 xv = xh[i];
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      dg[i].x += gh[i].x * xh[i].x; dg[i].y += gh[i].y * xh[i].y; dg[i].z += gh[i].z * xh[i].z; dg[i].w += gh[i].w * xh[i].w;
      db[i].x += gh[i].x; db[i].y += gh[i].y; db[i].z += gh[i].z; db[i].w += gh[i].w;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
      gh[i].x *= gm.x; gh[i].y *= gm.y; gh[i].z *= gm.z; gh[i].w *= gm.w;
      s1 += (gh[i].x + gh[i].y) + (gh[i].z + gh[i].w);
      s2 += (gh[i].x * xh[i].x + gh[i].y * xh[i].y) + (gh[i].z * xh[i].z + gh[i].w * xh[i].w);
    }
    s1 = qv_warp_sum(s1) * (1.0f / D);
    s2 = qv_warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      float4 o = make_float4(rs * (gh[i].x - s1 - xh[i].x * s2), rs * (gh[i].y - s1 - xh[i].y * s2),
                             rs * (gh[i].z - s1 - xh[i].z * s2), rs * (gh[i].w - s1 - xh[i].w * s2));
      if (g_res) { o.x += gr[i].x; o.y += gr[i].y; o.z += gr[i].z; o.w += gr[i].w; }
      *reinterpret_cast<float4*>(g_x + r * out_row_stride * D + c) = o;
      if (gq_.on) {
        bool in0, in1, in2, in3;
        qv_fq(yv[i].x, gq_.q, &in0, nullptr); qv_fq(yv[i].y, gq_.q, &in1, nullptr);
        qv_fq(yv[i].z, gq_.q, &in2, nullptr); qv_fq(yv[i].w, gq_.q, &in3, nullptr);
        o.x = in0 ? o.x : 0.f; o.y = in1 ? o.y : 0.f; o.z = in2 ? o.z : 0.f; o.w = in3 ? o.w : 0.f;
        dbias[i].x += o.x; dbias[i].y += o.y; dbias[i].z += o.z; dbias[i].w += o.w;
        const float4 ws = __ldg(reinterpret_cast<const float4*>(gp_wscale + c));
        store_planes4(gp_out, gp_out + gp_plane_stride, r * D + c, o.x * ws.x, o.y * ws.y, o.z * ws.z, o.w * ws.w);
      }
    }
  }
  // block reduce of the per-warp column partials (fixed order)
  __shared__ float4 sm[8][2][VPL * 32];
  if (partials) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) { sm[warp][0][i * 32 + lane] = dg[i]; sm[warp][1][i * 32 + lane] = db[i]; }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 2 * VPL * 32; idx += blockDim.x) {
      const int which = idx / (VPL * 32), c4 = idx % (VPL * 32);
      float4 a = sm[0][which][c4];
      for (int w = 1; w < nwarps; ++w) {
        const float4 b = sm[w][which][c4];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      *reinterpret_cast<float4*>(partials + (static_cast<int64_t>(blockIdx.x) * 2 + which) * D + c4 * 4) = a;
    }
  }
  if (gq_.on && gp_partials) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VPL; ++i) sm[warp][0][i * 32 + lane] = dbias[i];
    __syncthreads();
    for (int c4 = threadIdx.x; c4 < VPL * 32; c4 += blockDim.x) {
      float4 a = sm[0][0][c4];
      for (int w = 1; w < nwarps; ++w) {
        const float4 b = sm[w][0][c4];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      *reinterpret_cast<float4*>(gp_partials + static_cast<int64_t>(blockIdx.x) * D + c4 * 4) = a;
    }
  }
}

// out[c] (+)= sum_b partials[b][c].  Block = 32 columns (8 lanes x float4) x 128 row slices: every thread's loads are
// independent 16-byte reads (all in flight at once), rows are read as full 128-byte lines; fixed-order combine in shared
// memory so the result is deterministic.  ncols % 4 == 0 takes the vector path.
__global__ void __launch_bounds__(1024) colsum_reduce_kernel(const float* __restrict__ partials, int nblk, int64_t ncols,
                                                             float* __restrict__ out, int accumulate, int vec) {
  __shared__ float sm[128][33];
  const int c4 = threadIdx.x & 7, slice = threadIdx.x >> 3;
  const int64_t c = blockIdx.x * 32LL + c4 * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (vec) {
    if (c < ncols) {
#pragma unroll 4
      for (int b = slice; b < nblk; b += 128) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(partials + static_cast<int64_t>(b) * ncols + c));
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
  } else {
    for (int b = slice; b < nblk; b += 128) {
      const float* row = partials + static_cast<int64_t>(b) * ncols;
      if (c + 0 < ncols) s.x += __ldg(row + c + 0);
      if (c + 1 < ncols) s.y += __ldg(row + c + 1);
      if (c + 2 < ncols) s.z += __ldg(row + c + 2);
      if (c + 3 < ncols) s.w += __ldg(row + c + 3);
    }
  }
  sm[slice][c4 * 4 + 0] = s.x; sm[slice][c4 * 4 + 1] = s.y; sm[slice][c4 * 4 + 2] = s.z; sm[slice][c4 * 4 + 3] = s.w;
  __syncthreads();
  // 128 slices -> 4 partial sums per column (threads 0..127), then one thread per column adds the four in order
  if (threadIdx.x < 128) {
    const int col = threadIdx.x & 31, part = threadIdx.x >> 5;
    float t = 0.f;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) t += sm[part * 32 + k][col];
    sm[part * 32][col] = t;         // slot (part*32, col) was read only by this thread
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int64_t cc = blockIdx.x * 32LL + threadIdx.x;
    if (cc < ncols) {
      const float t = ((sm[0][threadIdx.x] + sm[32][threadIdx.x]) + sm[64][threadIdx.x]) + sm[96][threadIdx.x];
      out[cc] = accumulate ? out[cc] + t : t;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Backward "gradient planes":  gp'[r,n] = g[r',n] * [gelu'(FQ(y))] * mask(y[r,n]) * w_scale[n]  -> bf16 hi/lo planes
// and per-block partial column sums of the UNSCALED masked gradient (bias grad).
// Each thread owns 4 consecutive columns (blockDim.x <= 256 threads, blockIdx.y walks the column groups); a block walks
// rows_per_block rows.
// Row remap (patch embed): output row r = b*P + i reads g row b*T + i + 1 (drops the cls token).
// ------------------------------------------------------------------------------------------------
template <int U>   // rows in flight per thread: 2 U independent 16-byte loads before the first use
__global__ void __launch_bounds__(256) gp_planes_kernel(const float* __restrict__ g, const float* __restrict__ y_raw, const float* y_scale,
                                 const int32_t* y_zp, int qmin, int qmax, const float* __restrict__ w_scale,
                                 int w_scale_stride, int gelu, int64_t R, int64_t N, int remap_P, int remap_T,
                                 __nv_bfloat16* __restrict__ out, int64_t plane_stride, float* __restrict__ partials,
                                 int rows_per_block) {
  const int64_t c = (static_cast<int64_t>(blockIdx.y) * blockDim.x + threadIdx.x) * 4;
  if (c >= N) return;
  const OptQ oq = load_optq(y_scale, y_zp, qmin, qmax);
  float4 ws = make_float4(1.f, 1.f, 1.f, 1.f);
  if (w_scale) {
    if (w_scale_stride) ws = __ldg(reinterpret_cast<const float4*>(w_scale + c));
    else { const float s = __ldg(w_scale); ws = make_float4(s, s, s, s); }
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = min(R, r0 + rows_per_block);
  for (int64_t rb = r0; rb < r1; rb += U) {
    float4 gv[U], yv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = rb + u;
      if (r < r1) {
        const int64_t gr = remap_P ? (r / remap_P) * remap_T + (r % remap_P) + 1 : r;
        gv[u] = __ldg(reinterpret_cast<const float4*>(g + gr * N + c));
        if (y_raw) yv[u] = __ldg(reinterpret_cast<const float4*>(y_raw + r * N + c));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = rb + u;
      if (r >= r1) break;
      float4 gq = gv[u];
      if (y_raw) {
        float yq[4] = {yv[u].x, yv[u].y, yv[u].z, yv[u].w};
        float gg[4] = {gq.x, gq.y, gq.z, gq.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          bool in = true;
          float v = yq[j];
          if (oq.on) v = qv_fq(v, oq.q, &in, nullptr);
          float f = gg[j];
          if (gelu) f *= gelu_grad(v);
          gg[j] = in ? f : 0.f;
        }
        gq = make_float4(gg[0], gg[1], gg[2], gg[3]);
      }
      acc.x += gq.x; acc.y += gq.y; acc.z += gq.z; acc.w += gq.w;
      store_planes4(out, out + plane_stride, r * N + c, gq.x * ws.x, gq.y * ws.y, gq.z * ws.z, gq.w * ws.w);
    }
  }
  if (partials) *reinterpret_cast<float4*>(partials + static_cast<int64_t>(blockIdx.x) * N + c) = acc;
}

// ------------------------------------------------------------------------------------------------
// out planes = [GELU]( FQ(y_raw) )   (fc1 -> fc2 operand; qkv pre-pass with gelu = 0; teacher without FQ)
// ------------------------------------------------------------------------------------------------
// codes_only: write ONE plane holding the centred integer code (q - zp), exact in bf16 (|code| <= 255): the operand format
// of the integer-code attention kernels (FQ(x) = code * scale, the scale is applied by the consumer).
__global__ void __launch_bounds__(256) act_planes_kernel(const float* __restrict__ y_raw, const float* y_scale,
                                                         const int32_t* y_zp, int qmin, int qmax, int gelu, int codes_only,
                                                         int64_t n, __nv_bfloat16* __restrict__ out, int64_t plane_stride) {
  const OptQ oq = load_optq(y_scale, y_zp, qmin, qmax);
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  // GELU(FQ(y)) takes at most qmax - qmin + 1 <= 256 distinct values: a per-block table of packed (hi | lo << 16) bf16 pairs
  // indexed by the clamped code replaces erff on every element (same expression evaluated once per code -> same bits).
  __shared__ uint32_t lut[256];
  const bool use_lut = gelu && oq.on && !codes_only && (qmax - qmin) < 256;
  if (use_lut) {
    for (int k = threadIdx.x; k <= qmax - qmin; k += blockDim.x) {
      const float v = gelu_fwd(__fmul_rn(__fsub_rn(static_cast<float>(qmin + k), oq.q.zp), oq.q.scale));
      __nv_bfloat16 h, l;
      qv_split_bf16(v, h, l);
      lut[k] = static_cast<uint32_t>(__bfloat16_as_ushort(h)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l)) << 16);
    }
    __syncthreads();
  }
  auto emit = [&](int64_t i, const float4& v) {
    float a[4] = {v.x, v.y, v.z, v.w};
    if (use_lut) {
      uint32_t e[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float r = __fadd_rn(rintf(__fmul_rn(a[j], oq.q.inv)), oq.q.zp);
        e[j] = lut[static_cast<int>(fminf(fmaxf(r, oq.q.qmin), oq.q.qmax) - oq.q.qmin)];
      }
      *reinterpret_cast<uint2*>(out + i * 4) = make_uint2(__byte_perm(e[0], e[1], 0x5410), __byte_perm(e[2], e[3], 0x5410));
      *reinterpret_cast<uint2*>(out + plane_stride + i * 4) = make_uint2(__byte_perm(e[0], e[1], 0x7632), __byte_perm(e[2], e[3], 0x7632));
    } else if (codes_only) {
      __nv_bfloat16 c[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float cc;
        qv_fq(a[j], oq.q, nullptr, &cc);
        c[j] = __float2bfloat16_rn(cc);
      }
      *reinterpret_cast<uint2*>(out + i * 4) = *reinterpret_cast<uint2*>(c);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (oq.on) a[j] = qv_fq(a[j], oq.q, nullptr, nullptr);
        if (gelu) a[j] = gelu_fwd(a[j]);
      }
      store_planes4(out, out + plane_stride, i * 4, a[0], a[1], a[2], a[3]);
    }
  };
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {                 // four independent 16-byte loads in flight per thread
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(y_raw) + i);
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(y_raw) + i + stride);
    const float4 v2 = __ldg(reinterpret_cast<const float4*>(y_raw) + i + 2 * stride);
    const float4 v3 = __ldg(reinterpret_cast<const float4*>(y_raw) + i + 3 * stride);
    emit(i, v0); emit(i + stride, v1); emit(i + 2 * stride, v2); emit(i + 3 * stride, v3);
  }
  for (; i < n4; i += stride) emit(i, __ldg(reinterpret_cast<const float4*>(y_raw) + i));
}

// ------------------------------------------------------------------------------------------------
// x0[b, 0, :] = cls + pos[0];  x0[b, 1+i, :] = FQ(p_raw[b*P+i, :]) + pos[1+i]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_fwd_kernel(const float* __restrict__ p_raw, const float* p_scale,
                                                        const int32_t* p_zp, int qmin, int qmax,
                                                        const float* __restrict__ cls, const float* __restrict__ pos,
                                                        int64_t B, int P, int D, float* __restrict__ x0) {
  const OptQ oq = load_optq(p_scale, p_zp, qmin, qmax);
  const int T = P + 1;
  const int d4 = D >> 2;
  const int64_t total = B * T * d4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += stride) {
    const int c4 = static_cast<int>(i % d4);
    const int64_t row = i / d4;
    const int t = static_cast<int>(row % T);
    const int64_t b = row / T;
    const float4 pe = __ldg(reinterpret_cast<const float4*>(pos) + static_cast<int64_t>(t) * d4 + c4);
    float4 v;
    if (t == 0) {
      v = __ldg(reinterpret_cast<const float4*>(cls) + c4);
    } else {
      v = __ldg(reinterpret_cast<const float4*>(p_raw) + (b * P + (t - 1)) * d4 + c4);
      if (oq.on) {
        v.x = qv_fq(v.x, oq.q, nullptr, nullptr);
        v.y = qv_fq(v.y, oq.q, nullptr, nullptr);
        v.z = qv_fq(v.z, oq.q, nullptr, nullptr);
        v.w = qv_fq(v.w, oq.q, nullptr, nullptr);
      }
    }
    reinterpret_cast<float4*>(x0)[i] = make_float4(v.x + pe.x, v.y + pe.y, v.z + pe.z, v.w + pe.w);
  }
}

// ------------------------------------------------------------------------------------------------
// im2col of the fake-quantised input image for the 16x16/16 patch-embed conv:
//   out[b*P + py*G + px][c*ps*ps + ky*ps + kx] = bf16( clamp(rint(x*inv)+zp) - zp )   (exact integer)
// One thread per 4 consecutive kx (float4 load along the image row, 8-byte store).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col_fq_kernel(const float* __restrict__ img, const float* scale,
                                                        const int32_t* zp, int qmin, int qmax, int64_t B, int C, int HW,
                                                        int ps, __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out_lo, int fq_on) {
  QvQParams q;
  if (fq_on) q = qv_load_qparams(scale, zp, qmin, qmax);
  const int G = HW / ps;
  const int w4 = HW >> 2;
  const int64_t total = B * C * HW * w4;
  const int64_t K = static_cast<int64_t>(C) * ps * ps;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += stride) {
    const int xw = static_cast<int>(i % w4) * 4;
    int64_t rest = i / w4;
    const int yh = static_cast<int>(rest % HW);
    rest /= HW;
    const int c = static_cast<int>(rest % C);
    const int64_t b = rest / C;
    const float4 v = __ldg(reinterpret_cast<const float4*>(img) + i);
    float a[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 o[4], ol[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float cc = a[j];
      if (fq_on) qv_fq(a[j], q, nullptr, &cc);
      qv_split_bf16(cc, o[j], ol[j]);     // integer codes: lo == 0 and is not stored
    }
    const int py = yh / ps, ky = yh % ps, px = xw / ps, kx = xw % ps;
    const int64_t row = b * G * G + static_cast<int64_t>(py) * G + px;
    const int64_t col = static_cast<int64_t>(c) * ps * ps + ky * ps + kx;
    *reinterpret_cast<uint2*>(out + row * K + col) = *reinterpret_cast<uint2*>(o);
    if (out_lo) *reinterpret_cast<uint2*>(out_lo + row * K + col) = *reinterpret_cast<uint2*>(ol);
  }
}

// ------------------------------------------------------------------------------------------------
// softmax over the first T entries of each score row (scaled), written as bf16 hi/lo planes.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_planes_kernel(const float* __restrict__ S, int64_t ldS, int64_t rows, int T,
                                                             float scale, __nv_bfloat16* __restrict__ P, int64_t ldP,
                                                             int64_t plane_stride) {
  const int lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* s = S + r * ldS;
  float v[8];                                  // T <= 256
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int j = i * 32 + lane;
    v[i] = j < T ? __ldg(s + j) * scale : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  mx = qv_warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int j = i * 32 + lane;
    v[i] = j < T ? expf(v[i] - mx) : 0.f;
    sum += v[i];
  }
  sum = qv_warp_sum(sum);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int j = i * 32 + lane;
    if (j < ldP) {
      __nv_bfloat16 h, l;
      qv_split_bf16(j < T ? v[i] * inv : 0.f, h, l);
      P[r * ldP + j] = h;
      P[plane_stride + r * ldP + j] = l;
    }
  }
}

// dS = P * (dP - sum_j dP*P) * scale  -> planes  (softmax backward, one warp per row)
__global__ void __launch_bounds__(256) attn_ds_kernel(const __nv_bfloat16* __restrict__ P, int64_t ldP, int64_t p_plane_stride,
                                                      const float* __restrict__ dP, int64_t lddP, int64_t rows, int T,
                                                      float scale, __nv_bfloat16* __restrict__ dS, int64_t ldS,
                                                      int64_t s_plane_stride) {
  const int lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float p[8], d[8];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int j = i * 32 + lane;
    if (j < T) {
      p[i] = __bfloat162float(P[r * ldP + j]) + __bfloat162float(P[p_plane_stride + r * ldP + j]);
      d[i] = __ldg(dP + r * lddP + j);
    } else {
      p[i] = 0.f;
      d[i] = 0.f;
    }
    dot += p[i] * d[i];
  }
  dot = qv_warp_sum(dot);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int j = i * 32 + lane;
    if (j < ldS) {
      __nv_bfloat16 h, l;
      qv_split_bf16(j < T ? p[i] * (d[i] - dot) * scale : 0.f, h, l);
      dS[r * ldS + j] = h;
      dS[s_plane_stride + r * ldS + j] = l;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// classifier head (384 -> 10): tiny, exact fp32 FMA.  One warp per (b, n) output.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wq,
                                                       const float* __restrict__ bias, int B, int K, int N,
                                                       float* __restrict__ out, uint32_t* minmax) {
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= B * N) return;
  const int b = o / N, n = o % N;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s = fmaf(x[static_cast<int64_t>(b) * K + k], wq[static_cast<int64_t>(n) * K + k], s);
  s = qv_warp_sum(s);
  if (lane == 0) {
    s += bias ? bias[n] : 0.f;
    out[o] = s;
    if (minmax) {
      atomicMin(minmax, qv_f2ord(s));
      atomicMax(minmax + 1, qv_f2ord(s));
    }
  }
}

// gx[b,k] = sum_n g[b,n] wq[n,k];  gw[n,k] (+)= mask * sum_b g[b,n] x[b,k];  gb[n] (+)= sum_b g[b,n]
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                       const float* __restrict__ wq, const uint8_t* __restrict__ wmask,
                                                       int B, int K, int N, float* __restrict__ gx, float* __restrict__ gw,
                                                       float* __restrict__ gb, int accumulate) {
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t n_gx = static_cast<int64_t>(B) * K, n_gw = static_cast<int64_t>(N) * K;
  if (tid < n_gx) {
    const int b = static_cast<int>(tid / K), k = static_cast<int>(tid % K);
    float s = 0.f;
    for (int n = 0; n < N; ++n) s = fmaf(g[b * N + n], wq[static_cast<int64_t>(n) * K + k], s);
    gx[tid] = s;
  } else if (tid < n_gx + n_gw) {
    const int64_t i = tid - n_gx;
    const int n = static_cast<int>(i / K), k = static_cast<int>(i % K);
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(g[b * N + n], x[static_cast<int64_t>(b) * K + k], s);
    if (wmask && !wmask[i]) s = 0.f;
    gw[i] = accumulate ? gw[i] + s : s;
  } else if (tid < n_gx + n_gw + N) {
    const int n = static_cast<int>(tid - n_gx - n_gw);
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += g[b * N + n];
    gb[n] = accumulate ? gb[n] + s : s;
  }
}

// out[c] (+)= sum_r x[r][c]   (thread per column; for many columns / few rows: pos_embed & cls grads)
__global__ void __launch_bounds__(256) colsum_rows_kernel(const float* __restrict__ x, int64_t R, int64_t N, int64_t ld,
                                                          float* __restrict__ out, int accumulate) {
  const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (c >= N) return;
  float s = 0.f;
  for (int64_t r = 0; r < R; ++r) s += __ldg(x + r * ld + c);
  out[c] = accumulate ? out[c] + s : s;
}

// im2col of the QUANTISED input image (converted model: nnq.Quantize followed by nnq.Conv2d 16x16/16 as an int8 GEMM):
//   out[b*P + py*G + px][c*ps*ps + ky*ps + kx] = clamp(rint(x * (1/s)) + z, 0, 255)   (uint8)
__global__ void __launch_bounds__(256) im2col_u8_kernel(const float* __restrict__ img, const float* scale, const int32_t* zp,
                                                        int64_t B, int C, int HW, int ps, uint8_t* __restrict__ out) {
  const float inv = __fdiv_rn(1.0f, __ldg(scale));
  const float z = static_cast<float>(__ldg(zp));
  const int G = HW / ps;
  const int w4 = HW >> 2;
  const int64_t total = B * C * HW * w4;
  const int64_t K = static_cast<int64_t>(C) * ps * ps;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += stride) {
    const int xw = static_cast<int>(i % w4) * 4;
    int64_t rest = i / w4;
    const int yh = static_cast<int>(rest % HW);
    rest /= HW;
    const int c = static_cast<int>(rest % C);
    const int64_t b = rest / C;
    const float4 v = __ldg(reinterpret_cast<const float4*>(img) + i);
    const float a[4] = {v.x, v.y, v.z, v.w};
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f = __fadd_rn(nearbyintf(__fmul_rn(a[j], inv)), z);
      f = fminf(fmaxf(f, 0.0f), 255.0f);
      w |= static_cast<uint32_t>(f) << (8 * j);
    }
    const int py = yh / ps, ky = yh % ps, px = xw / ps, kx = xw % ps;
    const int64_t row = b * G * G + static_cast<int64_t>(py) * G + px;
    const int64_t col = static_cast<int64_t>(c) * ps * ps + ky * ps + kx;
    *reinterpret_cast<uint32_t*>(out + row * K + col) = w;
  }
}

// y = GELU(x) (exact erf), fp32 -> fp32, fused with the min / max of y (dynamic quantisation of the next int8 Linear)
__global__ void __launch_bounds__(256) gelu_minmax_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ y,
                                                          uint32_t* acc) {
  float mn = INFINITY, mx = -INFINITY;
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 o = make_float4(gelu_fwd(v.x), gelu_fwd(v.y), gelu_fwd(v.z), gelu_fwd(v.w));
    reinterpret_cast<float4*>(y)[i] = o;
    mn = fminf(mn, fminf(fminf(o.x, o.y), fminf(o.z, o.w)));
    mx = fmaxf(mx, fmaxf(fmaxf(o.x, o.y), fmaxf(o.z, o.w)));
  }
  if (acc) {
    mn = qv_warp_min(mn);
    mx = qv_warp_max(mx);
    if ((threadIdx.x & 31) == 0 && mn <= mx) {
      atomicMin(acc, qv_f2ord(mn));
      atomicMax(acc + 1, qv_f2ord(mx));
    }
  }
}

// Converted-student glue: q = quantize_u8(LayerNorm(x)) with the DYNAMIC qparams of the finished min / max accumulator, recomputing
// the LayerNorm output from the row statistics resid_ln_fwd_kernel saved instead of reading an fp32 copy of it (4 B/elt read + 1 written
// against 4 written + 4 read + 1 written).  Same arithmetic, operation for operation, as resid_ln_fwd_kernel's output expression
// (v - mean) * rstd * gamma + beta (subtract, multiply, fused multiply-add) and as qv_qparams_from_minmax + qv_quantize_u8, so the
// codes are bit-identical to the unfused chain (tests/test_int8_gpu.py).  One warp per row, persistent.
template <int VPL>
__global__ void __launch_bounds__(256) ln_quantize_u8_dyn_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                                 const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, int64_t R, const uint32_t* acc,
                                                                 float* scale_out, int32_t* zp_out, uint8_t* __restrict__ q) {
  constexpr int D = 128 * VPL;
  const int lane = threadIdx.x & 31;
  // torch/ao/quantization/observer.py:349-427 (affine quint8, 0..255), fp32 arithmetic -- as qv_qparams_from_minmax
  const float mn = fminf(qv_ord2f(acc[0]), 0.0f), mx = fmaxf(qv_ord2f(acc[1]), 0.0f);
  float s = __fdiv_rn(__fsub_rn(mx, mn), 255.0f);
  s = fmaxf(s, 1.1920928955078125e-07f);
  float z = __fsub_rn(0.0f, nearbyintf(__fdiv_rn(mn, s)));
  z = fminf(fmaxf(z, 0.0f), 255.0f);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *scale_out = s;
    *zp_out = static_cast<int32_t>(z);
  }
  const float inv = __fdiv_rn(1.0f, s);
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5); r < R; r += n_warps) {
    const float m = __ldg(mean + r), rs = __ldg(rstd + r);
    float4 v[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(x + r * D + (i * 32 + lane) * 4));
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      const float o[4] = {__fmaf_rn(__fmul_rn(__fsub_rn(v[i].x, m), rs), g.x, b.x), __fmaf_rn(__fmul_rn(__fsub_rn(v[i].y, m), rs), g.y, b.y),
                          __fmaf_rn(__fmul_rn(__fsub_rn(v[i].z, m), rs), g.z, b.z), __fmaf_rn(__fmul_rn(__fsub_rn(v[i].w, m), rs), g.w, b.w)};
      uint32_t w = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f = __fadd_rn(nearbyintf(__fmul_rn(o[j], inv)), z);
        f = fminf(fmaxf(f, 0.0f), 255.0f);
        w |= static_cast<uint32_t>(f) << (8 * j);
      }
      *reinterpret_cast<uint32_t*>(q + r * D + c) = w;
    }
  }
}

inline int ew_blocks(int64_t n_items, int per_sm = 8) {
  const int sms = qv_num_sms();
  int64_t b = (n_items + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sms > 0 ? sms : 1) * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

#define QV_NEED_GPU() QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)")

extern "C" int qv_resid_ln_fwd(const float* x_in, const float* y_raw, const float* y_scale, const int32_t* y_zp,
                               int32_t qmin, int32_t qmax, const float* gamma, const float* beta, float eps, int64_t R,
                               int32_t D, int64_t in_row_stride, float* x_out, uint16_t* h_planes, int64_t plane_stride,
                               float* h_f32, float* mean, float* rstd, uint32_t* minmax, int32_t plane_fmt, int32_t* sat_flag,
                               int32_t sat_bit, void* stream) {
  QV_REQUIRE((x_in || y_raw) && gamma && beta && R > 0, QV_ERR_INVALID, "bad resid_ln_fwd arguments");
  QV_REQUIRE(plane_fmt == 0 || plane_fmt == 1, QV_ERR_INVALID, "plane_fmt must be 0 (bf16 hi/lo) or 1 (mixed fp16 + fp8)");
  QV_REQUIRE(D % 128 == 0 && D >= 128 && D <= 1024, QV_ERR_UNSUPPORTED, "LayerNorm width must be a multiple of 128 <= 1024 (got %d)", D);
  QV_REQUIRE((y_scale == nullptr) == (y_zp == nullptr), QV_ERR_INVALID, "y_scale and y_zp go together");
  QV_NEED_GPU();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rows_per_block = 8;                       // one warp per row, a warp walks rows n_warps apart
  const int64_t blocks_needed = (R + rows_per_block - 1) / rows_per_block;
  const int64_t resident = static_cast<int64_t>(qv_num_sms()) * (D <= 512 ? 3 : 2);
  const unsigned grid = static_cast<unsigned>(blocks_needed < resident ? blocks_needed : resident);
  __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(h_planes);
  if (in_row_stride < 1) in_row_stride = 1;
#define LAUNCH(V)                                                                                                        \
  resid_ln_fwd_kernel<V><<<grid, 256, 0, st>>>(x_in, y_raw, y_scale, y_zp, qmin, qmax, gamma, beta, eps, R, in_row_stride, \
                                               x_out, hp, plane_stride, h_f32, mean, rstd, minmax, plane_fmt, sat_flag, sat_bit)
  switch (D / 128) {
    case 1: LAUNCH(1); break;
    case 2: LAUNCH(2); break;
    case 3: LAUNCH(3); break;
    case 4: LAUNCH(4); break;
    case 6: LAUNCH(6); break;
    case 8: LAUNCH(8); break;
    default: return qv_set_error(QV_ERR_UNSUPPORTED, "LayerNorm width %d not instantiated", D);
  }
#undef LAUNCH
  return qv_check_launch("qv_resid_ln_fwd");
}

namespace {
int ln_bwd_impl(const float* g_h, const float* x, const float* mean, const float* rstd, const float* gamma, const float* g_res,
                int64_t R, int32_t D, int64_t out_row_stride, float* g_x, float* partials, int32_t rows_per_block,
                const float* h_raw, const float* h_scale, const int32_t* h_zp, int32_t qmin, int32_t qmax, const float* gp_y,
                const float* gp_scale, const int32_t* gp_zp, int32_t gp_qmin, int32_t gp_qmax, const float* gp_wscale,
                uint16_t* gp_out, int64_t gp_plane_stride, float* gp_partials, void* stream) {
  QV_REQUIRE(g_h && x && mean && rstd && gamma && g_x && R > 0 && rows_per_block > 0, QV_ERR_INVALID, "bad ln_bwd arguments");
  QV_REQUIRE(D % 128 == 0 && D >= 128 && D <= 1024, QV_ERR_UNSUPPORTED, "LayerNorm width must be a multiple of 128 <= 1024 (got %d)", D);
  QV_REQUIRE(!h_raw || (h_scale && h_zp), QV_ERR_INVALID, "h_raw needs h_scale and h_zp");
  QV_NEED_GPU();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>((R + rows_per_block - 1) / rows_per_block);
  if (out_row_stride < 1) out_row_stride = 1;
#define LAUNCH(V) ln_bwd_kernel<V><<<grid, 256, 0, st>>>(g_h, x, mean, rstd, gamma, g_res, R, out_row_stride, g_x, partials, \
                                                         rows_per_block, h_raw, h_scale, h_zp, qmin, qmax, gp_y, gp_scale, gp_zp, \
                                                         gp_qmin, gp_qmax, gp_wscale, reinterpret_cast<__nv_bfloat16*>(gp_out),   \
                                                         gp_plane_stride, gp_partials)
  switch (D / 128) {
    case 1: LAUNCH(1); break;
    case 2: LAUNCH(2); break;
    case 3: LAUNCH(3); break;
    case 4: LAUNCH(4); break;
    case 6: LAUNCH(6); break;
    default: return qv_set_error(QV_ERR_UNSUPPORTED, "LayerNorm backward width %d not instantiated", D);
  }
#undef LAUNCH
  return qv_check_launch("qv_ln_bwd");
}
}  // namespace

extern "C" int qv_ln_bwd(const float* g_h, const float* x, const float* mean, const float* rstd, const float* gamma,
                         const float* g_res, int64_t R, int32_t D, int64_t out_row_stride, float* g_x, float* partials,
                         int32_t rows_per_block, const float* h_raw, const float* h_scale, const int32_t* h_zp, int32_t qmin,
                         int32_t qmax, void* stream) {
  return ln_bwd_impl(g_h, x, mean, rstd, gamma, g_res, R, D, out_row_stride, g_x, partials, rows_per_block, h_raw, h_scale, h_zp,
                     qmin, qmax, nullptr, nullptr, nullptr, 0, 0, nullptr, nullptr, 0, nullptr, stream);
}

extern "C" int qv_ln_bwd_gp(const float* g_h, const float* x, const float* mean, const float* rstd, const float* gamma,
                            const float* g_res, int64_t R, int32_t D, float* g_x, float* partials, int32_t rows_per_block,
                            const float* h_raw, const float* h_scale, const int32_t* h_zp, int32_t qmin, int32_t qmax,
                            const float* gp_y, const float* gp_scale, const int32_t* gp_zp, int32_t gp_qmin, int32_t gp_qmax,
                            const float* gp_wscale, uint16_t* gp_out, int64_t gp_plane_stride, float* gp_partials, void* stream) {
  QV_REQUIRE(gp_y && gp_scale && gp_zp && gp_wscale && gp_out, QV_ERR_INVALID, "bad ln_bwd_gp arguments");
  QV_REQUIRE(qv_aligned16(gp_y) && qv_aligned16(gp_wscale) && qv_aligned16(gp_out) && gp_plane_stride % 4 == 0, QV_ERR_INVALID,
             "gp_y / gp_wscale / gp_out must be 16-byte aligned");
  return ln_bwd_impl(g_h, x, mean, rstd, gamma, g_res, R, D, 1, g_x, partials, rows_per_block, h_raw, h_scale, h_zp, qmin, qmax,
                     gp_y, gp_scale, gp_zp, gp_qmin, gp_qmax, gp_wscale, gp_out, gp_plane_stride, gp_partials, stream);
}

extern "C" int qv_colsum_reduce(const float* partials, int32_t nblk, int64_t ncols, float* out, int32_t accumulate,
                                void* stream) {
  QV_REQUIRE(partials && out && nblk > 0 && ncols > 0, QV_ERR_INVALID, "bad colsum_reduce arguments");
  QV_NEED_GPU();
  colsum_reduce_kernel<<<static_cast<unsigned>((ncols + 31) / 32), 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      partials, nblk, ncols, out, accumulate, (ncols % 4 == 0 && qv_aligned16(partials)) ? 1 : 0);
  return qv_check_launch("qv_colsum_reduce");
}

extern "C" int qv_colsum_rows(const float* x, int64_t R, int64_t N, int64_t ld, float* out, int32_t accumulate,
                              void* stream) {
  QV_REQUIRE(x && out && R > 0 && N > 0 && ld >= N, QV_ERR_INVALID, "bad colsum_rows arguments");
  QV_NEED_GPU();
  colsum_rows_kernel<<<static_cast<unsigned>((N + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, R, N, ld,
                                                                                                           out, accumulate);
  return qv_check_launch("qv_colsum_rows");
}

extern "C" int qv_gp_planes(const float* g, const float* y_raw, const float* y_scale, const int32_t* y_zp, int32_t qmin,
                            int32_t qmax, const float* w_scale, int32_t w_scale_per_channel, int32_t gelu, int64_t R,
                            int64_t N, int32_t remap_P, int32_t remap_T, uint16_t* out_planes, int64_t plane_stride,
                            float* bias_partials, int32_t rows_per_block, void* stream) {
  QV_REQUIRE(g && out_planes && R > 0 && N > 0 && rows_per_block > 0, QV_ERR_INVALID, "bad gp_planes arguments");
  QV_REQUIRE(N % 4 == 0, QV_ERR_UNSUPPORTED, "gp_planes needs N %% 4 == 0 (got %lld)", (long long)N);
  QV_REQUIRE((y_scale == nullptr) == (y_zp == nullptr), QV_ERR_INVALID, "y_scale and y_zp go together");
  QV_NEED_GPU();
  const unsigned gridx = static_cast<unsigned>((R + rows_per_block - 1) / rows_per_block);
  const unsigned groups = static_cast<unsigned>(N / 4);
  const unsigned ny = (groups + 255) / 256;                              // column groups per row block, balanced
  const unsigned threads = ((groups + ny - 1) / ny + 31) / 32 * 32;
  const dim3 grid(gridx, ny);
  gp_planes_kernel<4><<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      g, y_raw, y_scale, y_zp, qmin, qmax, w_scale, w_scale_per_channel, gelu, R, N, remap_P, remap_T,
      reinterpret_cast<__nv_bfloat16*>(out_planes), plane_stride, bias_partials, rows_per_block);
  return qv_check_launch("qv_gp_planes");
}

extern "C" int qv_act_planes(const float* y_raw, const float* y_scale, const int32_t* y_zp, int32_t qmin, int32_t qmax,
                             int32_t gelu, int32_t codes_only, int64_t n, uint16_t* out_planes, int64_t plane_stride,
                             void* stream) {
  QV_REQUIRE(y_raw && out_planes && n > 0 && n % 4 == 0, QV_ERR_INVALID, "bad act_planes arguments (n must be a multiple of 4)");
  QV_REQUIRE((y_scale == nullptr) == (y_zp == nullptr), QV_ERR_INVALID, "y_scale and y_zp go together");
  QV_REQUIRE(!codes_only || (y_scale && !gelu), QV_ERR_INVALID, "codes_only needs quantisation parameters and no GELU");
  QV_NEED_GPU();
  act_planes_kernel<<<ew_blocks(n >> 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y_raw, y_scale, y_zp, qmin, qmax, gelu, codes_only, n, reinterpret_cast<__nv_bfloat16*>(out_planes), plane_stride);
  return qv_check_launch("qv_act_planes");
}

extern "C" int qv_embed_fwd(const float* p_raw, const float* p_scale, const int32_t* p_zp, int32_t qmin, int32_t qmax,
                            const float* cls, const float* pos, int64_t B, int32_t P, int32_t D, float* x0, void* stream) {
  QV_REQUIRE(p_raw && cls && pos && x0 && B > 0 && P > 0 && D > 0 && D % 4 == 0, QV_ERR_INVALID, "bad embed_fwd arguments");
  QV_REQUIRE((p_scale == nullptr) == (p_zp == nullptr), QV_ERR_INVALID, "p_scale and p_zp go together");
  QV_NEED_GPU();
  embed_fwd_kernel<<<ew_blocks(B * (P + 1) * (D / 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p_raw, p_scale, p_zp, qmin, qmax, cls, pos, B, P, D, x0);
  return qv_check_launch("qv_embed_fwd");
}

extern "C" int qv_im2col_fq(const float* img, const float* scale, const int32_t* zp, int32_t qmin, int32_t qmax, int64_t B,
                            int32_t C, int32_t HW, int32_t patch, uint16_t* out_plane, uint16_t* out_lo_plane, void* stream) {
  QV_REQUIRE(img && out_plane && B > 0 && C > 0 && HW > 0 && patch > 0, QV_ERR_INVALID, "bad im2col arguments");
  QV_REQUIRE(HW % patch == 0 && patch % 4 == 0, QV_ERR_UNSUPPORTED, "im2col needs HW %% patch == 0 and patch %% 4 == 0");
  QV_REQUIRE((scale == nullptr) == (zp == nullptr), QV_ERR_INVALID, "scale and zp go together");
  QV_NEED_GPU();
  const int64_t total = B * C * HW * (HW / 4);
  im2col_fq_kernel<<<ew_blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      img, scale, zp, qmin, qmax, B, C, HW, patch, reinterpret_cast<__nv_bfloat16*>(out_plane),
      reinterpret_cast<__nv_bfloat16*>(out_lo_plane), scale != nullptr);
  return qv_check_launch("qv_im2col_fq");
}

extern "C" int qv_im2col_u8(const float* img, const float* scale, const int32_t* zp, int64_t B, int32_t C, int32_t HW,
                            int32_t patch, uint8_t* out, void* stream) {
  QV_REQUIRE(img && scale && zp && out && B > 0 && C > 0 && HW > 0 && patch > 0, QV_ERR_INVALID, "bad im2col_u8 arguments");
  QV_REQUIRE(HW % patch == 0 && patch % 4 == 0, QV_ERR_UNSUPPORTED, "im2col needs HW %% patch == 0 and patch %% 4 == 0");
  QV_NEED_GPU();
  const int64_t total = B * C * HW * (HW / 4);
  im2col_u8_kernel<<<ew_blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(img, scale, zp, B, C, HW, patch, out);
  return qv_check_launch("qv_im2col_u8");
}

extern "C" int qv_gelu_minmax(const float* x, int64_t n, float* y, uint32_t* acc, void* stream) {
  QV_REQUIRE(x && y && n > 0 && n % 4 == 0, QV_ERR_INVALID, "bad gelu_minmax arguments (n must be a multiple of 4)");
  QV_NEED_GPU();
  gelu_minmax_kernel<<<ew_blocks(n >> 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, y, acc);
  return qv_check_launch("qv_gelu_minmax");
}

extern "C" int qv_softmax_planes(const float* S, int64_t ldS, int64_t rows, int32_t T, float scale, uint16_t* P, int64_t ldP,
                                 int64_t plane_stride, void* stream) {
  QV_REQUIRE(S && P && rows > 0 && T > 0 && T <= 256 && ldP >= T && ldP <= 256 && ldS >= T, QV_ERR_INVALID,
             "bad softmax_planes arguments (T <= 256)");
  QV_NEED_GPU();
  softmax_planes_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      S, ldS, rows, T, scale, reinterpret_cast<__nv_bfloat16*>(P), ldP, plane_stride);
  return qv_check_launch("qv_softmax_planes");
}

extern "C" int qv_attn_ds(const uint16_t* P, int64_t ldP, int64_t p_plane_stride, const float* dP, int64_t lddP, int64_t rows,
                          int32_t T, float scale, uint16_t* dS, int64_t ldS, int64_t s_plane_stride, void* stream) {
  QV_REQUIRE(P && dP && dS && rows > 0 && T > 0 && T <= 256 && ldS <= 256 && ldS >= T && ldP >= T && lddP >= T, QV_ERR_INVALID,
             "bad attn_ds arguments (T <= 256)");
  QV_NEED_GPU();
  attn_ds_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(P), ldP, p_plane_stride, dP, lddP, rows, T, scale,
      reinterpret_cast<__nv_bfloat16*>(dS), ldS, s_plane_stride);
  return qv_check_launch("qv_attn_ds");
}

extern "C" int qv_head_fwd(const float* x, const float* wq, const float* bias, int32_t B, int32_t K, int32_t N, float* out,
                           uint32_t* minmax, void* stream) {
  QV_REQUIRE(x && wq && out && B > 0 && K > 0 && N > 0, QV_ERR_INVALID, "bad head_fwd arguments");
  QV_NEED_GPU();
  head_fwd_kernel<<<(B * N + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, wq, bias, B, K, N, out, minmax);
  return qv_check_launch("qv_head_fwd");
}

extern "C" int qv_head_bwd(const float* g, const float* x, const float* wq, const uint8_t* wmask, int32_t B, int32_t K,
                           int32_t N, float* gx, float* gw, float* gb, int32_t accumulate, void* stream) {
  QV_REQUIRE(g && x && wq && gx && gw && gb && B > 0 && K > 0 && N > 0, QV_ERR_INVALID, "bad head_bwd arguments");
  QV_NEED_GPU();
  const int64_t total = static_cast<int64_t>(B) * K + static_cast<int64_t>(N) * K + N;
  head_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      g, x, wq, wmask, B, K, N, gx, gw, gb, accumulate);
  return qv_check_launch("qv_head_bwd");
}

extern "C" int qv_ln_quantize_u8_dyn(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                                     int64_t R, int32_t D, const uint32_t* acc, float* scale_out, int32_t* zero_point_out, uint8_t* q,
                                     void* stream) {
  QV_REQUIRE(x && mean && rstd && gamma && beta && acc && scale_out && zero_point_out && q && R > 0, QV_ERR_INVALID,
             "bad ln_quantize_u8_dyn arguments");
  QV_REQUIRE(D % 128 == 0 && D >= 128 && D <= 1024, QV_ERR_UNSUPPORTED, "LayerNorm width must be a multiple of 128 <= 1024 (got %d)", D);
  QV_REQUIRE(qv_aligned16(x) && qv_aligned16(gamma) && qv_aligned16(beta) && (reinterpret_cast<uintptr_t>(q) & 3u) == 0, QV_ERR_INVALID,
             "ln_quantize_u8_dyn needs 16-byte aligned fp32 buffers and a 4-byte aligned code buffer");
  QV_NEED_GPU();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t blocks_needed = (R + 7) / 8;
  const int64_t resident = static_cast<int64_t>(qv_num_sms()) * 4;
  const unsigned grid = static_cast<unsigned>(blocks_needed < resident ? blocks_needed : resident);
#define LAUNCH(V) ln_quantize_u8_dyn_kernel<V><<<grid, 256, 0, st>>>(x, mean, rstd, gamma, beta, R, acc, scale_out, zero_point_out, q)
  switch (D / 128) {
    case 1: LAUNCH(1); break;
    case 2: LAUNCH(2); break;
    case 3: LAUNCH(3); break;
    case 4: LAUNCH(4); break;
    case 6: LAUNCH(6); break;
    case 8: LAUNCH(8); break;
    default: return qv_set_error(QV_ERR_UNSUPPORTED, "LayerNorm width %d not instantiated", D);
  }
#undef LAUNCH
  return qv_check_launch("qv_ln_quantize_u8_dyn");
}
