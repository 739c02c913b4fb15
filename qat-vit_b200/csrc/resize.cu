// resize.cu -- the reference's input transform on the GPU (ref/src/training/qat_trainer.py:210-216):
//
//     transforms.Compose([Resize(224, BICUBIC), ToTensor(), Normalize(mean, std)])   on uint8 HWC images (CIFAR-10: 32x32x3)
//
// i.e. Pillow's two-pass 8-bit bicubic resample (src/libImaging/Resample.c: horizontal pass into a uint8 intermediate, then the
// vertical pass, 22-bit fixed-point coefficients, int32 accumulation from 1 << 21, arithmetic shift, clip to 0..255), then
// float(u8) / 255 and (x - mean[c]) / std[c] in fp32 -- all integer / IEEE operations, so the result is bit-identical to the CPU
// pipeline.  The coefficient tables depend only on the sizes; the host side computes them once in double precision exactly
// as Pillow's precompute_coeffs / normalize_coeffs_8bpc do (qatvit_b200/data.py) and passes them in.
//
// One block per image: the input (3 KB), both tap tables, the horizontally resampled planes [C][Hin][Wout] and the 3 x 256 table
// "uint8 level -> normalised float" (ToTensor + Normalize are a pointwise map, evaluated once per level with IEEE divide /
// subtract) live in shared memory; HBM traffic is the image in and the fp32 NCHW tensor out (602 KB per image, written as
// 16-byte vectors in full 128-byte lines).  A thread owns output columns: its horizontal taps sit in registers, and the vertical
// pass reads four neighbouring pixels as one 32-bit word -- no per-pixel index arithmetic in either loop.
#include <stdio.h>
#include <mutex>

#include "qv_common.cuh"

namespace {

constexpr int RS_PRECISION_BITS = 22;
constexpr int RS_MAX_TAPS = 8;        // taps held in registers on the fast path (bicubic up-scaling: 5)

__device__ __forceinline__ int clip8(int v) { return min(max(v, 0), 255); }

__global__ void __launch_bounds__(256, 2) qv_resize_normalize_kernel(const uint8_t* __restrict__ img, int Hin, int Win, int C,
                                                                     int Hout, int Wout, const int32_t* __restrict__ bounds_h,
                                                                     const int32_t* __restrict__ coef_h,
                                                                     const int32_t* __restrict__ bounds_v,
                                                                     const int32_t* __restrict__ coef_v, int ksize,
                                                                     const float* __restrict__ mean, const float* __restrict__ stdv,
                                                                     float* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int Wp = (Wout + 3) & ~3;
  float* s_lut = reinterpret_cast<float*>(smem);                    // [C][256]
  int32_t* s_bv = reinterpret_cast<int32_t*>(s_lut + 256 * C);      // [Hout][2]
  int32_t* s_kv = s_bv + 2 * Hout;                                  // [Hout][ksize]
  uint8_t* s_in = reinterpret_cast<uint8_t*>(s_kv + ksize * Hout);  // [Hin][Win][C]
  uint8_t* s_tmp = s_in + ((Hin * Win * C + 15) & ~15);             // [C][Hin][Wp]
  const int64_t b = blockIdx.x;
  for (int i = threadIdx.x; i < 256 * C; i += blockDim.x) {
    const int c = i >> 8;
    const float t = __fdiv_rn(static_cast<float>(i & 255), 255.0f);                                 // ToTensor
    s_lut[i] = __fdiv_rn(__fsub_rn(t, __ldg(mean + c)), __ldg(stdv + c));                            // Normalize
  }
  for (int i = threadIdx.x; i < 2 * Hout; i += blockDim.x) s_bv[i] = __ldg(bounds_v + i);
  for (int i = threadIdx.x; i < ksize * Hout; i += blockDim.x) s_kv[i] = __ldg(coef_v + i);
  const uint8_t* src = img + b * Hin * Win * C;
  for (int i = threadIdx.x; i < Hin * Win * C; i += blockDim.x) s_in[i] = __ldg(src + i);
  __syncthreads();
  // horizontal pass, written PLANAR: tmp[c][y][xx] = clip8((2^21 + sum_k in[y][xmin + k][c] * kh[xx][k]) >> 22)
  for (int xx = threadIdx.x; xx < Wout; xx += blockDim.x) {
    const int xmin = __ldg(bounds_h + 2 * xx), n = __ldg(bounds_h + 2 * xx + 1);
    int kh[RS_MAX_TAPS];
#pragma unroll
    for (int k = 0; k < RS_MAX_TAPS; ++k) kh[k] = (k < n && k < ksize) ? __ldg(coef_h + xx * ksize + k) : 0;
    for (int c = 0; c < C; ++c)
      for (int y = 0; y < Hin; ++y) {
        const uint8_t* p = s_in + (y * Win + xmin) * C + c;
        int ss = 1 << (RS_PRECISION_BITS - 1);
        if (n <= RS_MAX_TAPS) {
#pragma unroll
          for (int k = 0; k < RS_MAX_TAPS; ++k)
            if (k < n) ss += static_cast<int>(p[k * C]) * kh[k];
        } else {
          for (int k = 0; k < n; ++k) ss += static_cast<int>(p[k * C]) * __ldg(coef_h + xx * ksize + k);
        }
        s_tmp[(c * Hin + y) * Wp + xx] = static_cast<uint8_t>(clip8(ss >> RS_PRECISION_BITS));
      }
  }
  __syncthreads();
  // vertical pass + level table, NCHW output
  float* dst = out + b * C * Hout * Wout;
  if ((Wout & 3) == 0) {
    const int w4 = Wout >> 2;
    const int groups = max(1, static_cast<int>(blockDim.x) / w4);            // row groups working side by side
    for (int u = threadIdx.x; u < w4 * groups; u += blockDim.x) {
      const int x4 = u % w4, g = u / w4;
      for (int c = 0; c < C; ++c) {
        const float* lut = s_lut + (c << 8);
        const uint32_t* plane = reinterpret_cast<const uint32_t*>(s_tmp + c * Hin * Wp) + x4;
        float* drow = dst + static_cast<int64_t>(c) * Hout * Wout + 4 * x4;
        for (int yy = g; yy < Hout; yy += groups) {
          const int ymin = s_bv[2 * yy], n = s_bv[2 * yy + 1];
          int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0, s3 = s0;
          for (int k = 0; k < n; ++k) {
            const uint32_t w = plane[(ymin + k) * (Wp >> 2)];        // four neighbouring pixels of input row ymin + k
            const int kv = s_kv[yy * ksize + k];
            s0 += static_cast<int>(w & 0xffu) * kv;
            s1 += static_cast<int>((w >> 8) & 0xffu) * kv;
            s2 += static_cast<int>((w >> 16) & 0xffu) * kv;
            s3 += static_cast<int>(w >> 24) * kv;
          }
          *reinterpret_cast<float4*>(drow + static_cast<int64_t>(yy) * Wout) =
              make_float4(lut[clip8(s0 >> RS_PRECISION_BITS)], lut[clip8(s1 >> RS_PRECISION_BITS)],
                          lut[clip8(s2 >> RS_PRECISION_BITS)], lut[clip8(s3 >> RS_PRECISION_BITS)]);
        }
      }
    }
  } else {
    for (int xx = threadIdx.x; xx < Wout; xx += blockDim.x)
      for (int c = 0; c < C; ++c)
        for (int yy = 0; yy < Hout; ++yy) {
          const int ymin = s_bv[2 * yy], n = s_bv[2 * yy + 1];
          int ss = 1 << (RS_PRECISION_BITS - 1);
          for (int k = 0; k < n; ++k) ss += static_cast<int>(s_tmp[(c * Hin + ymin + k) * Wp + xx]) * s_kv[yy * ksize + k];
          dst[(static_cast<int64_t>(c) * Hout + yy) * Wout + xx] = s_lut[(c << 8) + clip8(ss >> RS_PRECISION_BITS)];
        }
  }
}

}  // namespace

extern "C" int qv_resize_normalize_u8(const uint8_t* img, int64_t B, int32_t Hin, int32_t Win, int32_t C, int32_t Hout,
                                      int32_t Wout, const int32_t* bounds_h, const int32_t* coef_h, const int32_t* bounds_v,
                                      const int32_t* coef_v, int32_t ksize, const float* mean, const float* stdv, float* out,
                                      void* stream) {
  QV_REQUIRE(B >= 0, QV_ERR_INVALID, "bad batch");
  if (B == 0) return QV_OK;
  QV_REQUIRE(img && out && bounds_h && coef_h && bounds_v && coef_v && mean && stdv, QV_ERR_INVALID, "null pointer");
  QV_REQUIRE(Hin > 0 && Win > 0 && C > 0 && Hout > 0 && Wout > 0 && ksize > 0 && ksize <= 64, QV_ERR_INVALID, "bad geometry");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  QV_REQUIRE(qv_aligned16(out), QV_ERR_INVALID, "out must be 16-byte aligned");
  const size_t smem = static_cast<size_t>(4) * (256 * C + (2 + ksize) * static_cast<size_t>(Hout)) +
                      ((static_cast<size_t>(Hin) * Win * C + 15) & ~size_t(15)) + static_cast<size_t>(C) * Hin * ((Wout + 3) & ~3);
  QV_REQUIRE(smem <= 200 * 1024, QV_ERR_UNSUPPORTED,
             "resize: a %dx%dx%d image and its %d-wide intermediate do not fit shared memory (%zu bytes; CIFAR-sized inputs do)",
             Hin, Win, C, Wout, smem);
  QV_REQUIRE(B < 2147483647LL, QV_ERR_UNSUPPORTED, "batch too large");
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(qv_resize_normalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  QV_REQUIRE(attr_err == cudaSuccess, QV_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  qv_resize_normalize_kernel<<<static_cast<unsigned>(B), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      img, Hin, Win, C, Hout, Wout, bounds_h, coef_h, bounds_v, coef_v, ksize, mean, stdv, out);
  return qv_check_launch("qv_resize_normalize_u8");
}
