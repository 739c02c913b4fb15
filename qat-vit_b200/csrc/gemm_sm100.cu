// gemm_sm100.cu -- persistent, warp-specialised tcgen05 GEMM for the fake-quant Linear family.
//
//   D[M,N] = sum_pairs A[pa] (bf16 planes) x B[pb]^T (bf16 planes), fp32 accumulation in TMEM.
//
// Replaces F.linear inside torch.ao.nn.qat.Linear.forward (torch/ao/nn/qat/modules/linear.py:50-51),
// its autograd dgrad / wgrad mm's, and the teacher's nn.Linear (ref qat_trainer.py:337-341).
// fp32 tensors reach the tensor cores as bf16 hi/lo plane stacks; fake-quantised weights as ONE exact
// plane of integer codes with the per-channel scale applied in the epilogue (SURVEY.md §7 "hard parts").
//
// Structure (one CTA per SM, static round-robin over 128 x BN output tiles):
//   warp 0      : TMA producer  -- cp.async.bulk.tensor into a STAGES-deep 128B-swizzled smem ring
//   warp 1      : tcgen05.mma issuer (one thread) -- accumulators double-buffered in TMEM
//   warps 2..5  : epilogue -- tcgen05.ld -> scale/bias -> fused observer min/max -> coalesced fp32 stores
// so the epilogue of tile i overlaps the main loop of tile i+1 (the student GEMMs have K = 384 and are
// close to HBM-bound on the fp32 output write).
#include <cuda.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <mutex>

#include "qv_common.cuh"
#include "qv_ptx.cuh"

using namespace qvptx;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;            // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int A_STAGE_BYTES = BM * BK * 2;

struct GemmKParams {
  int64_t M, N;
  int32_t kblocks;        // ceil(K / BK) per plane pair
  int32_t npairs;
  int32_t pair_a[4], pair_b[4];
  int32_t tiles_m, tiles_n, splits, kb_per_split;
  float* d;
  int64_t ldd;
  const float* col_scale;
  const float* col_rscale;
  const float* alpha;
  const float* bias;
  uint32_t* minmax;
  float* workspace;
  // batching: item -> (outer index bt) -> (bo, bi) = (bt / batch_inner, bt % batch_inner)
  int32_t nbatch, batch_inner;
  int32_t a_c2_outer, a_c2_inner, a_col0, a_col_inner;
  int32_t b_c2_outer, b_c2_inner, b_col0, b_col_inner;
  int64_t d_off_outer, d_off_inner;
};

template <int BN>
struct Cfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN == 128) ? 6 : 4;
  static constexpr int TMEM_COLS = 2 * BN;                 // 256 or 512 (power of two)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
qv_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const GemmKParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms need 1024-byte aligned tile bases
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * C::STAGES;      // [2]
  uint64_t* tmem_empty = bars + 2 * C::STAGES + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_items = p.tiles_m * p.tiles_n * (p.nbatch > 1 ? p.nbatch : p.splits);

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int n_blk = item % p.tiles_n;
        const int m_blk = (item / p.tiles_n) % p.tiles_m;
        const int outer = item / (p.tiles_n * p.tiles_m);
        const int z = p.nbatch > 1 ? 0 : outer;
        const int bt = p.nbatch > 1 ? outer : 0;
        const int bo = bt / p.batch_inner, bi = bt % p.batch_inner;
        const int a_c2 = bo * p.a_c2_outer + bi * p.a_c2_inner, a_col = p.a_col0 + bi * p.a_col_inner;
        const int b_c2 = bo * p.b_c2_outer + bi * p.b_c2_inner, b_col = p.b_col0 + bi * p.b_col_inner;
        const int kb0 = z * p.kb_per_split;
        const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
        for (int pr = 0; pr < p.npairs; ++pr) {
          const int pa = p.pair_a[pr], pb = p.pair_b[pr];
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
            uint8_t* sb = smem_b + stage * C::B_STAGE_BYTES;
            if (!A_MN) {
              tma_load_4d(sa, &map_a, &full_bar[stage], a_col + kb * BK, m_blk * BM, a_c2, pa);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_4d(sa + j * 8192, &map_a, &full_bar[stage], a_col + m_blk * BM + j * 64, kb * BK, a_c2, pa);
            }
            if (!B_MN) {
              tma_load_4d(sb, &map_b, &full_bar[stage], b_col + kb * BK, n_blk * BN, b_c2, pb);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_4d(sb + j * 8192, &map_b, &full_bar[stage], b_col + n_blk * BN + j * 64, kb * BK, b_c2, pb);
            }
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      constexpr uint32_t a_lbo = A_MN ? 8192u : 16u, b_lbo = B_MN ? 8192u : 16u;
      constexpr uint32_t a_kadv = A_MN ? (UMMA_K * 128u) : (UMMA_K * 2u);   // bytes per 16-deep k step
      constexpr uint32_t b_kadv = B_MN ? (UMMA_K * 128u) : (UMMA_K * 2u);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local) {
        const int z = p.nbatch > 1 ? 0 : item / (p.tiles_n * p.tiles_m);
        const int kb0 = z * p.kb_per_split;
        const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
        const int iters = p.npairs * (kb1 - kb0);
        const int buf = local & 1;
        const uint32_t use = static_cast<uint32_t>(local >> 1);      // n-th use of this TMEM buffer
        mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * BN);
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * C::B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = umma_smem_desc(sa + k * a_kadv, a_lbo, 1024u);
            const uint64_t db = umma_smem_desc(sb + k * b_kadv, b_lbo, 1024u);
            umma_bf16(d_tmem, da, db, idesc, (it | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);       // frees this smem stage once the MMAs have read it
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[buf]);           // accumulator complete -> epilogue
      }
    }
  } else {
    // =============================== epilogue (128 threads) ===============================
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int row_in_tile = q * 32 + lane;
    float mn = INFINITY, mx = -INFINITY;
    const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
    const bool vec_ok = p.splits > 1 ? (p.N % 4 == 0 && (reinterpret_cast<uintptr_t>(p.workspace) & 15) == 0)
                                     : (p.ldd % 4 == 0 && (reinterpret_cast<uintptr_t>(p.d) & 15) == 0);
    int local = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local) {
      const int n_blk = item % p.tiles_n;
      const int m_blk = (item / p.tiles_n) % p.tiles_m;
      const int outer = item / (p.tiles_n * p.tiles_m);
      const int z = p.nbatch > 1 ? 0 : outer;
      const int bt = p.nbatch > 1 ? outer : 0;
      const int buf = local & 1;
      const uint32_t use = static_cast<uint32_t>(local >> 1);
      mbar_wait(&tmem_full[buf], use & 1);
      tc_fence_after();
      const int64_t m = static_cast<int64_t>(m_blk) * BM + row_in_tile;
      float* out;
      int64_t ldo;
      const bool raw = p.splits > 1;
      if (raw) {
        out = p.workspace + static_cast<int64_t>(z) * p.M * p.N;
        ldo = p.N;
      } else {
        out = p.d + (bt / p.batch_inner) * p.d_off_outer + (bt % p.batch_inner) * p.d_off_inner;
        ldo = p.ldd;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN + c0), r);
        tmem_ld_wait();
        const int64_t n0 = static_cast<int64_t>(n_blk) * BN + c0;
        if (m < p.M && n0 < p.N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float a = __uint_as_float(r[j]);
            if (!raw) {
              const int64_t n = n0 + j;
              if (n < p.N) {
                float mult = alpha;
                if (p.col_scale) mult *= __ldg(p.col_scale + n);
                if (p.col_rscale) mult = __fdiv_rn(mult, __ldg(p.col_rscale + n));
                a *= mult;
                if (p.bias) a += __ldg(p.bias + n);
                mn = fminf(mn, a);
                mx = fmaxf(mx, a);
              }
            }
            v[j] = a;
          }
          float* dst = out + m * ldo + n0;
          if (vec_ok && n0 + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < p.N) dst[j] = v[j];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[buf]);
    }
    if (p.minmax && p.splits == 1) {
      mn = qv_warp_min(mn);
      mx = qv_warp_max(mx);
      if (lane == 0 && mn <= mx) {
        atomicMin(p.minmax, qv_f2ord(mn));
        atomicMax(p.minmax + 1, qv_f2ord(mx));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// split-K reduction (+ weight-FQ STE mask, 1/scale un-folding, accumulate into the gradient arena)
// ------------------------------------------------------------------------------------------------
__global__ void qv_splitk_reduce_kernel(const float* __restrict__ ws, int splits, int64_t M, int64_t N,
                                        const float* __restrict__ row_rscale, const float* __restrict__ alpha,
                                        const uint8_t* __restrict__ mask, float* __restrict__ out, int accumulate) {
  const int64_t total = M * N;
  const float al = alpha ? __ldg(alpha) : 1.0f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws[z * total + i];
    float mult = al;
    if (row_rscale) mult = __fdiv_rn(mult, __ldg(row_rscale + i / N));
    s *= mult;
    if (mask && !mask[i]) s = 0.f;
    out[i] = accumulate ? out[i] + s : s;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static std::once_flag once;
  static EncodeTiledFn fn = nullptr;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// bf16 plane stack [planes][nb][rows][ld] viewed as a 4-D tensor (cols, rows, nb, planes); box = (64, box_rows, 1, 1).
// Out-of-range rows / cols (per batch matrix) are zero-filled by TMA, which is what makes ragged M/N/K and
// per-(image, head) batching safe in the contraction dimension.
int make_map(CUtensorMap* m, const qv_operand& op, int planes, int box_rows) {
  EncodeTiledFn enc = get_encode();
  QV_REQUIRE(enc != nullptr, QV_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  QV_REQUIRE(op.ptr && qv_aligned16(op.ptr), QV_ERR_INVALID, "gemm operand base must be a 16-byte aligned device pointer");
  QV_REQUIRE(op.rows > 0 && op.cols > 0, QV_ERR_INVALID, "gemm operand extent must be positive");
  QV_REQUIRE(op.ld % 8 == 0 && op.ld >= op.cols, QV_ERR_INVALID, "gemm operand row pitch must be >= cols and a multiple of 8 bf16 (got %lld)", (long long)op.ld);
  int64_t nb = op.nb > 0 ? op.nb : 1;
  int64_t bstride = op.batch_stride > 0 ? op.batch_stride : op.rows * op.ld;
  int64_t pstride = op.plane_stride > 0 ? op.plane_stride : bstride * nb;
  QV_REQUIRE(bstride % 8 == 0 && pstride % 8 == 0, QV_ERR_INVALID, "batch / plane strides must be multiples of 8 bf16");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(op.cols), static_cast<cuuint64_t>(op.rows), static_cast<cuuint64_t>(nb),
                        static_cast<cuuint64_t>(planes)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(op.ld) * 2, static_cast<cuuint64_t>(bstride) * 2,
                           static_cast<cuuint64_t>(pstride) * 2};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(box_rows), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(op.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  QV_REQUIRE(r == CUDA_SUCCESS, QV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

template <int BN, bool A_MN, bool B_MN>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const GemmKParams& kp, int grid, cudaStream_t st) {
  using C = Cfg<BN>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(qv_gemm_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    C::SMEM_BYTES);
  });
  QV_REQUIRE(attr_err == cudaSuccess, QV_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  qv_gemm_kernel<BN, A_MN, B_MN><<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(ma, mb, kp);
  return qv_check_launch("qv_gemm_bf16");
}

}  // namespace

extern "C" int qv_gemm_bf16(const qv_gemm_args* a, void* stream) {
  QV_REQUIRE(a != nullptr, QV_ERR_INVALID, "null args");
  QV_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, QV_ERR_INVALID, "empty gemm (M=%lld N=%lld K=%lld)", (long long)a->M,
             (long long)a->N, (long long)a->K);
  QV_REQUIRE(a->npairs >= 1 && a->npairs <= 4, QV_ERR_INVALID, "npairs must be 1..4");
  const int splits = a->splits > 1 ? a->splits : 1;
  const int nbatch = a->nbatch > 1 ? a->nbatch : 1;
  QV_REQUIRE(!(splits > 1 && nbatch > 1), QV_ERR_UNSUPPORTED, "split-K and batching are mutually exclusive");
  if (splits > 1) QV_REQUIRE(a->workspace != nullptr, QV_ERR_INVALID, "split-K needs a workspace");
  else QV_REQUIRE(a->d != nullptr, QV_ERR_INVALID, "null output");
  int pa_max = 0, pb_max = 0;
  for (int i = 0; i < a->npairs; ++i) {
    QV_REQUIRE(a->pair_a[i] >= 0 && a->pair_b[i] >= 0 && a->pair_a[i] < 8 && a->pair_b[i] < 8, QV_ERR_INVALID,
               "bad plane index");
    pa_max = a->pair_a[i] > pa_max ? a->pair_a[i] : pa_max;
    pb_max = a->pair_b[i] > pb_max ? a->pair_b[i] : pb_max;
  }
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  const int BN = 128;
  CUtensorMap ma, mb;
  int rc = make_map(&ma, a->a, pa_max + 1, a->a.mn_major ? 64 : BM);
  if (rc) return rc;
  rc = make_map(&mb, a->b, pb_max + 1, a->b.mn_major ? 64 : BN);
  if (rc) return rc;

  GemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.M = a->M;
  kp.N = a->N;
  kp.kblocks = static_cast<int32_t>((a->K + BK - 1) / BK);
  kp.npairs = a->npairs;
  for (int i = 0; i < a->npairs; ++i) { kp.pair_a[i] = a->pair_a[i]; kp.pair_b[i] = a->pair_b[i]; }
  kp.tiles_m = static_cast<int32_t>((a->M + BM - 1) / BM);
  kp.tiles_n = static_cast<int32_t>((a->N + BN - 1) / BN);
  int sp = splits > kp.kblocks ? kp.kblocks : splits;
  kp.kb_per_split = (kp.kblocks + sp - 1) / sp;
  sp = (kp.kblocks + kp.kb_per_split - 1) / kp.kb_per_split;     // no empty split
  QV_REQUIRE(splits == 1 || sp == splits, QV_ERR_INVALID,
             "split-K: %d splits over %d k-blocks leaves empty splits (use %d)", splits, kp.kblocks, sp);
  kp.splits = sp;
  kp.d = a->d;
  kp.ldd = a->ldd;
  kp.col_scale = a->col_scale;
  kp.col_rscale = a->col_rscale;
  kp.alpha = a->alpha;
  kp.bias = a->bias;
  kp.minmax = a->minmax;
  kp.workspace = a->workspace;
  kp.nbatch = nbatch;
  kp.batch_inner = a->batch_inner > 0 ? a->batch_inner : 1;
  kp.a_c2_outer = a->a.c2_outer; kp.a_c2_inner = a->a.c2_inner; kp.a_col0 = a->a.col0; kp.a_col_inner = a->a.col_inner;
  kp.b_c2_outer = a->b.c2_outer; kp.b_c2_inner = a->b.c2_inner; kp.b_col0 = a->b.col0; kp.b_col_inner = a->b.col_inner;
  kp.d_off_outer = a->d_off_outer;
  kp.d_off_inner = a->d_off_inner;
  const int64_t items = static_cast<int64_t>(kp.tiles_m) * kp.tiles_n * (nbatch > 1 ? nbatch : kp.splits);
  QV_REQUIRE(items < (1LL << 31), QV_ERR_UNSUPPORTED, "too many tiles");
  const int sms = qv_num_sms();
  const int grid = static_cast<int>(items < sms ? items : sms);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool amn = a->a.mn_major != 0, bmn = a->b.mn_major != 0;
  if (!amn && !bmn) return launch<128, false, false>(ma, mb, kp, grid, st);
  if (amn && bmn) return launch<128, true, true>(ma, mb, kp, grid, st);
  if (amn && !bmn) return launch<128, true, false>(ma, mb, kp, grid, st);
  return launch<128, false, true>(ma, mb, kp, grid, st);
}

extern "C" int qv_splitk_reduce(const float* workspace, int32_t splits, int64_t M, int64_t N, const float* row_rscale,
                                const float* alpha, const uint8_t* mask, float* out, int32_t accumulate, void* stream) {
  QV_REQUIRE(workspace && out && splits >= 1 && M > 0 && N > 0, QV_ERR_INVALID, "bad splitk_reduce arguments");
  const int64_t total = M * N;
  int blocks = static_cast<int>((total + 255) / 256);
  const int cap = qv_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  qv_splitk_reduce_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(workspace, splits, M, N, row_rscale,
                                                                                 alpha, mask, out, accumulate);
  return qv_check_launch("qv_splitk_reduce");
}
