// gemm_sm100.cu -- persistent, warp-specialised tcgen05 GEMM for the fake-quant Linear family.
//
//   D[M,N] (fp32) = sum over plane pairs of A[pa] (bf16) x B[pb]^T (bf16), fp32 accumulation in TMEM.
//
// Replaces F.linear inside torch.ao.nn.qat.Linear.forward (torch/ao/nn/qat/modules/linear.py:50-51), its autograd
// dgrad / wgrad mm's, the teacher's nn.Linear (ref qat_trainer.py:337-341) and -- batched per (image, head) -- the
// matmuls of F.scaled_dot_product_attention and its backward.  fp32 tensors reach the tensor cores as bf16 hi/lo plane
// stacks; fake-quantised weights as ONE exact plane of integer codes with the per-channel scale in the epilogue.
//
// Structure (one CTA per SM, static round-robin over 128 x BN output tiles):
//   warp 0      : TMA producer.  One pipeline stage holds ALL planes of one 64-deep k-block (NA A-planes + NB B-planes),
//                 so the hi/lo passes re-use the B (or A) tile from shared memory instead of re-fetching it from L2.
//   warp 1      : tcgen05.mma issuer (one thread): for each stage, every (pa, pb) pair x 4 UMMA_K steps into the same
//                 TMEM accumulator; accumulators are double-buffered in TMEM.
//   warps 2..5  : epilogue: tcgen05.ld (32 lanes x 32 columns) -> scale / bias -> fused observer min/max ->
//                 128B-swizzled smem staging -> per-warp TMA store (cp.async.bulk.tensor), double-buffered, so the
//                 epilogue of tile i overlaps the main loop of tile i+1 and global writes are full 128-byte lines.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <mutex>

#include "qv_common.cuh"
#include "qv_ptx.cuh"
#include "qv_tma.cuh"
#include "qv_observer.cuh"

using namespace qvptx;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;            // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
// Threads: TMA warp, MMA warp, then the epilogue warps.  The plane-output epilogue (EPI 1: bias [+ GELU] + operand split, 13-23
// instructions per element with MUFU / conversion chains) is LATENCY-bound with two warps per scheduler -- ncu: issue slots 40 %,
// tensor pipe 57 % on the teacher's fc1, where a 256 x 256 pair tile's 12 k-blocks of MMAs are shorter than its epilogue -- so it runs on
// 16 warps (four per TMEM lane quarter and scheduler) over 32-column chunks; the lighter fp32 epilogue keeps 8.
__host__ __device__ constexpr int epi_warps(int epi) { return epi == 1 ? 16 : 8; }
__host__ __device__ constexpr int num_threads(int epi) { return 64 + 32 * epi_warps(epi); }
constexpr int A_PLANE_BYTES = BM * BK * 2;
// epilogue staging per warp: 32-row x 128-byte buffers.  fp32 output: one buffer per 32-column chunk; bf16 hi/lo plane
// output: a hi and a lo buffer per 64-column chunk.  Two warps alternate on each TMEM lane quarter, so per-warp single
// buffering already overlaps one warp's TMA store with the other's math.
// EPI 0: 32 x 32 fp32;  EPI 1: 32 rows x 32 columns of both planes (2 x 2 KB: bf16 hi | lo, or fp16 | hi8 + lo8);  EPI 2: y tile x2 + planes
constexpr int epi_warp_bytes(int epi) { return (epi == 2 ? 3 : 1) * 32 * 128; }
constexpr int epi_terms_bytes(int epi) { return epi >= 1 ? 0 : 8 * 2 * 32 * 4; }   // per-warp [mult][bias] column terms
constexpr int SMEM_LIMIT = 232448;             // 227 KB
#ifndef QV_GEMM_PAIR_DEFAULT
// bit 0: mixed-format (teacher) GEMMs with fp32 output, bit 4: ... with plane output, bit 1: (2,2) bf16 hi/lo plane GEMMs,
// bit 2: gradient-planes dgrad, bit 3: every (2,1) GEMM, bit 5: 256-wide tiles for the mixed-format pairs, bit 6: the (2,1) GEMMs with
// plane output or K >= 1024.  (Tried and dropped: the producer prefetching
// the A tiles 6 k-blocks ahead into L2 with cp.async.bulk.prefetch.tensor -- isolated qkv 280 -> 305 us, fc2 353 -> 376 us.)
// The (2,1) student GEMMs are MMA-bound with one CTA per tile already (56 KB per k-block against 8 MMAs) and gain nothing.
#define QV_GEMM_PAIR_DEFAULT 115
#endif

struct GemmKParams {
  int64_t M, N;
  int32_t kblocks;        // ceil(K / BK)
  int32_t tiles_m, tiles_n, splits, kb_per_split;
  const float* col_scale;
  const float* col_rscale;
  const float* alpha;
  const float* bias;
  uint32_t* minmax;
  // batching: outer index bt -> (bo, bi) = (bt / batch_inner, bt % batch_inner)
  int32_t nbatch, batch_inner;
  int32_t a_c2_outer, a_c2_inner, a_col0, a_col_inner;
  int32_t b_c2_outer, b_c2_inner, b_col0, b_col_inner;
  int32_t o_c2_outer, o_c2_inner, o_col0, o_col_inner;
  int32_t act;            // 0 = none, 1 = exact-erf GELU (applied after scale / bias)
  float acc_scale;        // accumulator pre-scale (1, or 2^-14 for the fp16 + fp8 "mixed" operand format)
  int32_t out_fmt;        // EPI 1: 0 = bf16 hi/lo planes, 1 = mixed planes (fp16 hi | e5m2 hi8 / lo8 blocks)
  // EPI 2 ("gradient planes"): gq = acc * [gelu'(FQ(y))] * STEmask(y), planes = split(gq * col_scale[n]), column sums of gq
  const float* ep_raw;    // y: raw output [M, N] of the Linear whose output fake-quant the gradient passes through
  int64_t ep_raw_ld;
  const float* ep_scale;
  const int32_t* ep_zp;
  int32_t ep_qmin, ep_qmax, ep_gelu;
  float* ep_colsum;       // [ceil(M/32)][N] per-32-row-slab column sums of gq (bias-grad partials), may be NULL
  // observer update in the kernel tail (needs minmax): the grid's last epilogue warp turns the merged min / max into the
  // module's running range and (scale, zero_point) -- qv_obs_update without a launch
  float* obs_min_val; float* obs_max_val; float* obs_scale; int32_t* obs_zero_point;
  const int64_t* obs_enabled; const int64_t* obs_fq_enabled;
  float obs_c; int32_t obs_qmin, obs_qmax, obs_symmetric;
  uint32_t* obs_ticket;
  int32_t* sat_flag;      // EPI 1, out_fmt 1: OR sat_bit into *sat_flag when an output leaves the mixed format's range
  int32_t sat_bit;
};

// SPLIT (mixed-format CTA pairs): one pipeline stage holds ONE region of a k-block -- the fp16 planes of A and B, or their fp8
// value / residual planes.  The fp16 product needs only the first, the two fp8 cross terms only the second, so the stages are
// half as large, twice as many fit (5-6 instead of 2-3) and each is released as soon as its four MMAs have read it.
template <int BN, int NA, int NB, int EPI = 0, int CG = 1, bool SPLIT = false>
struct Cfg {
  // CG = 2 (CTA pair, tcgen05 cta_group::2): a 256 x BN tile per pair; each CTA stages its 128 rows of A and BN / 2 rows of B
  static constexpr int B_PLANE_BYTES = (BN / CG) * BK * 2;
  static constexpr int STAGE_BYTES = SPLIT ? (A_PLANE_BYTES + B_PLANE_BYTES) : (NA * A_PLANE_BYTES + NB * B_PLANE_BYTES);
  static_assert(!SPLIT || (NA == 2 && NB == 2), "region-split stages: two regions per operand");
  static constexpr int EPI_WARP_BYTES = epi_warp_bytes(EPI);
  static constexpr int EPI_BYTES = epi_warps(EPI) * EPI_WARP_BYTES + epi_terms_bytes(EPI);
  static constexpr int MAX_STAGES = (SMEM_LIMIT - 1024 - 256 - EPI_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = MAX_STAGES > 6 ? 6 : MAX_STAGES;       // barrier block holds 2 x 6 + 13 words
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int NPAIRS = (NA == 2 && NB == 2) ? 3 : NA * NB;   // (hi,hi) (hi,lo) (lo,hi): lo*lo is dropped
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN must be a multiple of 32 up to 256");
  static_assert(EPI == 0 || BN % 64 == 0, "plane output works on 64-column chunks");
  static_assert(CG == 1 || (CG == 2 && BN % 16 == 0 && B_PLANE_BYTES % 1024 == 0), "CTA pair: N a multiple of 16, 1 KB aligned half tiles");
};

#ifdef QV_ATTN_DEBUG      // timeline-instrumented build (make debug): clock64 events of CTA 0, read back with qv_gemm_debug_read
__device__ unsigned long long qv_gemm_dbg[3][4096];
__device__ __forceinline__ void gdbg(int who, int& n, int tag) {
  if (blockIdx.x == 0 && n < 4096) qv_gemm_dbg[who][n++] = (static_cast<unsigned long long>(tag) << 48) | (clock64() & 0xffffffffffffULL);
}
#define GDBG(who, tag) gdbg(who, gdbg_n, tag)
#else
#define GDBG(who, tag)
#endif

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// EPI = 0: fp32 output;  EPI = 1: bf16 hi/lo plane output (the operand format of the next GEMM), optional GELU;
// EPI = 2: dgrad producing the NEXT layer's gradient planes: the accumulator (dL/d FQ(y) or dL/d GELU(FQ(y))) is multiplied by
// gelu'(FQ(y)) -- a 256-entry table over the integer codes, built per CTA -- and the STE mask recomputed from the raw y tile,
// folded with the per-channel weight scale and written as hi/lo planes; bias-grad column sums leave as per-slab partials.
template <int BN, int NA, int NB, bool A_MN, bool B_MN, int EPI, int MIX, int CG = 1>
__global__ void __launch_bounds__(num_threads(EPI), 1)
qv_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_o, const __grid_constant__ CUtensorMap map_y, const GemmKParams p) {
  constexpr bool SPLIT = MIX != 0 && CG == 2;
  using C = Cfg<BN, NA, NB, EPI, CG, SPLIT>;
  static_assert(!MIX || (NA == 2 && NB == 2 && !A_MN && !B_MN), "mixed operands: two K-major regions per operand");
  static_assert(CG == 1 || (!A_MN && !B_MN), "CTA pairs: K-major operands only");
  // CTA pair (CG = 2, launched as clusters of 2): p.tiles_m counts 256-row tiles; CTA `rank` owns rows [rank * 128, +128) of
  // the pair's tile and loads the B rows [rank * BN / 2, +BN / 2); rank 0 issues the MMAs for both.
  const int rank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const int item0 = static_cast<int>(blockIdx.x) / CG, item_step = static_cast<int>(gridDim.x) / CG;
  auto m_of = [&](int item) { return ((item / p.tiles_n) % p.tiles_m) * CG + rank; };
  if constexpr (CG == 2) cluster_sync_all();        // both CTAs of the pair are resident before anything crosses over
  constexpr int EPI_WARP_BYTES = C::EPI_WARP_BYTES;
  constexpr int EPI_BYTES = C::EPI_BYTES;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms need 1024-byte aligned tile bases
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_epi = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + EPI_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * C::STAGES;      // [2]
  uint64_t* tmem_empty = bars + 2 * C::STAGES + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
  uint64_t* raw_bar = bars + 2 * C::STAGES + 5;    // [8] EPI 2: one per epilogue warp (raw y tile landed)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // EPI 2: FQ(y) takes at most qmax - qmin + 1 <= 256 distinct values, so  gelu'(FQ(y)) * STEmask(y)  is ONE table lookup on the
  // code clamped to [qmin - 1, qmax + 1]: entry 0 and the last entry are the two out-of-range sides (mask = 0), entry k + 1 is
  // gelu'((qmin + k - zp) * scale) -- or 1 without GELU.  Same expression per code as the elementwise kernels -> same bits.
  // The table is kept in 8 interleaved copies (entry k, copy r at word 8 k + r; lane L reads copy L & 7): the codes of a warp's 32
  // rows are data-dependent, and with ONE copy the gather cost ~6 shared-memory wavefronts per load (ncu: 63 % of the kernel's
  // shared-load wavefronts were bank conflicts, LSU data pipe 69 % busy = the limiter); with 8 copies only lanes L, L + 8, L + 16,
  // L + 24 can still collide.
  constexpr int LUT_COPIES = 8;
  __shared__ float gelu_lut[EPI == 2 ? 260 * LUT_COPIES : 1];
  if constexpr (EPI == 2) {
    const QvQParams yq = qv_load_qparams(p.ep_scale, p.ep_zp, p.ep_qmin, p.ep_qmax);
    const int ncodes = p.ep_qmax - p.ep_qmin + 1;
    for (int k = threadIdx.x; k < ncodes + 2 && k < 260; k += num_threads(EPI)) {
      float v = 0.f;
      if (k >= 1 && k <= ncodes)
        v = p.ep_gelu ? qv_gelu_grad(__fmul_rn(__fsub_rn(static_cast<float>(p.ep_qmin + k - 1), yq.zp), yq.scale)) : 1.0f;
#pragma unroll
      for (int r = 0; r < LUT_COPIES; ++r) gelu_lut[k * LUT_COPIES + r] = v;
    }
  }

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    prefetch_tensormap(&map_o);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], epi_warps(EPI) * 32 * CG);   // every epilogue thread of the pair arrives on the leader's barrier
    }
    if constexpr (EPI == 2) {
      prefetch_tensormap(&map_y);
      for (int w = 0; w < 8; ++w) mbar_init(&raw_bar[w], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
    else tmem_alloc(tmem_slot, C::TMEM_COLS);
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();        // barrier inits of both CTAs are visible cluster-wide
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // epilogue threads release an accumulator buffer on the LEADER's barrier (the only CTA whose MMA thread waits on it)
  auto tmem_empty_arrive = [&](int buf) {
    if constexpr (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[buf]), 0));
    else mbar_arrive(&tmem_empty[buf]);
  };

  const int num_items = p.tiles_m * p.tiles_n * (p.nbatch > 1 ? p.nbatch : p.splits);

  if (warp == 0) {
    // =============================== TMA producer ===============================
    // Like the MMA issuer below, the whole warp walks the schedule and ONE elected lane issues the loads: coordinates, shared-memory
    // addresses and barrier addresses then live in uniform registers.  As a single thread (`if (lane == 0)`) every TMA instruction
    // paid a vector-to-uniform waterfall, and the timeline showed the producer BUSY 87 % of a teacher GEMM while the MMA side waited
    // for operands 21 % of it: the loads were late because their issue was slow, not because memory was.
    {
      int stage = 0;
      uint32_t phase = 0;
#ifdef QV_ATTN_DEBUG
      int gdbg_n = (lane == 0) ? 0 : 4096;
#endif
      for (int item = item0; item < num_items; item += item_step) {
        const int n_blk = item % p.tiles_n;
        const int m_blk = m_of(item);
        const int outer = item / (p.tiles_n * p.tiles_m);
        const int z = p.nbatch > 1 ? 0 : outer;
        const int bt = p.nbatch > 1 ? outer : 0;
        const int bo = bt / p.batch_inner, bi = bt % p.batch_inner;
        const int a_c2 = bo * p.a_c2_outer + bi * p.a_c2_inner, a_col = p.a_col0 + bi * p.a_col_inner;
        const int b_c2 = bo * p.b_c2_outer + bi * p.b_c2_inner, b_col = p.b_col0 + bi * p.b_col_inner;
        const int kb0 = z * p.kb_per_split;
        const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          GDBG(0, 1);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          GDBG(0, 2);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + NA * A_PLANE_BYTES;
          if constexpr (SPLIT) {
            // region r of this k-block -> its own stage: [A region r: 128 rows][B region r: BN / 2 rows]
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              if (r == 1) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                sa = smem + stage * C::STAGE_BYTES;
              }
              if (elect_one()) {
                if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
                const uint32_t lead_bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
                tma_load_4d_pair(sa, &map_a, lead_bar, a_col + kb * BK, m_blk * BM, a_c2, r);
                tma_load_4d_pair(sa + A_PLANE_BYTES, &map_b, lead_bar, b_col + kb * BK, n_blk * BN + rank * (BN / 2), b_c2, r);
              }
              __syncwarp();
              if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
            continue;
          } else if constexpr (CG == 2) {
            // both CTAs' bytes complete on the leader's full barrier (its MMA warp is the only consumer)
            if (elect_one()) {
              if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
              const uint32_t lead_bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
#pragma unroll
              for (int pa = 0; pa < NA; ++pa)
                tma_load_4d_pair(sa + pa * A_PLANE_BYTES, &map_a, lead_bar, a_col + kb * BK, m_blk * BM, a_c2, pa);
#pragma unroll
              for (int pb = 0; pb < NB; ++pb)
                tma_load_4d_pair(sb + pb * C::B_PLANE_BYTES, &map_b, lead_bar, b_col + kb * BK, n_blk * BN + rank * (BN / 2), b_c2, pb);
            }
            __syncwarp();
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          if (elect_one()) {
            mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
#pragma unroll
            for (int pa = 0; pa < NA; ++pa) {
              if (!A_MN) {
                tma_load_4d(sa + pa * A_PLANE_BYTES, &map_a, &full_bar[stage], a_col + kb * BK, m_blk * BM, a_c2, pa);
              } else {
#pragma unroll
                for (int j = 0; j < BM / 64; ++j)
                  tma_load_4d(sa + pa * A_PLANE_BYTES + j * 8192, &map_a, &full_bar[stage], a_col + m_blk * BM + j * 64,
                              kb * BK, a_c2, pa);
              }
            }
#pragma unroll
            for (int pb = 0; pb < NB; ++pb) {
              if (!B_MN) {
                tma_load_4d(sb + pb * C::B_PLANE_BYTES, &map_b, &full_bar[stage], b_col + kb * BK, n_blk * BN, b_c2, pb);
              } else {
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                  tma_load_4d(sb + pb * C::B_PLANE_BYTES + j * 8192, &map_b, &full_bar[stage], b_col + n_blk * BN + j * 64,
                              kb * BK, b_c2, pb);
              }
            }
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
#ifndef QV_GEMM_SINGLE_LANE_ISSUE
    // =============================== MMA issuer ===============================
    // The whole warp walks the tile / k-block schedule (barrier waits included) and ONE elected lane issues the MMAs and
    // commits.  With warp-uniform control flow and descriptors formed as "stage base + compile-time offset", ptxas keeps the
    // operands in uniform registers: ~4 instructions per tcgen05.mma.  Under an `if (lane == 0)` region every MMA costs ~21
    // (descriptor arithmetic in vector registers + an ELECT / 5 x R2UR waterfall loop): ~210 instructions per k-block on ONE
    // thread, as long as the k-block's 8 MMAs take to execute -- and that thread shares its scheduler with the epilogue warps of
    // TMEM lane quarter 1, so a heavy epilogue (GELU + operand split) slows the MMA stream itself (timeline: the issuing thread's
    // busy time grows from 212 to 279 us on the teacher's fc1 when the epilogue changes from fp32 to GELU + mixed planes, while it
    // waits for tmem_empty only 1 %).
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM * CG, BN, A_MN, B_MN);
      constexpr uint32_t a_lbo = A_MN ? 8192u : 16u, b_lbo = B_MN ? 8192u : 16u;
      constexpr uint32_t a_kadv = A_MN ? (UMMA_K * 128u) : (UMMA_K * 2u);   // bytes per 16-deep k step
      constexpr uint32_t b_kadv = B_MN ? (UMMA_K * 128u) : (UMMA_K * 2u);
      auto mma16 = [](uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
        if constexpr (CG == 2) umma_bf16_pair(d, da, db, id, acc); else umma_bf16(d, da, db, id, acc);
      };
      auto mma8 = [](uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
        if constexpr (CG == 2) umma_f8_pair(d, da, db, id, acc); else umma_f8(d, da, db, id, acc);
      };
      auto commit = [](uint64_t* bar) {
        if constexpr (CG == 2) umma_commit_pair(bar); else umma_commit(bar);
      };
      // descriptors of stage 0 (low 14 bits of the 64-bit descriptor = byte address >> 4: a stage / plane / k step is an ADD)
      const uint64_t dA0 = umma_smem_desc(smem_u32(smem), a_lbo, 1024u);
      const uint64_t dB0 = umma_smem_desc(smem_u32(smem), b_lbo, 1024u);
      const uint64_t dK0 = umma_smem_desc(smem_u32(smem), 16u, 1024u);       // K-major (mixed / split paths)
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int item = item0; item < num_items; item += item_step, ++local) {
        const int z = p.nbatch > 1 ? 0 : item / (p.tiles_n * p.tiles_m);
        const int kb0 = z * p.kb_per_split;
        const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
        const int buf = local & 1;
        const uint32_t use = static_cast<uint32_t>(local >> 1);      // n-th use of this TMEM buffer
        mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t so = static_cast<uint64_t>(static_cast<uint32_t>(stage) * static_cast<uint32_t>(C::STAGE_BYTES >> 4));
          const uint32_t first = (kb > kb0) ? 1u : 0u;
          if constexpr (SPLIT) {
            // stage = fp16 region: the main product; next stage = fp8 region: the two cross terms
            constexpr uint32_t idesc16 = umma_idesc_f16(BM * CG, BN);
            constexpr uint32_t idesc_hl = umma_idesc_f8(BM * CG, BN, 1u, 1u);
            constexpr uint32_t idesc_lh = umma_idesc_f8(BM * CG, BN, 1u, 0u);
            if (elect_one()) {
              const uint64_t da = dK0 + so, db = da + (A_PLANE_BYTES >> 4);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) mma16(d_tmem, da + 2 * k, db + 2 * k, idesc16, k > 0 ? 1u : first);
              commit(&empty_bar[stage]);
            }
            __syncwarp();
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t so8 = static_cast<uint64_t>(static_cast<uint32_t>(stage) * static_cast<uint32_t>(C::STAGE_BYTES >> 4));
            if (elect_one()) {
              const uint64_t da = dK0 + so8, db = da + (A_PLANE_BYTES >> 4);
#pragma unroll
              for (int k = 0; k < 2; ++k) mma8(d_tmem, da + 2 * k, db + 4 + 2 * k, idesc_hl, 1u);          // A hi8 x B lo8 (64 bytes on)
#pragma unroll
              for (int k = 0; k < 2; ++k) mma8(d_tmem, da + 4 + 2 * k, db + 2 * k, idesc_lh, 1u);          // A lo8 x B hi8
              commit(&empty_bar[stage]);
            }
            __syncwarp();
          } else if constexpr (MIX) {
            constexpr uint32_t idesc16 = umma_idesc_f16(BM * CG, BN);
            constexpr uint32_t idesc_hl = umma_idesc_f8(BM * CG, BN, 1u, 1u);    // A hi8 e5m2 x B lo8 e5m2
            constexpr uint32_t idesc_lh = umma_idesc_f8(BM * CG, BN, 1u, 0u);    // A lo8 e5m2 x B hi8 e4m3
            if (elect_one()) {
              const uint64_t da = dK0 + so, db = da + ((NA * A_PLANE_BYTES) >> 4);
              const uint64_t da8 = da + (A_PLANE_BYTES >> 4), db8 = db + (C::B_PLANE_BYTES >> 4);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) mma16(d_tmem, da + 2 * k, db + 2 * k, idesc16, k > 0 ? 1u : first);
#pragma unroll
              for (int k = 0; k < 2; ++k) mma8(d_tmem, da8 + 2 * k, db8 + 4 + 2 * k, idesc_hl, 1u);
#pragma unroll
              for (int k = 0; k < 2; ++k) mma8(d_tmem, da8 + 4 + 2 * k, db8 + 2 * k, idesc_lh, 1u);
              commit(&empty_bar[stage]);     // frees this smem stage (in both CTAs of a pair) once the MMAs have read it
            }
            __syncwarp();
          } else {
            if (elect_one()) {
              const uint64_t sa = dA0 + so, sb = dB0 + so + ((NA * A_PLANE_BYTES) >> 4);
#pragma unroll
              for (int pr = 0; pr < C::NPAIRS; ++pr) {
                // pair order: (0,0) [,(0,1)] [,(1,0)]  -- for NB == 1 the second pair is (1,0)
                const int pa = (NB == 1) ? pr : (pr == 2 ? 1 : 0);
                const int pb = (NB == 1) ? 0 : (pr == 1 ? 1 : 0);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                  const uint64_t da = sa + ((pa * A_PLANE_BYTES + k * a_kadv) >> 4);
                  const uint64_t db = sb + ((pb * C::B_PLANE_BYTES + k * b_kadv) >> 4);
                  mma16(d_tmem, da, db, idesc, (pr > 0 || k > 0) ? 1u : first);
                }
              }
              commit(&empty_bar[stage]);
            }
            __syncwarp();
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) commit(&tmem_full[buf]);        // accumulator complete -> epilogue (of both CTAs of a pair)
        __syncwarp();
      }
    }
#else
    // =============================== MMA issuer (A/B build: make alt) ===============================
    // One thread walks the schedule and issues: the default until round 2.  Every tcgen05.mma then pays descriptor arithmetic in
    // vector registers + an ELECT / R2UR waterfall (~21 instructions); measured against the warp-uniform issuer above on the teacher
    // shapes: fc1 + GELU + mixed planes 451 vs 425 us, fp32 output 378 vs 367 us (profiles/r02_summary.md).
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM * CG, BN, A_MN, B_MN);
      constexpr uint32_t a_lbo = A_MN ? 8192u : 16u, b_lbo = B_MN ? 8192u : 16u;
      constexpr uint32_t a_kadv = A_MN ? (UMMA_K * 128u) : (UMMA_K * 2u);   // bytes per 16-deep k step
      constexpr uint32_t b_kadv = B_MN ? (UMMA_K * 128u) : (UMMA_K * 2u);
      // CG = 2: one instruction drives both SMs (M = 256); descriptors are offsets valid in BOTH CTAs' shared memory
      auto mma16 = [](uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
        if constexpr (CG == 2) umma_bf16_pair(d, da, db, id, acc); else umma_bf16(d, da, db, id, acc);
      };
      auto mma8 = [](uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
        if constexpr (CG == 2) umma_f8_pair(d, da, db, id, acc); else umma_f8(d, da, db, id, acc);
      };
      auto commit = [](uint64_t* bar) {
        if constexpr (CG == 2) umma_commit_pair(bar); else umma_commit(bar);
      };
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
#ifdef QV_ATTN_DEBUG
      int gdbg_n = 0;
#endif
      for (int item = item0; item < num_items; item += item_step, ++local) {
        const int z = p.nbatch > 1 ? 0 : item / (p.tiles_n * p.tiles_m);
        const int kb0 = z * p.kb_per_split;
        const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
        const int buf = local & 1;
        const uint32_t use = static_cast<uint32_t>(local >> 1);      // n-th use of this TMEM buffer
        GDBG(1, 10);
        mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);
        GDBG(1, 11);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          GDBG(1, 12);
          mbar_wait(&full_bar[stage], phase);
          GDBG(1, 13);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + NA * A_PLANE_BYTES;
          if constexpr (SPLIT) {
            // stage = fp16 region: the main product; next stage = fp8 region: the two cross terms (see the MIX branch below)
            constexpr uint32_t idesc16 = umma_idesc_f16(BM * CG, BN);
            constexpr uint32_t idesc_hl = umma_idesc_f8(BM * CG, BN, 1u, 1u);
            constexpr uint32_t idesc_lh = umma_idesc_f8(BM * CG, BN, 1u, 0u);
            const uint32_t sb0 = sa + A_PLANE_BYTES;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              mma16(d_tmem, umma_smem_desc(sa + k * 32u, 16u, 1024u), umma_smem_desc(sb0 + k * 32u, 16u, 1024u), idesc16,
                    (kb > kb0 || k > 0) ? 1u : 0u);
            commit(&empty_bar[stage]);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            GDBG(1, 12);
            mbar_wait(&full_bar[stage], phase);
            GDBG(1, 13);
            tc_fence_after();
            const uint32_t sa8 = smem_u32(smem + stage * C::STAGE_BYTES), sb8 = sa8 + A_PLANE_BYTES;
#pragma unroll
            for (int k = 0; k < 2; ++k)
              mma8(d_tmem, umma_smem_desc(sa8 + k * 32u, 16u, 1024u), umma_smem_desc(sb8 + 64u + k * 32u, 16u, 1024u), idesc_hl, 1u);
#pragma unroll
            for (int k = 0; k < 2; ++k)
              mma8(d_tmem, umma_smem_desc(sa8 + 64u + k * 32u, 16u, 1024u), umma_smem_desc(sb8 + k * 32u, 16u, 1024u), idesc_lh, 1u);
          } else if constexpr (MIX) {
            // region 0 of each operand: fp16 (x 2^5 / x 2^9); region 1: per 64-deep k-block a 128-byte row = 64 fp8 of the
            // value (hi8) then 64 fp8 of the fp16 rounding residual (lo8).  fp32-grade product = hi16.hi16 (4 x K16, kind::f16)
            // + hi8.lo8 + lo8.hi8 (2 x K32 each, kind::f8f6f4 at twice the rate), every term scaled by 2^14 into ONE accumulator.
            constexpr uint32_t idesc16 = umma_idesc_f16(BM * CG, BN);
            constexpr uint32_t idesc_hl = umma_idesc_f8(BM * CG, BN, 1u, 1u);    // A hi8 e5m2 x B lo8 e5m2
            constexpr uint32_t idesc_lh = umma_idesc_f8(BM * CG, BN, 1u, 0u);    // A lo8 e5m2 x B hi8 e4m3
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t da = umma_smem_desc(sa + k * 32u, 16u, 1024u);
              const uint64_t db = umma_smem_desc(sb + k * 32u, 16u, 1024u);
              mma16(d_tmem, da, db, idesc16, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            const uint32_t sa8 = sa + A_PLANE_BYTES, sb8 = sb + C::B_PLANE_BYTES;
#pragma unroll
            for (int k = 0; k < 2; ++k)
              mma8(d_tmem, umma_smem_desc(sa8 + k * 32u, 16u, 1024u), umma_smem_desc(sb8 + 64u + k * 32u, 16u, 1024u), idesc_hl, 1u);
#pragma unroll
            for (int k = 0; k < 2; ++k)
              mma8(d_tmem, umma_smem_desc(sa8 + 64u + k * 32u, 16u, 1024u), umma_smem_desc(sb8 + k * 32u, 16u, 1024u), idesc_lh, 1u);
          } else {
#pragma unroll
          for (int pr = 0; pr < C::NPAIRS; ++pr) {
            // pair order: (0,0) [,(0,1)] [,(1,0)]  -- for NB == 1 the second pair is (1,0)
            const int pa = (NB == 1) ? pr : (pr == 2 ? 1 : 0);
            const int pb = (NB == 1) ? 0 : (pr == 1 ? 1 : 0);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t da = umma_smem_desc(sa + pa * A_PLANE_BYTES + k * a_kadv, a_lbo, 1024u);
              const uint64_t db = umma_smem_desc(sb + pb * C::B_PLANE_BYTES + k * b_kadv, b_lbo, 1024u);
              mma16(d_tmem, da, db, idesc, (kb > kb0 || pr > 0 || k > 0) ? 1u : 0u);
            }
          }
          }
          commit(&empty_bar[stage]);            // frees this smem stage (in both CTAs of a pair) once the MMAs have read it
          GDBG(1, 14);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        commit(&tmem_full[buf]);                // accumulator complete -> epilogue (of both CTAs of a pair)
      }
    }
#endif
  } else {
    // =============================== epilogue (8 warps) ===============================
    // Two warps per TMEM lane quarter: warp (q, par) takes the column chunks whose index has parity `par`, so every SM
    // sub-partition always has a second epilogue warp to switch to while the other waits on TMEM / TMA latencies.
    // Per chunk: tcgen05.ld (the next chunk's load is issued before this one is processed) -> per-column scale / bias
    // (staged once per chunk in smem, read back as broadcast LDS.128) -> [GELU] -> observer min/max ->
    // 128B-swizzled smem staging -> TMA store.
    const int ew = warp - 2;
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int par = ew >> 2;                     // chunk parity owned by this warp
    if constexpr (EPI == 2) {
      // ---- gradient-planes epilogue, in units of 32 columns (the two warps of a lane quarter take alternate units).
      // The raw y tile of a unit (32 rows x 32 fp32) arrives by TMA in the warp's staging buffer -- coalesced, off the LSU
      // path -- one unit ahead: as soon as the lanes have copied their row to registers the next unit's load is issued, for
      // an item's first unit before its accumulator is even complete.  Planes leave through 64-byte-swizzled staging and TMA.
      constexpr int NU = BN / 32;
      // y tile in: 2 x 4 KB (128B-swizzled rows of 32 fp32), alternating, so the TMA write of unit n+1 never lands in the
      // buffer whose ld.shared reads of unit n may still be in flight (observed as stale last chunks with one buffer)
      uint8_t* my_raw = smem_epi + ew * EPI_WARP_BYTES;
      uint8_t* my_out = my_raw + 8192;                            // 2 KB hi + 2 KB lo: 32 rows x 64 B, 64B swizzle
      uint64_t* my_bar = &raw_bar[ew];
      const QvQParams yq = qv_load_qparams(p.ep_scale, p.ep_zp, p.ep_qmin, p.ep_qmax);
      // clamp bounds in the magic domain (exact: small integers added to 1.5 * 2^23) and the table base folded with the bits of the
      // lower bound (index 0), all modulo 2^32
      const float t_lo = 12582912.0f + (yq.qmin - 1.0f - yq.zp), t_hi = 12582912.0f + (yq.qmax + 1.0f - yq.zp);
      const uint32_t lut_off = smem_u32(gelu_lut) + static_cast<uint32_t>(lane & 7) * 4u - (__float_as_uint(t_lo) << 5);
      auto issue_raw = [&](int item, int u, uint32_t slot) {      // lane 0 only
        const int n_blk = item % p.tiles_n;
        const int m_blk = m_of(item);
        mbar_expect_tx(my_bar, 4096);
        tma_load_3d(my_raw + slot * 4096u, &map_y, my_bar, n_blk * BN + u * 32, m_blk * BM + q * 32, 0);
      };
      uint32_t raw_phase = 0;             // also the slot of the tile being waited for (loads alternate slots)
      int local = 0;
#ifdef QV_ATTN_DEBUG
      int gdbg_n = (threadIdx.x == 64) ? 0 : 4096;
#endif
      if (item0 < num_items && lane == 0) issue_raw(item0, par, 0);
      for (int item = item0; item < num_items; item += item_step, ++local) {
        const int n_blk = item % p.tiles_n;
        const int m_blk = m_of(item);
        const int buf = local & 1;
        const uint32_t use = static_cast<uint32_t>(local >> 1);
        const int row0 = m_blk * BM + q * 32;
        const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN);
        bool acc_ready = false;
#pragma unroll 1
        for (int u = par; u < NU; u += 2) {
          // ---- this unit's y values: smem -> registers, then hand the buffer to the next unit's load ----
          GDBG(2, 30);
          mbar_wait(my_bar, raw_phase);
          GDBG(2, 31);
          float4 yv[8];
          {
            const uint32_t srow = smem_u32(my_raw) + raw_phase * 4096u + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t addr = srow + (static_cast<uint32_t>(j ^ (lane & 7)) << 4);
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(yv[j].x), "=f"(yv[j].y), "=f"(yv[j].z), "=f"(yv[j].w)
                           : "r"(addr) : "memory");
            }
          }
          raw_phase ^= 1;
          __syncwarp();
          const bool last = u + 2 >= NU;
          if (lane == 0) {
            if (!last) issue_raw(item, u + 2, raw_phase);
            else if (item + item_step < num_items) issue_raw(item + item_step, par, raw_phase);
          }
          // this unit's 32 weight scales (every lane the same 128 bytes: broadcast loads), issued before the accumulator wait
          float4 csv[8];
          {
            const int64_t n0s = static_cast<int64_t>(n_blk) * BN + u * 32;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              csv[j] = (p.col_scale && n0s < p.N) ? __ldg(reinterpret_cast<const float4*>(p.col_scale + n0s + 4 * j))     // N % 64 == 0: whole units
                                                  : make_float4(1.f, 1.f, 1.f, 1.f);
          }
          if (!acc_ready) {
            GDBG(2, 20);
            mbar_wait(&tmem_full[buf], use & 1);
            GDBG(2, 21);
            tc_fence_after();
            acc_ready = true;
          }
          uint32_t rr[32];
          tmem_ld_cols<32>(t_base + u * 32, rr);
          tmem_ld_wait();
          GDBG(2, 32);
          if (last) {                                             // this warp's last TMEM read of the buffer
            tc_fence_before();
            tmem_empty_arrive(buf);
          }
          const int64_t n0 = static_cast<int64_t>(n_blk) * BN + u * 32;
          if (n0 >= p.N || static_cast<int64_t>(row0) >= p.M) continue;          // warp-uniform
          if (lane == 0) tma_store_wait_read<0>();                // previous planes have left the staging buffer
          __syncwarp();
          GDBG(2, 33);
          // Per element (8 instructions + the operand split; was 24): t = y * inv + 1.5 * 2^23 is the magic-number round-to-nearest-
          // even (exact for |y * inv| < 2^22, still far out of range beyond; no FRND): its mantissa IS rint(y * inv), so the code is
          // clamped in that domain -- [qmin - 1 - zp, qmax + 1 - zp] + 1.5 * 2^23, floats compare like the integers they hold -- and
          // the clamped float's BITS, shifted left by 5 (modulo 2^32) plus a constant, are the address of the table entry (8 copies x
          // 4 bytes per code; no F2I, no subtract).  The table holds gelu' * STE mask, so there is no compare / select either.
          const uint32_t srow_hi = smem_u32(my_out) + lane * 64, srow_lo = srow_hi + 2048;
          // column-sum scratch: the y slot this unit has just emptied (its next TMA refill is two units away)
          const uint32_t cs_row = smem_u32(my_raw) + (raw_phase ^ 1u) * 4096u + lane * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {                           // 8 columns per step
            const float4 sa = csv[2 * j], sb = csv[2 * j + 1];
            const float4 ya = yv[2 * j], yb = yv[2 * j + 1];
            const float yy[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
            const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
            float a[8], gq[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float t = __fadd_rn(__fmul_rn(yy[e], yq.inv), 12582912.0f);
              const float tc = fminf(fmaxf(t, t_lo), t_hi);
              float l;
              asm("ld.shared.f32 %0, [%1];" : "=f"(l) : "r"((__float_as_uint(tc) << 5) + lut_off));
              const float f = __uint_as_float(rr[8 * j + e]) * l;
              gq[e] = f;
              a[e] = __fmaf_rn(f, sc[e], 0.0f);       // + 0: a masked element (f = -0 for a negative accumulator) leaves as +0, like the select did
            }
            if (p.ep_colsum) {
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cs_row + (static_cast<uint32_t>((2 * j) ^ (lane & 7)) << 4)),
                           "f"(gq[0]), "f"(gq[1]), "f"(gq[2]), "f"(gq[3]) : "memory");
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cs_row + (static_cast<uint32_t>((2 * j + 1) ^ (lane & 7)) << 4)),
                           "f"(gq[4]), "f"(gq[5]), "f"(gq[6]), "f"(gq[7]) : "memory");
            }
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a0 = a[2 * e], a1 = a[2 * e + 1];
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(a0, a1);
              const uint32_t hbits = *reinterpret_cast<const uint32_t*>(&h2);
              const float r0 = a0 - __uint_as_float(hbits << 16), r1 = a1 - __uint_as_float(hbits & 0xffff0000u);
              const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
              hi[e] = hbits;
              lo[e] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            // 64B swizzle: 16-byte chunk j of row `lane` lives at chunk j ^ ((lane >> 1) & 3)
            const uint32_t sw = (static_cast<uint32_t>(j ^ ((lane >> 1) & 3)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow_hi + sw), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]),
                         "r"(hi[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow_lo + sw), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]),
                         "r"(lo[3]) : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&map_o, my_out, static_cast<int>(n0), row0, 0, 0);
            tma_store_4d(&map_o, my_out + 2048, static_cast<int>(n0), row0, 0, 1);
            tma_store_commit();
          }
          GDBG(2, 34);
          if (p.ep_colsum) {
            // bias-grad partial of column `lane` over this slab's 32 rows: read the fp32 tile back column-wise (row i keeps its
            // 16-byte chunk k at k ^ (i & 7): the 32 lanes of one load hit 32 different banks), summed in row order -> deterministic
            const uint32_t cs_base = smem_u32(my_raw) + (raw_phase ^ 1u) * 4096u + static_cast<uint32_t>(lane & 3) * 4u;
            const uint32_t ck = static_cast<uint32_t>(lane >> 2);
            float c4[4] = {0.f, 0.f, 0.f, 0.f};             // four interleaved partial sums (fixed order): no 32-deep dependent chain
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float t;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(cs_base + i * 128 + ((ck ^ static_cast<uint32_t>(i & 7)) << 4)) : "memory");
              c4[i & 3] += t;
            }
            const float cs = (c4[0] + c4[1]) + (c4[2] + c4[3]);
            p.ep_colsum[(static_cast<int64_t>(m_blk) * 4 + q) * p.N + n0 + lane] = cs;
          }
          GDBG(2, 35);
        }
      }
      if (lane == 0) tma_store_wait_read<0>();
    } else if constexpr (EPI == 1) {
    // ---- plane output (the next GEMM's operand), 16 warps: warp (q, par) takes the 32-column chunks ch = par, par + 4, ... of its
    // TMEM lane quarter.  Per chunk: tcgen05.ld (the next chunk's load already in flight) -> bias [-> GELU] -> operand split ->
    // 64B-swizzled staging -> TMA stores (hi / lo plane, or fp16 region + the block's hi8 and lo8 byte runs).
    constexpr int NCHUNK = BN / 32;
    static_assert(NCHUNK >= 4, "every one of the four warps of a lane quarter needs a chunk (it arrives on tmem_empty once per tile)");
    const int par4 = ew >> 2;
    uint8_t* my_epi = smem_epi + ew * EPI_WARP_BYTES;
    __half2 amax2 = __floats2half2_rn(0.f, 0.f);  // mixed plane output: largest |fp16 image| (value x 2^7) this thread has written
    int local = 0;
#ifdef QV_ATTN_DEBUG
    int gdbg_n = (threadIdx.x == 64) ? 0 : 4096;
#endif
    for (int item = item0; item < num_items; item += item_step, ++local) {
      const int n_blk = item % p.tiles_n;
      const int m_blk = m_of(item);
      const int outer = item / (p.tiles_n * p.tiles_m);
      const int bt = p.nbatch > 1 ? outer : 0;
      const int bo = bt / p.batch_inner, bi = bt % p.batch_inner;
      const int o_c2 = bo * p.o_c2_outer + bi * p.o_c2_inner;
      const int o_col = p.o_col0 + bi * p.o_col_inner;
      const int buf = local & 1;
      const uint32_t use = static_cast<uint32_t>(local >> 1);
      GDBG(2, 20);
      mbar_wait(&tmem_full[buf], use & 1);
      GDBG(2, 21);
      tc_fence_after();
      const int row0 = m_blk * BM + q * 32;
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN);
      uint32_t rr[32], nxt[32];
      tmem_ld_cols<32>(t_base + par4 * 32, nxt);
#pragma unroll 1
      for (int ch = par4; ch < NCHUNK; ch += 4) {
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) rr[j] = nxt[j];
        if (ch + 4 < NCHUNK) {
          tmem_ld_cols<32>(t_base + (ch + 4) * 32, nxt);
        } else {                                                // this warp's last chunk is in registers
          tc_fence_before();
          tmem_empty_arrive(buf);
          GDBG(2, 22);
        }
        const int64_t n0 = static_cast<int64_t>(n_blk) * BN + ch * 32;
        if (n0 >= p.N || static_cast<int64_t>(row0) >= p.M) continue;      // warp-uniform: nothing to store
        if (lane == 0) tma_store_wait_read<0>();                // this warp's previous stores have read the staging buffer
        __syncwarp();
        const uint32_t srow0 = smem_u32(my_epi) + lane * 64;    // region 0 / hi plane: 32 rows x 64 bytes, 64B swizzle
        const uint32_t swz = static_cast<uint32_t>((lane >> 1) & 3);
        uint32_t h8w[8], l8w[8];                                // mixed: this row's 32 hi8 / lo8 bytes
#pragma unroll
        for (int j = 0; j < 4; ++j) {                           // 8 columns per step
          // plane output takes bias only (checked on the host); every lane reads the same 32 bytes: broadcast LDG.128
          float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bb = ba;
          if (p.bias) {
            ba = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + 8 * j));
            bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + 8 * j + 4));
          }
          const float as = p.acc_scale;
          float a[8] = {__uint_as_float(rr[8 * j]) * as + ba.x,     __uint_as_float(rr[8 * j + 1]) * as + ba.y,
                        __uint_as_float(rr[8 * j + 2]) * as + ba.z, __uint_as_float(rr[8 * j + 3]) * as + ba.w,
                        __uint_as_float(rr[8 * j + 4]) * as + bb.x, __uint_as_float(rr[8 * j + 5]) * as + bb.y,
                        __uint_as_float(rr[8 * j + 6]) * as + bb.z, __uint_as_float(rr[8 * j + 7]) * as + bb.w};
          const uint32_t sw = (static_cast<uint32_t>(j) ^ swz) << 4;
          if (p.out_fmt == 1) {                                 // mixed planes: the next GEMM's fp16 + fp8 operand
            uint32_t h16[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint64_t x2 = qv2_pack(a[2 * e], a[2 * e + 1]);
              const uint64_t s2 = (p.act == 1) ? qv_gelu128_pair(x2) : qv2_mul(x2, qv2_pack(QV_MIX_SCALE, QV_MIX_SCALE));
              uint32_t ph, pl;
              h16[e] = qv_mix_split2_scaled<QV_MIX_ACT>(s2, ph, pl);
              const __half2 ah = __habs2(*reinterpret_cast<const __half2*>(&h16[e]));      // range guard on the fp16 image (x 2^7)
              amax2 = __hmax2(amax2, ah);
              if (e & 1) { h8w[2 * j + (e >> 1)] |= ph << 16; l8w[2 * j + (e >> 1)] |= pl << 16; }
              else { h8w[2 * j + (e >> 1)] = ph; l8w[2 * j + (e >> 1)] = pl; }
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow0 + sw), "r"(h16[0]), "r"(h16[1]), "r"(h16[2]), "r"(h16[3]) : "memory");
            continue;
          }
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float a0 = a[2 * e], a1 = a[2 * e + 1];
            if (p.act == 1) { a0 = gelu_erf(a0); a1 = gelu_erf(a1); }
            // packed split: one F2FP for the hi pair, exact residuals, one F2FP for the lo pair
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(a0, a1);
            const uint32_t hbits = *reinterpret_cast<const uint32_t*>(&h2);
            const float r0 = a0 - __uint_as_float(hbits << 16), r1 = a1 - __uint_as_float(hbits & 0xffff0000u);
            const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
            hi[e] = hbits;
            lo[e] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow0 + sw), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow0 + 2048 + sw), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
        }
        if (p.out_fmt == 1) {
          // region 1 of this 64-column block is one 128-byte line per row = 64 hi8 | 64 lo8; this chunk is its half (ch & 1):
          // two 32-row x 32-byte tiles (linear), stored as 16-element boxes of the bf16-typed tensor
          const uint32_t r1 = smem_u32(my_epi) + 2048 + lane * 32;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(r1), "r"(h8w[0]), "r"(h8w[1]), "r"(h8w[2]), "r"(h8w[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(r1 + 16), "r"(h8w[4]), "r"(h8w[5]), "r"(h8w[6]), "r"(h8w[7]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(r1 + 1024), "r"(l8w[0]), "r"(l8w[1]), "r"(l8w[2]), "r"(l8w[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(r1 + 1040), "r"(l8w[4]), "r"(l8w[5]), "r"(l8w[6]), "r"(l8w[7]) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int c = o_col + static_cast<int>(n0);
          tma_store_4d(&map_o, my_epi, c, row0, o_c2, 0);
          if (p.out_fmt == 1) {
            const int blk = c & ~63, half = (c >> 5) & 1;
            tma_store_4d(&map_y, my_epi + 2048, blk + 16 * half, row0, o_c2, 1);
            tma_store_4d(&map_y, my_epi + 3072, blk + 32 + 16 * half, row0, o_c2, 1);
          } else {
            tma_store_4d(&map_o, my_epi + 2048, c, row0, o_c2, 1);
          }
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_read<0>();
    {
      // the fp16 image saturates at 65504 >= 448 * 2^7 = 57344, so a clamped value still trips the guard
      const float amax = fmaxf(__low2float(amax2), __high2float(amax2)) * (1.0f / QV_MIX_SCALE);
      if (p.sat_flag && __any_sync(0xffffffffu, amax > QV_MIX_ACT_MAX) && lane == 0) atomicOr(p.sat_flag, p.sat_bit);
    }
    } else {
    constexpr int CW = 32;                       // EPI 0: columns per chunk (one 128-byte row of fp32)
    constexpr int NCHUNK = BN / CW;
    uint8_t* my_epi = smem_epi + ew * EPI_WARP_BYTES;
    float* my_terms = reinterpret_cast<float*>(smem_epi + 8 * EPI_WARP_BYTES) + ew * 64;   // EPI 0: [mult 32][bias 32]
    float mn = INFINITY, mx = -INFINITY;
    const float alpha = p.alpha ? __ldg(p.alpha) : 1.0f;
    const bool raw = p.splits > 1;
    int local = 0;
#ifdef QV_ATTN_DEBUG
    int gdbg_n = (threadIdx.x == 64) ? 0 : 4096;
#endif
    for (int item = item0; item < num_items; item += item_step, ++local) {
      const int n_blk = item % p.tiles_n;
      const int m_blk = m_of(item);
      const int outer = item / (p.tiles_n * p.tiles_m);
      const int bt = p.nbatch > 1 ? outer : 0;
      const int bo = bt / p.batch_inner, bi = bt % p.batch_inner;
      const int o_c2 = raw ? outer : bo * p.o_c2_outer + bi * p.o_c2_inner;
      const int o_col = raw ? 0 : p.o_col0 + bi * p.o_col_inner;
      const int buf = local & 1;
      const uint32_t use = static_cast<uint32_t>(local >> 1);
      GDBG(2, 20);
      mbar_wait(&tmem_full[buf], use & 1);
      GDBG(2, 21);
      tc_fence_after();
      const int row0 = m_blk * BM + q * 32;                     // first row of this warp's 32-row slab
      const bool row_ok = static_cast<int64_t>(row0 + lane) < p.M;
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN);
      // register double buffer: `nxt` receives the next chunk's tcgen05.ld while `rr` is processed.  The chunk loop is NOT
      // unrolled (the body is ~3 K instructions; unrolling it thrashed the instruction cache: ncu `no_instruction` stalls).
      uint32_t rr[CW], nxt[CW];
      tmem_ld_cols<CW>(t_base + par * CW, nxt);
#pragma unroll 1
      for (int ch = par; ch < NCHUNK; ch += 2) {
        const int c0 = ch * CW;
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < CW; ++j) rr[j] = nxt[j];
        if (ch + 2 < NCHUNK) {
          tmem_ld_cols<CW>(t_base + (ch + 2) * CW, nxt);
        } else {                                                // this warp's last chunk is in registers
          tc_fence_before();
          tmem_empty_arrive(buf);
          GDBG(2, 22);
        }
        const int64_t n0 = static_cast<int64_t>(n_blk) * BN + c0;
        if (n0 >= p.N || static_cast<int64_t>(row0) >= p.M) continue;      // warp-uniform: nothing to store
        if (lane == 0) tma_store_wait_read<0>();                // this warp's previous store has read the staging buffer
        __syncwarp();
        if constexpr (EPI == 0) {
          if (!raw) {                                           // lane j fetches the terms of column n0 + j; smem broadcast
            const int64_t n = n0 + lane;
            float m_ = 1.0f, b_ = 0.0f;
            if (n < p.N) {
              m_ = alpha * p.acc_scale;
              if (p.col_scale) m_ *= __ldg(p.col_scale + n);
              if (p.col_rscale) m_ = __fdiv_rn(m_, __ldg(p.col_rscale + n));
              if (p.bias) b_ = __ldg(p.bias + n);
            }
            my_terms[lane] = m_;
            my_terms[CW + lane] = b_;
            __syncwarp();
          }
        }
        if constexpr (EPI == 0) {
          const int ncols = static_cast<int>(min(static_cast<int64_t>(32), p.N - n0));   // valid columns of this chunk
          const uint32_t srow = smem_u32(my_epi) + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 v = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]),
                                   __uint_as_float(rr[4 * j + 3]));
            if (!raw) {
              const float4 m4 = *reinterpret_cast<const float4*>(my_terms + 4 * j);
              const float4 b4 = *reinterpret_cast<const float4*>(my_terms + CW + 4 * j);
              v.x = v.x * m4.x + b4.x; v.y = v.y * m4.y + b4.y; v.z = v.z * m4.z + b4.z; v.w = v.w * m4.w + b4.w;
              if (p.minmax) {
                const float big = INFINITY;
                const float lo0 = (row_ok && 4 * j + 0 < ncols) ? v.x : big, hi0 = (row_ok && 4 * j + 0 < ncols) ? v.x : -big;
                const float lo1 = (row_ok && 4 * j + 1 < ncols) ? v.y : big, hi1 = (row_ok && 4 * j + 1 < ncols) ? v.y : -big;
                const float lo2 = (row_ok && 4 * j + 2 < ncols) ? v.z : big, hi2 = (row_ok && 4 * j + 2 < ncols) ? v.z : -big;
                const float lo3 = (row_ok && 4 * j + 3 < ncols) ? v.w : big, hi3 = (row_ok && 4 * j + 3 < ncols) ? v.w : -big;
                mn = fminf(mn, fminf(fminf(lo0, lo1), fminf(lo2, lo3)));
                mx = fmaxf(mx, fmaxf(fmaxf(hi0, hi1), fmaxf(hi2, hi3)));
              }
            }
            // 128B-swizzled staging: 16-byte chunk j of row `lane` lives at chunk (j ^ (lane & 7))
            const uint32_t addr = srow + (static_cast<uint32_t>(j ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&map_o, my_epi, o_col + static_cast<int>(n0), row0, o_c2);
            tma_store_commit();
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_read<0>();
    if (p.minmax && !raw) {
      mn = qv_warp_min(mn);
      mx = qv_warp_max(mx);
      if (lane == 0 && mn <= mx) {
        atomicMin(p.minmax, qv_f2ord(mn));
        atomicMax(p.minmax + 1, qv_f2ord(mx));
      }
      if (p.obs_ticket && lane == 0) {
        __threadfence();                                          // this warp's min / max are visible before its ticket
        const uint32_t t = atomicAdd(p.obs_ticket, 1u);
        if (t == gridDim.x * 8u - 1u) {                           // every epilogue warp of the grid has merged its range
          __threadfence();
          const uint32_t emn = *reinterpret_cast<volatile uint32_t*>(p.minmax);
          const uint32_t emx = *reinterpret_cast<volatile uint32_t*>(p.minmax + 1);
          qv_observer_step(emn, emx, p.obs_enabled, p.obs_fq_enabled, p.obs_min_val, p.obs_max_val, p.obs_scale,
                           p.obs_zero_point, p.obs_c, p.obs_qmin, p.obs_qmax, p.obs_symmetric);
          *p.obs_ticket = 0u;                                     // re-armed for the next launch on this stream
        }
      }
    }
    }   // EPI != 2
  }

  tc_fence_before();
  __syncwarp();
  if constexpr (CG == 2) cluster_sync_all();   // no MMA still reads the peer's tiles, no arrive is in flight to a CTA that exits
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
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
std::atomic<int64_t> g_pair_launches{0};       // qv_gemm_pair_launches(): GEMMs launched as CTA pairs

template <int BN, int NA, int NB, bool A_MN, bool B_MN, int EPI = 0, int MIX = 0, int CG = 1>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const GemmKParams& kp, int grid,
           cudaStream_t st, const CUtensorMap* my = nullptr) {
  using C = Cfg<BN, NA, NB, EPI, CG, (MIX != 0 && CG == 2)>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(qv_gemm_kernel<BN, NA, NB, A_MN, B_MN, EPI, MIX, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    C::SMEM_BYTES);
  });
  QV_REQUIRE(attr_err == cudaSuccess, QV_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  if constexpr (CG == 2) {       // CTA pairs: clusters of two CTAs (the two SMs of a TPC)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
    cfg.blockDim = dim3(num_threads(EPI), 1, 1);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, qv_gemm_kernel<BN, NA, NB, A_MN, B_MN, EPI, MIX, CG>, ma, mb, mo, my ? *my : mo, kp);
    QV_REQUIRE(e == cudaSuccess, QV_ERR_CUDA, "cudaLaunchKernelEx(cluster 2): %s", cudaGetErrorString(e));
    g_pair_launches.fetch_add(1, std::memory_order_relaxed);
  } else {
    qv_gemm_kernel<BN, NA, NB, A_MN, B_MN, EPI, MIX, CG><<<grid, num_threads(EPI), C::SMEM_BYTES, st>>>(ma, mb, mo, my ? *my : mo, kp);
  }
  return qv_check_launch("qv_gemm_bf16");
}

// CTA pairs (cta_group::2) for the big K-major GEMMs: QV_GEMM_PAIR=0 switches them off (A/B measurement)
int pair_mode() {
  const char* e = getenv("QV_GEMM_PAIR");     // read per launch: tests flip it inside one process
  return e ? atoi(e) : QV_GEMM_PAIR_DEFAULT;
}

template <int BN, int NA, int NB>
int launch_major(bool amn, bool bmn, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo,
                 const GemmKParams& kp, int grid, cudaStream_t st) {
  if (!amn && !bmn) return launch<BN, NA, NB, false, false>(ma, mb, mo, kp, grid, st);
  if (amn && bmn) return launch<BN, NA, NB, true, true>(ma, mb, mo, kp, grid, st);
  if (!amn && bmn) return launch<BN, NA, NB, false, true>(ma, mb, mo, kp, grid, st);
  return qv_set_error(QV_ERR_UNSUPPORTED, "operand layout (A MN-major, B K-major) is not instantiated");
}

template <int BN>
int launch_planes(int na, int nb, bool amn, bool bmn, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo,
                  const GemmKParams& kp, int grid, cudaStream_t st) {
  if (na == 1 && nb == 1) return launch_major<BN, 1, 1>(amn, bmn, ma, mb, mo, kp, grid, st);
  if (na == 2 && nb == 1) return launch_major<BN, 2, 1>(amn, bmn, ma, mb, mo, kp, grid, st);
  if (na == 2 && nb == 2) return launch_major<BN, 2, 2>(amn, bmn, ma, mb, mo, kp, grid, st);
  return qv_set_error(QV_ERR_UNSUPPORTED, "plane combination (%d A planes, %d B planes) is not instantiated", na, nb);
}

// N-tile choice: 192 divides every ViT-S/B projection width (384 .. 3072) with the best operand re-use; small or odd
// widths (attention head slices, score matrices) take 64 / 128.
int pick_bn(int64_t N) {
  if (N <= 64) return 64;
  // (A wave-aware choice -- 128-wide tiles where 192 leaves a partial last wave, e.g. N = 384: 5.3 -> 7.99 waves -- measured
  // SLOWER inside the step, 54.5 vs 53.7 ms: the step is power-bound, idle SMs of a partial wave cost nothing, and the narrower
  // tile re-reads the A operand 1.5x as often.)
  if (N % 192 == 0) return 192;
  if (N <= 128) return 128;
  if (N % 128 == 0) return 128;
  return (N > 1024) ? 192 : 128;
}

}  // namespace

extern "C" int qv_gemm_bf16(const qv_gemm_args* a, void* stream) {
  QV_REQUIRE(a != nullptr, QV_ERR_INVALID, "null args");
  QV_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, QV_ERR_INVALID, "empty gemm (M=%lld N=%lld K=%lld)", (long long)a->M,
             (long long)a->N, (long long)a->K);
  QV_REQUIRE(a->a_planes >= 1 && a->a_planes <= 2 && a->b_planes >= 1 && a->b_planes <= 2, QV_ERR_INVALID,
             "a_planes / b_planes must be 1 or 2");
  const int splits = a->splits > 1 ? a->splits : 1;
  const int nbatch = a->nbatch > 1 ? a->nbatch : 1;
  QV_REQUIRE(!(splits > 1 && nbatch > 1), QV_ERR_UNSUPPORTED, "split-K and batching are mutually exclusive");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  const bool mix = a->mix != 0;
  const bool planes_out = a->out_kind == 1 || a->out_kind == 2;
  int BN = a->tile_n > 0 ? a->tile_n : pick_bn(a->N);
  if (planes_out && a->tile_n <= 0 && BN < 128) BN = 128;      // plane output is instantiated for 128 / 192 wide tiles
  QV_REQUIRE(BN == 64 || BN == 128 || BN == 192, QV_ERR_UNSUPPORTED, "tile_n must be 64, 128 or 192");
  QV_REQUIRE(a->out_kind >= 0 && a->out_kind <= 2, QV_ERR_INVALID, "out_kind must be 0 (fp32), 1 (bf16 hi/lo planes) or 2 (mixed planes)");
  if (mix)
    QV_REQUIRE(a->a_planes == 2 && a->b_planes == 2 && !a->a.mn_major && !a->b.mn_major && splits == 1 && a->K % 64 == 0 &&
                   a->act != 2, QV_ERR_UNSUPPORTED,
               "mixed (fp16 + fp8) operands: unsplit K-major (2,2)-region GEMMs with K a multiple of 64 only");
  QV_REQUIRE(a->act == 0 || ((a->act == 1 || a->act == 2) && planes_out), QV_ERR_UNSUPPORTED,
             "act = GELU / gradient-planes epilogue needs out_kind = 1 (plane output)");
  const bool grad_epi = a->act == 2;
  if (grad_epi) {
    QV_REQUIRE(splits == 1 && nbatch == 1 && a->a_planes == 2 && a->b_planes == 1 && !a->a.mn_major && !a->b.mn_major,
               QV_ERR_UNSUPPORTED, "gradient-planes epilogue is instantiated for unsplit, unbatched K-major (2,1)-plane GEMMs");
    QV_REQUIRE(a->N % 64 == 0, QV_ERR_UNSUPPORTED, "plane output needs N to be a multiple of 64");
    QV_REQUIRE(a->ep_raw && a->ep_scale && a->ep_zp, QV_ERR_INVALID, "gradient-planes epilogue needs ep_raw, ep_scale and ep_zp");
    QV_REQUIRE(qv_aligned16(a->ep_raw) && a->ep_raw_ld >= a->N && a->ep_raw_ld % 4 == 0, QV_ERR_INVALID,
               "ep_raw must be 16-byte aligned with a row pitch >= N that is a multiple of 4 floats");
    QV_REQUIRE(a->ep_qmax >= a->ep_qmin && a->ep_qmax - a->ep_qmin <= 255, QV_ERR_INVALID, "ep_qmin / ep_qmax span more than 256 codes");
    QV_REQUIRE(!a->bias && !a->col_rscale && !a->alpha && !a->minmax, QV_ERR_UNSUPPORTED,
               "gradient-planes epilogue takes col_scale only (no bias / alpha / observer)");
    QV_REQUIRE(!a->col_scale || qv_aligned16(a->col_scale), QV_ERR_INVALID, "col_scale must be 16-byte aligned");
    QV_REQUIRE(a->out.nb <= 1 && a->out.col0 == 0, QV_ERR_UNSUPPORTED, "gradient-planes epilogue writes a plain [2][M][N] plane stack");
    if (a->tile_n <= 0 && BN < 128) BN = 128;
    QV_REQUIRE(BN >= 128, QV_ERR_UNSUPPORTED, "gradient-planes epilogue needs tile_n 128 / 192");
  } else if (planes_out) {
    QV_REQUIRE(splits == 1 && (a->a_planes == 2 || a->b_planes == 1) && !a->a.mn_major && !a->b.mn_major && BN >= 128, QV_ERR_UNSUPPORTED,
               "plane output is instantiated for unsplit K-major (2,1)-, (2,2)- and (1,1)-plane GEMMs with tile_n 128/192");
    QV_REQUIRE(a->N % 64 == 0, QV_ERR_UNSUPPORTED, "plane output needs N to be a multiple of 64");
    QV_REQUIRE(!a->col_scale && !a->col_rscale && !a->alpha && !a->minmax, QV_ERR_UNSUPPORTED,
               "plane output takes a bias term only (no scale / alpha / observer)");
    QV_REQUIRE(!a->bias || qv_aligned16(a->bias), QV_ERR_INVALID, "plane output needs a 16-byte aligned bias");
  }
  // CTA pairs: unsplit, unbatched K-major GEMMs on 192-wide tiles with at least one 256-row tile per SM pair
  // (teacher Linears, student forward / dgrad, dgrad + gradient planes); everything else stays one CTA per tile.
  const int sms_all = qv_num_sms();
  const bool pair = pair_mode() != 0 && BN == 192 && splits == 1 && nbatch == 1 && !a->a.mn_major && !a->b.mn_major &&
                    a->a_planes == 2 && (mix || planes_out || grad_epi || a->b_planes == 1 || a->b_planes == 2) &&
                    a->M >= 256LL * (sms_all / 2) &&
                    ((pair_mode() & (mix ? (planes_out ? 16 : 1) : grad_epi ? 4 : a->b_planes == 2 ? 2 : 8)) != 0 ||
                     // bit 6: the (2,1) student GEMMs that gain from pairing -- plane output (proj dgrad: 56 -> 46 us) and long K
                     // (>= 16 k-blocks: 1-2 %); the K = 384 forward shapes lose 5-10 % as pairs (6 k-blocks per tile)
                     ((pair_mode() & 64) != 0 && !mix && !grad_epi && a->b_planes == 1 && (planes_out || a->K >= 1024)));
  const int CGh = pair ? 2 : 1;
  // mixed-format pairs on 256-wide tiles (bit 5): 64 KB into each SM per k-block for 128 x 256 outputs (1 000 clk at 64 B/clk
  // against 1 024 clk of MMA), 4 / 8 epilogue chunks split evenly over the two warps of a lane quarter (192: 3 chunks, 2 + 1),
  // and 197 x N / 256 pair tiles fill 74 pairs in whole waves at the bench shapes
  if (pair && mix && a->tile_n <= 0 && a->N % 256 == 0 && (pair_mode() & 32) != 0) BN = 256;
  CUtensorMap ma, mb, mo, my;
  int rc = make_map(&ma, a->a, a->a_planes, a->a.mn_major ? 64 : BM);
  if (rc) return rc;
  rc = make_map(&mb, a->b, a->b_planes, a->b.mn_major ? 64 : BN / CGh);
  if (rc) return rc;
  if (splits > 1) {
    QV_REQUIRE(a->workspace != nullptr, QV_ERR_INVALID, "split-K needs a workspace");
    QV_REQUIRE(a->N % 4 == 0, QV_ERR_UNSUPPORTED, "split-K needs N to be a multiple of 4");
    rc = make_out_map(&mo, a->workspace, a->N, a->M, a->N, splits, a->M * a->N);
  } else {
    const qv_out& o = a->out;
    QV_REQUIRE(o.ptr != nullptr, QV_ERR_INVALID, "null output");
    // a tile edge that is not the tensor edge must fall on a 32-column store box
    QV_REQUIRE(a->N % 32 == 0 || (o.col_inner == 0 && o.col0 + a->N == o.cols), QV_ERR_UNSUPPORTED,
               "N must be a multiple of 32 unless the output tile ends at the tensor edge");
    if (grad_epi) {
      rc = make_out_planes_map(&mo, o.ptr, o.cols, o.rows, o.ld, o.nb, o.batch_stride, a->out_plane_stride, 32);
      if (rc) return rc;
      // the raw y tile comes in through the fp32 store-map geometry used as a load map: box = 32 cols x 32 rows, 128B swizzle
      rc = make_out_map(&my, const_cast<float*>(a->ep_raw), a->N, a->M, a->ep_raw_ld, 1, 0);
    } else if (planes_out) {
      rc = make_out_planes_map(&mo, o.ptr, o.cols, o.rows, o.ld, o.nb, o.batch_stride, a->out_plane_stride, 32);
      if (rc) return rc;
      // mixed planes: region 1 leaves as 32-byte runs (a 32-column chunk's hi8 / lo8 bytes) = 16-element boxes of the bf16-typed tensor
      rc = make_out_planes_map(&my, o.ptr, o.cols, o.rows, o.ld, o.nb, o.batch_stride, a->out_plane_stride, 16);
    } else
      rc = make_out_map(&mo, o.ptr, o.cols, o.rows, o.ld, o.nb, o.batch_stride);
  }
  if (rc) return rc;

  GemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.M = a->M;
  kp.N = a->N;
  kp.kblocks = static_cast<int32_t>((a->K + BK - 1) / BK);
  kp.tiles_m = static_cast<int32_t>((a->M + BM * CGh - 1) / (BM * CGh));
  kp.tiles_n = static_cast<int32_t>((a->N + BN - 1) / BN);
  int sp = splits > kp.kblocks ? kp.kblocks : splits;
  kp.kb_per_split = (kp.kblocks + sp - 1) / sp;
  sp = (kp.kblocks + kp.kb_per_split - 1) / kp.kb_per_split;     // no empty split
  QV_REQUIRE(splits == 1 || sp == splits, QV_ERR_INVALID,
             "split-K: %d splits over %d k-blocks leaves empty splits (use %d)", splits, kp.kblocks, sp);
  kp.splits = sp;
  kp.col_scale = a->col_scale;
  kp.col_rscale = a->col_rscale;
  kp.alpha = a->alpha;
  kp.bias = a->bias;
  kp.minmax = a->minmax;
  kp.nbatch = nbatch;
  kp.batch_inner = a->batch_inner > 0 ? a->batch_inner : 1;
  kp.a_c2_outer = a->a.c2_outer; kp.a_c2_inner = a->a.c2_inner; kp.a_col0 = a->a.col0; kp.a_col_inner = a->a.col_inner;
  kp.b_c2_outer = a->b.c2_outer; kp.b_c2_inner = a->b.c2_inner; kp.b_col0 = a->b.col0; kp.b_col_inner = a->b.col_inner;
  kp.o_c2_outer = a->out.c2_outer; kp.o_c2_inner = a->out.c2_inner; kp.o_col0 = a->out.col0; kp.o_col_inner = a->out.col_inner;
  kp.act = a->act;
  kp.acc_scale = mix ? QV_MIX_ACC_SCALE : 1.0f;
  kp.out_fmt = a->out_kind == 2 ? 1 : 0;
  kp.ep_raw = a->ep_raw; kp.ep_raw_ld = a->ep_raw_ld; kp.ep_scale = a->ep_scale; kp.ep_zp = a->ep_zp;
  kp.ep_qmin = a->ep_qmin; kp.ep_qmax = a->ep_qmax; kp.ep_gelu = a->ep_gelu; kp.ep_colsum = a->ep_colsum;
  if (a->obs_ticket) {
    QV_REQUIRE(a->minmax && splits == 1 && !planes_out, QV_ERR_INVALID, "a fused observer update needs minmax on an unsplit fp32-output GEMM");
    QV_REQUIRE(a->obs_min_val && a->obs_max_val && a->obs_scale && a->obs_zero_point && a->obs_enabled && a->obs_fq_enabled,
               QV_ERR_INVALID, "null observer state pointer");
    kp.obs_min_val = a->obs_min_val; kp.obs_max_val = a->obs_max_val; kp.obs_scale = a->obs_scale;
    kp.obs_zero_point = a->obs_zero_point; kp.obs_enabled = a->obs_enabled; kp.obs_fq_enabled = a->obs_fq_enabled;
    kp.obs_c = a->obs_c; kp.obs_qmin = a->obs_qmin; kp.obs_qmax = a->obs_qmax; kp.obs_symmetric = a->obs_symmetric;
    kp.obs_ticket = a->obs_ticket;
  }
  kp.sat_flag = (a->out_kind == 2) ? a->sat_flag : nullptr;
  kp.sat_bit = a->sat_bit;
  const int64_t items = static_cast<int64_t>(kp.tiles_m) * kp.tiles_n * (nbatch > 1 ? nbatch : kp.splits);
  QV_REQUIRE(items < (1LL << 31), QV_ERR_UNSUPPORTED, "too many tiles");
  const int sms = qv_num_sms();
  const int units = sms / CGh;                                   // CTAs, or CTA pairs
  const int grid = static_cast<int>(items < units ? items : units) * CGh;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool amn = a->a.mn_major != 0, bmn = a->b.mn_major != 0;
  if (grad_epi) {
    if (pair) return launch<192, 2, 1, false, false, 2, 0, 2>(ma, mb, mo, kp, grid, st, &my);
    if (BN == 128) return launch<128, 2, 1, false, false, 2>(ma, mb, mo, kp, grid, st, &my);
    return launch<192, 2, 1, false, false, 2>(ma, mb, mo, kp, grid, st, &my);
  }
  if (mix) {
    if (planes_out) {
      if (pair && BN == 256) return launch<256, 2, 2, false, false, 1, 1, 2>(ma, mb, mo, kp, grid, st, &my);
      if (pair) return launch<192, 2, 2, false, false, 1, 1, 2>(ma, mb, mo, kp, grid, st, &my);
      if (BN == 128) return launch<128, 2, 2, false, false, 1, 1>(ma, mb, mo, kp, grid, st, &my);
      return launch<192, 2, 2, false, false, 1, 1>(ma, mb, mo, kp, grid, st, &my);
    }
    if (pair && BN == 256) return launch<256, 2, 2, false, false, 0, 1, 2>(ma, mb, mo, kp, grid, st);
    if (pair) return launch<192, 2, 2, false, false, 0, 1, 2>(ma, mb, mo, kp, grid, st);
    if (BN == 64) return launch<64, 2, 2, false, false, 0, 1>(ma, mb, mo, kp, grid, st);
    if (BN == 128) return launch<128, 2, 2, false, false, 0, 1>(ma, mb, mo, kp, grid, st);
    return launch<192, 2, 2, false, false, 0, 1>(ma, mb, mo, kp, grid, st);
  }
  if (planes_out) {
    if (a->a_planes == 1) {      // single-pass (half-precision, pre-QAT --amp variant) GEMM whose output is the next operand
      if (BN == 128) return launch<128, 1, 1, false, false, 1>(ma, mb, mo, kp, grid, st, &my);
      return launch<192, 1, 1, false, false, 1>(ma, mb, mo, kp, grid, st, &my);
    }
    if (a->b_planes == 1) {
      if (pair) return launch<192, 2, 1, false, false, 1, 0, 2>(ma, mb, mo, kp, grid, st, &my);
      if (BN == 128) return launch<128, 2, 1, false, false, 1>(ma, mb, mo, kp, grid, st, &my);
      return launch<192, 2, 1, false, false, 1>(ma, mb, mo, kp, grid, st, &my);
    }
    if (pair) return launch<192, 2, 2, false, false, 1, 0, 2>(ma, mb, mo, kp, grid, st, &my);
    if (BN == 128) return launch<128, 2, 2, false, false, 1>(ma, mb, mo, kp, grid, st, &my);
    return launch<192, 2, 2, false, false, 1>(ma, mb, mo, kp, grid, st, &my);
  }
  if (pair) {
    if (a->b_planes == 1) return launch<192, 2, 1, false, false, 0, 0, 2>(ma, mb, mo, kp, grid, st);
    return launch<192, 2, 2, false, false, 0, 0, 2>(ma, mb, mo, kp, grid, st);
  }
  switch (BN) {
    case 64: return launch_planes<64>(a->a_planes, a->b_planes, amn, bmn, ma, mb, mo, kp, grid, st);
    case 128: return launch_planes<128>(a->a_planes, a->b_planes, amn, bmn, ma, mb, mo, kp, grid, st);
    default: return launch_planes<192>(a->a_planes, a->b_planes, amn, bmn, ma, mb, mo, kp, grid, st);
  }
}

extern "C" int64_t qv_gemm_pair_launches(void) { return g_pair_launches.load(std::memory_order_relaxed); }

extern "C" int qv_splitk_reduce(const float* workspace, int32_t splits, int64_t M, int64_t N, const float* row_rscale,
                                const float* alpha, const uint8_t* mask, float* out, int32_t accumulate, void* stream) {
  QV_REQUIRE(workspace && out && splits >= 1 && M > 0 && N > 0, QV_ERR_INVALID, "bad splitk_reduce arguments");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  const int64_t total = M * N;
  int blocks = static_cast<int>((total + 255) / 256);
  const int cap = qv_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  qv_splitk_reduce_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(workspace, splits, M, N, row_rscale,
                                                                                 alpha, mask, out, accumulate);
  return qv_check_launch("qv_splitk_reduce");
}

#ifdef QV_ATTN_DEBUG
extern "C" int qv_gemm_debug_read(unsigned long long* host_out) {   // host_out: [3][4096]
  return cudaMemcpyFromSymbol(host_out, qv_gemm_dbg, sizeof(unsigned long long) * 3 * 4096) == cudaSuccess ? 0 : -1;
}
extern "C" int qv_gemm_debug_clear() {
  static unsigned long long z[3 * 4096];
  return cudaMemcpyToSymbol(qv_gemm_dbg, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif
