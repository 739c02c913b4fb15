// int8_sm100.cu -- converted-student inference: quantized Linear on the tcgen05 integer tensor cores.
//
// Replaces torch.ops.quantized.linear, the forward of torch.ao.nn.quantized.Linear that stock convert() produces from
// every nnqat.Linear (ref/src/training/qat_trainer.py:377-388; torch/ao/nn/quantized/modules/linear.py:187-190):
//
//   acc[m,n] = sum_k (q_x[m,k] - z_x) * q_w[n,k]                                                (int32, exact)
//   q_y[m,n] = clamp(rint((float(acc) + b[n] / (s_x s_w[n])) * (s_x s_w[n] / s_y)) + z_y, 0, 255)   (x86 / fbgemm engine)
//   q_y[m,n] = clamp(rint(float(acc + rint(b[n] / (s_x s_w[n]))) * (s_x s_w[n] / s_y)) + z_y, 0, 255) (qnnpack engine)
//
// q_x is uint8, q_w int8: `tcgen05.mma.kind::i8` (u8 x s8 -> s32 accumulators in TMEM) computes sum_k q_x q_w; the
// zero-point term z_x * sum_k q_w[n,k] is removed in the epilogue from a per-channel weight row sum.  Results are
// bit-identical to the fbgemm / x86 engines of the reference (tests/test_int8_gpu.py; oracle: oracle/fq_oracle.c qo_int8_linear).
//
// Structure: persistent 128 x BN tiles; warp 0 TMA producer (128-byte rows of 128 int8, 128B swizzle, OOB zero fill),
// warp 1 MMA issuer (4 x K=32 steps per stage, double-buffered TMEM accumulators), warps 2-17 requantising epilogue (four per
// TMEM lane quarter, one 32-column chunk of a 128-wide tile each; the per-column terms of all N columns are computed once per CTA
// into a shared-memory table).  Outputs: dequantised fp32, quint8 codes, or the centred codes as one bf16 plane.
#include <cuda.h>
#include <stdio.h>
#include <string.h>

#include "qv_common.cuh"
#include "qv_ptx.cuh"
#include "qv_tma.cuh"

using namespace qvptx;

namespace {

constexpr int I8_BM = 128;
constexpr int I8_BK = 128;          // 128 int8 = one 128-byte swizzle row
constexpr int I8_UMMA_K = 32;
// TMA warp, MMA warp, 16 requantising warps: four per TMEM lane quarter (= per scheduler partition), one 32-column chunk of a
// 128-wide tile each.  ncu on the 8-warp version: issue slots 44 % busy with stall reasons `wait` 1.5 and `long scoreboard` 1.1 per
// issued instruction -- two warps per scheduler cannot hide the dependent-issue and TMEM / shared-memory latencies of a chunk.
constexpr int I8_EPI_WARPS = 16;
constexpr int I8_THREADS = 64 + 32 * I8_EPI_WARPS;
// Output staging per epilogue warp: 4 KB = one 32 x 32 fp32 box (128B swizzle), or two alternating 32 x 32 uint8 boxes.  A warp
// stages one chunk per tile, so the previous box has a whole tile time to leave before its buffer is written again.  When BOTH
// output kinds are requested (tests only) the codes leave by plain stores.
constexpr int I8_STAGE_WARP_BYTES = 4096;
constexpr int I8_STAGE_OUT_BYTES = I8_EPI_WARPS * I8_STAGE_WARP_BYTES;
constexpr int I8_TERM_BYTES = I8_EPI_WARPS * 3 * 32 * 4;   // per epilogue warp: [add int32 x 32][bias term x 32][multiplier x 32] of the chunk's columns
// The per-column requantisation terms depend on the column alone: for N <= I8_TABLE_N they are computed ONCE per CTA into a
// shared-memory table when the kernel starts ([add][bias term][multiplier], each padded to a whole last tile), instead of once per
// 32-column chunk of every tile (three dependent global loads and two IEEE divides on the critical path of a chunk whose tile has
// only 12 MMAs in front of it).
constexpr int I8_TABLE_N = 2048;
constexpr int I8_TABLE_LEN = I8_TABLE_N + 128;
constexpr int I8_TABLE_BYTES = 3 * I8_TABLE_LEN * 4;
constexpr int I8_A_BYTES = I8_BM * I8_BK;

template <int BN>
struct I8Cfg {
  static constexpr int B_BYTES = BN * I8_BK;
  static constexpr int STAGE_BYTES = I8_A_BYTES + B_BYTES;
  static constexpr int STAGES = 4;        // a K = 384 tile is 3 k-blocks, K = 1536 twelve: four 32 KB stages keep the producer ahead
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + I8_STAGE_OUT_BYTES + I8_TERM_BYTES + I8_TABLE_BYTES + 1024 + 256;
  static_assert(SMEM_BYTES <= 232448, "int8 linear: shared memory over the 227 KB limit");
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : 256;
};

struct I8Params {
  int64_t M, N, K;
  int32_t kblocks, tiles_m, tiles_n;
  const float* sx;        // device scalar: input scale
  const int32_t* zx;      // device scalar: input zero point
  const float* sw;        // [N] (per channel) or [1]
  int32_t per_channel;
  const int32_t* wsum;    // [N] sum_k q_w[n,k]
  const float* bias;      // [N] fp32 or null
  float sy;
  int32_t zy;
  int32_t bias_int;       // 0: float bias added to float(acc) (x86 / fbgemm); 1: bias pre-quantised to int32 (qnnpack)
  uint8_t* qy;            // [M,N] quint8 codes (may be null)
  float* y;               // [M,N] dequantised (q_y - z_y) * s_y (may be null)
  uint16_t* codes16;      // [M,N] bf16(q_y - z_y): the one-plane integer operand of the fused attention (codes-only kernel, N % 8 == 0)
  int32_t tma_y, tma_q;   // outputs leave through shared-memory staging + TMA stores (full 128-byte lines) when their pitch allows
};

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// kind::i8 instruction descriptor: A unsigned 8-bit, B signed 8-bit, D int32, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_u8s8(int M, int N) {
  return (2u << 4) | (0u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

__device__ __forceinline__ void tma_store_2d(const void* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}

// WANT_Y: the dequantised fp32 output is requested (codes optional); otherwise codes only -- the shape the compact executor uses
// for qkv and fc1, with the rounding / clamp / pack done in the integer domain (no float -> int conversions on the XU pipe).
template <int BN, bool WANT_Y>
__global__ void __launch_bounds__(I8_THREADS, 1)
qv_int8_linear_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_q, const I8Params p) {
  using C = I8Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_out = smem + C::STAGES * C::STAGE_BYTES;          // 1 KB aligned: STAGE_BYTES is a multiple of 1024
  uint8_t* smem_terms = smem_out + I8_STAGE_OUT_BYTES;
  int32_t* tab_add = reinterpret_cast<int32_t*>(smem_terms + I8_TERM_BYTES);
  float* tab_badd = reinterpret_cast<float*>(tab_add + I8_TABLE_LEN);
  float* tab_mult = tab_badd + I8_TABLE_LEN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_terms + I8_TERM_BYTES + I8_TABLE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tmem_full = bars + 2 * C::STAGES;
  uint64_t* tmem_empty = bars + 2 * C::STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    if (p.tma_y) prefetch_tensormap(&map_y);
    if (p.tma_q || p.codes16) prefetch_tensormap(&map_q);
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], I8_EPI_WARPS * 32); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  const bool use_table = p.N <= I8_TABLE_N;
  if (use_table) {
    const float sx = __ldg(p.sx);
    const int32_t zx = __ldg(p.zx);
    const int n_pad = p.tiles_n * BN;                 // <= N + BN - 1 <= I8_TABLE_LEN
    for (int n = threadIdx.x; n < n_pad; n += I8_THREADS) {
      int32_t my_add = 0;          // - z_x * sum_k q_w  (+ rint(b / (s_x s_w)) in qnnpack mode)
      float my_badd = 0.f;         // b / (s_x s_w)  (x86 / fbgemm mode)
      float my_mult = 0.f;         // s_x s_w / s_y
      if (n < p.N) {
        const float bs = __fmul_rn(sx, __ldg(p.sw + (p.per_channel ? n : 0)));
        my_mult = __fdiv_rn(bs, p.sy);
        my_add = -zx * __ldg(p.wsum + n);
        if (p.bias) {
          const float bq = __fdiv_rn(__ldg(p.bias + n), bs);
          if (p.bias_int) my_add += static_cast<int32_t>(nearbyintf(bq));
          else my_badd = bq;
        }
      }
      tab_add[n] = my_add;
      tab_badd[n] = my_badd;
      tab_mult[n] = my_mult;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_items = p.tiles_m * p.tiles_n;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int n_blk = item % p.tiles_n, m_blk = item / p.tiles_n;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          tma_load_2d(sa, &map_a, &full_bar[stage], kb * I8_BK, m_blk * I8_BM);
          tma_load_2d(sa + I8_A_BYTES, &map_b, &full_bar[stage], kb * I8_BK, n_blk * BN);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // whole warp walks the schedule, one elected lane issues: descriptors stay in uniform registers (a 128 x BN x 32 int8 MMA
    // executes in ~64 clk, far less than the ~21-instruction ELECT / R2UR sequence a single issuing thread pays per MMA)
    {
      constexpr uint32_t idesc = umma_idesc_u8s8(I8_BM, BN);
      const uint64_t d0 = umma_smem_desc(smem_u32(smem), 16u, 1024u);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local) {
        const int buf = local & 1;
        mbar_wait(&tmem_empty[buf], ((local >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * BN);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = d0 + static_cast<uint64_t>(static_cast<uint32_t>(stage) * static_cast<uint32_t>(C::STAGE_BYTES >> 4));
            const uint64_t db = da + (I8_A_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < I8_BK / I8_UMMA_K; ++k) umma_i8(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(&tmem_full[buf]);
        __syncwarp();
      }
    }
  } else {
    // =============================== requantising epilogue ===============================
    // A K = 384 tile is 12 MMAs (~0.8 k clk at the int8 rate): the epilogue, not the tensor pipe, bounds this kernel, so it runs
    // on 16 warps, fetches the per-column terms as broadcast 16-byte shared-memory loads (they were three shuffles per ELEMENT) and
    // rounds with the magic-number add (round-to-nearest-even like nearbyintf, exact for |v| < 2^22 and clamped the same beyond).
    const int ew = warp - 2;
    const int q = warp & 3;
    const int par = ew >> 2;
    const float sx = __ldg(p.sx);
    const int32_t zx = __ldg(p.zx);
    const float zyf = static_cast<float>(p.zy);
    const float rint_c = 12582912.0f - zyf;                  // (v + 1.5 * 2^23) - rint_c = rint(v) + z_y
    const bool vec_ok = (p.N % 16 == 0);
    int32_t* t_add = reinterpret_cast<int32_t*>(smem_terms) + ew * 96;
    float* t_badd = reinterpret_cast<float*>(t_add + 32);
    float* t_mult = t_badd + 32;
    // Row-per-lane registers written straight to global memory touch 32 different lines per store instruction (one 16-byte
    // piece of a sector each: the LSU, not HBM, bounded the kernel): outputs go through swizzled staging and leave as TMA boxes.
    uint8_t* st_base = smem_out + ew * I8_STAGE_WARP_BYTES;
    const bool any_tma = p.tma_y || p.tma_q || (!WANT_Y && p.codes16 != nullptr);
    const int32_t code_bias = p.zy - 0x4B400000;                      // (v + 1.5 * 2^23) as an integer -> rint(v) + z_y
    const bool codes16 = !WANT_Y && p.codes16 != nullptr;
    const int32_t centre_bias = 0x4B400000 - p.zy;                    // bits(1.5 * 2^23 + (q - z_y)) = q + centre_bias for |q - z_y| <= 255
    uint32_t st_ctr = 0;                                              // chunks this warp has staged so far
    int local = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local) {
      const int n_blk = item % p.tiles_n, m_blk = item / p.tiles_n;
      const int buf = local & 1;
      mbar_wait(&tmem_full[buf], (local >> 1) & 1);
      tc_fence_after();
      const int64_t row = static_cast<int64_t>(m_blk) * I8_BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      if (par * 32 >= BN) {            // BN = 64: the third and fourth warp of a lane quarter have no chunk, only the hand-back
        tc_fence_before();
        mbar_arrive(&tmem_empty[buf]);
      }
#pragma unroll 1
      for (int c0 = par * 32; c0 < BN; c0 += 8 * I8_EPI_WARPS) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN + c0), r);
        const int64_t n0 = static_cast<int64_t>(n_blk) * BN + c0;
        // per-column terms: from the CTA's table, or (N > I8_TABLE_N) computed while the TMEM load is in flight: lane j owns column n0 + j
        const int32_t* c_add = tab_add + n0;
        const float* c_badd = tab_badd + n0;
        const float* c_mult = tab_mult + n0;
        if (!use_table) {
          c_add = t_add; c_badd = t_badd; c_mult = t_mult;
          int32_t my_add = 0;          // - z_x * sum_k q_w  (+ rint(b / (s_x s_w)) in qnnpack mode)
          float my_badd = 0.f;         // b / (s_x s_w)  (x86 / fbgemm mode)
          float my_mult = 0.f;         // s_x s_w / s_y
          const int64_t n = n0 + lane;
          if (n < p.N) {
            const float bs = __fmul_rn(sx, __ldg(p.sw + (p.per_channel ? n : 0)));
            my_mult = __fdiv_rn(bs, p.sy);
            my_add = -zx * __ldg(p.wsum + n);
            if (p.bias) {
              const float bq = __fdiv_rn(__ldg(p.bias + n), bs);
              if (p.bias_int) my_add += static_cast<int32_t>(nearbyintf(bq));
              else my_badd = bq;
            }
          }
          __syncwarp();                // the previous chunk's reads of this warp's term buffer are done
          t_add[lane] = my_add;
          t_badd[lane] = my_badd;
          t_mult[lane] = my_mult;
          __syncwarp();
        }
        tmem_ld_wait();
        if (c0 + 8 * I8_EPI_WARPS >= BN) {           // this warp's last chunk of the tile is in registers
          tc_fence_before();
          mbar_arrive(&tmem_empty[buf]);
        }
        if (n0 >= p.N) continue;
        const int ncols = static_cast<int>(min(static_cast<int64_t>(32), p.N - n0));
        uint32_t packed[8];
        uint32_t packed16[WANT_Y ? 1 : 16];
        float deq[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const int4 a4 = *reinterpret_cast<const int4*>(c_add + 4 * j4);
          const float4 b4 = *reinterpret_cast<const float4*>(c_badd + 4 * j4);
          const float4 m4 = *reinterpret_cast<const float4*>(c_mult + 4 * j4);
          const int32_t aj[4] = {a4.x, a4.y, a4.z, a4.w};
          const float bj[4] = {b4.x, b4.y, b4.z, b4.w};
          const float mj[4] = {m4.x, m4.y, m4.z, m4.w};
          uint32_t w = 0;
          if constexpr (WANT_Y) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int32_t acc = static_cast<int32_t>(r[4 * j4 + e]) + aj[e];
              const float v = __fmul_rn(__fadd_rn(static_cast<float>(acc), bj[e]), mj[e]);
              float f = __fsub_rn(__fadd_rn(v, 12582912.0f), rint_c);
              f = fminf(fmaxf(f, 0.0f), 255.0f);
              w |= static_cast<uint32_t>(f) << (8 * e);
              deq[4 * j4 + e] = __fmul_rn(__fsub_rn(f, zyf), p.sy);
            }
          } else {
            // codes only: v + 1.5 * 2^23 holds rint(v) in its low mantissa bits (round-to-nearest-even, |v| < 2^22); the float's bit
            // pattern is monotonic in v beyond that range too, so  bits - 0x4B400000 + z_y  clamped to [0, 255] as an INTEGER gives
            // the same code as the float route above for every finite v; four clamped codes are gathered with byte permutes
            int32_t c[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int32_t acc = static_cast<int32_t>(r[4 * j4 + e]) + aj[e];
              const float v = __fmul_rn(__fadd_rn(static_cast<float>(acc), bj[e]), mj[e]);
              const int32_t i = __float_as_int(__fadd_rn(v, 12582912.0f)) + code_bias;
              c[e] = min(max(i, 0), 255);
            }
            w = __byte_perm(__byte_perm(c[0], c[1], 0x0040), __byte_perm(c[2], c[3], 0x0040), 0x5410);
            if (codes16) {            // centred codes as bf16: |q - z_y| <= 255 is exact in 8 significant bits = the float's upper half
              // int -> float without the XU pipe (ncu: 92 % busy with one I2F per element more): 1.5 * 2^23 + n as a bit pattern, minus 1.5 * 2^23
              const uint32_t f0 = __float_as_uint(__fsub_rn(__int_as_float(c[0] + centre_bias), 12582912.0f));
              const uint32_t f1 = __float_as_uint(__fsub_rn(__int_as_float(c[1] + centre_bias), 12582912.0f));
              const uint32_t f2 = __float_as_uint(__fsub_rn(__int_as_float(c[2] + centre_bias), 12582912.0f));
              const uint32_t f3 = __float_as_uint(__fsub_rn(__int_as_float(c[3] + centre_bias), 12582912.0f));
              packed16[2 * j4] = __byte_perm(f0, f1, 0x7632);
              packed16[2 * j4 + 1] = __byte_perm(f2, f3, 0x7632);
            }
          }
          packed[j4] = w;
        }
        const int row0 = m_blk * I8_BM + q * 32;
        if (static_cast<int64_t>(row0) >= p.M) continue;          // warp-uniform: the whole slab is padding
        if (any_tma) {
          const uint32_t sb = st_ctr & 1u;
          uint8_t* st_y = st_base;
          uint8_t* st_q = st_base + sb * 1024;
          ++st_ctr;
          if (lane == 0) {                                        // bulk groups retire in order: with two code buffers the store of
            if (WANT_Y) tma_store_wait_read<0>();                 // the chunk BEFORE the previous one has left the buffer reused now
            else tma_store_wait_read<1>();
          }
          __syncwarp();
          if (!WANT_Y && codes16) {                               // 32 rows x 64 bytes (32 bf16), linear; two 2 KB buffers
            st_q = st_base + sb * 2048;
            const uint32_t a = smem_u32(st_q) + lane * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + 16 * j), "r"(packed16[WANT_Y ? 0 : 4 * j]),
                           "r"(packed16[WANT_Y ? 0 : 4 * j + 1]), "r"(packed16[WANT_Y ? 0 : 4 * j + 2]), "r"(packed16[WANT_Y ? 0 : 4 * j + 3]) : "memory");
          } else if (p.tma_q && p.qy) {                           // 32 rows x 32 bytes, linear
            const uint32_t a = smem_u32(st_q) + lane * 32;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(packed[0]), "r"(packed[1]), "r"(packed[2]), "r"(packed[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + 16), "r"(packed[4]), "r"(packed[5]), "r"(packed[6]), "r"(packed[7]) : "memory");
          }
          if (WANT_Y && p.tma_y && p.y) {                         // 32 rows x 128 bytes, 16-byte chunk j of row r at j ^ (r & 7)
            const uint32_t a = smem_u32(st_y) + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a + (static_cast<uint32_t>(j ^ (lane & 7)) << 4)), "f"(deq[4 * j]),
                           "f"(deq[4 * j + 1]), "f"(deq[4 * j + 2]), "f"(deq[4 * j + 3]) : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if ((p.tma_q && p.qy) || codes16) tma_store_2d(&map_q, st_q, static_cast<int>(n0), row0);
            if (WANT_Y && p.tma_y && p.y) tma_store_3d(&map_y, st_y, static_cast<int>(n0), row0, 0);
            tma_store_commit();
          }
        }
        if (!row_ok) continue;
        if (p.qy && !p.tma_q) {
          uint8_t* dst = p.qy + row * p.N + n0;
          if (vec_ok && ncols == 32) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            *reinterpret_cast<uint4*>(dst + 16) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) dst[j] = static_cast<uint8_t>((packed[j >> 2] >> (8 * (j & 3))) & 0xffu);
          }
        }
        if (WANT_Y && p.y && !p.tma_y) {
          float* dst = p.y + row * p.N + n0;
          if (p.N % 4 == 0 && ncols == 32) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(deq[4 * j], deq[4 * j + 1], deq[4 * j + 2], deq[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) dst[j] = deq[j];
          }
        }
      }
    }
    if (any_tma && lane == 0) tma_store_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// uint8 / int8 matrix [rows][K] as a 2-D tensor (K, rows); box = (128 bytes, box_rows), 128B swizzle, OOB zero fill
int make_map_i8(CUtensorMap* m, const void* ptr, int64_t rows, int64_t K, int box_rows) {
  EncodeTiledFn enc = get_encode();
  QV_REQUIRE(enc != nullptr, QV_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  QV_REQUIRE(ptr && qv_aligned16(ptr), QV_ERR_INVALID, "int8 operand base must be a 16-byte aligned device pointer");
  QV_REQUIRE(K % 16 == 0, QV_ERR_UNSUPPORTED, "int8 linear needs K to be a multiple of 16 (got %lld)", (long long)K);
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(K)};
  cuuint32_t box[2] = {128, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  QV_REQUIRE(r == CUDA_SUCCESS, QV_ERR_CUDA, "cuTensorMapEncodeTiled(int8 operand) failed (%d)", (int)r);
  return 0;
}

// quint8 output [rows][N] as a 2-D tensor (N, rows); store box = (32 bytes, 32 rows), no swizzle
int make_map_q_out(CUtensorMap* m, uint8_t* ptr, int64_t rows, int64_t N) {
  EncodeTiledFn enc = get_encode();
  QV_REQUIRE(enc != nullptr, QV_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(N)};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  QV_REQUIRE(r == CUDA_SUCCESS, QV_ERR_CUDA, "cuTensorMapEncodeTiled(quint8 output) failed (%d)", (int)r);
  return 0;
}

// bf16 code output [rows][N] as a 2-D tensor (N, rows); store box = (32 codes = 64 bytes, 32 rows), no swizzle
int make_map_codes_out(CUtensorMap* m, uint16_t* ptr, int64_t rows, int64_t N) {
  EncodeTiledFn enc = get_encode();
  QV_REQUIRE(enc != nullptr, QV_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(N) * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  QV_REQUIRE(r == CUDA_SUCCESS, QV_ERR_CUDA, "cuTensorMapEncodeTiled(bf16 code output) failed (%d)", (int)r);
  return 0;
}

template <int BN, bool WANT_Y>
int launch_i8(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& my, const CUtensorMap& mq, const I8Params& kp, int grid,
              cudaStream_t st) {
  using C = I8Cfg<BN>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(qv_int8_linear_kernel<BN, WANT_Y>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  });
  QV_REQUIRE(attr_err == cudaSuccess, QV_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  qv_int8_linear_kernel<BN, WANT_Y><<<grid, I8_THREADS, C::SMEM_BYTES, st>>>(ma, mb, my, mq, kp);
  return qv_check_launch("qv_int8_linear");
}

// x -> quint8 codes with the scale / zero point in device scalars: clamp(rint(x * (1/s)) + z, 0, 255)
// (torch.quantize_per_tensor; the nnq.Quantize module convert() puts in place of QuantStub)
__global__ void __launch_bounds__(256) quantize_u8_kernel(const float* __restrict__ x, int64_t n, const float* scale,
                                                          const int32_t* zp, uint8_t* __restrict__ q) {
  const float inv = __fdiv_rn(1.0f, __ldg(scale));
  const float z = static_cast<float>(__ldg(zp));
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float a[4] = {v.x, v.y, v.z, v.w};
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f = __fadd_rn(nearbyintf(__fmul_rn(a[j], inv)), z);
      f = fminf(fmaxf(f, 0.0f), 255.0f);
      w |= static_cast<uint32_t>(f) << (8 * j);
    }
    reinterpret_cast<uint32_t*>(q)[i] = w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = n4 * 4; i < n; ++i) {
      float f = __fadd_rn(nearbyintf(__fmul_rn(x[i], inv)), z);
      q[i] = static_cast<uint8_t>(fminf(fmaxf(f, 0.0f), 255.0f));
    }
}

// Affine quint8 qparams from an observed min / max (torch/ao/quantization/observer.py:349-427, the convert-time Python
// formula; also what the float-glue executor uses for dynamically quantised inputs): fp32 arithmetic.
__device__ __forceinline__ void qparams_affine(const uint32_t* acc, int32_t qmin, int32_t qmax, float& s, float& z) {
  const float mn = fminf(qv_ord2f(acc[0]), 0.0f), mx = fmaxf(qv_ord2f(acc[1]), 0.0f);
  s = __fdiv_rn(__fsub_rn(mx, mn), static_cast<float>(qmax - qmin));
  s = fmaxf(s, 1.1920928955078125e-07f);
  z = __fsub_rn(static_cast<float>(qmin), nearbyintf(__fdiv_rn(mn, s)));
  z = fminf(fmaxf(z, static_cast<float>(qmin)), static_cast<float>(qmax));
}

__global__ void qparams_from_minmax_kernel(const uint32_t* acc, int32_t qmin, int32_t qmax, float* scale, int32_t* zp) {
  float s, z;
  qparams_affine(acc, qmin, qmax, s, z);
  *scale = s;
  *zp = static_cast<int32_t>(z);
}

__device__ __forceinline__ uint32_t quant4_u8(float a0, float a1, float a2, float a3, float inv, float z) {
  const float a[4] = {a0, a1, a2, a3};
  uint32_t w = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float f = __fadd_rn(nearbyintf(__fmul_rn(a[j], inv)), z);
    f = fminf(fmaxf(f, 0.0f), 255.0f);
    w |= static_cast<uint32_t>(f) << (8 * j);
  }
  return w;
}

// qv_qparams_from_minmax + qv_quantize_u8 in one launch: every block derives the dynamic (scale, zero point) from the finished
// min / max accumulator itself (a dozen flops), block 0 publishes them for the int8 Linear that consumes the codes.
__global__ void __launch_bounds__(256) quantize_u8_dyn_kernel(const float* __restrict__ x, int64_t n, const uint32_t* acc, float* scale_out,
                                                              int32_t* zp_out, uint8_t* __restrict__ q) {
  float s, z;
  qparams_affine(acc, 0, 255, s, z);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *scale_out = s;
    *zp_out = static_cast<int32_t>(z);
  }
  const float inv = __fdiv_rn(1.0f, s);
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    reinterpret_cast<uint32_t*>(q)[i] = quant4_u8(v.x, v.y, v.z, v.w, inv, z);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = n4 * 4; i < n; ++i) {
      float f = __fadd_rn(nearbyintf(__fmul_rn(x[i], inv)), z);
      q[i] = static_cast<uint8_t>(fminf(fmaxf(f, 0.0f), 255.0f));
    }
}

// The output of a converted Linear is (q - z_y) * s_y with q in 0..255: an integer code times ONE scale.  Consumers that can work on
// codes take them instead of the dequantised fp32 tensor (1 byte instead of 4 out of the GEMM, no fp32 round trip):
//   * attention: the centred codes q - z_y as ONE bf16 plane (|q - z_y| <= 255 is exact in bf16), s_y applied by the kernel --
//     the operand format of the QAT student's fused attention (qv_attn_fwd with one plane);
//   * GELU + dynamic re-quantisation (fc1 -> fc2): GELU((q - z_y) s_y) takes at most 256 values, so its min / max and the
//     re-quantised codes are table lookups on q.  Every table entry is computed with the expressions of the elementwise path
//     (qv_int8_linear's dequantisation, qv_gelu_minmax, qv_quantize_u8), so the codes are bit-identical to it.
__global__ void __launch_bounds__(256) codes_from_u8_kernel(const uint8_t* __restrict__ q, int64_t n16, int32_t zp,
                                                            __nv_bfloat16* __restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n16; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(q) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c0 = static_cast<int>((w[j >> 1] >> (16 * (j & 1))) & 0xffu) - zp;
      const int c1 = static_cast<int>((w[j >> 1] >> (16 * (j & 1) + 8)) & 0xffu) - zp;
      const __nv_bfloat162 b = __floats2bfloat162_rn(static_cast<float>(c0), static_cast<float>(c1));
      o[j] = *reinterpret_cast<const uint32_t*>(&b);
    }
    uint4* dst = reinterpret_cast<uint4*>(out) + 2 * i;
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
}

__device__ __forceinline__ float gelu_of_code(int c, float zyf, float sy) {
  return qv_gelu_fwd(__fmul_rn(__fsub_rn(static_cast<float>(c), zyf), sy));       // dequantise as qv_int8_linear does, then exact-erf GELU
}

// min / max of GELU(dequant(q)) merged into acc: one table lookup per element (blockDim.x == 256 == table size)
__global__ void __launch_bounds__(256) gelu_u8_minmax_kernel(const uint8_t* __restrict__ q, int64_t n16, float sy, int32_t zy, uint32_t* acc) {
  __shared__ float lut[256];
  lut[threadIdx.x] = gelu_of_code(threadIdx.x, static_cast<float>(zy), sy);
  __syncthreads();
  float mn = INFINITY, mx = -INFINITY;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n16; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(q) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float g = lut[(w[j >> 2] >> (8 * (j & 3))) & 0xffu];
      mn = fminf(mn, g);
      mx = fmaxf(mx, g);
    }
  }
  mn = qv_warp_min(mn);
  mx = qv_warp_max(mx);
  if ((threadIdx.x & 31) == 0 && mn <= mx) {
    atomicMin(acc, qv_f2ord(mn));
    atomicMax(acc + 1, qv_f2ord(mx));
  }
}

// out = quantize_u8(GELU(dequant(q))) with the dynamic qparams of the finished accumulator: a 256-byte code -> code table
__global__ void __launch_bounds__(256) gelu_u8_requant_kernel(const uint8_t* __restrict__ q, int64_t n16, float sy, int32_t zy,
                                                              const uint32_t* acc, float* scale_out, int32_t* zp_out,
                                                              uint8_t* __restrict__ out) {
  __shared__ uint8_t lut[256];
  float s, z;
  qparams_affine(acc, 0, 255, s, z);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *scale_out = s;
    *zp_out = static_cast<int32_t>(z);
  }
  {
    const float inv = __fdiv_rn(1.0f, s);
    float f = __fadd_rn(nearbyintf(__fmul_rn(gelu_of_code(threadIdx.x, static_cast<float>(zy), sy), inv)), z);
    f = fminf(fmaxf(f, 0.0f), 255.0f);
    lut[threadIdx.x] = static_cast<uint8_t>(f);
  }
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n16; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(q) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t r = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) r |= static_cast<uint32_t>(lut[(w[k] >> (8 * j)) & 0xffu]) << (8 * j);
      o[k] = r;
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace

static int int8_linear_impl(const uint8_t* qx, int64_t M, int64_t K, const float* sx, const int32_t* zx, const int8_t* qw,
                            int64_t N, const float* sw, int32_t per_channel, const int32_t* wsum, const float* bias, float sy,
                            int32_t zy, int32_t bias_int, uint8_t* qy, float* y, uint16_t* codes16, void* stream) {
  QV_REQUIRE(qx && qw && sx && zx && sw && wsum && (qy || y || codes16), QV_ERR_INVALID, "null pointer in int8_linear");
  QV_REQUIRE(M > 0 && N > 0 && K > 0 && sy > 0.f, QV_ERR_INVALID, "bad int8_linear shape (M=%lld N=%lld K=%lld)", (long long)M,
             (long long)N, (long long)K);
  QV_REQUIRE(!codes16 || (N % 8 == 0 && qv_aligned16(codes16) && zy >= 0 && zy <= 255), QV_ERR_INVALID,
             "int8_linear_codes needs N %% 8 == 0, a 16-byte aligned output and a quint8 zero point");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  const int BN = N <= 64 ? 64 : 128;
  CUtensorMap ma, mb;
  int rc = make_map_i8(&ma, qx, M, K, I8_BM);
  if (rc) return rc;
  rc = make_map_i8(&mb, qw, N, K, BN);
  if (rc) return rc;
  I8Params kp;
  memset(&kp, 0, sizeof(kp));
  kp.M = M; kp.N = N; kp.K = K;
  kp.kblocks = static_cast<int32_t>((K + I8_BK - 1) / I8_BK);
  kp.tiles_m = static_cast<int32_t>((M + I8_BM - 1) / I8_BM);
  kp.tiles_n = static_cast<int32_t>((N + BN - 1) / BN);
  kp.sx = sx; kp.zx = zx; kp.sw = sw; kp.per_channel = per_channel; kp.wsum = wsum; kp.bias = bias;
  kp.sy = sy; kp.zy = zy; kp.bias_int = bias_int; kp.qy = qy; kp.y = y; kp.codes16 = codes16;
  // TMA-stored outputs need 16-byte aligned bases and row pitches (N % 4 floats / N % 16 bytes); the 10-class head keeps plain stores
  CUtensorMap my = ma, mq = ma;
  kp.tma_y = (y && N % 4 == 0 && qv_aligned16(y)) ? 1 : 0;
  kp.tma_q = (qy && !y && N % 16 == 0 && qv_aligned16(qy)) ? 1 : 0;    // one staged output kind per launch (4 KB per warp)
  if (kp.tma_y) {
    rc = make_out_map(&my, y, N, M, N, 1, 0);
    if (rc) return rc;
  }
  if (codes16) {
    rc = make_map_codes_out(&mq, codes16, M, N);
    if (rc) return rc;
  } else if (kp.tma_q) {
    rc = make_map_q_out(&mq, qy, M, N);
    if (rc) return rc;
  }
  const int64_t items = static_cast<int64_t>(kp.tiles_m) * kp.tiles_n;
  const int sms = qv_num_sms();
  const int grid = static_cast<int>(items < sms ? items : sms);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (BN == 64) return y ? launch_i8<64, true>(ma, mb, my, mq, kp, grid, st) : launch_i8<64, false>(ma, mb, my, mq, kp, grid, st);
  return y ? launch_i8<128, true>(ma, mb, my, mq, kp, grid, st) : launch_i8<128, false>(ma, mb, my, mq, kp, grid, st);
}

extern "C" int qv_int8_linear(const uint8_t* qx, int64_t M, int64_t K, const float* sx, const int32_t* zx, const int8_t* qw,
                              int64_t N, const float* sw, int32_t per_channel, const int32_t* wsum, const float* bias, float sy,
                              int32_t zy, int32_t bias_int, uint8_t* qy, float* y, void* stream) {
  QV_REQUIRE(qy || y, QV_ERR_INVALID, "null pointer in int8_linear");
  return int8_linear_impl(qx, M, K, sx, zx, qw, N, sw, per_channel, wsum, bias, sy, zy, bias_int, qy, y, nullptr, stream);
}

extern "C" int qv_int8_linear_codes(const uint8_t* qx, int64_t M, int64_t K, const float* sx, const int32_t* zx, const int8_t* qw,
                                    int64_t N, const float* sw, int32_t per_channel, const int32_t* wsum, const float* bias, float sy,
                                    int32_t zy, int32_t bias_int, uint16_t* codes, void* stream) {
  QV_REQUIRE(codes, QV_ERR_INVALID, "null pointer in int8_linear_codes");
  return int8_linear_impl(qx, M, K, sx, zx, qw, N, sw, per_channel, wsum, bias, sy, zy, bias_int, nullptr, nullptr, codes, stream);
}

extern "C" int qv_quantize_u8(const float* x, int64_t n, const float* scale, const int32_t* zero_point, uint8_t* q, void* stream) {
  QV_REQUIRE(x && scale && zero_point && q && n > 0, QV_ERR_INVALID, "bad quantize_u8 arguments");
  QV_REQUIRE(qv_aligned16(x) && (reinterpret_cast<uintptr_t>(q) & 3u) == 0, QV_ERR_INVALID, "quantize_u8 needs aligned buffers");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  int64_t blocks = ((n >> 2) + 255) / 256;
  const int64_t cap = static_cast<int64_t>(qv_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  quantize_u8_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, scale, zero_point, q);
  return qv_check_launch("qv_quantize_u8");
}

extern "C" int qv_qparams_from_minmax(const uint32_t* acc, int32_t qmin, int32_t qmax, float* scale, int32_t* zero_point,
                                      void* stream) {
  QV_REQUIRE(acc && scale && zero_point && qmax > qmin, QV_ERR_INVALID, "bad qparams_from_minmax arguments");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  qparams_from_minmax_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(acc, qmin, qmax, scale, zero_point);
  return qv_check_launch("qv_qparams_from_minmax");
}

static int i8_ew_blocks(int64_t items) {
  int64_t blocks = (items + 255) / 256;
  const int64_t cap = static_cast<int64_t>(qv_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

extern "C" int qv_quantize_u8_dyn(const float* x, int64_t n, const uint32_t* acc, float* scale_out, int32_t* zero_point_out,
                                  uint8_t* q, void* stream) {
  QV_REQUIRE(x && acc && scale_out && zero_point_out && q && n > 0, QV_ERR_INVALID, "bad quantize_u8_dyn arguments");
  QV_REQUIRE(qv_aligned16(x) && (reinterpret_cast<uintptr_t>(q) & 3u) == 0, QV_ERR_INVALID, "quantize_u8_dyn needs aligned buffers");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  quantize_u8_dyn_kernel<<<i8_ew_blocks(n >> 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, acc, scale_out, zero_point_out, q);
  return qv_check_launch("qv_quantize_u8_dyn");
}

extern "C" int qv_codes_from_u8(const uint8_t* q, int64_t n, int32_t zero_point, uint16_t* codes, void* stream) {
  QV_REQUIRE(q && codes && n > 0 && n % 16 == 0, QV_ERR_INVALID, "bad codes_from_u8 arguments (n must be a multiple of 16)");
  QV_REQUIRE(zero_point >= 0 && zero_point <= 255, QV_ERR_INVALID, "codes_from_u8: quint8 zero point out of range");
  QV_REQUIRE(qv_aligned16(q) && qv_aligned16(codes), QV_ERR_INVALID, "codes_from_u8 needs 16-byte aligned buffers");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  codes_from_u8_kernel<<<i8_ew_blocks(n >> 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      q, n >> 4, zero_point, reinterpret_cast<__nv_bfloat16*>(codes));
  return qv_check_launch("qv_codes_from_u8");
}

extern "C" int qv_gelu_u8_minmax(const uint8_t* q, int64_t n, float sy, int32_t zy, uint32_t* acc, void* stream) {
  QV_REQUIRE(q && acc && n > 0 && n % 16 == 0 && sy > 0.f, QV_ERR_INVALID, "bad gelu_u8_minmax arguments (n must be a multiple of 16)");
  QV_REQUIRE(qv_aligned16(q), QV_ERR_INVALID, "gelu_u8_minmax needs a 16-byte aligned buffer");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  gelu_u8_minmax_kernel<<<i8_ew_blocks(n >> 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(q, n >> 4, sy, zy, acc);
  return qv_check_launch("qv_gelu_u8_minmax");
}

extern "C" int qv_gelu_u8_requant(const uint8_t* q, int64_t n, float sy, int32_t zy, const uint32_t* acc, float* scale_out,
                                  int32_t* zero_point_out, uint8_t* out, void* stream) {
  QV_REQUIRE(q && acc && scale_out && zero_point_out && out && n > 0 && n % 16 == 0 && sy > 0.f, QV_ERR_INVALID,
             "bad gelu_u8_requant arguments (n must be a multiple of 16)");
  QV_REQUIRE(qv_aligned16(q) && qv_aligned16(out), QV_ERR_INVALID, "gelu_u8_requant needs 16-byte aligned buffers");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  gelu_u8_requant_kernel<<<i8_ew_blocks(n >> 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(q, n >> 4, sy, zy, acc, scale_out,
                                                                                           zero_point_out, out);
  return qv_check_launch("qv_gelu_u8_requant");
}
