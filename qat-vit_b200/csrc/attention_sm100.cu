// attention_sm100.cu -- fused softmax-attention forward for one (image, head) per work item, on tcgen05 / TMEM.
//
// Replaces F.scaled_dot_product_attention inside timm's Attention.forward (reference call path
// ref/src/training/qat_trainer.py:337-341 -> timm Block -> Attention) for both the frozen ViT-B/16 teacher and the
// ViT-S/16 student:  O = softmax(Q K^T * scale) V  with T <= 224 tokens and head_dim 64, so the whole key range fits
// one tile and no online-softmax rescaling is needed.
//
//   operands   bf16 plane stacks inside the [tokens, 3*D] qkv tensor (Q | K | V column blocks, 64 columns per head):
//              NPL = 2: fp32 values as hi/lo planes (teacher; 3 tensor-core products per GEMM);
//              NPL = 1: exact integer fake-quant codes (student: FQ(x) = code * s, so Q K^T = s^2 * codes codes^T needs ONE
//              product and P V two) -- the per-tensor scale s is read from the observer's device buffer.
//   S = Q K^T  SS-mode tcgen05.mma (Q, K tiles staged by TMA, 128B swizzle) into TMEM, 128 query rows per tile, 2 tiles.
//   softmax    one thread per query row (TMEM lane == row): tcgen05.ld 32 columns at a time, row max, exp2, row sum; the
//              un-normalised probabilities are written back IN PLACE over S as bf16 hi/lo pairs (tcgen05.st), laid out so
//              that each 16-key k-step of the next MMA finds its A operand in 8 consecutive TMEM columns.
//   O = P V    TS-mode tcgen05.mma: A = P from TMEM, B = V from shared memory (MN-major), fp32 accumulate in TMEM.
//   epilogue   O * (1 / rowsum) [* s] -> bf16 hi/lo planes (the operand format of the proj GEMM), optional logsumexp.
//
// Scores and probabilities never touch HBM (the unfused path moved ~16 B per score: 0.5 GB per layer at batch 256).
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 softmax/epilogue of query tile 0,
// warps 6-9 of query tile 1.  TMEM (512 columns): S0/P0 [0,224) | S1/P1 [224,448) | O0 [448,512) ; O1 re-uses [0,64)
// once P0 has been consumed.
#include <cuda.h>
#include <stdio.h>
#include <string.h>

#include "qv_common.cuh"
#include "qv_ptx.cuh"
#include "qv_tma.cuh"

using namespace qvptx;

namespace {

constexpr int AT_THREADS = 320;
constexpr int HD = 64;                       // head dim
constexpr int Q_TILE_BYTES = 128 * HD * 2;   // 16 KB: 128 query rows x 64 bf16
constexpr int K_PLANE_BYTES = 224 * HD * 2;  // 28 KB: up to 224 keys x 64 bf16
constexpr int V_BOX_BYTES = 64 * HD * 2;     // 8 KB: 64 keys x 64 bf16
constexpr int V_PLANE_BYTES = 4 * V_BOX_BYTES;
constexpr int S_COLS = 224;                  // TMEM columns reserved per score tile
constexpr int O0_COL = 448;

template <int NPL>
struct AttnCfg {
  static constexpr int Q_BYTES = NPL * 2 * Q_TILE_BYTES;
  static constexpr int K_BYTES = NPL * K_PLANE_BYTES;
  static constexpr int V_BYTES = NPL * V_PLANE_BYTES;
  static constexpr int SMEM_BYTES = Q_BYTES + K_BYTES + V_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int NPAIRS_S = (NPL == 2) ? 3 : 1;    // (hi,hi) (hi,lo) (lo,hi)  |  codes x codes
  static constexpr int NPAIRS_PV = (NPL == 2) ? 3 : 2;   // P is always hi/lo;  V hi/lo or exact codes
};

struct AttnParams {
  int32_t B, T, H;
  int32_t n_keys;        // T rounded up to 16 (MMA N of the score GEMM, K of the PV GEMM)
  int32_t m_tiles;       // ceil(T / 128)
  float scale;           // softmax scale (head_dim^-0.5)
  const float* qk_scale; // optional device scalar s: logits are multiplied by s*s (integer-code operands)
  const float* v_scale;  // optional device scalar s: output is multiplied by s
  __nv_bfloat16* out;    // [2][B*T][out_ld] hi/lo planes; head h writes columns h*64 .. h*64+63 (may be null)
  int64_t out_plane_stride, out_ld;
  float* out_f32;        // optional fp32 copy of the output, [B*T][H*64]
  float* lse;            // optional [B*H*T]: scale' * rowmax + ln(rowsum)   (natural-log logsumexp of the scaled logits)
};

// A operand from TMEM (TS form): D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// pack (a0, a1) as bf16 hi pair and the exact residuals as bf16 lo pair (element 0 in the low half)
__device__ __forceinline__ void split_pack2(float a0, float a1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(a0, a1);
  hi = *reinterpret_cast<const uint32_t*>(&h2);
  const float r0 = a0 - __uint_as_float(hi << 16), r1 = a1 - __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
  lo = *reinterpret_cast<const uint32_t*>(&l2);
}

template <int NPL>
__global__ void __launch_bounds__(AT_THREADS, 1)
qv_attn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const AttnParams p) {
  using C = AttnCfg<NPL>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                       // [NPL][2 tiles][128 x 64]
  uint8_t* sK = sQ + C::Q_BYTES;            // [NPL][224 x 64]
  uint8_t* sV = sK + C::K_BYTES;            // [NPL][4 boxes][64 keys x 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + C::V_BYTES);
  uint64_t* qk_full = bars + 0;
  uint64_t* qk_empty = bars + 1;
  uint64_t* v_full = bars + 2;
  uint64_t* v_empty = bars + 3;
  uint64_t* s_full = bars + 4;              // [2]
  uint64_t* p_ready = bars + 6;             // [2]
  uint64_t* o_full = bars + 8;              // [2]
  uint64_t* tmem_free = bars + 10;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int D = p.H * HD;
  const int num_items = p.B * p.H;
  const int n_keys = p.n_keys;
  const int nkb = (n_keys + 63) >> 6;       // 64-key V boxes
  const int mt = p.m_tiles;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_q);
    prefetch_tensormap(&map_k);
    prefetch_tensormap(&map_v);
    mbar_init(qk_full, 1);
    mbar_init(qk_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_ready[g], 128);
      mbar_init(&o_full[g], 1);
      mbar_init(&tmem_free[g], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      const uint32_t qk_bytes = static_cast<uint32_t>(NPL * (mt * Q_TILE_BYTES + n_keys * HD * 2));
      const uint32_t v_bytes = static_cast<uint32_t>(NPL * nkb * V_BOX_BYTES);
      int local = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local) {
        const int b = item / p.H, h = item % p.H;
        const uint32_t ph = static_cast<uint32_t>(local & 1);
        mbar_wait(qk_empty, ph ^ 1);
        mbar_expect_tx(qk_full, qk_bytes);
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl) {
          for (int g = 0; g < mt; ++g)
            tma_load_4d(sQ + (pl * 2 + g) * Q_TILE_BYTES, &map_q, qk_full, h * HD, g * 128, b, pl);
          tma_load_4d(sK + pl * K_PLANE_BYTES, &map_k, qk_full, D + h * HD, 0, b, pl);
        }
        mbar_wait(v_empty, ph ^ 1);
        mbar_expect_tx(v_full, v_bytes);
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl)
          for (int kb = 0; kb < nkb; ++kb)
            tma_load_4d(sV + (pl * 4 + kb) * V_BOX_BYTES, &map_v, v_full, 2 * D + h * HD, kb * 64, b, pl);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, n_keys, false, false);
      const uint32_t idesc_pv = umma_idesc_bf16(128, HD, false, true);
      const int ksteps_pv = n_keys >> 4;
      int local = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local) {
        const uint32_t ph = static_cast<uint32_t>(local & 1);
        // TMEM of the previous item fully drained by both softmax groups
        mbar_wait(&tmem_free[0], ph ^ 1);
        if (mt == 2) mbar_wait(&tmem_free[1], ph ^ 1);
        mbar_wait(qk_full, ph);
        tc_fence_after();
        // ---- S_g = Q_g K^T ----
        for (int g = 0; g < mt; ++g) {
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(g * S_COLS);
#pragma unroll
          for (int pr = 0; pr < C::NPAIRS_S; ++pr) {
            const int pa = (pr == 2) ? 1 : 0;
            const int pb = (pr == 1) ? 1 : 0;
            const uint32_t a_base = smem_u32(sQ + (pa * 2 + g) * Q_TILE_BYTES);
            const uint32_t b_base = smem_u32(sK + pb * K_PLANE_BYTES);
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) {
              const uint64_t da = umma_smem_desc(a_base + k * 32, 16u, 1024u);
              const uint64_t db = umma_smem_desc(b_base + k * 32, 16u, 1024u);
              umma_bf16(d_tmem, da, db, idesc_s, (pr > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&s_full[g]);
        }
        umma_commit(qk_empty);                 // Q / K smem may be refilled once the score MMAs have read it
        // ---- O_g = P_g V ----
        mbar_wait(v_full, ph);
        for (int g = 0; g < mt; ++g) {
          mbar_wait(&p_ready[g], ph);
          if (g == 1) mbar_wait(&o_full[0], ph);   // O1 lives in [0,64): P0 must have been consumed by the PV0 MMAs
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(g == 0 ? O0_COL : 0);
          const uint32_t p_tmem = tmem_base + static_cast<uint32_t>(g * S_COLS);
#pragma unroll
          for (int pr = 0; pr < C::NPAIRS_PV; ++pr) {
            // NPL == 2: (P_hi,V_hi) (P_hi,V_lo) (P_lo,V_hi);  NPL == 1: (P_hi,V) (P_lo,V)
            const int pa = (NPL == 2) ? (pr == 2 ? 1 : 0) : pr;
            const int pb = (NPL == 2) ? (pr == 1 ? 1 : 0) : 0;
            for (int kk = 0; kk < ksteps_pv; ++kk) {
              // 16 keys of P: 8 TMEM columns inside the 32-column chunk kk/2 -- [hi even | hi odd | lo even | lo odd]
              const uint32_t a_tmem = p_tmem + static_cast<uint32_t>((kk >> 1) * 32 + pa * 16 + (kk & 1) * 8);
              const uint64_t db = umma_smem_desc(smem_u32(sV + (pb * 4 + (kk >> 2)) * V_BOX_BYTES) + (kk & 3) * 2048, 8192u, 1024u);
              umma_bf16_ts(d_tmem, a_tmem, db, idesc_pv, (pr > 0 || kk > 0) ? 1u : 0u);
            }
          }
          umma_commit(&o_full[g]);
        }
        umma_commit(v_empty);
      }
    }
  } else {
    // =============================== softmax + output (one thread per query row) ===============================
    const int g = (warp - 2) >> 2;               // query tile of this warp group
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    if (g < mt) {
      const int row_in_tile = q * 32 + lane;
      const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
      const uint32_t s_tmem = tmem_base + lane_addr + static_cast<uint32_t>(g * S_COLS);
      const uint32_t o_tmem = tmem_base + lane_addr + static_cast<uint32_t>(g == 0 ? O0_COL : 0);
      const int nch = (n_keys + 31) >> 5;        // 32-column chunks (the tail chunk may hold 16 stale columns: masked)
      float sc = p.scale;
      float vs = 1.0f;
      if (p.qk_scale) { const float s = __ldg(p.qk_scale); sc *= s * s; }
      if (p.v_scale) vs = __ldg(p.v_scale);
      const float c2 = sc * 1.4426950408889634f;  // logits -> base-2 exponent
      int local = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local) {
        const int b = item / p.H, h = item % p.H;
        const uint32_t ph = static_cast<uint32_t>(local & 1);
        mbar_wait(&s_full[g], ph);
        tc_fence_after();
        uint32_t rr[32], nxt[32];
        // ---- pass 1: row max over the T valid keys ----
        float mx = -INFINITY;
        tmem_ld_32x32(s_tmem, nxt);
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) rr[j] = nxt[j];
          if (c + 1 < nch) tmem_ld_32x32(s_tmem + (c + 1) * 32, nxt);
          const int nvalid = p.T - c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = (j < nvalid) ? fmaxf(mx, __uint_as_float(rr[j])) : mx;
        }
        // ---- pass 2: e = exp2((s - max) * c2); row sum; P hi/lo written in place ----
        const float mxc = mx * c2;
        float sum = 0.f;
        tmem_ld_32x32(s_tmem, nxt);
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) rr[j] = nxt[j];
          if (c + 1 < nch) tmem_ld_32x32(s_tmem + (c + 1) * 32, nxt);
          const int nvalid = p.T - c * 32;
          uint32_t pk[32];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float e0 = (2 * j < nvalid) ? ex2_approx(fmaf(__uint_as_float(rr[2 * j]), c2, -mxc)) : 0.f;
            const float e1 = (2 * j + 1 < nvalid) ? ex2_approx(fmaf(__uint_as_float(rr[2 * j + 1]), c2, -mxc)) : 0.f;
            sum += e0 + e1;
            split_pack2(e0, e1, pk[j], pk[16 + j]);   // keys 32c+2j, +1: hi pairs in words 0..15, lo pairs in 16..31
          }
          tmem_st_32x32(s_tmem + c * 32, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_ready[g]);
        // ---- output: O / rowsum (* s) -> bf16 hi/lo planes ----
        mbar_wait(&o_full[g], ph);
        tc_fence_after();
        uint32_t o[64];
        tmem_ld_32x64(o_tmem, o);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&tmem_free[g]);
        const int t = g * 128 + row_in_tile;
        if (t < p.T) {
          const float inv = vs / sum;
          const int64_t row = static_cast<int64_t>(b) * p.T + t;
          if (p.out) {
            __nv_bfloat16* dst_hi = p.out + row * p.out_ld + h * HD;
            __nv_bfloat16* dst_lo = dst_hi + p.out_plane_stride;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                split_pack2(__uint_as_float(o[8 * j + 2 * e]) * inv, __uint_as_float(o[8 * j + 2 * e + 1]) * inv, hi[e], lo[e]);
              *reinterpret_cast<uint4*>(dst_hi + 8 * j) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(dst_lo + 8 * j) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
          if (p.out_f32) {
            float* dst = p.out_f32 + row * (static_cast<int64_t>(p.H) * HD) + h * HD;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(__uint_as_float(o[4 * j]) * inv, __uint_as_float(o[4 * j + 1]) * inv,
                                                                    __uint_as_float(o[4 * j + 2]) * inv, __uint_as_float(o[4 * j + 3]) * inv);
          }
          if (p.lse) p.lse[(static_cast<int64_t>(b) * p.H + h) * p.T + t] = mx * sc + logf(sum);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int NPL>
int launch_attn(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const AttnParams& ap, int grid,
                cudaStream_t st) {
  using C = AttnCfg<NPL>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(qv_attn_fwd_kernel<NPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  });
  QV_REQUIRE(attr_err == cudaSuccess, QV_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  qv_attn_fwd_kernel<NPL><<<grid, AT_THREADS, C::SMEM_BYTES, st>>>(mq, mk, mv, ap);
  return qv_check_launch("qv_attn_fwd");
}


// ====================================================================================================================
// Fused attention BACKWARD for integer-code operands (the QAT student: Q, K, V = s * codes), one (image, head) per item.
//
// Given dO (bf16 hi/lo planes), the codes and the forward's logsumexp, recomputes the probabilities on the tensor cores and
// produces dQ, dK, dV (fp32) without ever writing scores to HBM.  Replaces the autograd of F.scaled_dot_product_attention
// (4 batched GEMMs + softmax-backward passes in the unfused path).  With z = scale * Q K^T, P = softmax(z):
//     dP = dO V^T,  delta_i = sum_j P_ij dP_ij,  dz = P o (dP - delta),  dQ = scale dz K,  dK = scale dz^T Q,  dV = P^T dO.
// TMEM accumulators are row-per-lane, so the kernel runs two kinds of sub-pass per 128-row tile:
//   pass A (lanes = queries):  R0 = Q K^T, R1 = dO V^T  -> threads: delta_i, dz (bf16 hi/lo, in place over R1) -> dQ = dz K
//   pass B (lanes = keys):     R0 = K Q^T, R1 = V dO^T  -> threads: P^T, dz^T (in place)  -> dV = P^T dO, dK = dz^T Q
// (recomputing S / dP transposed costs two small extra MMAs and saves staging dz through shared memory).  All second-stage
// products are TS-mode MMAs (A from TMEM); the same [tokens x 64] shared-memory tiles serve as K-major operands of the first
// stage and MN-major operands of the second.
// TMEM: R0 [0,224) | R1 [224,448) | ACC0 [448,512) (dQ / dV) ; dK re-uses [0,64) once P^T has been consumed.
// Warps: 0 TMA, 1 MMA, 2-9 compute (two warps per TMEM lane quarter, alternating 32-column chunks).
// ====================================================================================================================
constexpr int BW_TILE_BYTES = 256 * HD * 2;     // 32 KB: up to 256 token rows x 64 bf16 (rows >= T zero-filled by TMA)
constexpr int BW_SMEM_BYTES = 5 * BW_TILE_BYTES + 4096 /*lse, delta, partials*/ + 1024 /*align*/ + 256 /*barriers*/;

struct AttnBwdParams {
  int32_t B, T, H;
  int32_t n_keys, m_tiles;
  float scale;
  const float* qscale;      // device scalar s (Q = K = V scale); may be null (= 1)
  const float* lse;         // [B*H*T] from the forward
  float* g_qkv;             // [B*T][3*D]: dQ | dK | dV column blocks
  // FUSED: the qkv Linear's backward prologue (qv_gp_planes) applied on the way out -- gq = g * STEmask(y_raw),
  // planes = hi/lo split of gq * w_scale[col], per-slab column sums of gq (bias grad partials)
  const float* y_raw;       // [B*T][3*D] raw qkv output
  const float* y_scale;
  const int32_t* y_zp;
  int32_t qmin, qmax;
  const float* w_scale;     // [3*D]
  __nv_bfloat16* gp;        // [2][B*T][3*D]
  int64_t gp_plane_stride;
  float* colsum;            // [B * m_tiles * 4][3*D]
};

template <bool FUSED>
__global__ void __launch_bounds__(AT_THREADS, 1)
qv_attn_bwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                   const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + BW_TILE_BYTES;
  uint8_t* sV = sK + BW_TILE_BYTES;
  uint8_t* sDOh = sV + BW_TILE_BYTES;
  uint8_t* sDOl = sDOh + BW_TILE_BYTES;
  float* lse2_s = reinterpret_cast<float*>(sDOl + BW_TILE_BYTES);   // [256] lse * log2(e)
  float* delta_s = lse2_s + 256;                                     // [256] sum_j P_ij dP_raw_ij
  float* part_s = delta_s + 256;                                     // [2 sub-pass parities][2 warps of a pair][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(part_s + 512);
  uint64_t* ld_full = bars + 0;
  uint64_t* ld_empty = bars + 1;
  uint64_t* mma1_done = bars + 2;
  uint64_t* cmp_done = bars + 3;
  uint64_t* acc_done = bars + 4;
  uint64_t* acc2_done = bars + 5;
  uint64_t* epi_done = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int D = p.H * HD;
  const int num_items = p.B * p.H;
  const int n_keys = p.n_keys;
  const int mt = p.m_tiles;
  const int ksteps = n_keys >> 4;
  constexpr uint32_t R1_COL = S_COLS, ACC0_COL = O0_COL;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_qkv);
    prefetch_tensormap(&map_do);
    mbar_init(ld_full, 1);
    mbar_init(ld_empty, 1);
    mbar_init(mma1_done, 1);
    mbar_init(cmp_done, 256);
    mbar_init(acc_done, 1);
    mbar_init(acc2_done, 1);
    mbar_init(epi_done, 256);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int local = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local) {
        const int b = item / p.H, h = item % p.H;
        mbar_wait(ld_empty, static_cast<uint32_t>(local & 1) ^ 1);
        mbar_expect_tx(ld_full, 5 * BW_TILE_BYTES);
        tma_load_4d(sQ, &map_qkv, ld_full, h * HD, 0, b, 0);
        tma_load_4d(sK, &map_qkv, ld_full, D + h * HD, 0, b, 0);
        tma_load_4d(sV, &map_qkv, ld_full, 2 * D + h * HD, 0, b, 0);
        tma_load_4d(sDOh, &map_do, ld_full, h * HD, 0, b, 0);
        tma_load_4d(sDOl, &map_do, ld_full, h * HD, 0, b, 1);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc_ss = umma_idesc_bf16(128, n_keys, false, false);
      const uint32_t idesc_ts = umma_idesc_bf16(128, HD, false, true);
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aDh = smem_u32(sDOh), aDl = smem_u32(sDOl);
      const uint32_t r0 = tmem_base, r1 = tmem_base + R1_COL, acc0 = tmem_base + ACC0_COL, acc1 = tmem_base;
      // first-stage SS product: D[tmem] (+)= A[rows tile*128.., K-major] * B[rows 0..n_keys, K-major]^T over d = 64
      auto ss = [&](uint32_t d_tmem, uint32_t a_tile, int tile, uint32_t b_tile, bool accumulate) {
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint64_t da = umma_smem_desc(a_tile + tile * (128 * 128) + k * 32, 16u, 1024u);
          const uint64_t db = umma_smem_desc(b_tile + k * 32, 16u, 1024u);
          umma_bf16(d_tmem, da, db, idesc_ss, (accumulate || k > 0) ? 1u : 0u);
        }
      };
      // second-stage TS product: D[tmem] (+)= A[tmem region, plane] * B[[tokens x 64] tile, MN-major] over the tokens
      auto ts = [&](uint32_t d_tmem, uint32_t a_region, int plane, uint32_t b_tile, bool accumulate) {
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint32_t a_tmem = a_region + static_cast<uint32_t>((kk >> 1) * 32 + plane * 16 + (kk & 1) * 8);
          const uint64_t db = umma_smem_desc(b_tile + kk * 2048, 8192u, 1024u);
          umma_bf16_ts(d_tmem, a_tmem, db, idesc_ts, (accumulate || kk > 0) ? 1u : 0u);
        }
      };
      uint32_t sp = 0;          // running sub-pass counter (phase of mma1_done / cmp_done / acc_done / epi_done)
      uint32_t spb = 0;         // running pass-B counter (phase of acc2_done)
      int local = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local) {
        mbar_wait(ld_full, static_cast<uint32_t>(local & 1));
        tc_fence_after();
        for (int g = 0; g < mt; ++g, ++sp) {                     // ---- pass A: lanes = queries of tile g ----
          mbar_wait(epi_done, (sp & 1) ^ 1);
          tc_fence_after();
          ss(r0, aQ, g, aK, false);                              // S = Q K^T
          ss(r1, aDh, g, aV, false);                             // dP = dO V^T (hi + lo)
          ss(r1, aDl, g, aV, true);
          umma_commit(mma1_done);
          mbar_wait(cmp_done, sp & 1);
          tc_fence_after();
          ts(acc0, r1, 0, aK, false);                            // dQ = dz K
          ts(acc0, r1, 1, aK, true);
          umma_commit(acc_done);
        }
        for (int kt = 0; kt < mt; ++kt, ++sp, ++spb) {           // ---- pass B: lanes = keys of tile kt ----
          mbar_wait(epi_done, (sp & 1) ^ 1);
          tc_fence_after();
          ss(r0, aK, kt, aQ, false);                             // S^T = K Q^T
          ss(r1, aV, kt, aDh, false);                            // dP^T = V dO^T (hi + lo)
          ss(r1, aV, kt, aDl, true);
          umma_commit(mma1_done);
          mbar_wait(cmp_done, sp & 1);
          tc_fence_after();
          ts(acc0, r0, 0, aDh, false);                           // dV = P^T dO : (hi,hi) (hi,lo) (lo,hi)
          ts(acc0, r0, 0, aDl, true);
          ts(acc0, r0, 1, aDh, true);
          umma_commit(acc_done);
          mbar_wait(acc_done, sp & 1);                           // P^T consumed: [0,64) may now hold dK
          tc_fence_after();
          ts(acc1, r1, 0, aQ, false);                            // dK = dz^T Q
          ts(acc1, r1, 1, aQ, true);
          umma_commit(acc2_done);
        }
        umma_commit(ld_empty);                                   // every MMA that reads this item's tiles has completed
      }
    }
  } else {
    // =============================== compute warps ===============================
    const int cw = warp - 2;
    const int q = warp & 3;
    const int par = cw >> 2;
    const int row = q * 32 + lane;               // row inside the 128-row tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t r0 = tmem_base + lane_addr, r1 = r0 + R1_COL, acc0 = r0 + ACC0_COL, acc1 = r0;
    const int nch = (n_keys + 31) >> 5;
    const float s = p.qscale ? __ldg(p.qscale) : 1.0f;
    const float c2 = p.scale * s * s * 1.4426950408889634f;
    const float gscale = p.scale * s * s;        // dQ, dK factor (see header comment)
    const int ctid = threadIdx.x - 64;           // 0..255
    QvQParams yq;
    if constexpr (FUSED) yq = qv_load_qparams(p.y_scale, p.y_zp, p.qmin, p.qmax);
    const int D3 = 3 * D;
    // FUSED output of one 32-row x 32-column block: lane = token `tok` of image b, columns col .. col+31 of the [., 3D] row
    auto load_y = [&](float4 (&yv)[8], int b, int tok, int col) {
      const float4* yp = reinterpret_cast<const float4*>(p.y_raw + (static_cast<int64_t>(b) * p.T + tok) * D3 + col);
#pragma unroll
      for (int j = 0; j < 8; ++j) yv[j] = (tok < p.T) ? __ldg(yp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto emit = [&](const uint32_t (&o)[32], const float4 (&yv)[8], float mult, int b, int tok, int col, int slab) {
      const bool ok = tok < p.T;
      float gq[32];
      __nv_bfloat16* dst = p.gp + (static_cast<int64_t>(b) * p.T + tok) * D3 + col;
#pragma unroll
      for (int j = 0; j < 4; ++j) {                                // 8 columns per step: one 16-byte store per plane
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          const float4 ws = __ldg(reinterpret_cast<const float4*>(p.w_scale + col) + 2 * j + hlf);
          const float4 y4 = yv[2 * j + hlf];
          const float yy[4] = {y4.x, y4.y, y4.z, y4.w};
          const float wv[4] = {ws.x, ws.y, ws.z, ws.w};
          float a[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float r = __fadd_rn(rintf(__fmul_rn(yy[e], yq.inv)), yq.zp);
            const bool in = (yq.qmin <= r) && (r <= yq.qmax);
            const float f = (ok && in) ? __uint_as_float(o[8 * j + 4 * hlf + e]) * mult : 0.f;
            gq[8 * j + 4 * hlf + e] = f;
            a[e] = f * wv[e];
          }
          split_pack2(a[0], a[1], hi[2 * hlf], lo[2 * hlf]);
          split_pack2(a[2], a[3], hi[2 * hlf + 1], lo[2 * hlf + 1]);
        }
        if (ok) {
          *reinterpret_cast<uint4*>(dst + 8 * j) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(dst + p.gp_plane_stride + 8 * j) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      const float cs = qv_warp_colsum32(gq, lane);
      p.colsum[static_cast<int64_t>(slab) * D3 + col + lane] = cs;
    };
    uint32_t sp = 0, spb = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int b = item / p.H, h = item % p.H;
      const float* lse_bh = p.lse + (static_cast<int64_t>(b) * p.H + h) * p.T;
      asm volatile("bar.sync 9, 256;" ::: "memory");            // previous item's pass B has finished reading lse2_s / delta_s
      lse2_s[ctid] = (ctid < p.T) ? __ldg(lse_bh + ctid) * 1.4426950408889634f : 0.0f;
      float* grow_base = p.g_qkv + static_cast<int64_t>(b) * p.T * (3 * D) + h * HD + par * 32;
      // ------------------------------ pass A ------------------------------
      for (int g = 0; g < mt; ++g, ++sp) {
        const int i = g * 128 + row;
        const float Li = (i < p.T) ? __ldg(lse_bh + i) * 1.4426950408889634f : 0.0f;
        mbar_wait(mma1_done, sp & 1);
        tc_fence_after();
        float dpart = 0.f;
#pragma unroll 1
        for (int c = par; c < nch; c += 2) {
          uint32_t sv[32], dv[32];
          tmem_ld_32x32(r0 + c * 32, sv);
          tmem_ld_32x32(r1 + c * 32, dv);
          tmem_ld_wait();
          const int nvalid = p.T - c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float pj = (j < nvalid) ? ex2_approx(fmaf(__uint_as_float(sv[j]), c2, -Li)) : 0.f;
            dpart = fmaf(pj, __uint_as_float(dv[j]), dpart);
          }
        }
        float* part = part_s + (sp & 1) * 256;
        part[par * 128 + row] = dpart;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        const float delta = part[row] + part[128 + row];
        if (par == 0) delta_s[g * 128 + row] = delta;
#pragma unroll 1
        for (int c = par; c < nch; c += 2) {
          uint32_t sv[32], dv[32], pk[32];
          tmem_ld_32x32(r0 + c * 32, sv);
          tmem_ld_32x32(r1 + c * 32, dv);
          tmem_ld_wait();
          const int nvalid = p.T - c * 32;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float p0 = (2 * j < nvalid) ? ex2_approx(fmaf(__uint_as_float(sv[2 * j]), c2, -Li)) : 0.f;
            const float p1 = (2 * j + 1 < nvalid) ? ex2_approx(fmaf(__uint_as_float(sv[2 * j + 1]), c2, -Li)) : 0.f;
            split_pack2(p0 * (__uint_as_float(dv[2 * j]) - delta), p1 * (__uint_as_float(dv[2 * j + 1]) - delta), pk[j], pk[16 + j]);
          }
          tmem_st_32x32(r1 + c * 32, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(cmp_done);
        // dQ tile: this warp stores columns par*32 .. par*32+31 of its 32 rows
        float4 yv[FUSED ? 8 : 1];
        if constexpr (FUSED) load_y(yv, b, i, h * HD + par * 32);       // in flight while the dQ MMAs finish
        mbar_wait(acc_done, sp & 1);
        tc_fence_after();
        uint32_t o[32];
        tmem_ld_32x32(acc0 + par * 32, o);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(epi_done);
        if constexpr (FUSED) {
          emit(o, yv, gscale, b, i, h * HD + par * 32, (b * mt + g) * 4 + q);
        } else if (i < p.T) {
          float* dst = grow_base + static_cast<int64_t>(i) * (3 * D);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(__uint_as_float(o[4 * j]) * gscale, __uint_as_float(o[4 * j + 1]) * gscale,
                                                                  __uint_as_float(o[4 * j + 2]) * gscale, __uint_as_float(o[4 * j + 3]) * gscale);
        }
      }
      asm volatile("bar.sync 9, 256;" ::: "memory");            // lse2_s and delta_s complete for every query row
      // ------------------------------ pass B ------------------------------
      for (int kt = 0; kt < mt; ++kt, ++sp, ++spb) {
        const int jrow = kt * 128 + row;                          // key index of this lane
        mbar_wait(mma1_done, sp & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = par; c < nch; c += 2) {
          uint32_t sv[32], dv[32], pp[32], pz[32];
          tmem_ld_32x32(r0 + c * 32, sv);
          tmem_ld_32x32(r1 + c * 32, dv);
          tmem_ld_wait();
          const int nvalid = p.T - c * 32;                        // valid query columns of this chunk
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 L = *reinterpret_cast<const float2*>(lse2_s + c * 32 + 2 * j);      // broadcast reads
            const float2 dl = *reinterpret_cast<const float2*>(delta_s + c * 32 + 2 * j);
            const float p0 = (2 * j < nvalid) ? ex2_approx(fmaf(__uint_as_float(sv[2 * j]), c2, -L.x)) : 0.f;
            const float p1 = (2 * j + 1 < nvalid) ? ex2_approx(fmaf(__uint_as_float(sv[2 * j + 1]), c2, -L.y)) : 0.f;
            split_pack2(p0, p1, pp[j], pp[16 + j]);
            split_pack2(p0 * (__uint_as_float(dv[2 * j]) - dl.x), p1 * (__uint_as_float(dv[2 * j + 1]) - dl.y), pz[j], pz[16 + j]);
          }
          tmem_st_32x32(r0 + c * 32, pp);
          tmem_st_32x32(r1 + c * 32, pz);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(cmp_done);
        uint32_t o[32];
        float4 yv[FUSED ? 8 : 1];
        if constexpr (FUSED) load_y(yv, b, jrow, 2 * D + h * HD + par * 32);
        mbar_wait(acc_done, sp & 1);                              // dV
        tc_fence_after();
        tmem_ld_32x32(acc0 + par * 32, o);
        tmem_ld_wait();
        if constexpr (FUSED) {
          emit(o, yv, 1.0f, b, jrow, 2 * D + h * HD + par * 32, (b * mt + kt) * 4 + q);
          load_y(yv, b, jrow, D + h * HD + par * 32);
        } else if (jrow < p.T) {
          float* dst = grow_base + static_cast<int64_t>(jrow) * (3 * D) + 2 * D;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(__uint_as_float(o[4 * j]), __uint_as_float(o[4 * j + 1]),
                                                                  __uint_as_float(o[4 * j + 2]), __uint_as_float(o[4 * j + 3]));
        }
        mbar_wait(acc2_done, spb & 1);                            // dK
        tc_fence_after();
        tmem_ld_32x32(acc1 + par * 32, o);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(epi_done);
        if constexpr (FUSED) {
          emit(o, yv, gscale, b, jrow, D + h * HD + par * 32, (b * mt + kt) * 4 + q);
        } else if (jrow < p.T) {
          float* dst = grow_base + static_cast<int64_t>(jrow) * (3 * D) + D;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(__uint_as_float(o[4 * j]) * gscale, __uint_as_float(o[4 * j + 1]) * gscale,
                                                                  __uint_as_float(o[4 * j + 2]) * gscale, __uint_as_float(o[4 * j + 3]) * gscale);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

extern "C" int qv_attn_fwd(const uint16_t* qkv_planes, int32_t n_planes, int64_t plane_stride, int64_t ld, int32_t B,
                           int32_t T, int32_t H, float scale, const float* qk_scale, const float* v_scale,
                           uint16_t* out_planes, int64_t out_plane_stride, int64_t out_ld, float* out_f32, float* lse,
                           void* stream) {
  QV_REQUIRE(qkv_planes && (out_planes || out_f32) && B > 0 && T > 0 && H > 0, QV_ERR_INVALID, "bad attn_fwd arguments");
  QV_REQUIRE(n_planes == 1 || n_planes == 2, QV_ERR_INVALID, "n_planes must be 1 (integer codes) or 2 (fp32 hi/lo)");
  QV_REQUIRE(T <= 224, QV_ERR_UNSUPPORTED, "fused attention holds all keys in one tile: T <= 224 (got %d)", T);
  QV_REQUIRE(ld >= 3LL * H * HD, QV_ERR_INVALID, "qkv row pitch must cover Q | K | V (3 * H * 64 columns)");
  QV_REQUIRE(!out_planes || (out_ld >= static_cast<int64_t>(H) * HD && out_ld % 8 == 0 && out_plane_stride % 8 == 0 &&
                             qv_aligned16(out_planes)),
             QV_ERR_INVALID, "output planes must be 16-byte aligned with pitches that are multiples of 8 bf16");
  QV_REQUIRE(!out_f32 || qv_aligned16(out_f32), QV_ERR_INVALID, "fp32 output must be 16-byte aligned");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  AttnParams ap;
  memset(&ap, 0, sizeof(ap));
  ap.B = B; ap.T = T; ap.H = H;
  ap.n_keys = (T + 15) / 16 * 16;
  ap.m_tiles = (T + 127) / 128;
  ap.scale = scale;
  ap.qk_scale = qk_scale;
  ap.v_scale = v_scale;
  ap.out = reinterpret_cast<__nv_bfloat16*>(out_planes);
  ap.out_plane_stride = out_plane_stride;
  ap.out_ld = out_ld;
  ap.out_f32 = out_f32;
  ap.lse = lse;
  // one tensor, three box shapes: per-image matrices [T rows, ld cols]; rows >= T are zero-filled by TMA
  qv_operand op;
  memset(&op, 0, sizeof(op));
  op.ptr = qkv_planes; op.ld = ld; op.plane_stride = plane_stride; op.rows = T; op.cols = 3LL * H * HD;
  op.nb = B; op.batch_stride = static_cast<int64_t>(T) * ld;
  CUtensorMap mq, mk, mv;
  int rc = make_map(&mq, op, n_planes, 128);
  if (rc) return rc;
  rc = make_map(&mk, op, n_planes, ap.n_keys);
  if (rc) return rc;
  rc = make_map(&mv, op, n_planes, 64);
  if (rc) return rc;
  const int items = B * H;
  const int sms = qv_num_sms();
  const int grid = items < sms ? items : sms;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_planes == 2) return launch_attn<2>(mq, mk, mv, ap, grid, st);
  return launch_attn<1>(mq, mk, mv, ap, grid, st);
}

namespace {
int attn_bwd_impl(const uint16_t* qkv_codes, int64_t ld, const float* qscale, const uint16_t* do_planes, int64_t do_plane_stride,
                  int64_t do_ld, const float* lse, int32_t B, int32_t T, int32_t H, float scale, float* g_qkv,
                  const AttnBwdParams* fused, void* stream);
}

extern "C" int qv_attn_bwd(const uint16_t* qkv_codes, int64_t ld, const float* qscale, const uint16_t* do_planes,
                           int64_t do_plane_stride, int64_t do_ld, const float* lse, int32_t B, int32_t T, int32_t H, float scale,
                           float* g_qkv, void* stream) {
  QV_REQUIRE(g_qkv && qv_aligned16(g_qkv), QV_ERR_INVALID, "g_qkv must be a 16-byte aligned device pointer");
  return attn_bwd_impl(qkv_codes, ld, qscale, do_planes, do_plane_stride, do_ld, lse, B, T, H, scale, g_qkv, nullptr, stream);
}

extern "C" int qv_attn_bwd_gp(const uint16_t* qkv_codes, int64_t ld, const float* qscale, const uint16_t* do_planes,
                              int64_t do_plane_stride, int64_t do_ld, const float* lse, int32_t B, int32_t T, int32_t H,
                              float scale, const float* y_raw, const float* y_scale, const int32_t* y_zp, int32_t qmin,
                              int32_t qmax, const float* w_scale, uint16_t* gp_planes, int64_t gp_plane_stride,
                              float* colsum, void* stream) {
  QV_REQUIRE(y_raw && y_scale && y_zp && w_scale && gp_planes && colsum, QV_ERR_INVALID, "bad attn_bwd_gp arguments");
  QV_REQUIRE(qv_aligned16(y_raw) && qv_aligned16(w_scale) && qv_aligned16(gp_planes) && gp_plane_stride % 8 == 0, QV_ERR_INVALID,
             "y_raw / w_scale / gp_planes must be 16-byte aligned (plane stride a multiple of 8 bf16)");
  AttnBwdParams f;
  memset(&f, 0, sizeof(f));
  f.y_raw = y_raw; f.y_scale = y_scale; f.y_zp = y_zp; f.qmin = qmin; f.qmax = qmax; f.w_scale = w_scale;
  f.gp = reinterpret_cast<__nv_bfloat16*>(gp_planes); f.gp_plane_stride = gp_plane_stride; f.colsum = colsum;
  return attn_bwd_impl(qkv_codes, ld, qscale, do_planes, do_plane_stride, do_ld, lse, B, T, H, scale, nullptr, &f, stream);
}

namespace {
int attn_bwd_impl(const uint16_t* qkv_codes, int64_t ld, const float* qscale, const uint16_t* do_planes, int64_t do_plane_stride,
                  int64_t do_ld, const float* lse, int32_t B, int32_t T, int32_t H, float scale, float* g_qkv,
                  const AttnBwdParams* fused, void* stream) {
  QV_REQUIRE(qkv_codes && do_planes && lse && (g_qkv || fused) && B > 0 && T > 0 && H > 0, QV_ERR_INVALID, "bad attn_bwd arguments");
  QV_REQUIRE(T <= 224, QV_ERR_UNSUPPORTED, "fused attention holds all keys in one tile: T <= 224 (got %d)", T);
  QV_REQUIRE(ld >= 3LL * H * HD && do_ld >= static_cast<int64_t>(H) * HD, QV_ERR_INVALID, "row pitches too small");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  AttnBwdParams ap;
  memset(&ap, 0, sizeof(ap));
  if (fused) ap = *fused;
  ap.B = B; ap.T = T; ap.H = H;
  ap.n_keys = (T + 15) / 16 * 16;
  ap.m_tiles = (T + 127) / 128;
  ap.scale = scale;
  ap.qscale = qscale;
  ap.lse = lse;
  ap.g_qkv = g_qkv;
  qv_operand op;
  memset(&op, 0, sizeof(op));
  op.ptr = qkv_codes; op.ld = ld; op.plane_stride = 0; op.rows = T; op.cols = 3LL * H * HD;
  op.nb = B; op.batch_stride = static_cast<int64_t>(T) * ld;
  CUtensorMap mq, md;
  int rc = make_map(&mq, op, 1, 256);
  if (rc) return rc;
  op.ptr = do_planes; op.ld = do_ld; op.plane_stride = do_plane_stride; op.cols = static_cast<int64_t>(H) * HD;
  op.batch_stride = static_cast<int64_t>(T) * do_ld;
  rc = make_map(&md, op, 2, 256);
  if (rc) return rc;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(qv_attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM_BYTES);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(qv_attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM_BYTES);
  });
  QV_REQUIRE(attr_err == cudaSuccess, QV_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  const int items = B * H;
  const int sms = qv_num_sms();
  const int grid = items < sms ? items : sms;
  if (fused) qv_attn_bwd_kernel<true><<<grid, AT_THREADS, BW_SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(mq, md, ap);
  else qv_attn_bwd_kernel<false><<<grid, AT_THREADS, BW_SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(mq, md, ap);
  return qv_check_launch("qv_attn_bwd");
}
}  // namespace
