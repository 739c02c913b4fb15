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

constexpr int AT_THREADS = 576;                // forward: warp 0 TMA, warp 1 MMA, warps 2-17 softmax / output (two per row)
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
  // per softmax warp: the output staging the TMA stores read.  One 2 KB buffer (32 rows x 64 B) where shared memory is full
  // (hi/lo operand planes: 224 KB), two for the one-plane code operands (hi and lo leave without an intermediate wait)
  static constexpr int STAGE_PER_WARP = (NPL == 1) ? 4096 : 2048;
  static constexpr int STAGE_BYTES = 16 * STAGE_PER_WARP;
  static constexpr int SMEM_BYTES = Q_BYTES + K_BYTES + V_BYTES + 8192 /*row max / exponent / sum exchange*/ + STAGE_BYTES + 1024 /*align slack*/ +
                                    256 /*barriers*/;
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
  int32_t out_fmt;       // 0 = bf16 hi/lo planes, 1 = mixed fp16 + fp8 planes (the teacher's proj operand)
  float* lse;            // optional [B*H*T]: scale' * rowmax + ln(rowsum)   (natural-log logsumexp of the scaled logits)
  int32_t* sat_flag;     // out_fmt 1: OR sat_bit into *sat_flag when an output leaves the mixed format's range (may be null)
  int32_t sat_bit;
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

// 32 lanes x 16 consecutive columns (one 16-key piece: the register budget of the two-warps-per-lane-quarter layouts)
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// Packed fp32 pairs (FFMA2 / FADD2 / FMUL2: one issue slot for two elements -- the chunk / softmax warps are issue- and
// latency-bound, not pipe-bound).  Same IEEE round-to-nearest results as the scalar forms, element by element.
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t f2_pack_u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// (a0, a1) -> bf16 hi pair + bf16 pair of the exact residuals (same values as split_pack2; the residual is one packed subtract)
__device__ __forceinline__ void split_pack2_f2(uint64_t a, uint32_t& hi, uint32_t& lo) {
  float a0, a1;
  f2_unpack(a, a0, a1);
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(a0, a1);
  hi = *reinterpret_cast<const uint32_t*>(&h2);
  float r0, r1;
  f2_unpack(f2_sub(a, f2_pack_u(hi << 16, hi & 0xffff0000u)), r0, r1);
  const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
  lo = *reinterpret_cast<const uint32_t*>(&l2);
}

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

#ifdef QV_ATTN_DEBUG
__device__ unsigned long long qv_dbg_buf[3][8192];
__device__ __forceinline__ void dbg_event(int who, int& n, int tag) {
  if (blockIdx.x == 0 && n < 8192) qv_dbg_buf[who][n++] = (static_cast<unsigned long long>(tag) << 48) | (clock64() & 0xffffffffffffULL);
}
#define DBG(who, tag) dbg_event(who, dbg_n, tag)
#else
#define DBG(who, tag)
#endif

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// Softmax of one 16-key unit for one query row (TMEM lane): 16 fp32 logits (already in registers) in, un-normalised
// probabilities 2^(s * c2 - r) out IN PLACE as bf16 pairs [hi 8 words | lo 8 words] (the A operand of the P V product: each
// 16-key k-step reads 8 consecutive columns).  c2p = (scale * log2 e) twice, nrp = -r twice (r: the row's reference exponent);
// returns the two partial sums packed.
template <bool MASKED>
__device__ __forceinline__ uint64_t fw_softmax_unit(uint32_t taddr, const uint32_t (&sv)[16], uint64_t c2p, uint64_t nrp,
                                                    uint64_t sum2, int nvalid) {
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float e0, e1;
    f2_unpack(f2_fma(f2_pack_u(sv[2 * j], sv[2 * j + 1]), c2p, nrp), e0, e1);
    e0 = ex2_approx(e0);
    e1 = ex2_approx(e1);
    if constexpr (MASKED) {
      e0 = (2 * j < nvalid) ? e0 : 0.f;
      e1 = (2 * j + 1 < nvalid) ? e1 : 0.f;
    }
    const uint64_t e = f2_pack(e0, e1);
    sum2 = f2_add(sum2, e);
    split_pack2_f2(e, pk[j], pk[8 + j]);
  }
  tmem_st_32x16(taddr, pk);
  return sum2;
}
// max of a unit's valid logits
__device__ __forceinline__ float fw_unit_max(const uint32_t (&sv)[16], int nvalid) {
  float m0 = -INFINITY, m1 = -INFINITY;
  if (nvalid >= 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      m0 = fmax3(m0, __uint_as_float(sv[4 * j]), __uint_as_float(sv[4 * j + 1]));
      m1 = fmax3(m1, __uint_as_float(sv[4 * j + 2]), __uint_as_float(sv[4 * j + 3]));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) m0 = (j < nvalid) ? fmaxf(m0, __uint_as_float(sv[j])) : m0;
  }
  return fmaxf(m0, m1);
}
// RARE PATH: multiply the probabilities already written for units [u0, u1) of this lane's row by f, an exact power of two <= 1
// (the row's reference exponent moved up): hi and lo bf16 parts scale exactly (underflow flushes towards zero, as it should).
__device__ __noinline__ void fw_rescale_units(uint32_t s_tmem, int u0, int u1, float f) {
  for (int u = u0; u < u1; ++u) {
    uint32_t w[16];
    tmem_ld_32x16(s_tmem + u * 16, w);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(w[j] << 16) * f, __uint_as_float(w[j] & 0xffff0000u) * f);
      w[j] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    tmem_st_32x16(s_tmem + u * 16, w);
  }
  tmem_st_wait();
}

// Each lane of a warp holds NCH 16-byte chunks (NCH * 16 contiguous bytes) of ITS OWN row; the 32 rows are `row_pitch` bytes
// apart in global memory.  Written straight from the registers, one store instruction would touch 32 different 128-byte lines
// (one LSU wavefront per lane: the forward's output phase was bound by exactly that).  Instead the tile goes through a 2 KB
// swizzled shared-memory buffer and leaves row-run by row-run: a store instruction then covers 32 / NCH rows x (NCH * 16) bytes.
// `base` = global address of the warp's first row's run; rows >= n_rows are skipped.  Whole warp calls; NCH = 2 or 4.
template <int NCH>
__device__ __forceinline__ void warp_store_row_runs(uint8_t* stage, int lane, const uint4 (&v)[NCH], uint8_t* base, int64_t row_pitch,
                                                    int n_rows) {
  const uint32_t sbase = smem_u32(stage);
  __syncwarp();                                             // the previous pass through this buffer has been read out
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const uint32_t pos = static_cast<uint32_t>(c ^ ((lane >> 1) & (NCH - 1)));
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + lane * (NCH * 16) + pos * 16), "r"(v[c].x), "r"(v[c].y),
                 "r"(v[c].z), "r"(v[c].w) : "memory");
  }
  __syncwarp();
  constexpr int RPI = 32 / NCH;                             // rows per instruction
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int r = k * RPI + lane / NCH, c = lane % NCH;
    const uint32_t pos = static_cast<uint32_t>(c ^ ((r >> 1) & (NCH - 1)));
    uint4 w;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w)
                 : "r"(sbase + r * (NCH * 16) + pos * 16) : "memory");
    if (r < n_rows) *reinterpret_cast<uint4*>(base + r * row_pitch + c * 16) = w;
  }
}

// A warp's 32 rows x (NCH * 16) bytes (lane = row) written to a staging buffer in the layout a TMA store box expects: NCH = 4 ->
// 64-byte rows under the 64B swizzle (16-byte chunk j of row r at j ^ ((r >> 1) & 3)); NCH = 2 -> plain 32-byte rows.
template <int NCH>
__device__ __forceinline__ void stage_rows_tma(uint8_t* stage, int lane, const uint4 (&v)[NCH]) {
  const uint32_t sbase = smem_u32(stage) + static_cast<uint32_t>(lane) * (NCH * 16);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const uint32_t pos = (NCH == 4) ? static_cast<uint32_t>(c ^ ((lane >> 1) & 3)) : static_cast<uint32_t>(c);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + pos * 16), "r"(v[c].x), "r"(v[c].y), "r"(v[c].z), "r"(v[c].w)
                 : "memory");
  }
}

template <int NPL>
__global__ void __launch_bounds__(AT_THREADS, 1)
qv_attn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o,
                   const __grid_constant__ CUtensorMap map_o8, const AttnParams p) {
  using C = AttnCfg<NPL>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                       // [NPL][2 tiles][128 x 64]
  uint8_t* sK = sQ + C::Q_BYTES;            // [NPL][224 x 64]
  uint8_t* sV = sK + C::K_BYTES;            // [NPL][4 boxes][64 keys x 64]
  float* xchg = reinterpret_cast<float*>(sV + C::V_BYTES);      // [3 kinds: first-unit max, reference exponent, sum][2 tiles][2 key halves][128 rows]
  uint8_t* stage_all = reinterpret_cast<uint8_t*>(xchg + 2048); // [16 softmax warps][2 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_all + C::STAGE_BYTES);
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
    if (p.out) {
      prefetch_tensormap(&map_o);
      prefetch_tensormap(&map_o8);
    }
    mbar_init(qk_full, 1);
    mbar_init(qk_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_ready[g], 256);
      mbar_init(&o_full[g], 1);
      mbar_init(&tmem_free[g], 256);
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
    // whole warp walks the schedule, one elected lane issues (operands stay in uniform registers; see gemm_sm100.cu)
    {
      const uint32_t idesc_s = umma_idesc_bf16(128, n_keys, false, false);
      const uint32_t idesc_pv = umma_idesc_bf16(128, HD, false, true);
      const int ksteps_pv = n_keys >> 4;
      const uint64_t dQ0 = umma_smem_desc(smem_u32(sQ), 16u, 1024u);          // K-major Q tiles / K planes
      const uint64_t dK0 = umma_smem_desc(smem_u32(sK), 16u, 1024u);
      const uint64_t dV0 = umma_smem_desc(smem_u32(sV), 8192u, 1024u);        // MN-major V boxes
      int local = 0;
#ifdef QV_ATTN_DEBUG
      int dbg_n = 0;
#endif
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local) {
        const uint32_t ph = static_cast<uint32_t>(local & 1);
        DBG(0, 21);
        mbar_wait(qk_full, ph);
        DBG(0, 22);
        // ---- S_g = Q_g K^T (tile g's TMEM region must have been drained: O1 lives inside S0's columns) ----
        for (int g = 0; g < mt; ++g) {
          if (g == 0) {
            mbar_wait(&tmem_free[0], ph ^ 1);
            if (mt == 2) mbar_wait(&tmem_free[1], ph ^ 1);
          }
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(g * S_COLS);
          if (elect_one()) {
#pragma unroll
            for (int pr = 0; pr < C::NPAIRS_S; ++pr) {
              const int pa = (pr == 2) ? 1 : 0;
              const int pb = (pr == 1) ? 1 : 0;
              const uint64_t da = dQ0 + static_cast<uint64_t>((pa * 2 + g) * (Q_TILE_BYTES >> 4));
              const uint64_t db = dK0 + static_cast<uint64_t>(pb * (K_PLANE_BYTES >> 4));
#pragma unroll
              for (int k = 0; k < HD / 16; ++k) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc_s, (pr > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&s_full[g]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(qk_empty);  // Q / K smem may be refilled once the score MMAs have read it
        __syncwarp();
        DBG(0, 23);
        // ---- O_g = P_g V ----
        mbar_wait(v_full, ph);
        for (int g = 0; g < mt; ++g) {
          mbar_wait(&p_ready[g], ph);
          DBG(0, 24 + 3 * g);
          if (g == 1) mbar_wait(&o_full[0], ph);   // O1 lives in [0,64): P0 must have been consumed by the PV0 MMAs
          tc_fence_after();
          DBG(0, 25 + 3 * g);
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(g == 0 ? O0_COL : 0);
          const uint32_t p_tmem = tmem_base + static_cast<uint32_t>(g * S_COLS);
          if (elect_one()) {
#pragma unroll
            for (int pr = 0; pr < C::NPAIRS_PV; ++pr) {
              // NPL == 2: (P_hi,V_hi) (P_hi,V_lo) (P_lo,V_hi);  NPL == 1: (P_hi,V) (P_lo,V)
              const int pa = (NPL == 2) ? (pr == 2 ? 1 : 0) : pr;
              const int pb = (NPL == 2) ? (pr == 1 ? 1 : 0) : 0;
              const uint64_t dv = dV0 + static_cast<uint64_t>(pb * 4 * (V_BOX_BYTES >> 4));
              for (int kk = 0; kk < ksteps_pv; ++kk) {
                // 16 keys of P = 16 TMEM columns: [hi pairs 8 | lo pairs 8]
                const uint32_t a_tmem = p_tmem + static_cast<uint32_t>(kk * 16 + pa * 8);
                // key step kk: box kk/4 (8 KB each), 2048 bytes per step inside the box -> contiguous: kk * 2048 bytes
                umma_bf16_ts(d_tmem, a_tmem, dv + static_cast<uint64_t>(kk) * 128u, idesc_pv, (pr > 0 || kk > 0) ? 1u : 0u);
              }
            }
            umma_commit(&o_full[g]);
          }
          __syncwarp();
          DBG(0, 26 + 3 * g);
        }
        if (elect_one()) umma_commit(v_empty);
        __syncwarp();
      }
    }
  } else {
    // =============================== softmax + output: TWO threads per query row ===============================
    // Warp (g, q, hf): query tile g, TMEM lane quarter q (= warp id mod 4, the hardware's access rule), key half hf.  The two
    // warps of a (g, q) pair split the row's keys in 16-key units (7 + 6 of 13 at T = 197) and its 64 output columns in halves;
    // row max and row sum cross between them through shared memory and a 64-thread named barrier.  Four softmax warps per
    // scheduler partition instead of two: the phase was latency-bound (issue slots 36 % busy, tensor pipe 15 %).
    const int j4 = (warp - 2) >> 2;              // 0..3
    const int g = j4 >> 1;                       // query tile
    const int hf = j4 & 1;                       // key half / output column half
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    if (g < mt) {
      const int row_in_tile = q * 32 + lane;
      const bool warp_has_rows = g * 128 + q * 32 < p.T;          // else: all 32 rows are padding -- only keep the barrier protocol
      const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
      const uint32_t s_tmem = tmem_base + lane_addr + static_cast<uint32_t>(g * S_COLS);
      const uint32_t o_tmem = tmem_base + lane_addr + static_cast<uint32_t>(g == 0 ? O0_COL : 0) + static_cast<uint32_t>(hf * 32);
      const int nu = n_keys >> 4;                // 16-key units
      const int u_lo = hf == 0 ? 0 : (nu + 1) >> 1, u_hi = hf == 0 ? (nu + 1) >> 1 : nu;
      float* x_max = xchg + (g * 2 + hf) * 128 + row_in_tile;      // mine; the partner's is 128 floats away
      float* x_sum = x_max + 1024;                                 // x_max + 512: the reference exponents
      const int partner = (hf == 0) ? 128 : -128;
      const int bar_id = 1 + g * 4 + q;
      constexpr int SPW = C::STAGE_PER_WARP;
      uint8_t* my_stage = stage_all + (warp - 2) * SPW;
      float sc = p.scale;
      float vs = 1.0f;
      if (p.qk_scale) { const float s = __ldg(p.qk_scale); sc *= s * s; }
      if (p.v_scale) vs = __ldg(p.v_scale);
      const float c2 = sc * 1.4426950408889634f;  // logits -> base-2 exponent
      const uint64_t c2p = f2_pack(c2, c2);
      float amax = 0.f;                           // mixed plane output: largest |value| written by this thread
      int local = 0;
#ifdef QV_ATTN_DEBUG
      int dbg_n = (threadIdx.x == 128 || threadIdx.x == 384) ? 0 : 8192;     // warp 4 = (g 0, q 0, hf 0); warp 12 = (g 1, q 0, hf 0)
      const int dbg_who = threadIdx.x == 128 ? 1 : 2;
#endif
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local) {
        const int b = item / p.H, h = item % p.H;
        const uint32_t ph = static_cast<uint32_t>(local & 1);
        DBG(dbg_who, 30);
        mbar_wait(&s_full[g], ph);
        tc_fence_after();
        DBG(dbg_who, 31);
        // ---- ONE pass over the scores (TMEM reads, 64 B/clk per SM, bound this phase: a separate max pass read every score
        // twice).  P_j = 2^(s_j c2 - r) for a per-row reference exponent r -- ANY r gives the same normalised result as long as
        // nothing overflows, so r = ceil(c2 * max over the first 16-key unit of both halves) (exchanged), and a later unit whose
        // maximum exceeds r by more than 2^64 moves r up by an integer: what that lane has already written is rescaled by the
        // exact power of two (rare: it needs a key > 44 nats above the best of 32 sampled keys).  The halves reconcile their r
        // the same way before P is released.  lse = r ln 2 + ln(sum). ----
        float r = -INFINITY, sum = 0.f;            // r in log2 units, integer-valued
        if (warp_has_rows) {
          uint32_t sv[16];
          float m0 = -INFINITY;
          if (u_lo < u_hi) {
            tmem_ld_32x16(s_tmem + u_lo * 16, sv);
            tmem_ld_wait();
            m0 = fw_unit_max(sv, p.T - u_lo * 16);
          }
          *x_max = m0;
          asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
          r = ceilf(fmaxf(m0, x_max[partner]) * c2);
          DBG(dbg_who, 32);
          uint64_t sum2 = 0ull;
#pragma unroll 1
          for (int u = u_lo; u < u_hi; ++u) {
            const int nvalid = p.T - u * 16;
            if (u > u_lo) {
              tmem_ld_32x16(s_tmem + u * 16, sv);
              tmem_ld_wait();
              const float need = fmaf(fw_unit_max(sv, nvalid), c2, -r);
              if (__any_sync(0xffffffffu, need > 64.f)) {              // rare: re-base this lane's row
                const float r_new = need > 64.f ? r + ceilf(need) : r;
                const float f = ex2_approx(r - r_new);                    // exact: integer argument (1.0 for untouched lanes)
                fw_rescale_units(s_tmem, u_lo, u, f);
                sum2 = f2_mul(sum2, f2_pack(f, f));
                r = r_new;
              }
            }
            const float nr = -r;
            if (nvalid >= 16) sum2 = fw_softmax_unit<false>(s_tmem + u * 16, sv, c2p, f2_pack(nr, nr), sum2, nvalid);
            else sum2 = fw_softmax_unit<true>(s_tmem + u * 16, sv, c2p, f2_pack(nr, nr), sum2, nvalid);
          }
          float s0, s1;
          f2_unpack(sum2, s0, s1);
          sum = s0 + s1;
          x_max[512] = r;                                               // third exchange array: [1024, 1536)
          *x_sum = sum;
          asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
          const float r_p = x_max[512 + partner], sum_p = x_sum[partner];
          const float R = fmaxf(r, r_p);
          const float f = ex2_approx(r - R), f_p = ex2_approx(r_p - R);   // exact powers of two; 2^-inf = 0 for an empty half
          if (__any_sync(0xffffffffu, f != 1.0f) && u_lo < u_hi) fw_rescale_units(s_tmem, u_lo, u_hi, f);
          sum = sum * f + sum_p * f_p;
          r = R;
          tmem_st_wait();
        }
        tc_fence_before();
        mbar_arrive(&p_ready[g]);
        DBG(dbg_who, 33);
        // ---- output: O / rowsum (* s) -> bf16 hi/lo planes, this thread's 32 of the 64 columns ----
        mbar_wait(&o_full[g], ph);
        tc_fence_after();
        DBG(dbg_who, 34);
        uint32_t o[32];
        if (warp_has_rows) {
          tmem_ld_32x32(o_tmem, o);
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive(&tmem_free[g]);
        DBG(dbg_who, 36);
        const int t = g * 128 + row_in_tile;
        if (warp_has_rows) {
          // rows of this warp: t0w .. t0w + 31, the first n_rows of them real
          const int t0w = g * 128 + q * 32;
          const int n_rows = min(32, p.T - t0w);
          const float inv = vs / sum;                               // padded rows: finite garbage, never stored
          const int64_t row0 = static_cast<int64_t>(b) * p.T + t0w;
          if (p.out) {
            if (p.out_fmt == 1) {      // a head's 64 columns are exactly one block of the mixed format: 64 hi8 | 64 lo8
              uint4 h16v[4], h8v[2], l8v[2];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint32_t h16[4], ph8[4], pl8[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float a0 = __uint_as_float(o[8 * j + 2 * e]) * inv, a1 = __uint_as_float(o[8 * j + 2 * e + 1]) * inv;
                  if (t < p.T) amax = fmaxf(amax, fmaxf(fabsf(a0), fabsf(a1)));
                  h16[e] = qv_mix_split2<QV_MIX_ACT>(a0, a1, ph8[e], pl8[e]);
                }
                h16v[j] = make_uint4(h16[0], h16[1], h16[2], h16[3]);
                const uint32_t w0 = ph8[0] | (ph8[1] << 16), w1 = ph8[2] | (ph8[3] << 16);
                const uint32_t x0 = pl8[0] | (pl8[1] << 16), x1 = pl8[2] | (pl8[3] << 16);
                if (j & 1) { h8v[j >> 1].z = w0; h8v[j >> 1].w = w1; l8v[j >> 1].z = x0; l8v[j >> 1].w = x1; }
                else { h8v[j >> 1].x = w0; h8v[j >> 1].y = w1; l8v[j >> 1].x = x0; l8v[j >> 1].y = x1; }
              }
              // fp16 region: a 32-column box; region 1: this half's 32 hi8 and 32 lo8 bytes = two 16-"element" boxes of the block
              const int col = h * HD + hf * 32, col8 = h * HD + hf * 16;
              if (lane == 0) tma_store_wait_read<0>();            // the previous item's stores have read the staging buffer
              __syncwarp();
              stage_rows_tma<4>(my_stage, lane, h16v);
              uint8_t* st8 = my_stage + (SPW >= 4096 ? 2048 : 0);
              if constexpr (SPW >= 4096) {
                stage_rows_tma<2>(st8, lane, h8v);
                stage_rows_tma<2>(st8 + 1024, lane, l8v);
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_4d(&map_o, my_stage, col, t0w, b, 0);
                if constexpr (SPW >= 4096) {
                  tma_store_4d(&map_o8, st8, col8, t0w, b, 1);
                  tma_store_4d(&map_o8, st8 + 1024, col8 + 32, t0w, b, 1);
                }
                tma_store_commit();
                if constexpr (SPW < 4096) tma_store_wait_read<0>();
              }
              if constexpr (SPW < 4096) {
                __syncwarp();
                stage_rows_tma<2>(st8, lane, h8v);
                stage_rows_tma<2>(st8 + 1024, lane, l8v);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  tma_store_4d(&map_o8, st8, col8, t0w, b, 1);
                  tma_store_4d(&map_o8, st8 + 1024, col8 + 32, t0w, b, 1);
                  tma_store_commit();
                }
              }
            } else {
              uint4 hv[4], lv[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  split_pack2(__uint_as_float(o[8 * j + 2 * e]) * inv, __uint_as_float(o[8 * j + 2 * e + 1]) * inv, hi[e], lo[e]);
                hv[j] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                lv[j] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              }
              DBG(dbg_who, 37);
              const int col = h * HD + hf * 32;
              if (lane == 0) tma_store_wait_read<0>();            // the previous item's stores have read the staging buffer
              __syncwarp();
              stage_rows_tma<4>(my_stage, lane, hv);
              if constexpr (SPW >= 4096) stage_rows_tma<4>(my_stage + 2048, lane, lv);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_4d(&map_o, my_stage, col, t0w, b, 0);
                if constexpr (SPW >= 4096) tma_store_4d(&map_o, my_stage + 2048, col, t0w, b, 1);
                tma_store_commit();
                if constexpr (SPW < 4096) tma_store_wait_read<0>();
              }
              DBG(dbg_who, 38);
              if constexpr (SPW < 4096) {
                __syncwarp();
                stage_rows_tma<4>(my_stage, lane, lv);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  tma_store_4d(&map_o, my_stage, col, t0w, b, 1);
                  tma_store_commit();
                }
              }
              DBG(dbg_who, 39);
            }
          }
          if (p.out_f32) {
            if (lane == 0) tma_store_wait_read<0>();              // the plane stores above read the same staging buffer
            __syncwarp();
            uint8_t* f0 = reinterpret_cast<uint8_t*>(p.out_f32 + row0 * (static_cast<int64_t>(p.H) * HD) + h * HD + hf * 32);
            const int64_t pitch = static_cast<int64_t>(p.H) * HD * 4;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint4 fv[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int k = 16 * half + 4 * j;
                fv[j] = make_uint4(__float_as_uint(__uint_as_float(o[k]) * inv), __float_as_uint(__uint_as_float(o[k + 1]) * inv),
                                   __float_as_uint(__uint_as_float(o[k + 2]) * inv), __float_as_uint(__uint_as_float(o[k + 3]) * inv));
              }
              warp_store_row_runs<4>(my_stage, lane, fv, f0 + half * 64, pitch, n_rows);
            }
          }
          if (p.lse && hf == 0 && t < p.T) p.lse[(static_cast<int64_t>(b) * p.H + h) * p.T + t] = r * 0.69314718055994531f + logf(sum);
        }
        DBG(dbg_who, 35);
      }
      if (p.sat_flag && __any_sync(0xffffffffu, amax > QV_MIX_ACT_MAX) && lane == 0) atomicOr(p.sat_flag, p.sat_bit);
      if (lane == 0) tma_store_wait_read<0>();                    // staging buffers are read before the CTA's shared memory goes away
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
int launch_attn(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mo, const CUtensorMap& mo8,
                const AttnParams& ap, int grid, cudaStream_t st) {
  using C = AttnCfg<NPL>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(qv_attn_fwd_kernel<NPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  });
  QV_REQUIRE(attr_err == cudaSuccess, QV_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  qv_attn_fwd_kernel<NPL><<<grid, AT_THREADS, C::SMEM_BYTES, st>>>(mq, mk, mv, mo, mo8, ap);
  return qv_check_launch("qv_attn_fwd");
}


// ====================================================================================================================
// Fused attention BACKWARD for integer-code operands (the QAT student: Q, K, V = s * codes), one (image, head) per item.
//
// Given dO (bf16 hi/lo planes), O (the forward's output planes), the codes and the forward's logsumexp, recomputes the
// probabilities on the tensor cores and produces dQ, dK, dV without ever writing scores to HBM.  Replaces the autograd of
// F.scaled_dot_product_attention (4 batched GEMMs + softmax-backward passes in the unfused path).  With z = scale * Q K^T,
// P = softmax(z):
//     dP = dO V^T,  delta_i = sum_j P_ij dP_ij = sum_d dO_id O_id,  dz = P o (dP - delta),
//     dQ = scale dz K,  dK = scale dz^T Q,  dV = P^T dO.
// TMEM accumulators are row-per-lane, so the kernel runs two kinds of sub-pass per 128-row tile:
//   pass A (lanes = queries):  S = Q K^T, dP = dO V^T  -> threads: dz (bf16 hi/lo, in place over dP)   -> dQ += dz K
//   pass B (lanes = keys):     S^T = K Q^T, dP^T = V dO^T -> threads: P^T, dz^T (in place)  -> dV += P^T dO, dK += dz^T Q
// (recomputing S / dP transposed costs two small extra MMAs and saves staging dz through shared memory).
// Every sub-pass is cut into CHUNKS of 64 columns (keys in pass A, queries in pass B) with the S / dP accumulators
// double-buffered in TMEM: while the 8 compute warps turn chunk c into dz / P^T, the tensor pipe already runs the first-stage
// products of chunk c+1 and the second-stage products of chunk c-1.  delta comes from dO . O (computed per item from global
// memory while the item's tiles are still in flight), so a chunk needs no row-wide reduction before it can be processed.
// All second-stage products are TS-mode MMAs (A from TMEM); the same [tokens x 64] shared-memory tiles serve as K-major
// operands of the first stage and MN-major operands of the second.
// TMEM: buffer b in {0,1}: S_b [128 b, +64) | R_b [128 b + 64, +64) ;  accumulator set p in {0,1} (sub-pass parity):
// ACC0 [256 + 128 p, +64) (dQ / dV), ACC1 [320 + 128 p, +64) (dK).
// Warps: 0 TMA, 1 MMA, 2-5 CHUNK warps (one per TMEM lane quarter: S / dP chunk -> dz / P^T), 6-9 OUTPUT warps (one per lane
// quarter: drain accumulator set p of sub-pass n -- mask, hi/lo split, column sums, staging, TMA -- while the chunk warps and
// the tensor pipe are already in sub-pass n + 1; the chunk warps never touch an accumulator, so nothing on the MMA thread's
// critical path waits for an output stage any more).
// ====================================================================================================================
// 28 KB tiles: 224 token rows x 64 bf16 (T <= 224; rows >= T zero-filled by TMA).  The second 128-row MMA tile reads 32 rows
// past its end (the next tile's head): those lanes are tokens >= 224 > T, whose results are never stored.
constexpr int BW_TILE_ROWS = 224;
constexpr int BW_TILE_BYTES = BW_TILE_ROWS * HD * 2;
constexpr int BW_STAGE_BYTES = 8 * 2 * 4096;    // per compute warp: two 4 KB staging buffers (y tile in, gradient tile out)
constexpr int BW_SMEM_BYTES = 5 * BW_TILE_BYTES + 4096 /*lse, delta*/ + BW_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int BW_THREADS = 576;                  // warps: 0 TMA, 1 MMA, 2-9 chunk, 10-17 output (two of each per TMEM lane quarter)
constexpr uint32_t BW_ACC_COL = 256;             // accumulator set p (sub-pass parity): ACC0 at 256 + 128 p (dQ / dV), ACC1 64 further (dK)

struct AttnBwdParams {
  int32_t B, T, H;
  int32_t n_keys, m_tiles;
  float scale;
  const float* qscale;      // device scalar s (Q = K = V scale); may be null (= 1)
  const float* lse;         // [B*H*T] from the forward
  const __nv_bfloat16* o_planes;   // [2][B*T][o_ld]: the forward's output O (hi/lo), head h in columns h*64..
  int64_t o_plane_stride, o_ld;
  const __nv_bfloat16* do_planes;  // [2][B*T][do_ld]: dL/dO (also the TMA source of the dO tiles)
  int64_t do_plane_stride, do_ld;
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


__device__ __forceinline__ float bf16lo_f(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi_f(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// One 16-column piece of a chunk for one TMEM lane (row): S / dP in, dz (pass A) or P^T and dz^T (pass B) out, in place, as
// bf16 pairs [hi 8 words | lo 8 words].  nLp / ndp: this row's -lse*log2e and -delta (pass A, packed twice); lse_sa / delta_sa:
// shared-memory addresses of the 16 columns' -lse*log2e and -delta (pass B, broadcast reads).  MASKED: the piece straddles T.
template <bool PASS_A, bool MASKED>
__device__ __forceinline__ void bw_chunk_piece(uint32_t S, uint32_t R, uint64_t c2p, uint64_t nLp, uint64_t ndp,
                                               uint32_t lse_sa, uint32_t delta_sa, int nvalid) {
  uint32_t sv[16], dv[16];
  tmem_ld_32x16(S, sv);
  tmem_ld_32x16(R, dv);
  uint64_t nL[8], nd[8];
  if constexpr (!PASS_A) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(nL[2 * i]), "=l"(nL[2 * i + 1]) : "r"(lse_sa + 16 * i));
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(nd[2 * i]), "=l"(nd[2 * i + 1]) : "r"(delta_sa + 16 * i));
    }
  }
  tmem_ld_wait();
  uint32_t pz[16], pp[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float e0, e1;
    f2_unpack(f2_fma(f2_pack_u(sv[2 * j], sv[2 * j + 1]), c2p, PASS_A ? nLp : nL[j]), e0, e1);
    float p0 = ex2_approx(e0), p1 = ex2_approx(e1);
    if constexpr (MASKED && !PASS_A) {
      p0 = (2 * j < nvalid) ? p0 : 0.f;
      p1 = (2 * j + 1 < nvalid) ? p1 : 0.f;
    }
    const uint64_t pr = f2_pack(p0, p1);
    uint64_t z = f2_mul(pr, f2_add(f2_pack_u(dv[2 * j], dv[2 * j + 1]), PASS_A ? ndp : nd[j]));
    if constexpr (MASKED && PASS_A) {
      float z0, z1;
      f2_unpack(z, z0, z1);
      z = f2_pack((2 * j < nvalid) ? z0 : 0.f, (2 * j + 1 < nvalid) ? z1 : 0.f);
    }
    split_pack2_f2(z, pz[j], pz[8 + j]);
    if constexpr (!PASS_A) split_pack2_f2(pr, pp[j], pp[8 + j]);
  }
  if constexpr (!PASS_A) tmem_st_32x16(S, pp);
  tmem_st_32x16(R, pz);
}

template <bool FUSED>
__global__ void __launch_bounds__(BW_THREADS, 1)
qv_attn_bwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                   const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_o,
                   const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + BW_TILE_BYTES;
  uint8_t* sV = sK + BW_TILE_BYTES;
  uint8_t* sDOh = sV + BW_TILE_BYTES;
  uint8_t* sDOl = sDOh + BW_TILE_BYTES;
  float* lse2_s = reinterpret_cast<float*>(sDOl + BW_TILE_BYTES);   // [256] -lse * log2(e)   (negated: the packed fma adds it)
  float* delta_all = lse2_s + 256;                                   // [2 item parities][256] -(dO . O) / s  (negated likewise)
  uint8_t* stage_s = reinterpret_cast<uint8_t*>(lse2_s + 1024);     // [8 output warps][2][4 KB], 1024-byte aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_s + BW_STAGE_BYTES);
  uint64_t* ld_full = bars + 0;
  uint64_t* ld_empty = bars + 1;
  uint64_t* mma1_done = bars + 2;    // [2] first-stage products of chunk buffer b landed in TMEM
  uint64_t* cmp_done = bars + 4;     // [2] chunk warps have turned buffer b into dz / P^T
  uint64_t* acc_done = bars + 6;     // [2] accumulator set p complete
  uint64_t* epi_done = bars + 8;     // [2] accumulator set p drained by the output warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  uint64_t* y_bar = bars + 13;       // [8 output warps][2]: y tile landed in staging buffer k

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int D = p.H * HD;
  const int num_items = p.B * p.H;
  const int n_keys = p.n_keys;
  const int mt = p.m_tiles;
  const int nch = (n_keys + 63) >> 6;          // 64-column chunks per sub-pass
  const int nsub = 2 * mt;                     // sub-passes per item: pass A tiles, then pass B tiles

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_qkv);
    prefetch_tensormap(&map_do);
    mbar_init(ld_full, 1);
    mbar_init(ld_empty, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&mma1_done[b], 1);
      mbar_init(&cmp_done[b], 256);
      mbar_init(&acc_done[b], 1);
      mbar_init(&epi_done[b], 256);
    }
    if constexpr (FUSED) {
      prefetch_tensormap(&map_y);
      for (int w = 0; w < 16; ++w) mbar_init(&y_bar[w], 1);      // [8 output warps][2]
    }
    prefetch_tensormap(&map_o);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // dst[r] = -(dO_r . O_r) / s for token rows of [row_lo, row_lo + 128) of item (b, h) from global memory: 8 lanes per row, 8
  // columns each, 16 independent 16-byte loads in flight per lane; `w4` = this warp's index 0..3 inside its group of four; the
  // group covers the 64-row halves [half_lo, half_hi) of the range (chunk warps: one half per group of four; output warps: both).
  auto delta_rows = [&](float* dst, int b, int h, int row_lo, int w4, float inv_s, int half_lo, int half_hi) {
#pragma unroll 1
    for (int half = half_lo; half < half_hi; ++half) {
      uint4 oh[4], ol[4], dh[4], dl[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = row_lo + (half * 4 + u) * 16 + w4 * 4 + (lane >> 3);
        if (r < p.T) {
          const int64_t grow = static_cast<int64_t>(b) * p.T + r;
          const int col = h * HD + (lane & 7) * 8;
          oh[u] = __ldg(reinterpret_cast<const uint4*>(p.o_planes + grow * p.o_ld + col));
          ol[u] = __ldg(reinterpret_cast<const uint4*>(p.o_planes + p.o_plane_stride + grow * p.o_ld + col));
          dh[u] = __ldg(reinterpret_cast<const uint4*>(p.do_planes + grow * p.do_ld + col));
          dl[u] = __ldg(reinterpret_cast<const uint4*>(p.do_planes + p.do_plane_stride + grow * p.do_ld + col));
        } else {
          oh[u] = ol[u] = dh[u] = dl[u] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = row_lo + (half * 4 + u) * 16 + w4 * 4 + (lane >> 3);
        const uint32_t ohw[4] = {oh[u].x, oh[u].y, oh[u].z, oh[u].w}, olw[4] = {ol[u].x, ol[u].y, ol[u].z, ol[u].w};
        const uint32_t dhw[4] = {dh[u].x, dh[u].y, dh[u].z, dh[u].w}, dlw[4] = {dl[u].x, dl[u].y, dl[u].z, dl[u].w};
        float dot = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          dot = fmaf(bf16lo_f(ohw[e]) + bf16lo_f(olw[e]), bf16lo_f(dhw[e]) + bf16lo_f(dlw[e]), dot);
          dot = fmaf(bf16hi_f(ohw[e]) + bf16hi_f(olw[e]), bf16hi_f(dhw[e]) + bf16hi_f(dlw[e]), dot);
        }
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        dot += __shfl_xor_sync(0xffffffffu, dot, 4);
        if ((lane & 7) == 0) dst[r] = -dot * inv_s;
      }
    }
  };

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
        // The tiles are single-buffered (140 KB of the 227 KB), so the next item's loads start only when this item's last MMA
        // has read them -- and every CTA reaches that point together (20 MB in one burst).  Pull the next item's tiles into L2
        // now, while this item computes: the real loads then hit L2.
        const int nitem = item + static_cast<int>(gridDim.x);
        if (nitem < num_items) {
          const int nb = nitem / p.H, nh = nitem % p.H;
          tma_prefetch_l2_4d(&map_qkv, nh * HD, 0, nb, 0);
          tma_prefetch_l2_4d(&map_qkv, D + nh * HD, 0, nb, 0);
          tma_prefetch_l2_4d(&map_qkv, 2 * D + nh * HD, 0, nb, 0);
          tma_prefetch_l2_4d(&map_do, nh * HD, 0, nb, 0);
          tma_prefetch_l2_4d(&map_do, nh * HD, 0, nb, 1);
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // The WHOLE warp walks the schedule (waits included) and one elected lane issues: with warp-uniform control flow ptxas
    // keeps descriptors / TMEM addresses in uniform registers instead of a per-instruction R2UR + ELECT waterfall loop.
    {
      const uint32_t idesc_ts = umma_idesc_bf16(128, HD, false, true);
      // Descriptors are built ONCE per tile and mode; every MMA then only adds its byte offset (>> 4) to the start-address
      // field (tiles sit below 256 KB, so the 14-bit field never carries) -- the single issuing thread is on the critical
      // path of every chunk-step, and rebuilding a descriptor per instruction cost ~50 ns each.
      const uint64_t kQ = umma_smem_desc(smem_u32(sQ), 16u, 1024u), kK = umma_smem_desc(smem_u32(sK), 16u, 1024u),
                     kV = umma_smem_desc(smem_u32(sV), 16u, 1024u), kDh = umma_smem_desc(smem_u32(sDOh), 16u, 1024u),
                     kDl = umma_smem_desc(smem_u32(sDOl), 16u, 1024u);                       // K-major views
      const uint64_t mQ = umma_smem_desc(smem_u32(sQ), 8192u, 1024u), mK = umma_smem_desc(smem_u32(sK), 8192u, 1024u),
                     mDh = umma_smem_desc(smem_u32(sDOh), 8192u, 1024u), mDl = umma_smem_desc(smem_u32(sDOl), 8192u, 1024u);  // MN-major
      const int w_tail = n_keys - 64 * (nch - 1);
      const uint32_t idesc_full = umma_idesc_bf16(128, 64, false, false);
      const uint32_t idesc_tail = umma_idesc_bf16(128, w_tail, false, false);
      auto mma1 = [&](int sub, int c, uint32_t buf) {
        const uint32_t idesc = (c == nch - 1) ? idesc_tail : idesc_full;
        const uint32_t S = tmem_base + buf * 128u, R = S + 64u;
        const uint64_t rows = static_cast<uint64_t>(c) * 512u;           // 64 token rows x 128 B, in 16-byte units
        // S and dP are independent accumulators: their MMAs are issued alternately so that consecutive instructions never
        // accumulate into the same TMEM tile (back-to-back dependent accumulation at N = 64 runs at half rate)
        const bool pa = sub < mt;                                        // pass A: lanes = queries; pass B: lanes = keys
        const uint64_t t = static_cast<uint64_t>(pa ? sub : sub - mt) * 1024u;      // 128-row tile: 16384 B
        const uint64_t sa = (pa ? kQ : kK) + t, sb = (pa ? kK : kQ) + rows;         // S = Q K^T        | S^T = K Q^T
        const uint64_t r1a = (pa ? kDh : kV) + t, r1b = (pa ? kV : kDh) + rows;     // dP = dO V^T (hi) | dP^T = V dO^T (hi)
        const uint64_t r2a = (pa ? kDl : kV) + t, r2b = (pa ? kV : kDl) + rows;     //            (lo) |              (lo)
#pragma unroll
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) {
            umma_bf16(R, r1a + 2 * k, r1b + 2 * k, idesc, k > 0 ? 1u : 0u);
            umma_bf16(S, sa + 2 * k, sb + 2 * k, idesc, k > 0 ? 1u : 0u);
            umma_bf16(R, r2a + 2 * k, r2b + 2 * k, idesc, 1u);
          }
          umma_commit(&mma1_done[buf]);
        }
        __syncwarp();
      };
      auto mma2 = [&](int sub, int c, uint32_t buf, uint32_t aset) {
        const int ks = (c == nch - 1) ? (w_tail >> 4) : 4;
        const uint32_t S = tmem_base + buf * 128u, R = S + 64u;
        const uint64_t rows = static_cast<uint64_t>(c) * 512u;
        const uint32_t acc0 = tmem_base + BW_ACC_COL + aset * 128u, acc1 = acc0 + 64u;
        const bool pa = sub < mt;
        const uint64_t bK = mK + rows, bQ = mQ + rows, bDh = mDh + rows, bDl = mDl + rows;   // operands formed warp-uniformly
        const uint32_t first0 = c > 0 ? 1u : 0u;
        if (elect_one()) {
          if (pa) {                                                      // dQ += dz K (hi, lo)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              if (kk < ks) {
                const uint32_t off = static_cast<uint32_t>(kk * 16);     // 16 keys = 16 columns: [hi pairs 8 | lo pairs 8]
                umma_bf16_ts(acc0, R + off, bK + 128 * kk, idesc_ts, kk > 0 ? 1u : first0);
                umma_bf16_ts(acc0, R + off + 8, bK + 128 * kk, idesc_ts, 1u);
              }
            }
          } else {                                                       // dV and dK chains interleaved (independent tiles)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              if (kk < ks) {
                const uint32_t off = static_cast<uint32_t>(kk * 16);
                const uint32_t first = kk > 0 ? 1u : first0;
                umma_bf16_ts(acc0, S + off, bDh + 128 * kk, idesc_ts, first);            // dV += P^T dO : (hi,hi)
                umma_bf16_ts(acc1, R + off, bQ + 128 * kk, idesc_ts, first);             // dK += dz^T Q : hi
                umma_bf16_ts(acc0, S + off, bDl + 128 * kk, idesc_ts, 1u);               //               (hi,lo)
                umma_bf16_ts(acc1, R + off + 8, bQ + 128 * kk, idesc_ts, 1u);            //                lo
                umma_bf16_ts(acc0, S + off + 8, bDh + 128 * kk, idesc_ts, 1u);           //               (lo,hi)
              }
            }
          }
        }
      };
      uint32_t st = 0;          // running chunk-step counter: buffer = st & 1, barrier phase = (st >> 1) & 1
      uint32_t nsp = 0;         // running sub-pass counter (phase of acc_done / epi_done)
      int local = 0;
#ifdef QV_ATTN_DEBUG
      int dbg_n = 0;
#endif
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local) {
        DBG(0, 1);
        mbar_wait(ld_full, static_cast<uint32_t>(local & 1));
        tc_fence_after();
        DBG(0, 2);
        mma1(0, 0, st & 1);
        for (int sub = 0; sub < nsub; ++sub) {
          for (int c = 0; c < nch; ++c, ++st) {
            // first stage of the NEXT chunk-step goes out before this step's second stage waits for the compute warps
            if (c + 1 < nch) mma1(sub, c + 1, (st + 1) & 1);
            else if (sub + 1 < nsub) mma1(sub + 1, 0, (st + 1) & 1);
            DBG(0, 3);
            mbar_wait(&cmp_done[st & 1], (st >> 1) & 1);
            DBG(0, 4);
            // accumulator set nsp & 1 was last used by sub-pass nsp - 2: the output warps must have drained it
            if (c == 0) mbar_wait(&epi_done[nsp & 1], ((nsp >> 1) & 1) ^ 1);
            tc_fence_after();
            DBG(0, 5);
            mma2(sub, c, st & 1, nsp & 1);
            __syncwarp();
            DBG(0, 6);
            if (c == nch - 1) {
              if (elect_one()) umma_commit(&acc_done[nsp & 1]);
              __syncwarp();
              ++nsp;
            }
          }
        }
        if (elect_one()) umma_commit(ld_empty);                          // every MMA that reads this item's tiles has completed
        __syncwarp();
      }
    }
  } else if (warp < 10) {
    // =============================== chunk warps (TWO per TMEM lane quarter) ===============================
    // Warps 2-5 convert columns [0, 32) of every 64-column chunk, warps 6-9 columns [32, 64): each scheduler partition
    // then holds two chunk warps (one alone sat at IPC ~0.25, latency-bound -- the chunk math was the critical path of an
    // item: 28 us against 9.6 us of MMA work).  A 32-column half goes through the registers as two 16-column pieces
    // (tcgen05.ld / st .x16) with packed-pair fp32 math; validity masks only in a piece that straddles T.
    const int q = warp & 3;
    const int hf = (warp - 2) >> 2;              // which 32-column half of a chunk
    const int row = q * 32 + lane;               // row inside the 128-row tile == TMEM lane
    const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(hf * 32);
    const float s = p.qscale ? __ldg(p.qscale) : 1.0f;
    const float inv_s = 1.0f / s;
    const float c2 = p.scale * s * s * 1.4426950408889634f;
    const uint64_t c2p = f2_pack(c2, c2);
    const int ctid = threadIdx.x - 64;           // 0..255
    uint32_t st = 0;
#ifdef QV_ATTN_DEBUG
    int dbg_n = (threadIdx.x == 64) ? 0 : 8192;
#endif
    int local = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local) {
      const int b = item / p.H, h = item % p.H;
      const float* lse_bh = p.lse + (static_cast<int64_t>(b) * p.H + h) * p.T;
      float* delta_s = delta_all + (local & 1) * 256;
      const uint32_t lse_s32 = smem_u32(lse2_s), delta_s32 = smem_u32(delta_s);
      DBG(1, 10);
      asm volatile("bar.sync 9, 256;" ::: "memory");            // previous item's pass B has finished reading lse2_s
      lse2_s[ctid] = (ctid < p.T) ? -__ldg(lse_bh + ctid) * 1.4426950408889634f : 0.0f;
      delta_rows(delta_s, b, h, hf * 128, q, inv_s, 0, 2);       // all 256 rows, 32 per warp, from global memory while the tiles land
      asm volatile("bar.sync 9, 256;" ::: "memory");
      DBG(1, 11);
      for (int sub = 0; sub < nsub; ++sub) {
        const bool pass_a = sub < mt;
        const int tile = pass_a ? sub : sub - mt;
        const int tok = tile * 128 + row;                         // query (pass A) / key (pass B) of this lane
        const float nLi = lse2_s[tok & 255], ndi = delta_s[tok & 255];
        const uint64_t nLp = f2_pack(nLi, nLi), ndp = f2_pack(ndi, ndi);
        for (int c = 0; c < nch; ++c, ++st) {
          const uint32_t buf = st & 1;
          DBG(1, 12);
          mbar_wait(&mma1_done[buf], (st >> 1) & 1);
          tc_fence_after();
          DBG(1, 13);
#pragma unroll 1
          for (int u = 0; u < 2; ++u) {
            const int col0 = 64 * c + 32 * hf + 16 * u;           // first column (key in pass A, query in pass B) of this piece
            if (col0 >= n_keys) break;
            const uint32_t S = t0 + buf * 128u + static_cast<uint32_t>(u * 16), R = S + 64u;
            const int nvalid = p.T - col0;
            const uint32_t lse_sa = lse_s32 + static_cast<uint32_t>(col0) * 4u, delta_sa = delta_s32 + static_cast<uint32_t>(col0) * 4u;
            if (pass_a) {
              if (nvalid >= 16) bw_chunk_piece<true, false>(S, R, c2p, nLp, ndp, lse_sa, delta_sa, nvalid);
              else bw_chunk_piece<true, true>(S, R, c2p, nLp, ndp, lse_sa, delta_sa, nvalid);
            } else {
              if (nvalid >= 16) bw_chunk_piece<false, false>(S, R, c2p, nLp, ndp, lse_sa, delta_sa, nvalid);
              else bw_chunk_piece<false, true>(S, R, c2p, nLp, ndp, lse_sa, delta_sa, nvalid);
            }
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&cmp_done[buf]);
          DBG(1, 14);
        }
      }
    }
  } else {
    // =============================== output warps (one per TMEM lane quarter) ===============================
    // Drain accumulator set p of sub-pass n while the chunk warps and the tensor pipe are already in sub-pass n + 1.
    // Per warp: 32 rows x 64 columns of dQ (pass A) or dV and dK (pass B), as 32 x 32 tiles through four 4 KB staging buffers.
    const int q = warp & 3;
    const int wo = warp - 10;                    // 0..7
    const int ch = wo >> 2;                      // which 32-column half of every 64-column accumulator this warp drains
    const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(ch * 32);
    const float s = p.qscale ? __ldg(p.qscale) : 1.0f;
    const float gscale = p.scale * s * s;        // dQ, dK factor (see header comment)
    QvQParams yq;
    if constexpr (FUSED) yq = qv_load_qparams(p.y_scale, p.y_zp, p.qmin, p.qmax);
    const int D3 = 3 * D;
    uint8_t* my_stage = stage_s + wo * 8192;
    uint64_t* my_ybar = y_bar + wo * 2;
    uint32_t yph = 0;                            // bit k: phase of my_ybar[k]
    auto request_y = [&](int k, int b, int tok0, int col) {       // lane 0, after tma_store_wait_read<0>()
      mbar_expect_tx(&my_ybar[k], 4096);
      tma_load_3d(my_stage + k * 4096, &map_y, &my_ybar[k], col, tok0, b);
    };
    // Output of one 32-row x 32-column block, staged through shared memory so that global traffic is whole 128-byte lines moved
    // by TMA: the y tile (FUSED: the qkv Linear's raw output under these gradients) comes IN through buffer k, the lanes take
    // their rows to registers, and the finished tile (fp32, or bf16 hi/lo planes) leaves through the same buffer.  Rows >= T
    // are zero-filled on load and clipped on store by the per-image tensor maps.
    auto emit = [&](const uint32_t (&o)[32], int k, float mult, int b, int tok0, int col, int slab) {
      uint8_t* buf = my_stage + k * 4096;
      const bool ok = tok0 + lane < p.T;
      if constexpr (FUSED) {
        mbar_wait(&my_ybar[k], (yph >> k) & 1u);
        yph ^= 1u << k;
        float4 yv[8];
        const uint32_t srow = smem_u32(buf) + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t addr = srow + (static_cast<uint32_t>(j ^ (lane & 7)) << 4);
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(yv[j].x), "=f"(yv[j].y), "=f"(yv[j].z), "=f"(yv[j].w)
                       : "r"(addr) : "memory");
        }
        float gq[32];
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 ws = __ldg(reinterpret_cast<const float4*>(p.w_scale + col) + j);
          const float yy[4] = {yv[j].x, yv[j].y, yv[j].z, yv[j].w};
          const float wv[4] = {ws.x, ws.y, ws.z, ws.w};
          float a[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float r = __fadd_rn(rintf(__fmul_rn(yy[e], yq.inv)), yq.zp);
            const bool in = (yq.qmin <= r) && (r <= yq.qmax);
            const float f = (ok && in) ? __uint_as_float(o[4 * j + e]) * mult : 0.f;
            gq[4 * j + e] = f;
            a[e] = f * wv[e];
          }
          split_pack2(a[0], a[1], hi[2 * j], lo[2 * j]);
          split_pack2(a[2], a[3], hi[2 * j + 1], lo[2 * j + 1]);
        }
        __syncwarp();                                             // every lane holds its y row: the buffer may be overwritten
        const uint32_t orow = smem_u32(buf) + lane * 64;          // planes: 32 rows x 64 B (hi) | + 2 KB (lo), 64B swizzle
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t sw = (static_cast<uint32_t>(j ^ ((lane >> 1) & 3)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(orow + sw), "r"(hi[4 * j]), "r"(hi[4 * j + 1]),
                       "r"(hi[4 * j + 2]), "r"(hi[4 * j + 3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(orow + 2048 + sw), "r"(lo[4 * j]), "r"(lo[4 * j + 1]),
                       "r"(lo[4 * j + 2]), "r"(lo[4 * j + 3]) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&map_o, buf, col, tok0, b, 0);
          tma_store_4d(&map_o, buf + 2048, col, tok0, b, 1);
          tma_store_commit();
        }
        const float cs = qv_warp_colsum32(gq, lane);
        p.colsum[static_cast<int64_t>(slab) * D3 + col + lane] = cs;
      } else {
        const uint32_t srow = smem_u32(buf) + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t addr = srow + (static_cast<uint32_t>(j ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(__uint_as_float(o[4 * j]) * mult),
                       "f"(__uint_as_float(o[4 * j + 1]) * mult), "f"(__uint_as_float(o[4 * j + 2]) * mult),
                       "f"(__uint_as_float(o[4 * j + 3]) * mult) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&map_o, buf, col, tok0, b);
          tma_store_commit();
        }
      }
    };
    uint32_t nsp = 0;
#ifdef QV_ATTN_DEBUG
    int dbg_n = (threadIdx.x == 320) ? 0 : 8192;
#endif
    int local = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++local) {
      const int b = item / p.H, h = item % p.H;
      for (int sub = 0; sub < nsub; ++sub, ++nsp) {
        const bool pass_a = sub < mt;
        const int tile = pass_a ? sub : sub - mt;
        const int tok0 = tile * 128 + q * 32;                     // first token of this warp's 32-row slab
        const int slab = (b * mt + tile) * 4 + q;
        const int col_h = h * HD + ch * 32;                       // this warp's 32 columns of the head
        const uint32_t aset = nsp & 1;
        if (lane == 0) {
          tma_store_wait_read<0>();                               // the previous sub-pass's tiles have left the staging buffers
          if constexpr (FUSED) {                                  // y tiles under this sub-pass's outputs
            request_y(0, b, tok0, (pass_a ? 0 : 2 * D) + col_h);
            if (!pass_a) request_y(1, b, tok0, D + col_h);
          }
        }
        __syncwarp();
        mbar_wait(&acc_done[aset], (nsp >> 1) & 1);
        tc_fence_after();
        DBG(2, 15);
        const uint32_t acc0 = t0 + BW_ACC_COL + aset * 128u;
        uint32_t o[32];
        tmem_ld_32x32(acc0, o);
        tmem_ld_wait();
        if (pass_a) {
          tc_fence_before();
          mbar_arrive(&epi_done[aset]);
        }
        emit(o, 0, pass_a ? gscale : 1.0f, b, tok0, (pass_a ? 0 : 2 * D) + col_h, slab);               // dQ | dV
        if (!pass_a) {
          tmem_ld_32x32(acc0 + 64u, o);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&epi_done[aset]);
          emit(o, 1, gscale, b, tok0, D + col_h, slab);                                                 // dK
        }
        DBG(2, 16);
      }
    }
    if (lane == 0) tma_store_wait_read<0>();
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
                           int32_t out_fmt, int32_t* sat_flag, int32_t sat_bit, void* stream) {
  QV_REQUIRE(qkv_planes && (out_planes || out_f32) && B > 0 && T > 0 && H > 0, QV_ERR_INVALID, "bad attn_fwd arguments");
  QV_REQUIRE(out_fmt == 0 || (out_fmt == 1 && out_planes && out_ld % 64 == 0), QV_ERR_INVALID,
             "out_fmt must be 0 (bf16 hi/lo) or 1 (mixed fp16 + fp8 planes, row pitch a multiple of 64)");
  QV_REQUIRE(n_planes == 1 || n_planes == 2, QV_ERR_INVALID, "n_planes must be 1 (integer codes) or 2 (fp32 hi/lo)");
  QV_REQUIRE(scale > 0.f, QV_ERR_INVALID, "the softmax scale must be positive (the row reference exponent is a maximum)");
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
  ap.out_fmt = out_fmt;
  ap.sat_flag = out_fmt == 1 ? sat_flag : nullptr;
  ap.sat_bit = sat_bit;
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
  // output planes as [plane][image][token][column]: a store box that runs past an image's last token is clipped by the descriptor
  // (the second query tile's padding rows), 32-column boxes for the 2-byte planes, 16-"element" boxes for the mixed format's byte runs
  CUtensorMap mo = mq, mo8 = mq;
  if (out_planes) {
    rc = make_out_planes_map(&mo, out_planes, static_cast<int64_t>(H) * HD, T, out_ld, B, static_cast<int64_t>(T) * out_ld, out_plane_stride, 32);
    if (rc) return rc;
    rc = make_out_planes_map(&mo8, out_planes, static_cast<int64_t>(H) * HD, T, out_ld, B, static_cast<int64_t>(T) * out_ld, out_plane_stride, 16);
    if (rc) return rc;
  }
  const int items = B * H;
  const int sms = qv_num_sms();
  const int grid = items < sms ? items : sms;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_planes == 2) return launch_attn<2>(mq, mk, mv, mo, mo8, ap, grid, st);
  return launch_attn<1>(mq, mk, mv, mo, mo8, ap, grid, st);
}

namespace {
int attn_bwd_impl(const uint16_t* qkv_codes, int64_t ld, const float* qscale, const uint16_t* o_planes, int64_t o_plane_stride,
                  int64_t o_ld, const uint16_t* do_planes, int64_t do_plane_stride, int64_t do_ld, const float* lse, int32_t B,
                  int32_t T, int32_t H, float scale, float* g_qkv, const AttnBwdParams* fused, void* stream) {
  QV_REQUIRE(qkv_codes && o_planes && do_planes && lse && (g_qkv || fused) && B > 0 && T > 0 && H > 0, QV_ERR_INVALID,
             "bad attn_bwd arguments");
  QV_REQUIRE(T <= 224, QV_ERR_UNSUPPORTED, "fused attention holds all keys in one tile: T <= 224 (got %d)", T);
  QV_REQUIRE(ld >= 3LL * H * HD && do_ld >= static_cast<int64_t>(H) * HD && o_ld >= static_cast<int64_t>(H) * HD, QV_ERR_INVALID,
             "row pitches too small");
  QV_REQUIRE(qv_aligned16(o_planes) && o_ld % 8 == 0 && o_plane_stride % 8 == 0 && do_ld % 8 == 0 && do_plane_stride % 8 == 0,
             QV_ERR_INVALID, "O / dO planes must be 16-byte aligned with pitches that are multiples of 8 bf16");
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
  ap.o_planes = reinterpret_cast<const __nv_bfloat16*>(o_planes); ap.o_plane_stride = o_plane_stride; ap.o_ld = o_ld;
  ap.do_planes = reinterpret_cast<const __nv_bfloat16*>(do_planes); ap.do_plane_stride = do_plane_stride; ap.do_ld = do_ld;
  ap.g_qkv = g_qkv;
  qv_operand op;
  memset(&op, 0, sizeof(op));
  op.ptr = qkv_codes; op.ld = ld; op.plane_stride = 0; op.rows = T; op.cols = 3LL * H * HD;
  op.nb = B; op.batch_stride = static_cast<int64_t>(T) * ld;
  CUtensorMap mq, md;
  int rc = make_map(&mq, op, 1, BW_TILE_ROWS);
  if (rc) return rc;
  op.ptr = do_planes; op.ld = do_ld; op.plane_stride = do_plane_stride; op.cols = static_cast<int64_t>(H) * HD;
  op.batch_stride = static_cast<int64_t>(T) * do_ld;
  rc = make_map(&md, op, 2, BW_TILE_ROWS);
  if (rc) return rc;
  // staging maps (per-image: rows >= T are zero-filled on load / clipped on store)
  CUtensorMap my, mo;
  const int64_t D3 = 3LL * H * HD;
  if (fused) {
    rc = make_out_map(&my, const_cast<float*>(fused->y_raw), D3, T, D3, B, static_cast<int64_t>(T) * D3);
    if (rc) return rc;
    rc = make_out_planes_map(&mo, fused->gp, D3, T, D3, B, static_cast<int64_t>(T) * D3, fused->gp_plane_stride, 32);
  } else {
    rc = make_out_map(&mo, g_qkv, D3, T, D3, B, static_cast<int64_t>(T) * D3);
    my = mo;
  }
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
  if (fused) qv_attn_bwd_kernel<true><<<grid, BW_THREADS, BW_SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(mq, md, my, mo, ap);
  else qv_attn_bwd_kernel<false><<<grid, BW_THREADS, BW_SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(mq, md, my, mo, ap);
  return qv_check_launch("qv_attn_bwd");
}
}  // namespace

extern "C" int qv_attn_bwd(const uint16_t* qkv_codes, int64_t ld, const float* qscale, const uint16_t* o_planes,
                           int64_t o_plane_stride, int64_t o_ld, const uint16_t* do_planes, int64_t do_plane_stride,
                           int64_t do_ld, const float* lse, int32_t B, int32_t T, int32_t H, float scale, float* g_qkv,
                           void* stream) {
  QV_REQUIRE(g_qkv && qv_aligned16(g_qkv), QV_ERR_INVALID, "g_qkv must be a 16-byte aligned device pointer");
  return attn_bwd_impl(qkv_codes, ld, qscale, o_planes, o_plane_stride, o_ld, do_planes, do_plane_stride, do_ld, lse, B, T, H,
                       scale, g_qkv, nullptr, stream);
}

extern "C" int qv_attn_bwd_gp(const uint16_t* qkv_codes, int64_t ld, const float* qscale, const uint16_t* o_planes,
                              int64_t o_plane_stride, int64_t o_ld, const uint16_t* do_planes, int64_t do_plane_stride,
                              int64_t do_ld, const float* lse, int32_t B, int32_t T, int32_t H, float scale, const float* y_raw,
                              const float* y_scale, const int32_t* y_zp, int32_t qmin, int32_t qmax, const float* w_scale,
                              uint16_t* gp_planes, int64_t gp_plane_stride, float* colsum, void* stream) {
  QV_REQUIRE(y_raw && y_scale && y_zp && w_scale && gp_planes && colsum, QV_ERR_INVALID, "bad attn_bwd_gp arguments");
  QV_REQUIRE(qv_aligned16(y_raw) && qv_aligned16(w_scale) && qv_aligned16(gp_planes) && gp_plane_stride % 8 == 0, QV_ERR_INVALID,
             "y_raw / w_scale / gp_planes must be 16-byte aligned (plane stride a multiple of 8 bf16)");
  AttnBwdParams f;
  memset(&f, 0, sizeof(f));
  f.y_raw = y_raw; f.y_scale = y_scale; f.y_zp = y_zp; f.qmin = qmin; f.qmax = qmax; f.w_scale = w_scale;
  f.gp = reinterpret_cast<__nv_bfloat16*>(gp_planes); f.gp_plane_stride = gp_plane_stride; f.colsum = colsum;
  return attn_bwd_impl(qkv_codes, ld, qscale, o_planes, o_plane_stride, o_ld, do_planes, do_plane_stride, do_ld, lse, B, T, H,
                       scale, nullptr, &f, stream);
}

#ifdef QV_ATTN_DEBUG
extern "C" int qv_debug_read(unsigned long long* host_out) {   // host_out: [3][8192]
  return cudaMemcpyFromSymbol(host_out, qv_dbg_buf, sizeof(unsigned long long) * 3 * 8192) == cudaSuccess ? 0 : -1;
}
#endif
