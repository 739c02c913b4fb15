// qv_tma.cuh -- host-side TMA descriptor helpers shared by the GEMM and attention kernels.
#pragma once
#include <cuda.h>
#include <mutex>

#include "qv_common.cuh"

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
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
inline int make_map(CUtensorMap* m, const qv_operand& op, int planes, int box_rows) {
  EncodeTiledFn enc = get_encode();
  QV_REQUIRE(enc != nullptr, QV_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  QV_REQUIRE(op.ptr && qv_aligned16(op.ptr), QV_ERR_INVALID, "gemm operand base must be a 16-byte aligned device pointer");
  QV_REQUIRE(op.rows > 0 && op.cols > 0, QV_ERR_INVALID, "gemm operand extent must be positive");
  QV_REQUIRE(op.ld % 8 == 0 && op.ld >= op.cols, QV_ERR_INVALID,
             "gemm operand row pitch must be >= cols and a multiple of 8 bf16 (got %lld)", (long long)op.ld);
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
  QV_REQUIRE(r == CUDA_SUCCESS, QV_ERR_CUDA, "cuTensorMapEncodeTiled(operand) failed (%d)", (int)r);
  return 0;
}


// fp32 output [nb][rows][ld] as a 3-D tensor (cols, rows, nb); store box = (32 cols = 128 B, 32 rows, 1), 128B swizzle.
// box_cols = 16: 64-byte rows, 64B swizzle
inline int make_out_map(CUtensorMap* m, float* ptr, int64_t cols, int64_t rows, int64_t ld, int64_t nb, int64_t bstride,
                        int box_cols = 32) {
  EncodeTiledFn enc = get_encode();
  QV_REQUIRE(enc != nullptr, QV_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  QV_REQUIRE(ptr && qv_aligned16(ptr), QV_ERR_INVALID, "gemm output base must be a 16-byte aligned device pointer");
  QV_REQUIRE(rows > 0 && cols > 0 && ld >= cols && ld % 4 == 0, QV_ERR_INVALID,
             "gemm output row pitch must be >= cols and a multiple of 4 floats (got %lld)", (long long)ld);
  if (nb < 1) nb = 1;
  if (bstride <= 0) bstride = rows * ld;
  QV_REQUIRE(bstride % 4 == 0, QV_ERR_INVALID, "gemm output batch stride must be a multiple of 4 floats");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(nb)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 4, static_cast<cuuint64_t>(bstride) * 4};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  QV_REQUIRE(r == CUDA_SUCCESS, QV_ERR_CUDA, "cuTensorMapEncodeTiled(output) failed (%d)", (int)r);
  return 0;
}

// bf16 hi/lo plane output [2][nb][rows][ld] as a 4-D tensor (cols, rows, nb, plane); store box = (64 cols = 128 B, 32 rows).
// box_cols = 32: 64-byte rows, 64B swizzle (the plane-output and gradient-planes epilogues work in 32-column units);
// box_cols = 16: 32-byte rows, no swizzle (the hi8 / lo8 byte runs of the mixed format's region 1).
inline int make_out_planes_map(CUtensorMap* m, void* ptr, int64_t cols, int64_t rows, int64_t ld, int64_t nb, int64_t bstride,
                        int64_t pstride, int box_cols = 64) {
  EncodeTiledFn enc = get_encode();
  QV_REQUIRE(enc != nullptr, QV_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  QV_REQUIRE(ptr && qv_aligned16(ptr), QV_ERR_INVALID, "gemm plane output base must be a 16-byte aligned device pointer");
  QV_REQUIRE(rows > 0 && cols > 0 && ld >= cols && ld % 8 == 0, QV_ERR_INVALID,
             "gemm plane output row pitch must be >= cols and a multiple of 8 bf16 (got %lld)", (long long)ld);
  if (nb < 1) nb = 1;
  if (bstride <= 0) bstride = rows * ld;
  if (pstride <= 0) pstride = bstride * nb;
  QV_REQUIRE(bstride % 8 == 0 && pstride % 8 == 0, QV_ERR_INVALID, "plane output batch / plane strides must be multiples of 8 bf16");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(nb), 2};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(bstride) * 2,
                           static_cast<cuuint64_t>(pstride) * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_cols), 32, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_NONE : box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  QV_REQUIRE(r == CUDA_SUCCESS, QV_ERR_CUDA, "cuTensorMapEncodeTiled(plane output) failed (%d)", (int)r);
  return 0;
}


}  // namespace
