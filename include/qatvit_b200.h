/*
 * qatvit_b200.h -- C ABI of libqatvit_b200.so (hand-written sm_100a CUDA for the QAT-distillation
 * hot path of bdina9/qat-vit).  Plain pointers and sizes only; no torch types.
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch caching allocator), fp32 data
 *     contiguous and 16-byte aligned unless noted; `stream` is a cudaStream_t passed as void*;
 *   - the library never allocates device memory, never synchronises, never changes the device;
 *   - return value 0 = enqueued OK, negative = error (message: qv_last_error(), thread-local);
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Each entry point cites the reference interface it replaces:
 *   ref/...   = /root/reference (bdina9/qat-vit);  torch/... = torch 2.11.0 site-packages.
 */
#ifndef QATVIT_B200_H_
#define QATVIT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QV_OK 0
#define QV_ERR_INVALID (-1)
#define QV_ERR_CUDA (-2)
#define QV_ERR_UNSUPPORTED (-3)

/* ---- library ------------------------------------------------------------------------------- */
int qv_version(void);                 /* ABI version, currently 1 */
const char* qv_last_error(void);      /* thread-local message of the last failing call */
int qv_device_sm_count(void);         /* SMs of the current device, <0 on error (no device) */
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
int64_t qv_launch_count(void);

/* ---- observer + fake-quant (replaces torch.fused_moving_avg_obs_fake_quant,
 *      torch/ao/quantization/fake_quantize.py:423-438, called by the hooks prepare_qat installs:
 *      ref/src/training/qat_trainer.py:306-307) ------------------------------------------------ */

/* Reset a 2-word ordered-min/max accumulator ([0]=min, [1]=max encodings). */
int qv_minmax_reset(uint32_t* acc, int count, void* stream);
/* acc <- min/max(acc, x[0..n)).  Per-tensor activation observer, phase 1 (aten::aminmax). */
int qv_minmax_accumulate(const float* x, int64_t n, uint32_t* acc, void* stream);

/* Per-tensor observer phase 2: EMA of running min/max + ChooseQuantizationParams, on device, gated
 * by the module's int64 enable flags (read on device, no host sync).
 * State buffers are the module's own: min_val,max_val fp32[1]; scale fp32[1]; zero_point int32[1]. */
int qv_obs_update(const uint32_t* acc, const int64_t* observer_enabled, const int64_t* fake_quant_enabled,
                  float* min_val, float* max_val, float* scale, int32_t* zero_point,
                  float averaging_const, int32_t qmin, int32_t qmax, int32_t symmetric, void* stream);

/* y = fake_quant(x) with per-tensor (scale, zp); mask (uint8, may be NULL) = STE mask;
 * identity copy when *fake_quant_enabled == 0.  (FakeQuantizeCore, torch K4.) */
int qv_fq_apply(const float* x, int64_t n, const float* scale, const int32_t* zero_point,
                const int64_t* fake_quant_enabled, int32_t qmin, int32_t qmax, float* y, uint8_t* mask,
                void* stream);

/* Weight flavour, one launch: per-row (per_channel=1, ch_axis 0) or whole-tensor (per_channel=0; two
 * internal phases) min/max -> EMA -> qparams -> fake-quant of W[rows, cols].
 * Outputs (each may be NULL): y fp32 [rows,cols]; mask uint8; codes bf16 [rows,cols] holding the
 * centred integer code (q - zp), exact in bf16; codes_t bf16 [cols,rows] (transposed copy).
 * scratch: uint32[2] (only used when per_channel == 0). */
int qv_fq_weight(const float* w, int64_t rows, int64_t cols, int32_t per_channel,
                 const int64_t* observer_enabled, const int64_t* fake_quant_enabled, float* min_val,
                 float* max_val, float* scale, int32_t* zero_point, float averaging_const, int32_t qmin,
                 int32_t qmax, int32_t symmetric, float* y, uint8_t* mask, uint16_t* codes, uint16_t* codes_t,
                 uint32_t* scratch, void* stream);

/* STE backward gx = gy * mask (FusedMovingAvgObsFqHelperBackward0). */
int qv_fq_bwd(const float* gy, const uint8_t* mask, int64_t n, float* gx, void* stream);

/* fp32 -> bf16 hi/lo planes (x ~= hi + lo, |err| <= 2^-17 |x|): how fp32 operands reach the bf16 tensor cores. */
int qv_split_planes(const float* x, int64_t n, uint16_t* hi, uint16_t* lo, void* stream);

/* ---- distillation loss (replaces ref/src/training/qat_trainer.py:343-349) ---------------------
 * loss = alpha*T^2*KL(softmax(t/T) || softmax(s/T))_batchmean + (1-alpha)*CE_labelsmooth(s, y).
 * s_raw: student logits [B,C]; if s_scale != NULL the logits are fake-quantised on load with
 * (s_scale, s_zp, qmin, qmax) and the returned gradient is STE-masked (fused output observer path).
 * out3: fp32[3] = {loss, loss_kd*T^2, loss_ce}; grad: dL/ds_raw [B,C] (may be NULL). */
int qv_kd_ce_loss(const float* s_raw, const float* t, const int64_t* labels, int32_t B, int32_t C, float T,
                  float alpha, float eps, const float* s_scale, const int32_t* s_zp, int32_t qmin,
                  int32_t qmax, float* out3, float* grad, void* stream);

/* ---- tcgen05 GEMM family (replaces F.linear inside torch.ao.nn.qat.Linear.forward,
 *      torch/ao/nn/qat/modules/linear.py:50-51, its autograd dgrad/wgrad, and the teacher's nn.Linear)
 *
 * D[M,N] (fp32) = sum over `npairs` plane pairs (pa, pb) of  A[pa] * B[pb]^T , then the epilogue
 *   d = acc * (col_scale ? col_scale[n] : 1) * (alpha ? *alpha : 1) * (col_rscale ? 1/col_rscale[n] : 1)
 *       + (bias ? bias[n] : 0)
 * A, B are bf16 "plane stacks": plane p starts at base + p*plane_stride elements.  An fp32 tensor is
 * represented as hi/lo planes (x ~= hi + lo); integer codes as one exact plane.
 *   a_mn_major = 0: A plane is [M rows, K cols] (row pitch lda), K contiguous.
 *   a_mn_major = 1: A plane is [K rows, M cols] (row pitch lda), M contiguous  (wgrad: A = gy^T).
 *   same for B with N in place of M.
 * splits > 1 (split-K): partial sums go to workspace[splits][M][N] fp32 and `d` is not written; call
 * qv_splitk_reduce afterwards.  minmax (uint32[2], may be NULL): ordered min/max of the stored d
 * values are atomically merged (fused output observer, phase 1).                                  */
typedef struct qv_gemm_args {
  const void* a; int64_t lda; int64_t a_plane_stride; int32_t a_mn_major;
  const void* b; int64_t ldb; int64_t b_plane_stride; int32_t b_mn_major;
  int32_t npairs; int32_t pair_a[4]; int32_t pair_b[4];
  int64_t M, N, K;
  float* d; int64_t ldd;
  const float* col_scale; const float* col_rscale; const float* alpha; const float* bias;
  uint32_t* minmax;
  int32_t splits; float* workspace;
} qv_gemm_args;
int qv_gemm_bf16(const qv_gemm_args* args, void* stream);

/* out[M,N] (+)= sum_z workspace[z][M][N] * (row_rscale ? 1/row_rscale[m] : 1) * (alpha ? *alpha : 1)
 * masked by mask[M,N] (uint8, may be NULL) -- the weight fake-quant STE on wgrad.  accumulate != 0 adds
 * into out (gradient arena). */
int qv_splitk_reduce(const float* workspace, int32_t splits, int64_t M, int64_t N, const float* row_rscale,
                     const float* alpha, const uint8_t* mask, float* out, int32_t accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QATVIT_B200_H_ */
