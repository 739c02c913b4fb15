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
int qv_version(void);                 /* ABI version, currently 3 */
const char* qv_last_error(void);      /* thread-local message of the last failing call */
int qv_device_sm_count(void);         /* SMs of the current device, <0 on error (no device) */
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
int64_t qv_launch_count(void);
/* how many of those were CTA-pair GEMMs (tcgen05 cta_group::2, clusters of two CTAs; see qv_gemm_bf16) */
int64_t qv_gemm_pair_launches(void);
/* Stream-ordered clear of `bytes` bytes (cudaMemsetAsync): the engine's only whole-buffer initialisation -- the residual-stream
 * gradient whose non-cls rows are zero after the final LayerNorm's backward (ref: autograd's zero-filled SelectBackward0 of
 * `x[:, 0]` in timm VisionTransformer.forward_head). */
int qv_zero(void* ptr, int64_t bytes, void* stream);

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

/* Replaces: the `self.weight_fake_quant(self.weight)` call inside every torch.ao.nn.qat.Linear / Conv2d forward
 * (torch/ao/nn/qat/modules/linear.py:50-51, conv.py:55-56 -> fake_quantize.py:423-438), 49 per student forward
 * (ref/src/training/qat_trainer.py:341, `student_out = ddp_model(images)`).
 * Grouped per-channel weight fake-quant: every weight of the model in ONE launch (same arithmetic as qv_fq_weight with
 * per_channel = 1; outputs codes, codes_t and mask are all required).  `descs` is a DEVICE array of n_desc descriptors sorted by
 * block_start; weight i owns blocks [block_start_i, block_start_i + ceil(rows_i / 16)), total_blocks in all; rows must be
 * 16-byte aligned (cols % 4 == 0) and rows_i % 8 == 0 keeps the transposed stores aligned; max_cols = the largest cols. */
typedef struct qv_fqw_desc {
  const float* w; float* min_val; float* max_val; float* scale; int32_t* zero_point;
  const int64_t* observer_enabled; const int64_t* fake_quant_enabled;
  uint8_t* mask; uint16_t* codes; uint16_t* codes_t;
  int32_t rows, cols, block_start, reserved;
} qv_fqw_desc;
int qv_fq_weight_grouped(const qv_fqw_desc* descs_device, int32_t n_desc, int32_t total_blocks, int32_t max_cols,
                         float averaging_const, int32_t qmin, int32_t qmax, int32_t symmetric, void* stream);

/* Weight flavour, one launch: per-row (per_channel=1, ch_axis 0) or whole-tensor (per_channel=0; two
 * internal phases) min/max -> EMA -> qparams -> fake-quant of W[rows, cols].
 * Outputs (each may be NULL): y fp32 [rows,cols]; mask uint8; codes bf16 [rows,cols] holding the
 * centred integer code (q - zp), exact in bf16; codes_t bf16 [cols,rows] (transposed copy).
 * scratch: uint32[2] (only used when per_channel == 0).
 * scale_vec (may be NULL): fp32 [rows] <- the scale applied to each row (the per-tensor scale broadcast, or a copy of the
 * per-channel scales): the per-output-channel vector the GEMM epilogues and gradient-plane kernels take. */
int qv_fq_weight(const float* w, int64_t rows, int64_t cols, int32_t per_channel,
                 const int64_t* observer_enabled, const int64_t* fake_quant_enabled, float* min_val,
                 float* max_val, float* scale, int32_t* zero_point, float averaging_const, int32_t qmin,
                 int32_t qmax, int32_t symmetric, float* y, uint8_t* mask, uint16_t* codes, uint16_t* codes_t,
                 uint32_t* scratch, float* scale_vec, void* stream);

/* Learnable per-channel fake-quant (replaces torch._fake_quantize_learnable_per_channel_affine and its autograd node, called by
 * torch/ao/quantization/_learnable_fake_quantize.py:158-196; channel axis 0 of x[rows][cols], scale / zero_point fp32 [rows]).
 * OPT-IN: the reference has no learnable scale (its scale / zero_point are buffers); nothing on the default
 * path calls these.  forward: zr = clamp(rint(zero_point)), y = (clamp(zr + rint(x / scale)) - zr) * scale.
 * backward (north_star kernel 2, "the per-channel scale gradient done as a warp-shuffle reduction"): dx = gy * [in range];
 * dscale[r] / dzero_point[r] = per-channel sums of ATen's per-element terms times grad_factor (oracle/fq_oracle.c
 * qo_fq_learnable_bwd), one warp per channel, shuffle butterfly, deterministic.  dx / dscale / dzero_point may each be NULL. */
int qv_fq_learnable_fwd(const float* x, int64_t rows, int64_t cols, const float* scale, const float* zero_point,
                        int32_t qmin, int32_t qmax, float* y, void* stream);
int qv_fq_learnable_bwd(const float* gy, const float* x, int64_t rows, int64_t cols, const float* scale,
                        const float* zero_point, int32_t qmin, int32_t qmax, float grad_factor, float* dx, float* dscale,
                        float* dzero_point, void* stream);

/* STE backward gx = gy * mask (FusedMovingAvgObsFqHelperBackward0). */
int qv_fq_bwd(const float* gy, const uint8_t* mask, int64_t n, float* gx, void* stream);

/* fp32 -> bf16 hi/lo planes (x ~= hi + lo, |err| <= 2^-17 |x|): how fp32 operands reach the bf16 tensor cores. */
int qv_split_planes(const float* x, int64_t n, uint16_t* hi, uint16_t* lo, void* stream);
/* fp32 [rows, cols] (cols % 64 == 0) -> the "mixed" GEMM operand format (qv_gemm_args.mix): region0 = fp16(x * 2^s) [rows][cols];
 * region1 (rows * cols * 2 bytes): per row and 64-column block 64 fp8 of the value, then 64 e5m2 of the fp16 rounding residual.
 * kind 0 = activation scales (A operand), 1 = weight scales (B operand; the frozen teacher's nn.Linear weights, split once). */
int qv_split_planes_mix(const float* x, int64_t rows, int64_t cols, int32_t kind, uint16_t* region0, uint16_t* region1, void* stream);

/* ---- distillation loss (replaces ref/src/training/qat_trainer.py:343-349) ---------------------
 * loss = alpha*T^2*KL(softmax(t/T) || softmax(s/T))_batchmean + (1-alpha)*CE_labelsmooth(s, y).
 * s_raw: student logits [B,C]; if s_scale != NULL the logits are fake-quantised on load with
 * (s_scale, s_zp, qmin, qmax) and the returned gradient is STE-masked (fused output observer path).
 * out3: fp32[3] = {loss, loss_kd*T^2, loss_ce}; grad: dL/ds_raw [B,C] (may be NULL). */
int qv_kd_ce_loss(const float* s_raw, const float* t, const int64_t* labels, int32_t B, int32_t C, float T,
                  float alpha, float eps, const float* s_scale, const int32_t* s_zp, int32_t qmin,
                  int32_t qmax, float* out3, float* grad, void* stream);
/* The same loss for many classes / large batches (the reference's lines 343-349 are shape-agnostic; ImageNet-width heads): one warp
 * per row over a persistent grid, 16-byte loads when C % 4 == 0, every element read from HBM once and its gradient written once
 * (12 B per element), fixed-order (deterministic) reduction.  workspace: qv_kd_ce_rows_workspace_floats(B) floats owned by the
 * caller, ZERO before the first use (the kernel leaves its ticket word zero again). */
int64_t qv_kd_ce_rows_workspace_floats(int32_t B);
int qv_kd_ce_loss_rows(const float* s_raw, const float* t, const int64_t* labels, int32_t B, int32_t C, float T,
                       float alpha, float eps, const float* s_scale, const int32_t* s_zp, int32_t qmin,
                       int32_t qmax, float* out3, float* grad, float* workspace, void* stream);

/* ---- tcgen05 GEMM family (replaces F.linear inside torch.ao.nn.qat.Linear.forward,
 *      torch/ao/nn/qat/modules/linear.py:50-51, its autograd dgrad/wgrad, the teacher's nn.Linear, and --
 *      batched per (image, head) -- the matmuls inside F.scaled_dot_product_attention and its backward)
 *
 * D[M,N] (fp32) = sum over plane pairs (pa, pb) of  A[pa] * B[pb]^T  (see a_planes / b_planes), then the epilogue
 *   d = acc * (col_scale ? col_scale[n] : 1) * (alpha ? *alpha : 1) * (col_rscale ? 1/col_rscale[n] : 1)
 *       + (bias ? bias[n] : 0)
 * A, B are bf16 "plane stacks": an fp32 tensor is represented as hi/lo planes (x ~= hi + lo), integer
 * fake-quant codes as ONE exact plane.  Per operand (qv_operand):
 *   mn_major = 0: each matrix is [rows = M (or N), cols >= K], K contiguous (row pitch ld);
 *   mn_major = 1: each matrix is [rows = K, cols >= M (or N)], M/N contiguous (wgrad: A = gy^T);
 *   rows/cols: extent of ONE matrix -- reads outside are zero-filled by TMA (ragged tiles, batch edges);
 *   nb, batch_stride: number of / element distance between consecutive matrices of the tensor;
 *   batch item bt -> (bo, bi) = (bt / batch_inner, bt % batch_inner) selects matrix bo*c2_outer + bi*c2_inner
 *   and column offset col0 + bi*col_inner (a head's 64 columns inside a [tokens, 3*D] tensor);
 *   plane p starts p*plane_stride elements after ptr.
 * splits > 1 (split-K, unbatched only): raw partial sums go to workspace[splits][M][N] fp32, `d` and the
 * epilogue terms are ignored; call qv_splitk_reduce afterwards.  minmax (uint32[2], may be NULL): ordered
 * min/max of the stored d values are atomically merged (fused output observer, phase 1).            */
typedef struct qv_operand {
  const void* ptr; int64_t ld; int64_t plane_stride; int32_t mn_major;
  int64_t rows, cols;
  int64_t nb, batch_stride;
  int32_t c2_outer, c2_inner, col0, col_inner;
} qv_operand;
/* fp32 output tensor [nb][rows][ld]; batch item (bo, bi) is written at matrix bo*c2_outer + bi*c2_inner, columns
 * col0 + bi*col_inner + [0, N); rows >= `rows` and columns >= `cols` are clipped by the TMA store. */
typedef struct qv_out {
  float* ptr; int64_t ld;
  int64_t rows, cols;
  int64_t nb, batch_stride;
  int32_t c2_outer, c2_inner, col0, col_inner;
} qv_out;
typedef struct qv_gemm_args {
  qv_operand a, b;
  int32_t a_planes, b_planes;   /* (1,1): A*B ; (2,1): (A0+A1)*B ; (2,2): A0*B0 + A0*B1 + A1*B0 */
  int64_t M, N, K;
  qv_out out;
  const float* col_scale; const float* col_rscale; const float* alpha; const float* bias;
  uint32_t* minmax;
  int32_t splits; float* workspace;
  int32_t nbatch, batch_inner;
  int32_t tile_n;               /* 0 = auto; else 64 / 128 / 192 */
  /* out_kind = 1: `out.ptr` is a bf16 hi/lo plane stack [2][nb][rows][ld] (ld in bf16 elements, planes out_plane_stride
   * elements apart) -- the operand format of the next GEMM -- written straight from the epilogue, so a producer/consumer
   * pair of Linears with no observer in between (the teacher) never round-trips fp32 through HBM.
   * act = 1: exact-erf GELU after scale/bias (timm Mlp.act fused into fc1's epilogue); needs out_kind = 1. */
  int32_t out_kind; int32_t act; int64_t out_plane_stride;
  /* act = 2 (needs out_kind = 1, (2,1) planes): the dgrad GEMM that produces dL/d(out of a fake-quantised Linear [+ GELU])
   * applies that Linear's backward prologue in its epilogue -- qv_gp_planes fused (STE of torch's
   * FusedMovingAvgObsFqHelperBackward0 and GeluBackward0):
   *   gq[m,n] = acc[m,n] * (ep_gelu ? gelu'(FQ(y[m,n])) : 1) * STEmask(y[m,n]),   y = ep_raw (row pitch ep_raw_ld floats),
   *   out planes = hi/lo split of gq * col_scale[n];  ep_colsum (may be NULL): fp32 [ceil(M/32)][N], row s = column sums
   *   of gq over token rows 32 s .. 32 s + 31 (bias-grad partials; reduce with qv_colsum_reduce). */
  const float* ep_raw; int64_t ep_raw_ld; const float* ep_scale; const int32_t* ep_zp;
  int32_t ep_qmin, ep_qmax, ep_gelu; float* ep_colsum;
  /* obs_ticket != NULL (needs minmax, fp32 output, no split-K): the output observer's update -- qv_obs_update on (minmax ->
   * obs_min_val / obs_max_val EMA with obs_c -> obs_scale / obs_zero_point) -- runs in the kernel's tail, done by the last
   * epilogue warp of the grid, so the consumer of the fake-quantised output can be launched straight after the GEMM.
   * obs_ticket: a zero-initialised uint32 the kernel leaves at zero (one per stream). */
  float* obs_min_val; float* obs_max_val; float* obs_scale; int32_t* obs_zero_point;
  const int64_t* obs_enabled; const int64_t* obs_fq_enabled;
  float obs_c; int32_t obs_qmin, obs_qmax, obs_symmetric;
  uint32_t* obs_ticket;
  /* mix != 0 (needs a_planes = b_planes = 2, K-major operands, K % 64 == 0, no split-K): both operands are in the "mixed"
   * format written by qv_split_planes_mix / out_kind = 2 / the plane_fmt = 1 producers: region 0 (plane 0) holds fp16(x * 2^s),
   * region 1 (plane 1, same byte size) holds per 64-column block 64 fp8 of the value then 64 fp8 (e5m2) of the fp16 rounding
   * residual.  The product is fp16.fp16 + hi8.lo8 + lo8.hi8 in ONE fp32 accumulator (kind::f16 + 2 x kind::f8f6f4 at twice the
   * rate: the cost of two bf16 passes instead of three, same ~2^-16 accuracy); the kernel rescales by 2^-14 before the epilogue.
   * A must be an activation-kind tensor and B a weight-kind one.  out_kind = 2: like out_kind = 1, but the output planes are
   * written in the mixed activation format (the next mixed GEMM's A operand). */
  int32_t mix;
  /* out_kind = 2 only (may be NULL): range guard of the mixed activation format.  When an output value leaves the format's range
   * (|x| > 448: its fp8 e5m2 copy saturates) the kernel ORs sat_bit into *sat_flag (device int32; one atomic per warp, only
   * then).  The host checks the flag and routes that tensor to bf16 hi/lo planes (qatvit_b200/engine.py TeacherEngine). */
  int32_t* sat_flag; int32_t sat_bit;
} qv_gemm_args;
/* Scheduling (no ABI surface): unsplit, unbatched K-major GEMMs with M >= 256 x (SMs / 2) run as CTA PAIRS -- clusters of two
 * CTAs on the two SMs of a TPC computing one 256 x tile_n tile with tcgen05 cta_group::2, each SM staging 128 rows of A and
 * half of the B tile (the 4-byte-per-element operand formats are otherwise bound by the L2 -> SM fill rate, not the tensor
 * pipe).  Results are bit-identical to one CTA per tile.  Environment variable QV_GEMM_PAIR (bit mask, default 51; 0 = off)
 * selects the GEMM kinds; qv_gemm_pair_launches() counts them.  DESIGN.md section 3. */
int qv_gemm_bf16(const qv_gemm_args* args, void* stream);

/* out[M,N] (+)= sum_z workspace[z][M][N] * (row_rscale ? 1/row_rscale[m] : 1) * (alpha ? *alpha : 1)
 * masked by mask[M,N] (uint8, may be NULL) -- the weight fake-quant STE on wgrad.  accumulate != 0 adds
 * into out (gradient arena). */
int qv_splitk_reduce(const float* workspace, int32_t splits, int64_t M, int64_t N, const float* row_rscale,
                     const float* alpha, const uint8_t* mask, float* out, int32_t accumulate, void* stream);

/* ---- row / elementwise kernels around the GEMMs (LayerNorm, residual, GELU, embeddings, attention softmax) ----
 * They replace ATen native_layer_norm / add / gelu / softmax and their backward nodes inside the timm blocks the
 * reference trains; the activation fake-quant of the PRODUCING Linear is applied on load from
 * its raw output and (scale, zero_point), so the hook of torch/ao/quantization/quantize.py:150-152 costs no pass. */

/* x_out = x_in + FQ(y_raw) ; h = LayerNorm(x_out)*gamma+beta.  Row r reads input row r*in_row_stride.  Any of
 * x_in / y_raw / x_out / h_planes (bf16 [2][R][D]) / h_f32 / mean / rstd may be NULL; y_scale NULL = no fake-quant.
 * minmax (uint32[2], may be NULL): ordered min / max of h merged atomically -- the output observer of an OBSERVED LayerNorm
 * (plain nn.LayerNorm under prepare_qat), phase 1.
 * plane_fmt: 0 = h_planes are bf16 hi/lo planes; 1 = the mixed fp16 + fp8 operand format (qv_split_planes_mix, activation kind).
 * sat_flag / sat_bit (plane_fmt 1, may be NULL): range guard of the mixed format, see qv_gemm_args.sat_flag. */
int qv_resid_ln_fwd(const float* x_in, const float* y_raw, const float* y_scale, const int32_t* y_zp, int32_t qmin,
                    int32_t qmax, const float* gamma, const float* beta, float eps, int64_t R, int32_t D,
                    int64_t in_row_stride, float* x_out, uint16_t* h_planes, int64_t plane_stride, float* h_f32,
                    float* mean, float* rstd, uint32_t* minmax, int32_t plane_fmt, int32_t* sat_flag, int32_t sat_bit,
                    void* stream);
/* g_x[r*out_row_stride] = g_res[r] + LayerNormBackward(g_h, x, mean, rstd, gamma)[r]; partials: fp32
 * [ceil(R/rows_per_block)][2][D] per-block dgamma / dbeta sums (reduce with qv_colsum_reduce).
 * h_raw (may be NULL): the raw LayerNorm output of an observed LayerNorm; g_h then passes the STE mask of its fake-quant
 * (h_scale, h_zp, qmin, qmax) on load. */
int qv_ln_bwd(const float* g_h, const float* x, const float* mean, const float* rstd, const float* gamma,
              const float* g_res, int64_t R, int32_t D, int64_t out_row_stride, float* g_x, float* partials,
              int32_t rows_per_block, const float* h_raw, const float* h_scale, const int32_t* h_zp, int32_t qmin, int32_t qmax,
              void* stream);
/* Replaces (inside `loss.backward()`, ref qat_trainer.py:359): NativeLayerNormBackward0 of timm Block.norm1 / norm2 + the residual
 * AddBackward0 + FusedMovingAvgObsFqHelperBackward0 / bias reduction of the Linear feeding that residual (attn.proj, mlp.fc2).
 * qv_ln_bwd (out_row_stride 1) that also emits the gradient planes of the Linear whose fake-quantised output gp_y was added
 * into the residual stream this LayerNorm reads (attn.proj for norm2, the previous block's mlp.fc2 for norm1):
 * gp_out = hi/lo planes [2][R][D] of g_x * STEmask(gp_y) * gp_wscale[col]; gp_partials (may be NULL) fp32
 * [ceil(R/rows_per_block)][D]: per-block column sums of g_x * mask (that Linear's bias grad; reduce with qv_colsum_reduce). */
int qv_ln_bwd_gp(const float* g_h, const float* x, const float* mean, const float* rstd, const float* gamma, const float* g_res,
                 int64_t R, int32_t D, float* g_x, float* partials, int32_t rows_per_block, const float* h_raw,
                 const float* h_scale, const int32_t* h_zp, int32_t qmin, int32_t qmax, const float* gp_y, const float* gp_scale,
                 const int32_t* gp_zp, int32_t gp_qmin, int32_t gp_qmax, const float* gp_wscale, uint16_t* gp_out,
                 int64_t gp_plane_stride, float* gp_partials, void* stream);
int qv_colsum_reduce(const float* partials, int32_t nblk, int64_t ncols, float* out, int32_t accumulate, void* stream);
int qv_colsum_rows(const float* x, int64_t R, int64_t N, int64_t ld, float* out, int32_t accumulate, void* stream);
/* gp'[r,n] = g[r',n] * [gelu'(FQ(y))] * STEmask(y_raw[r,n]) * w_scale[n] -> bf16 hi/lo planes [2][R][N], plus per-block
 * column sums of the unscaled masked gradient (bias grad partials [ceil(R/rows_per_block)][N]).
 * remap_P/T != 0: output row b*P+i reads g row b*T+i+1 (patch-embed: drop the cls row). */
int qv_gp_planes(const float* g, const float* y_raw, const float* y_scale, const int32_t* y_zp, int32_t qmin, int32_t qmax,
                 const float* w_scale, int32_t w_scale_per_channel, int32_t gelu, int64_t R, int64_t N, int32_t remap_P,
                 int32_t remap_T, uint16_t* out_planes, int64_t plane_stride, float* bias_partials, int32_t rows_per_block,
                 void* stream);
/* out planes = [GELU](FQ(y_raw)) elementwise (n % 4 == 0).  codes_only != 0: ONE plane of centred integer codes (q - zp),
 * exact in bf16 -- FQ(y) = code * scale with the scale applied by the consumer (integer-code attention kernels). */
int qv_act_planes(const float* y_raw, const float* y_scale, const int32_t* y_zp, int32_t qmin, int32_t qmax, int32_t gelu,
                  int32_t codes_only, int64_t n, uint16_t* out_planes, int64_t plane_stride, void* stream);
/* x0 = cat(cls, FQ(p_raw)) + pos  ->  [B*(P+1), D] fp32 (timm VisionTransformer._pos_embed). */
int qv_embed_fwd(const float* p_raw, const float* p_scale, const int32_t* p_zp, int32_t qmin, int32_t qmax, const float* cls,
                 const float* pos, int64_t B, int32_t P, int32_t D, float* x0, void* stream);
/* im2col for the patch x patch / patch conv (nnqat.Conv2d, torch/ao/nn/qat/modules/conv.py:55-56, as a GEMM):
 * rows = B*(HW/patch)^2 patches, cols = C*patch*patch.  With (scale, zp): the image is fake-quantised on load and
 * out_plane holds the centred integer codes (one exact bf16 plane; out_lo_plane must be NULL).  With scale NULL
 * (teacher): raw fp32 pixels as bf16 hi (out_plane) / lo (out_lo_plane) planes. */
int qv_im2col_fq(const float* img, const float* scale, const int32_t* zp, int32_t qmin, int32_t qmax, int64_t B, int32_t C,
                 int32_t HW, int32_t patch, uint16_t* out_plane, uint16_t* out_lo_plane, void* stream);
/* P = softmax(S[:, :T] * scale) per row -> bf16 hi/lo planes (row pitch ldP, zero padded); T <= 256. */
int qv_softmax_planes(const float* S, int64_t ldS, int64_t rows, int32_t T, float scale, uint16_t* P, int64_t ldP,
                      int64_t plane_stride, void* stream);
/* dS = P * (dP - rowsum(dP*P)) * scale -> bf16 hi/lo planes. */
int qv_attn_ds(const uint16_t* P, int64_t ldP, int64_t p_plane_stride, const float* dP, int64_t lddP, int64_t rows, int32_t T,
               float scale, uint16_t* dS, int64_t ldS, int64_t s_plane_stride, void* stream);
/* Fused softmax attention forward, one (image, head) per work item on tcgen05 / TMEM (replaces
 * F.scaled_dot_product_attention in timm Attention.forward):  O = softmax(Q K^T * scale) V, head_dim 64,
 * T <= 224 tokens; scores / probabilities stay in tensor memory.
 * qkv_planes: bf16 plane stack [n_planes][B*T][ld] holding Q | K | V column blocks (H*64 columns each).
 *   n_planes = 2: fp32 values as hi/lo planes;  n_planes = 1: exact integer fake-quant codes, with the observer's scale
 *   passed as device scalars: logits *= (*qk_scale)^2, output *= *v_scale (either may be NULL).
 * out_planes: bf16 hi/lo planes [2][B*T][out_ld]; head h fills columns h*64..h*64+63 (the proj GEMM's A operand);
 * out_f32: fp32 [B*T][H*64] copy of the output (either output may be NULL, not both).
 * lse (may be NULL): fp32 [B*H*T] natural-log logsumexp of the scaled logits (saved for a recomputing backward).
 * out_fmt: 0 = out_planes are bf16 hi/lo planes; 1 = the mixed fp16 + fp8 operand format (activation kind; out_ld % 64 == 0).
 * sat_flag / sat_bit (out_fmt 1, may be NULL): range guard of the mixed format, see qv_gemm_args.sat_flag. */
int qv_attn_fwd(const uint16_t* qkv_planes, int32_t n_planes, int64_t plane_stride, int64_t ld, int32_t B, int32_t T,
                int32_t H, float scale, const float* qk_scale, const float* v_scale, uint16_t* out_planes,
                int64_t out_plane_stride, int64_t out_ld, float* out_f32, float* lse, int32_t out_fmt, int32_t* sat_flag,
                int32_t sat_bit, void* stream);
/* Fused attention backward for integer-code operands (the QAT student; autograd of F.scaled_dot_product_attention):
 * recomputes P from the codes and the forward's lse on the tensor cores and writes dQ | dK | dV (gradients w.r.t. the
 * fake-quantised q, k, v = s * codes) into g_qkv fp32 [B*T][3*H*64].  qkv_codes: ONE bf16 plane [B*T][ld]; o_planes: the
 * forward's output, bf16 hi/lo planes [2][B*T][o_ld] (qv_attn_fwd out_planes; gives delta = dO . O); do_planes: bf16 hi/lo
 * planes [2][B*T][do_ld] of dL/dO; lse: fp32 [B*H*T] (qv_attn_fwd); qscale: device scalar s (NULL = 1).  T <= 224. */
int qv_attn_bwd(const uint16_t* qkv_codes, int64_t ld, const float* qscale, const uint16_t* o_planes, int64_t o_plane_stride,
                int64_t o_ld, const uint16_t* do_planes, int64_t do_plane_stride, int64_t do_ld, const float* lse, int32_t B,
                int32_t T, int32_t H, float scale, float* g_qkv, void* stream);
/* Replaces (inside `loss.backward()`, ref/src/training/qat_trainer.py:359): ScaledDotProductAttention backward of timm
 * Attention.forward + FusedMovingAvgObsFqHelperBackward0 of the qkv output hook (torch/ao/quantization/quantize.py:150-152) + the
 * bias reduction of the qkv AddmmBackward0.
 * qv_attn_bwd with the qkv Linear's backward prologue (qv_gp_planes) fused into its output stage: instead of fp32 dQ | dK | dV,
 * writes gp = g * STEmask(y_raw) * w_scale[col] as bf16 hi/lo planes [2][B*T][3*H*64] (the A operand of the qkv dgrad / wgrad
 * GEMMs) and colsum fp32 [B * ceil(T/128) * 4][3*H*64]: per 32-token slab column sums of g * mask (bias-grad partials, reduce
 * with qv_colsum_reduce).  y_raw: the qkv Linear's raw output fp32 [B*T][3*H*64]; (y_scale, y_zp, qmin, qmax): its output
 * fake-quant; w_scale: fp32 [3*H*64] per-output-channel weight scale. */
int qv_attn_bwd_gp(const uint16_t* qkv_codes, int64_t ld, const float* qscale, const uint16_t* o_planes, int64_t o_plane_stride,
                   int64_t o_ld, const uint16_t* do_planes, int64_t do_plane_stride, int64_t do_ld, const float* lse, int32_t B,
                   int32_t T, int32_t H, float scale, const float* y_raw, const float* y_scale, const int32_t* y_zp,
                   int32_t qmin, int32_t qmax, const float* w_scale, uint16_t* gp_planes, int64_t gp_plane_stride,
                   float* colsum, void* stream);
/* classifier head (D -> num_classes), exact fp32: out = x wq^T + bias (+ fused output-observer min/max). */
int qv_head_fwd(const float* x, const float* wq, const float* bias, int32_t B, int32_t K, int32_t N, float* out,
                uint32_t* minmax, void* stream);
int qv_head_bwd(const float* g, const float* x, const float* wq, const uint8_t* wmask, int32_t B, int32_t K, int32_t N,
                float* gx, float* gw, float* gb, int32_t accumulate, void* stream);

/* ---- converted int8 student (replaces torch.ops.quantized.linear behind the nnq.Linear / nnq.Conv2d modules stock
 *      convert() builds: ref/src/training/qat_trainer.py:377-388; torch/ao/nn/quantized/modules/linear.py:187-190) ----
 * acc = sum_k (q_x - z_x) q_w on tcgen05 kind::i8 (u8 x s8 -> s32 in TMEM), then the engine's requantisation:
 *   bias_int = 0 (x86 / fbgemm, torch's default CPU engine): q_y = clamp(rint((float(acc) + b/(s_x s_w)) * (s_x s_w/s_y)) + z_y, 0, 255)
 *   bias_int = 1 (qnnpack engine):                           q_y = clamp(rint(float(acc + rint(b/(s_x s_w))) * (s_x s_w/s_y)) + z_y, 0, 255)  qx uint8 [M,K]; qw int8 [N,K] (symmetric: weight zero point 0); sx / zx: DEVICE scalars (the
 * input's qparams, possibly computed on device); sw fp32 [N] (per_channel) or [1]; wsum int32 [N] = sum_k qw[n,k];
 * bias fp32 [N] or NULL; (sy, zy) the module's output qparams.  Outputs (either may be NULL): qy uint8 [M,N] codes,
 * y fp32 [M,N] = (qy - zy) * sy (what DeQuantize / the float glue consumes).  K % 16 == 0. */
int qv_int8_linear(const uint8_t* qx, int64_t M, int64_t K, const float* sx, const int32_t* zx, const int8_t* qw, int64_t N,
                   const float* sw, int32_t per_channel, const int32_t* wsum, const float* bias, float sy, int32_t zy,
                   int32_t bias_int, uint8_t* qy, float* y, void* stream);
/* Same product and requantisation; the output is the CENTRED code q_y - z_y as bf16 [M,N] (|q_y - z_y| <= 255 is exact): the
 * one-plane integer operand qv_attn_fwd takes for q, k, v, with s_y passed to it as qk_scale / v_scale.  N % 8 == 0. */
int qv_int8_linear_codes(const uint8_t* qx, int64_t M, int64_t K, const float* sx, const int32_t* zx, const int8_t* qw, int64_t N,
                         const float* sw, int32_t per_channel, const int32_t* wsum, const float* bias, float sy, int32_t zy,
                         int32_t bias_int, uint16_t* codes, void* stream);
/* q = clamp(rint(x * (1/scale)) + zero_point, 0, 255) with device-scalar qparams (torch.quantize_per_tensor / nnq.Quantize). */
int qv_quantize_u8(const float* x, int64_t n, const float* scale, const int32_t* zero_point, uint8_t* q, void* stream);
/* Affine qparams from an ordered min/max accumulator (qv_minmax_accumulate), Python-observer formula
 * (torch/ao/quantization/observer.py:349-427): scale = max((max+ - min-)/(qmax-qmin), eps), zp = clamp(qmin - rint(min-/scale)). */
int qv_qparams_from_minmax(const uint32_t* acc, int32_t qmin, int32_t qmax, float* scale, int32_t* zero_point, void* stream);
/* Quantise the input image and gather 16x16 patches in one pass (nnq.Quantize + the im2col of nnq.Conv2d): uint8 [B*P, C*p*p]. */
int qv_im2col_u8(const float* img, const float* scale, const int32_t* zp, int64_t B, int32_t C, int32_t HW, int32_t patch,
                 uint8_t* out, void* stream);
/* y = GELU_erf(x) in fp32, with the min / max of y merged into acc (may be NULL): the float glue between fc1 and fc2. */
int qv_gelu_minmax(const float* x, int64_t n, float* y, uint32_t* acc, void* stream);
/* Compact float glue of the converted student (same arithmetic as the fp32 glue above, on the Linear's quint8 OUTPUT codes:
 * (q - z_y) * s_y is an integer code times one scale -- torch/ao/nn/quantized/modules/linear.py:187-190 returns exactly that
 * quantized tensor; the reference's DeQuantStub / float modules then see its dequantised values):
 *   qv_quantize_u8_dyn  = qv_qparams_from_minmax(acc, 0, 255) + qv_quantize_u8 in one launch (qparams also stored for the consumer);
 *   qv_codes_from_u8    : codes[i] = bf16(q[i] - zero_point), the one-plane integer operand of qv_attn_fwd (n % 16 == 0);
 *   qv_gelu_u8_minmax   : min / max of GELU_erf((q - zy) * sy) merged into acc (a 256-entry table of the qv_gelu_minmax values);
 *   qv_gelu_u8_requant  : out = quantize_u8(GELU_erf((q - zy) * sy)) with the dynamic qparams of acc (a 256-byte code -> code
 *                         table), qparams stored for the consuming Linear.  Bit-identical to qv_int8_linear(y) -> qv_gelu_minmax
 *                         -> qv_qparams_from_minmax -> qv_quantize_u8 (tests/test_int8_gpu.py). */
int qv_quantize_u8_dyn(const float* x, int64_t n, const uint32_t* acc, float* scale_out, int32_t* zero_point_out, uint8_t* q,
                       void* stream);
int qv_codes_from_u8(const uint8_t* q, int64_t n, int32_t zero_point, uint16_t* codes, void* stream);
/* q = quantize_u8(LayerNorm(x) * gamma + beta) with the dynamic qparams of acc, the LayerNorm output recomputed from the row
 * statistics qv_resid_ln_fwd saved (mean, rstd: fp32 [R]) instead of read back as fp32; x fp32 [R][D] is that call's x_out (or its
 * x_in when there was no residual).  Bit-identical to qv_resid_ln_fwd(h_f32) -> qv_quantize_u8_dyn.  D as for qv_resid_ln_fwd. */
int qv_ln_quantize_u8_dyn(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, int64_t R,
                          int32_t D, const uint32_t* acc, float* scale_out, int32_t* zero_point_out, uint8_t* q, void* stream);
int qv_gelu_u8_minmax(const uint8_t* q, int64_t n, float sy, int32_t zy, uint32_t* acc, void* stream);
int qv_gelu_u8_requant(const uint8_t* q, int64_t n, float sy, int32_t zy, const uint32_t* acc, float* scale_out,
                       int32_t* zero_point_out, uint8_t* out, void* stream);

/* ---- clip_grad_norm_ + AdamW on flat arenas (replaces ref/src/training/qat_trainer.py:360-361; torch.optim.AdamW foreach
 *      arithmetic, single parameter group) ----
 * params / grads / exp_avg / exp_avg_sq: fp32 [n] arenas in the same order; partials: fp32 [n_partials] scratch.
 * total = ||grads * grad_scale||_2 (written to norm_out if not NULL); grads are scaled by grad_scale * min(1, max_norm /
 * (total + 1e-6)) (max_norm <= 0: no clipping) before the AdamW update of step `step` (1-based); write_back_grad != 0 stores the
 * clipped gradients back (what clip_grad_norm_ leaves in .grad).  Deterministic; two launches; no host sync. */
int qv_clip_adamw(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float* partials, int32_t n_partials,
                  float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                  float* norm_out, int32_t write_back_grad, void* stream);

/* ---- input transform: ref/src/training/qat_trainer.py:210-216 -------------------------------------
 * transforms.Compose([Resize(size, BICUBIC), ToTensor(), Normalize(mean, std)]) on a batch of uint8 HWC images
 * img [B][Hin][Win][C] -> out fp32 [B][C][Hout][Wout].  Pillow's two-pass 8-bit bicubic resample (horizontal, then vertical;
 * 22-bit fixed-point taps), /255, (x - mean[c]) / std[c]: bit-identical to the CPU pipeline.  bounds_* int32 [n_out][2] =
 * (first input index, tap count), coef_* int32 [n_out][ksize]: Pillow's precompute_coeffs + normalize_coeffs_8bpc for the
 * horizontal (Win -> Wout) and vertical (Hin -> Hout) pass (computed by the host side, qatvit_b200/data.py). */
int qv_resize_normalize_u8(const uint8_t* img, int64_t B, int32_t Hin, int32_t Win, int32_t C, int32_t Hout, int32_t Wout,
                           const int32_t* bounds_h, const int32_t* coef_h, const int32_t* bounds_v, const int32_t* coef_v,
                           int32_t ksize, const float* mean, const float* stdv, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QATVIT_B200_H_ */
