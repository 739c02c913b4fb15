// optim.cu -- gradient clipping + AdamW on flat arenas: two launches instead of the ~300 foreach launches of
//   torch.nn.utils.clip_grad_norm_(params, 1.0); optimizer.step()        (ref/src/training/qat_trainer.py:360-361)
// over the 152 parameter tensors of the student.  Parameters, gradients and the two moment buffers are flat fp32 arenas in
// the same (parameter) order; the gradient arena is the buffer the NCCL all-reduce runs over, so the 1/world of the DDP
// gradient mean is folded into the same pass (grad_scale).
//
//   total = || g * grad_scale ||_2 ;  coef = min(1, max_norm / (total + 1e-6))          (clip_grad_norm_, error_if_nonfinite=False)
//   g' = g * grad_scale * coef
//   p *= 1 - lr * wd ;  m += (g' - m) * (1 - b1) ;  v = v * b2 + (1 - b2) * g'^2          (torch.optim.AdamW, foreach path)
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
//
// Kernel 1 writes one partial sum of squares per block (fixed order inside the block); kernel 2 re-reduces those partials in
// a fixed order in every block, so the result is deterministic and there is no host sync.
#include "qv_common.cuh"

namespace {

constexpr int OPT_THREADS = 256;

__global__ void __launch_bounds__(OPT_THREADS) sumsq_partials_kernel(const float* __restrict__ g, int64_t n,
                                                                    float* __restrict__ partials) {
  float acc = 0.f;
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = n4 * 4; i < n; ++i) acc += g[i] * g[i];
  acc = qv_warp_sum(acc);
  __shared__ float sm[OPT_THREADS / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < OPT_THREADS / 32; ++w) t += sm[w];
    partials[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(OPT_THREADS) clip_adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                                float* __restrict__ v, int64_t n,
                                                                const float* __restrict__ partials, int n_partials,
                                                                float grad_scale, float max_norm, float lr, float beta1,
                                                                float beta2, float eps, float weight_decay,
                                                                float bias_corr1, float bias_corr2_sqrt, float* norm_out,
                                                                int write_back_grad) {
  // every block reduces the partials in the same fixed order -> identical coefficient everywhere
  __shared__ float sm[OPT_THREADS / 32];
  __shared__ float s_coef;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n_partials; i += OPT_THREADS) acc += partials[i];
  acc = qv_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < OPT_THREADS / 32; ++w) t += sm[w];
    const float total = sqrtf(t) * grad_scale;
    float coef = max_norm > 0.f ? fminf(max_norm / (total + 1e-6f), 1.0f) : 1.0f;
    s_coef = coef * grad_scale;
    if (blockIdx.x == 0 && norm_out) *norm_out = total;
  }
  __syncthreads();
  const float gmul = s_coef;
  const float decay = 1.0f - lr * weight_decay;
  const float step_size = lr / bias_corr1;
  const float w1 = 1.0f - beta1, w2 = 1.0f - beta2;
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 gv = reinterpret_cast<float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float pp[4] = {pv.x, pv.y, pv.z, pv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w}, mm[4] = {mv.x, mv.y, mv.z, mv.w},
          vq[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = gg[j] * gmul;
      gg[j] = gj;
      pp[j] *= decay;
      mm[j] = mm[j] + (gj - mm[j]) * w1;
      vq[j] = vq[j] * beta2 + w2 * gj * gj;
      const float denom = sqrtf(vq[j]) / bias_corr2_sqrt + eps;
      pp[j] -= step_size * (mm[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(vq[0], vq[1], vq[2], vq[3]);
    if (write_back_grad) reinterpret_cast<float4*>(g)[i] = make_float4(gg[0], gg[1], gg[2], gg[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = n4 * 4; i < n; ++i) {
      const float gj = g[i] * gmul;
      float pj = p[i] * decay;
      const float mj = m[i] + (gj - m[i]) * w1;
      const float vj = v[i] * beta2 + w2 * gj * gj;
      pj -= step_size * (mj / (sqrtf(vj) / bias_corr2_sqrt + eps));
      p[i] = pj; m[i] = mj; v[i] = vj;
      if (write_back_grad) g[i] = gj;
    }
  }
}

}  // namespace

extern "C" int qv_clip_adamw(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float* partials,
                             int32_t n_partials, float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps,
                             float weight_decay, int64_t step, float* norm_out, int32_t write_back_grad, void* stream) {
  QV_REQUIRE(params && grads && exp_avg && exp_avg_sq && partials && n > 0 && n_partials > 0 && step >= 1, QV_ERR_INVALID,
             "bad clip_adamw arguments");
  QV_REQUIRE(qv_aligned16(params) && qv_aligned16(grads) && qv_aligned16(exp_avg) && qv_aligned16(exp_avg_sq), QV_ERR_INVALID,
             "clip_adamw arenas must be 16-byte aligned");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  sumsq_partials_kernel<<<n_partials, OPT_THREADS, 0, st>>>(grads, n, partials);
  int rc = qv_check_launch("qv_clip_adamw(sumsq)");
  if (rc) return rc;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
  int64_t blocks = ((n >> 2) + OPT_THREADS - 1) / OPT_THREADS;
  const int64_t cap = static_cast<int64_t>(qv_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  clip_adamw_kernel<<<static_cast<unsigned>(blocks), OPT_THREADS, 0, st>>>(
      params, grads, exp_avg, exp_avg_sq, n, partials, n_partials, grad_scale, max_norm, lr, beta1, beta2, eps, weight_decay,
      static_cast<float>(bc1), static_cast<float>(sqrt(bc2)), norm_out, write_back_grad);
  return qv_check_launch("qv_clip_adamw");
}
