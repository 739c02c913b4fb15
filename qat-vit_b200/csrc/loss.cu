// loss.cu -- fused teacher/student KL-divergence + label-smoothed cross-entropy, forward AND gradient in
// one launch.  Replaces the ~12 ATen launches of ref/src/training/qat_trainer.py:343-349
//   loss_ce = CrossEntropyLoss(label_smoothing)(s, y)
//   loss_kd = KLDivLoss('batchmean')(log_softmax(s/T), softmax(t/T)) * T^2
//   loss    = alpha*loss_kd + (1-alpha)*loss_ce
// and their autograd (dL/ds = alpha*T*(p_s - p_t)/B + (1-alpha)*(softmax(s) - q)/B).
// [B, C] is tiny (256 x 10): the kernel is latency-bound, so it is ONE block, one warp per sample row,
// with a fixed-order (deterministic) reduction.  Optionally fake-quantises the raw student logits on
// load (the head's output observer) and applies the STE mask to the returned gradient.
#include "qv_common.cuh"

namespace {

constexpr int LOSS_THREADS = 1024;

__global__ void __launch_bounds__(LOSS_THREADS) qv_kd_ce_kernel(const float* __restrict__ s_raw,
                                                               const float* __restrict__ t,
                                                               const int64_t* __restrict__ labels, int B, int C, float T,
                                                               float alpha, float eps, const float* s_scale,
                                                               const int32_t* s_zp, int qmin, int qmax,
                                                               float* __restrict__ out3, float* __restrict__ grad) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const bool fq = (s_scale != nullptr);
  QvQParams q;
  if (fq) q = qv_load_qparams(s_scale, s_zp, qmin, qmax);
  const float invT = 1.0f / T;
  const float invB = 1.0f / (float)B;
  float kd_acc = 0.f, ce_acc = 0.f;   // per-warp partial sums (lane 0 holds them)
  for (int b = warp; b < B; b += nwarps) {
    const float* sr = s_raw + (int64_t)b * C;
    const float* tr = t + (int64_t)b * C;
    // pass 1: maxima
    float ms = -INFINITY, mt = -INFINITY;
    for (int c = lane; c < C; c += 32) {
      float sv = sr[c];
      if (fq) sv = qv_fq(sv, q, nullptr, nullptr);
      ms = fmaxf(ms, sv);
      mt = fmaxf(mt, tr[c]);
    }
    ms = qv_warp_max(ms);
    mt = qv_warp_max(mt);
    // pass 2: partition functions at temperature T (student, teacher) and 1 (student)
    float zs = 0.f, zt = 0.f, z1 = 0.f;
    for (int c = lane; c < C; c += 32) {
      float sv = sr[c];
      if (fq) sv = qv_fq(sv, q, nullptr, nullptr);
      zs += expf((sv - ms) * invT);
      zt += expf((tr[c] - mt) * invT);
      z1 += expf(sv - ms);
    }
    zs = qv_warp_sum(zs);
    zt = qv_warp_sum(zt);
    z1 = qv_warp_sum(z1);
    const float lzs = logf(zs), lzt = logf(zt), lz1 = logf(z1);
    const int64_t y = labels[b];
    float kd = 0.f, ce = 0.f;
    for (int c = lane; c < C; c += 32) {
      float sv = sr[c];
      bool in = true;
      if (fq) sv = qv_fq(sv, q, &in, nullptr);
      const float lps = (sv - ms) * invT - lzs;        // log softmax(s/T)
      const float lpt = (tr[c] - mt) * invT - lzt;     // log softmax(t/T)
      const float lp1 = (sv - ms) - lz1;               // log softmax(s)
      const float pt = expf(lpt);
      const float qy = (c == y ? 1.0f - eps : 0.0f) + eps / (float)C;
      if (pt > 0.f) kd += pt * (lpt - lps);
      ce -= qy * lp1;
      if (grad) {
        const float g = (alpha * T * (expf(lps) - pt) + (1.0f - alpha) * (expf(lp1) - qy)) * invB;
        grad[(int64_t)b * C + c] = in ? g : 0.f;
      }
    }
    kd_acc += qv_warp_sum(kd);
    ce_acc += qv_warp_sum(ce);
  }
  __shared__ float skd[32], sce[32];
  if (lane == 0) { skd[warp] = kd_acc; sce[warp] = ce_acc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float kd = 0.f, ce = 0.f;
    for (int w = 0; w < nwarps; ++w) { kd += skd[w]; ce += sce[w]; }
    kd = kd * invB * T * T;
    ce = ce * invB;
    out3[0] = alpha * kd + (1.0f - alpha) * ce;
    out3[1] = kd;
    out3[2] = ce;
  }
}

}  // namespace

extern "C" int qv_kd_ce_loss(const float* s_raw, const float* t, const int64_t* labels, int32_t B, int32_t C, float T,
                             float alpha, float eps, const float* s_scale, const int32_t* s_zp, int32_t qmin,
                             int32_t qmax, float* out3, float* grad, void* stream) {
  QV_REQUIRE(s_raw && t && labels && out3, QV_ERR_INVALID, "null pointer in kd_ce_loss");
  QV_REQUIRE(B > 0 && C > 0 && T > 0.f, QV_ERR_INVALID, "bad kd_ce_loss shape (B=%d C=%d T=%f)", B, C, (double)T);
  QV_REQUIRE((s_scale == nullptr) == (s_zp == nullptr), QV_ERR_INVALID, "s_scale and s_zp go together");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  qv_kd_ce_kernel<<<1, LOSS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(s_raw, t, labels, B, C, T, alpha, eps,
                                                                             s_scale, s_zp, qmin, qmax, out3, grad);
  return qv_check_launch("qv_kd_ce_loss");
}
