// qv_observer.cuh -- the per-tensor observer step shared by the standalone kernel (fakequant.cu, qv_obs_update) and the GEMM
// tail (gemm_sm100.cu: the last epilogue warp of the grid runs it, so no separate launch sits between a fake-quant Linear and
// its consumer).  Arithmetic contract: SURVEY.md App. A / oracle/fq_oracle.c (MovingAverageMinMax EMA in fp32 mul-then-add,
// fbgemm ChooseQuantizationParams: float scale, double zero-point math).
#pragma once
#include "qv_common.cuh"

namespace {

// ---------------- ChooseQuantizationParams (fbgemm flavour; see oracle/fq_oracle.c) ----------------
__device__ void qv_choose_qparams(float mn, float mx, int qmin, int qmax, bool preserve_sparsity, float* scale_out,
                                  int32_t* zp_out) {
  if (mn < 0.f && mx > 0.f && preserve_sparsity) {
    const int sqmin = -((qmax - qmin) / 2 + 1);
    const int sqmax = (qmax - qmin) / 2;
    const float a = __fdiv_rn(mn, (float)sqmin);
    const float b = __fdiv_rn(mx, (float)sqmax);
    const double ms = fmax(fabs((double)a), fabs((double)b));
    mn = (float)__dmul_rn(ms, (double)sqmin);
    mx = (float)__dmul_rn(ms, (double)sqmax);
  }
  mn = fminf(mn, 0.f);
  mx = fmaxf(mx, 0.f);
  float scale = (float)__ddiv_rn(__dsub_rn((double)mx, (double)mn), (double)(qmax - qmin));
  if (scale == 0.0f || isinf(__fdiv_rn(1.0f, scale))) scale = 0.1f;
  const float kSmall = 6.1e-5f;
  if (scale < kSmall) {
    const float org = scale;
    scale = kSmall;
    if (mn == 0.0f) {
      mx = __fmul_rn(kSmall, (float)(qmax - qmin));
    } else if (mx == 0.0f) {
      mn = __fmul_rn(-kSmall, (float)(qmax - qmin));
    } else {
      const float amp = __fdiv_rn(kSmall, org);
      mn = __fmul_rn(mn, amp);
      mx = __fmul_rn(mx, amp);
    }
  }
  const double ds = (double)scale;
  const double mn_s = __ddiv_rn((double)mn, ds), mx_s = __ddiv_rn((double)mx, ds);
  const double zp_from_min = __dsub_rn((double)qmin, mn_s);
  const double zp_from_max = __dsub_rn((double)qmax, mx_s);
  const double err_min = __dadd_rn((double)abs(qmin), fabs(mn_s));
  const double err_max = __dadd_rn((double)abs(qmax), fabs(mx_s));
  double zp0 = err_min < err_max ? zp_from_min : zp_from_max;
  if (mn < 0.f && mx > 0.f && preserve_sparsity) zp0 = (double)(qmin + qmax) / 2.0;
  int32_t zp;
  if (zp0 < (double)qmin) zp = qmin;
  else if (zp0 > (double)qmax) zp = qmax;
  else zp = (int32_t)rint(zp0);
  *scale_out = scale;
  *zp_out = zp;
}

__device__ __forceinline__ float qv_ema(float r, float cur, float c) {
  if (isinf(r)) return cur;
  return __fadd_rn(r, __fmul_rn(c, __fsub_rn(cur, r)));
}

// One observer update from the encoded (ordered-uint) batch min / max: EMA of the running range, then qparams.
__device__ __forceinline__ void qv_observer_step(uint32_t enc_min, uint32_t enc_max, const int64_t* obs_on, const int64_t* fq_on,
                                                 float* min_val, float* max_val, float* scale, int32_t* zp, float c, int qmin,
                                                 int qmax, int symmetric) {
  if (*obs_on == 0) return;
  if (enc_min == QV_ORD_MIN_INIT && enc_max == QV_ORD_MAX_INIT) return;   // empty tensor: nothing observed
  const float cur_min = qv_ord2f(enc_min), cur_max = qv_ord2f(enc_max);
  const float rmin = qv_ema(*min_val, cur_min, c);
  const float rmax = qv_ema(*max_val, cur_max, c);
  *min_val = rmin;
  *max_val = rmax;
  if (*fq_on != 0) {
    float s;
    int32_t z;
    qv_choose_qparams(rmin, rmax, qmin, qmax, symmetric != 0, &s, &z);
    *scale = s;
    *zp = z;
  }
}

}  // namespace
