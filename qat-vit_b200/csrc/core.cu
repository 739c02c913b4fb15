// core.cu -- library plumbing: version, thread-local error text, launch accounting.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "qv_common.cuh"

namespace {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
}  // namespace

int qv_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int qv_check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return qv_set_error(QV_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return QV_OK;
}

int qv_num_sms() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 0;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  if (dev != cached_dev) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      cudaGetLastError();
      return -1;
    }
    cached_dev = dev;
    cached_sms = sms;
  }
  return cached_sms;
}

extern "C" int qv_version(void) { return 3; }
extern "C" const char* qv_last_error(void) { return g_err; }
extern "C" int qv_device_sm_count(void) {
  int n = qv_num_sms();
  if (n <= 0) return qv_set_error(QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  return n;
}
extern "C" int64_t qv_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int qv_zero(void* ptr, int64_t bytes, void* stream) {
  QV_REQUIRE(ptr != nullptr && bytes >= 0, QV_ERR_INVALID, "bad qv_zero arguments");
  QV_REQUIRE(qv_num_sms() > 0, QV_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  cudaError_t e = cudaMemsetAsync(ptr, 0, static_cast<size_t>(bytes), static_cast<cudaStream_t>(stream));
  QV_REQUIRE(e == cudaSuccess, QV_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
  return QV_OK;
}
