// Probe: register mapping of tcgen05.ld.sync.aligned.16x256b.{x1,x2} (which TMEM lane / column each thread's registers hold).
// TMEM is filled through the row-per-lane 32x32b store with value = lane * 1000 + column; warp 0 then loads with 16x256b.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I qat-vit_b200/csrc -o /tmp/tmem_probe tools/ubench/tmem_ld_shape_probe.cu
#include <cstdio>
#include <cstdint>
#include "qv_ptx.cuh"
using namespace qvptx;

__global__ void probe(uint32_t* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  uint32_t v[32];
  for (int c = 0; c < 32; ++c) v[c] = (warp * 32 + lane) * 1000 + c;
  const uint32_t taddr = base + (static_cast<uint32_t>(warp * 32) << 16);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {     // warp 1 reads ITS quarter (lanes 32..63)
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(base + (32u << 16) + 8u) : "memory");       // lane base 32, column base 8
    tmem_ld_wait();
    for (int k = 0; k < 8; ++k) out[lane * 8 + k] = r[k];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(base + (48u << 16) + 0u) : "memory");       // second 16-lane half of the quarter
    tmem_ld_wait();
    for (int k = 0; k < 8; ++k) out[256 + lane * 8 + k] = r[k];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(base, 32); }
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 512 * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  uint32_t h[512];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int half = 0; half < 2; ++half) {
    printf("--- load at lane base %d, column base %d: thread: (lane,col) of r0..r7\n", half ? 48 : 32, half ? 0 : 8);
    for (int t = 0; t < 32; ++t) {
      printf("t%02d:", t);
      for (int k = 0; k < 8; ++k) printf(" (%u,%u)", h[half * 256 + t * 8 + k] / 1000, h[half * 256 + t * 8 + k] % 1000);
      printf("\n");
    }
  }
  return 0;
}
