// Microbenchmark: per-SMSP reciprocal throughput (cycles per warp-instruction) of the instructions the softmax / epilogue
// code paths are made of, on this GPU.  One CTA of NW warps per SM; every thread runs UNROLL independent chains.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_rates pipe_rates.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define ITERS 2048
template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[8];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; u[i] = threadIdx.x * 7 + i; }
  unsigned long long p[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(seed));
      if (OP == 2) { if (i < 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(p[(i + 1) & 3])); }
      if (OP == 3) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      if (OP == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7]));
      if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      if (OP == 6) asm volatile("and.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      if (OP == 7) { if (i < 4) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(p[(i + 1) & 3])); }
      if (OP == 8) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed));
      if (OP == 9) asm volatile("shl.b32 %0, %0, 16;" : "+r"(u[i]));
      if (OP == 10) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      if (OP == 11) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed));
      if (OP == 12) { // mixed: ex2 + cvt (do they share a pipe?)
        if (i & 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        else asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 2) & 7]));
      }
      if (OP == 13) { // mixed: fma + and (fma pipe + alu pipe)
        if (i & 1) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(seed));
        else asm volatile("and.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 2) & 7]));
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); s += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int nw : {4, 8, 16}) {
    k<OP><<<148, nw * 32>>>(out, cyc, 0.5f);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double warp_instr_per_smsp = (double)ITERS * per_iter * (nw / 4.0);
    printf("%-28s nw=%2d  %8lld cyc  -> %.2f cyc per warp-instr per SMSP\n", name, nw, h, h / warp_instr_per_smsp);
  }
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("MUFU.EX2", 8); run<1>("FFMA", 8); run<2>("FFMA2", 4); run<3>("F2FP.BF16.PACK_AB", 8); run<10>("F2FP.F16.PACK_AB", 8);
  run<4>("FMNMX3", 8); run<5>("PRMT", 8); run<6>("LOP3", 8); run<7>("FADD2", 4); run<8>("FADD", 8); run<11>("FMUL", 8); run<9>("SHL", 8);
  run<12>("mix EX2 + F2FP.BF16", 8); run<13>("mix FFMA + LOP3", 8);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
