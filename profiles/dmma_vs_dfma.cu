// microbenchmark: DFMA vs DMMA (mma.sync m8n8k4 f64) throughput on one GPU, alone and mixed
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE> __global__ void k(double* out, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-3, b = 1.0 - seed;
  double f[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = a + i; c[i] = b * i; }
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fma(f[i], a, b);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) dmma(c[i], c[i + 1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i] + c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 1024 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    for (int mode = 0; mode < 3; ++mode) {
      float best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148, threads>>>(out, iters, 0.5);
        if (mode == 1) k<1><<<148, threads>>>(out, iters, 0.5);
        if (mode == 2) k<2><<<148, threads>>>(out, iters, 0.5);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      double warps = threads / 32.0;
      double dfma_instr = (mode != 1) ? 8.0 * iters * warps : 0;   // warp instr per SM
      double dmma_instr = (mode != 0) ? 4.0 * iters * warps : 0;
      printf("threads %4d mode %d (%s): %.3f ms; per SM: DFMA warp-instr/us %.1f DMMA warp-instr/us %.1f\n", threads, mode,
             mode == 0 ? "dfma" : mode == 1 ? "dmma" : "both", best, dfma_instr / (best * 1e3), dmma_instr / (best * 1e3));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
