// Microbenchmark (tuning aid): FP64 throughput of scalar DFMA against mma.sync.m8n8k4.f64 (DMMA) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu && ./dmma_bench
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>   // 0: 32 independent DFMA chains per thread   1: 8 independent DMMA accumulators per warp
__global__ void k(double* out, int iters, double x, double y) {
  double acc[32];
  for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fma(acc[i], x, y);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                     : "+d"(acc[2 * i]), "+d"(acc[2 * i + 1]) : "d"(x), "d"(y));
    }
  }
  double s = 0;
  for (int i = 0; i < 32; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int warps, double flop_per_thread_iter) {
  double* out; cudaMalloc(&out, 148 * 1024 * sizeof(double));
  const int iters = 4000, threads = warps * 32;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, threads>>>(out, 10, 1.0000001, 0.5);
  cudaEventRecord(e0); k<MODE><<<148, threads>>>(out, iters, 1.0000001, 0.5); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-8s warps/SM %2d  %.3f ms  %.2f TFLOP/s\n", name, warps, ms, flop_per_thread_iter * threads * 148.0 * iters / (ms * 1e-3) / 1e12);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8, 16, 32}) {
    run<0>("DFMA", w, 64.0);                 // 32 FMAs
    run<1>("DMMA", w, 8 * 2.0 * 8 * 8 * 4 / 32.0);   // 8 mma of 8x8x4 per warp
  }
  return 0;
}
