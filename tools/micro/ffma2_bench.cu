// Microbenchmark (tuning aid): issue / pipe throughput of scalar FFMA against packed fma.rn.f32x2 on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
  float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}
template <int MODE>   // 0: 16 FFMA   1: 8 FFMA2   2: 8 FFMA + 4 FFMA2   3: 16 FFMA + 8 FFMA2   4: 16 FFMA + 8 LOP3   5: 8 FFMA2 + 8 LOP3
__global__ void k(float* out, int iters, float x, float y) {
  float a[16]; uint64_t q[8]; int z[8];
  for (int i = 0; i < 8; ++i) z[i] = threadIdx.x + i;
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
  for (int i = 0; i < 8; ++i) q[i] = ((uint64_t)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i + 8]);
  const uint64_t X = ((uint64_t)__float_as_uint(x) << 32) | __float_as_uint(x), Y = ((uint64_t)__float_as_uint(y) << 32) | __float_as_uint(y);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = ffma1(a[i], x, y);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = ffma2(q[i], X, Y);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[2 * i] = ffma1(a[2 * i], x, y); q[i] = ffma2(q[i], X, Y); a[2 * i + 1] = ffma1(a[2 * i + 1], x, y); }
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[2 * i] = ffma1(a[2 * i], x, y); q[i] = ffma2(q[i], X, Y); a[2 * i + 1] = ffma1(a[2 * i + 1], x, y); }
    } else if (MODE == 4) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[2 * i] = ffma1(a[2 * i], x, y); a[2 * i + 1] = ffma1(a[2 * i + 1], x, y);
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"(it), "r"(iters));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        q[i] = ffma2(q[i], X, Y);
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"(it), "r"(iters));
      }
    }
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  for (int i = 0; i < 8; ++i) s += (float)z[i];
  for (int i = 0; i < 8; ++i) s += __uint_as_float((uint32_t)q[i]) + __uint_as_float((uint32_t)(q[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int warps_per_sm, double fma_per_iter) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4 * sizeof(float));
  const int iters = 20000, threads = warps_per_sm * 32;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, threads>>>(out, 100, 1.0001f, 0.5f);
  cudaEventRecord(e0); k<MODE><<<148, threads>>>(out, iters, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-26s warps/SM %2d  %.3f ms  FMA lanes/clk/SM %.1f  (cycles/iter/SM-quadrant-warp %.2f)\n", name, warps_per_sm, ms,
         fma_per_iter * threads * iters / cycles, cycles / iters / (warps_per_sm / 4.0));
  cudaFree(out);
}
int main() {
  for (int w : {4, 16, 32}) {
    run<0>("16 FFMA", w, 16); run<1>("8 FFMA2", w, 16); run<2>("8 FFMA + 4 FFMA2", w, 16); run<3>("16 FFMA + 8 FFMA2", w, 32);
    run<4>("16 FFMA + 8 LOP3", w, 16); run<5>("8 FFMA2 + 8 LOP3", w, 16);
  }
  return 0;
}
