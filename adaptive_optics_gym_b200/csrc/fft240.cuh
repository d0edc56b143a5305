// 240-point complex DFT (sign +, unnormalised: X[k] = sum_n x[n] exp(+2 pi i n k / 240)) as 16 x 15 Cooley-Tukey
// with thread-local sub-transforms -- the fine scale of the von-Karman screen synthesis is a 240 x 240 inverse DFT
// (AO_env.py:76-77 -> hcipy FiniteAtmosphericLayer; tables.py: W1 is a shifted DFT matrix), which the GEMM form pays
// 41.5 M FP64 FMAs per env for and this form 4.9 M.
//   n = 15 n1 + n2, k = k1 + 16 k2:
//   X[k1 + 16 k2] = sum_n2 w15^(n2 k2) [ w240^(n2 k1) sum_n1 w16^(n1 k1) x[15 n1 + n2] ]
// Stage 1: 15 DFT-16 (radix 2, registers) + twiddle, in place on slots 15 k1 + n2;  stage 2: 16 DFT-15 (Good-Thomas
// 3 x 5, no twiddles) on the contiguous slots 15 k1 .. 15 k1 + 14.  With FFT240_HOST_TEST the helpers are
// __host__ __device__: the index maps and constants are unit-tested on the CPU (tests/test_host_logic.py builds
// tools/micro/fft240_host.cu).
#pragma once
#include <cuda_runtime.h>
#ifdef FFT240_HOST_TEST          // tools/micro/fft240_host.cu: the same code on the CPU
#define FFT240_FN __host__ __device__ __forceinline__
#else
#define FFT240_FN __device__ __forceinline__
#endif

namespace fft240 {
FFT240_FN double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
FFT240_FN double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
FFT240_FN double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
FFT240_FN double2 cfma(double2 a, double wr, double wi, double2 acc) {   // acc + a (wr + i wi)
  return make_double2(acc.x + a.x * wr - a.y * wi, acc.y + a.x * wi + a.y * wr);
}

// in-place DFT-16, natural order in and out
FFT240_FN void dft16(double2 (&a)[16]) {
  constexpr double C1 = 0.92387953251128674, S1 = 0.38268343236508977, R = 0.70710678118654752;
  constexpr double WR[8] = {1.0, C1, R, S1, 0.0, -S1, -R, -C1};
  constexpr double WI[8] = {0.0, S1, R, C1, 1.0, C1, R, S1};
  constexpr int REV[16] = {0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15};
  double2 t[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) t[i] = a[REV[i]];
#pragma unroll
  for (int m = 2; m <= 16; m <<= 1) {
#pragma unroll
    for (int k = 0; k < 16; k += m) {
#pragma unroll
      for (int j = 0; j < m / 2; ++j) {
        const int w = j * (16 / m);
        const double2 u = t[k + j], v = t[k + j + m / 2];
        const double2 vt = make_double2(v.x * WR[w] - v.y * WI[w], v.x * WI[w] + v.y * WR[w]);
        t[k + j] = cadd(u, vt);
        t[k + j + m / 2] = csub(u, vt);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = t[i];
}

// in-place DFT-15 (Good-Thomas: n = (5 n1 + 3 n2) mod 15, k = (10 k1 + 6 k2) mod 15)
FFT240_FN void dft15(double2 (&a)[15]) {
  constexpr double W5R[5] = {1.0, 0.30901699437494742, -0.80901699437494742, -0.80901699437494742, 0.30901699437494742};
  constexpr double W5I[5] = {0.0, 0.95105651629515357, 0.58778525229247313, -0.58778525229247313, -0.95105651629515357};
  constexpr double W3R[3] = {1.0, -0.5, -0.5};
  constexpr double W3I[3] = {0.0, 0.86602540378443865, -0.86602540378443865};
  double2 t[3][5];
#pragma unroll
  for (int n1 = 0; n1 < 3; ++n1)
#pragma unroll
    for (int k2 = 0; k2 < 5; ++k2) {
      double2 s = a[(5 * n1) % 15];
#pragma unroll
      for (int n2 = 1; n2 < 5; ++n2) s = cfma(a[(5 * n1 + 3 * n2) % 15], W5R[(n2 * k2) % 5], W5I[(n2 * k2) % 5], s);
      t[n1][k2] = s;
    }
#pragma unroll
  for (int k1 = 0; k1 < 3; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 5; ++k2) {
      double2 s = t[0][k2];
#pragma unroll
      for (int n1 = 1; n1 < 3; ++n1) s = cfma(t[n1][k2], W3R[(n1 * k1) % 3], W3I[(n1 * k1) % 3], s);
      a[(10 * k1 + 6 * k2) % 15] = s;
    }
}

// the two stages on a 240-slot buffer (element i at buf[i * stride]); tw[j] = exp(2 pi i j / 240), j < 240
FFT240_FN void stage1(double2* buf, int stride, int n2, const double2* tw) {
  double2 a[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) a[n1] = buf[(15 * n1 + n2) * stride];
  dft16(a);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) buf[(15 * k1 + n2) * stride] = cmul(a[k1], tw[n2 * k1]);
}
// returns the 15 outputs X[k1 + 16 k2], k2 = 0 .. 14, in a[k2]
FFT240_FN void stage2(const double2* buf, int stride, int k1, double2 (&a)[15]) {
#pragma unroll
  for (int n2 = 0; n2 < 15; ++n2) a[n2] = buf[(15 * k1 + n2) * stride];
  dft15(a);
}
}  // namespace fft240
