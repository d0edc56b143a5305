// CPU check of csrc/fft240.cuh (the helpers are __host__ __device__): prints the max error of the 16 x 15 Cooley-Tukey
// 240-point transform against a direct O(N^2) DFT.  Built and run by tests/test_host_logic.py.
#include <cmath>
#include <cstdio>
#include <vector>
#define FFT240_HOST_TEST
#include "../../adaptive_optics_gym_b200/csrc/fft240.cuh"
int main() {
  const int N = 240;
  const double PI = 3.14159265358979323846;
  std::vector<double2> x(N), buf(N), tw(N), ref(N), out(N);
  unsigned long long s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0 - 0.5; };
  for (int i = 0; i < N; ++i) { x[i] = make_double2(rnd(), rnd()); tw[i] = make_double2(std::cos(2 * PI * i / N), std::sin(2 * PI * i / N)); }
  for (int k = 0; k < N; ++k) {
    double re = 0, im = 0;
    for (int n = 0; n < N; ++n) {
      const double a = 2 * PI * ((n * k) % N) / N;
      re += x[n].x * std::cos(a) - x[n].y * std::sin(a);
      im += x[n].x * std::sin(a) + x[n].y * std::cos(a);
    }
    ref[k] = make_double2(re, im);
  }
  buf = x;
  for (int n2 = 0; n2 < 15; ++n2) fft240::stage1(buf.data(), 1, n2, tw.data());
  for (int k1 = 0; k1 < 16; ++k1) {
    double2 a[15];
    fft240::stage2(buf.data(), 1, k1, a);
    for (int k2 = 0; k2 < 15; ++k2) out[k1 + 16 * k2] = a[k2];
  }
  double err = 0, mx = 0;
  for (int k = 0; k < N; ++k) {
    err = std::fmax(err, std::hypot(out[k].x - ref[k].x, out[k].y - ref[k].y));
    mx = std::fmax(mx, std::hypot(ref[k].x, ref[k].y));
  }
  std::printf("%.3e\n", err / mx);
  return err / mx < 1e-13 ? 0 : 1;
}
