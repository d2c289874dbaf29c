// Host check of the two-pass block transform of fdr2_stage.cuh against a direct evaluation in double.
// nvcc -std=c++17 -O1 -o host_test host_test.cu && ./host_test
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "fdr2_stage.cuh"
using namespace b2;
int main() {
  typedef Fdr2 F;
  std::vector<float2> x(F::NIN), E(F::E_ELEMS), SP(F::NIN), y(F::NOUT);
  std::vector<float> H(F::NOUT);
  std::vector<std::vector<float2>> vv(48);
  srand(1);
  for (auto &v : x) v = make_float2(rand() / float(RAND_MAX) - .5f, rand() / float(RAND_MAX) - .5f);
  for (auto &h : H) h = rand() / float(RAND_MAX);
  const double PI = 3.14159265358979323846;
  for (int g = 0; g < 64; ++g) {
    float2 wb[6];
    for (int i = 0; i < 6; ++i) { const double a = -2 * PI * double((long long)g * (1 << i) % 2560) / 2560; wb[i] = make_float2(float(cos(a)), float(sin(a))); }
    F::fwd_a(g, E.data(), [&](int n) { return x[n]; }, wb);
  }
  for (int j = 0; j < 40; ++j) F::fwd_b(j, E.data());
  for (int k = 0; k < 2560; ++k) SP[k] = E[F::sp_at(k)];
  // check the forward spectrum
  double worst = 0;
  for (int k = 0; k < 2560; k += 37) {
    double re = 0, im = 0;
    for (int n = 0; n < 2560; ++n) { const double a = -2 * PI * double((long long)n * k % 2560) / 2560; re += x[n].x * cos(a) - x[n].y * sin(a); im += x[n].x * sin(a) + x[n].y * cos(a); }
    worst = fmax(worst, hypot(re - SP[k].x, im - SP[k].y));
  }
  printf("forward: worst |X - SP| over sampled bins %.3e\n", worst);
  for (int t = 0; t < 48; ++t) {
    float2 wb[6];
    for (int i = 0; i < 6; ++i) { const double a = -2 * PI * double((long long)t * (1 << i) % 2352) / 2352; wb[i] = make_float2(float(cos(a)), float(sin(a))); }
    vv[t].resize(49); float2 (&v)[49] = *reinterpret_cast<float2 (*)[49]>(vv[t].data());
    F::inv_1_compute(t, E.data(), [&](int q) { return H[t + 48 * q]; }, wb, v);
  }
  for (int t = 0; t < 48; ++t) F::inv_1_store(t, E.data(), *reinterpret_cast<float2 (*)[49]>(vv[t].data()));
  for (int m = 0; m < 49; ++m) F::inv_2(m, E.data(), [&](int o, float2 v) { y[o] = v; });
  worst = 0;
  double scale = 0;
  for (int o = 0; o < 2352; o += 29) {
    double re = 0, im = 0;
    for (int Fq = 0; Fq < 2352; ++Fq) {
      const int k = Fq < 1176 ? Fq : Fq + 208;
      const double a = 2 * PI * double((long long)o * Fq % 2352) / 2352;
      const double xr = H[Fq] * SP[k].x, xi = H[Fq] * SP[k].y;
      re += xr * cos(a) - xi * sin(a); im += xr * sin(a) + xi * cos(a);
    }
    worst = fmax(worst, hypot(re - y[o].x, im - y[o].y));
    scale = fmax(scale, hypot(re, im));
  }
  printf("inverse: worst |y - direct| over sampled outputs %.3e (values up to %.3e)\n", worst, scale);
  return 0;
}
