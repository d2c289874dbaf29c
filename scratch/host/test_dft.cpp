// host unit test of the register DFTs against the definition
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "../../python-soxr_b200/csrc/fft_regs.cuh"
using namespace b2;
template <int N, bool INV> double check() {
  double2 v[N], x[N];
  for (int i = 0; i < N; ++i) { x[i] = v[i] = make_double2(sin(1.0 + 3.7 * i), cos(0.3 + 2.1 * i * i)); }
  Dft<N>::template run<INV>(v);
  double err = 0;
  for (int k = 0; k < N; ++k) {
    double re = 0, im = 0;
    for (int n = 0; n < N; ++n) {
      double a = (INV ? 2 : -2) * M_PI * n * k / N;
      re += x[n].x * cos(a) - x[n].y * sin(a);
      im += x[n].x * sin(a) + x[n].y * cos(a);
    }
    err = fmax(err, fmax(fabs(re - v[k].x), fabs(im - v[k].y)));
  }
  return err;
}
#define T(N) printf("N=%2d fwd %.2e inv %.2e\n", N, check<N, false>(), check<N, true>());
int main() { T(2) T(3) T(4) T(5) T(6) T(7) T(8) T(10) T(12) T(14) T(15) T(16) T(20) T(21) T(28) }
