// Feasibility probe for a two-pass (single-exchange) k_fdr: register-resident 40/48/49/64-point DFTs on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v -c probe.cu
#include "../../python-soxr_b200/csrc/fft_regs.cuh"
#include "unit_roots.h"
#include <utility>
using namespace b2;
namespace b2 {
// Cooley-Tukey composition in registers with compile-time twiddles: n = N2 n1 + n2, k = k1 + N1 k2
template <int N1, int N2> struct CtDft {
  static constexpr int N = N1 * N2;
  template <bool INV, int E, typename C> static __host__ __device__ __forceinline__ C tw(C x) {
    if constexpr (E == 0) return x;
    else {
      constexpr float wr = UnitRoots<N>::re[E];
      constexpr float wi = INV ? -UnitRoots<N>::im[E] : UnitRoots<N>::im[E];
      return cmul(x, C{wr, wi});
    }
  }
  template <bool INV, int K1, typename C, int... N2s>
  static __host__ __device__ __forceinline__ void col(C (&a)[N2][N1], C (&v)[N1 * N2], std::integer_sequence<int, N2s...>) {
    C b[N2];
    ((b[N2s] = tw<INV, (N2s * K1) % N>(a[N2s][K1])), ...);
    Dft<N2>::template run<INV>(b);
    ((v[K1 + N1 * N2s] = b[N2s]), ...);
  }
  template <bool INV, typename C, int... K1s>
  static __host__ __device__ __forceinline__ void cols(C (&a)[N2][N1], C (&v)[N1 * N2], std::integer_sequence<int, K1s...>) {
    (col<INV, K1s>(a, v, std::make_integer_sequence<int, N2>()), ...);
  }
  template <bool INV, typename C> static __host__ __device__ __forceinline__ void run(C (&v)[N1 * N2]) {
    C a[N2][N1];
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2) {
#pragma unroll
      for (int n1 = 0; n1 < N1; ++n1) a[n2][n1] = v[N2 * n1 + n2];
      Dft<N1>::template run<INV>(a[n2]);
    }
    cols<INV>(a, v, std::make_integer_sequence<int, N1>());
  }
};
template <> struct Dft<64> : CtDft<8, 8> {};
template <> struct Dft<49> : CtDft<7, 7> {};
template <> struct Dft<40> : PfaDft<8, 5> {};
template <> struct Dft<48> : PfaDft<16, 3> {};
}
template <int R, int STRIDE, int MAXREG>
__global__ void __launch_bounds__(64) __maxnreg__(MAXREG) k_probe(float2 *buf, const float2 *tw, int iters) {
  extern __shared__ float2 sm[];
  const int t = threadIdx.x;
  for (int i = t; i < R * STRIDE; i += 64) sm[i] = buf[i];
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    float2 v[R];
#pragma unroll
    for (int n = 0; n < R; ++n) v[n] = sm[STRIDE * n + t];
    Dft<R>::template run<false>(v);
    // twiddle chain: w^k from the powers of two of the thread's base root
    float2 w[R];
#pragma unroll
    for (int m = 1; m < R; m *= 2) w[m] = tw[t + 64 * m];
#pragma unroll
    for (int m = 3; m < R; ++m) {
      int hb = 1; while (hb * 2 <= m) hb *= 2;
      if (hb != m) w[m] = cmul(w[hb], w[m - hb]);
    }
#pragma unroll
    for (int m = 1; m < R; ++m) v[m] = cmul(v[m], w[m]);
    __syncthreads();
#pragma unroll
    for (int n = 0; n < R; ++n) sm[STRIDE * n + t] = v[n];
    __syncthreads();
  }
  for (int i = t; i < R * STRIDE; i += 64) buf[i] = sm[i];
}
#include <cstdio>
template <int R, int MAXREG> void bench(int ctas_per_sm) {
  const int grid = 148 * ctas_per_sm, iters = 400;
  float2 *buf, *tw;
  cudaMalloc(&buf, sizeof(float2) * R * 65); cudaMalloc(&tw, sizeof(float2) * 64 * 64);
  cudaMemset(buf, 0, sizeof(float2) * R * 65); cudaMemset(tw, 0, sizeof(float2) * 64 * 64);
  auto kern = k_probe<R, 65, MAXREG>;
  const size_t smem = sizeof(float2) * R * 65;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 64, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<grid, 64, smem>>>(buf, tw, iters);
  cudaEventRecord(e0);
  kern<<<grid, 64, smem>>>(buf, tw, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = ms * 1e-3 * 1.965e9 / iters / ctas_per_sm;   // SM cycles per (64 tasks of R points), amortised
  printf("R=%2d maxreg %3d: %d CTAs/SM asked (occupancy allows %d): %.0f SM-cycles per 64-task pass (%s)\n", R, MAXREG, ctas_per_sm, occ, cyc,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(buf); cudaFree(tw);
}
int main() {
  for (int c : {1, 2, 4, 6}) { bench<40, 128>(c); bench<48, 168>(c); bench<49, 168>(c); bench<64, 168>(c); }
  return 0;
}
