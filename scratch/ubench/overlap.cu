// Microbenchmark: can packed-FP32 work (FFMA2) and shared-memory traffic (LDS.64 / STS.64) overlap on sm_100a?
//   mode 0: FP only      mode 1: smem only      mode 2: both in every warp (interleaved)
//   mode 3: both, but even warps do FP and odd warps do smem (half the warps each, same total work as 0 + 1)
// Work per thread per iteration: NF packed FMAs on 16 independent accumulators, NS (LDS.64 + STS.64) pairs, conflict-free.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NF = 64, NS = 16;
__device__ __forceinline__ float2 lds2(const float2 *p) {
  float2 v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts2(float2 *p, float2 v) {
  unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y));
}
template <int MODE>
__global__ void __launch_bounds__(192, 5) k(float2 *out, float2 a, float2 b, int iters) {
  extern __shared__ float2 sm[];
  float2 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  float2 *p = sm + threadIdx.x;
  const bool do_fp = MODE == 0 || MODE == 2 || (MODE == 3 && !((threadIdx.x >> 5) & 1));
  const bool do_sm = MODE == 1 || MODE == 2 || (MODE == 3 && ((threadIdx.x >> 5) & 1));
  const int rep = MODE == 3 ? 2 : 1;   // half the warps do each kind: twice the iterations, same total work
  float2 acc = make_float2(0, 0);
  for (int it = 0; it < iters * rep; ++it) {
    if (do_sm) {
      float2 v[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) v[i] = lds2(p + i * 192);
#pragma unroll
      for (int i = 0; i < NS; ++i) { acc.x += v[i].x; }
      if (do_fp) {
#pragma unroll
        for (int r = 0; r < NF / 16; ++r)
#pragma unroll
          for (int i = 0; i < 16; ++i) x[i] = __ffma2_rn(x[i], a, b);
      }
#pragma unroll
      for (int i = 0; i < NS; ++i) sts2(p + i * 192, v[i]);
    } else if (do_fp) {
#pragma unroll
      for (int r = 0; r < NF / 16; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = __ffma2_rn(x[i], a, b);
    }
  }
  float2 s = acc;
#pragma unroll
  for (int i = 0; i < 16; ++i) { s.x += x[i].x; s.y += x[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> float run(const char *name) {
  float2 *out; cudaMalloc(&out, 148 * 5 * 192 * sizeof(float2));
  const size_t smem = NS * 192 * sizeof(float2);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000;
  k<MODE><<<148 * 5, 192, smem>>>(out, make_float2(1.0001f, 0.9999f), make_float2(1e-3f, 2e-3f), 16);
  cudaEventRecord(e0);
  k<MODE><<<148 * 5, 192, smem>>>(out, make_float2(1.0001f, 0.9999f), make_float2(1e-3f, 2e-3f), iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = ms * 1e-3 * 1.965e9;                       // per SM
  const double warps = 5 * 6, fp_instr = warps * iters * NF, sm_wavefronts = warps * iters * NS * 2 * 2;
  printf("%-34s %.3f ms  %.0f cycles/SM | FFMA2 per SMSP-cycle %.3f | smem wavefronts per SM-cycle %.3f\n", name, ms, cyc,
         MODE == 1 ? 0.0 : fp_instr / 4 / cyc, MODE == 0 ? 0.0 : sm_wavefronts / cyc);
  cudaFree(out);
  return ms;
}
int main() {
  const float a = run<0>("FP only (64 FFMA2 / iter)");
  const float b = run<1>("smem only (16 LDS.64 + 16 STS.64)");
  const float c = run<2>("both, interleaved in each warp");
  const float d = run<3>("both, split across warps");
  printf("sum %.3f  max %.3f  interleaved %.3f  split %.3f\n", a + b, a > b ? a : b, c, d);
  return 0;
}
