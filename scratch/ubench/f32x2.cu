// Microbenchmark: FFMA vs FFMA2 (packed f32x2) issue/throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float2 *out, float2 a, float2 b, int iters) {
  float2 x[8];
  for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { x[i].x = fmaf(x[i].x, a.x, b.x); x[i].y = fmaf(x[i].y, a.y, b.y); }
      else if (MODE == 1) x[i] = __ffma2_rn(x[i], a, b);
      else if (MODE == 2) { x[i].x = x[i].x + a.x; x[i].y = x[i].y + a.y; }
      else if (MODE == 3) x[i] = __fadd2_rn(x[i], a);
      else if (MODE == 4) { float2 bb = make_float2(b.x, b.x); x[i] = __ffma2_rn(x[i], a, bb); a.x += 1e-9f; }   // broadcast operand
      else if (MODE == 5) {  // complex multiply by (a.x + i a.y) with packed ops: x*a.x + (−x.y, x.x)*a.y
        float2 sw = make_float2(-x[i].y, x[i].x);
        x[i] = __ffma2_rn(sw, make_float2(a.y, a.y), __fmul2_rn(x[i], make_float2(a.x, a.x)));
      } else if (MODE == 6) {  // scalar complex multiply
        float2 v = x[i]; x[i].x = v.x * a.x - v.y * a.y; x[i].y = v.x * a.y + v.y * a.x;
      }
    }
  }
  float2 s = make_float2(0, 0);
  for (int i = 0; i < 8; ++i) { s.x += x[i].x; s.y += x[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, double flops_per_iter_thread) {
  float2 *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096;
  k<MODE><<<148 * 8, 256>>>(out, make_float2(1.0001f, 0.9999f), make_float2(1e-3f, 2e-3f), 16);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(out, make_float2(1.0001f, 0.9999f), make_float2(1e-3f, 2e-3f), iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double lane_ops = 148.0 * 8 * 256 * iters * flops_per_iter_thread;
  printf("%-28s %.3f ms  %.1f T scalar-ops/s  (%.2f per SM per clk @1.965GHz)\n", name, ms, lane_ops / ms / 1e9,
         lane_ops / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}
int main() {
  run<0>("FFMA scalar x2", 16); run<1>("FFMA2", 16); run<2>("FADD scalar x2", 16); run<3>("FADD2", 16);
  run<4>("FFMA2 + broadcast pack", 16); run<5>("cmul packed (FMUL2+FFMA2)", 32); run<6>("cmul scalar (2 FMUL+2 FFMA)", 32);
  return 0;
}
