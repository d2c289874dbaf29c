// Microbenchmark: cost of tcgen05.ld (TMEM -> registers, 32x32b shapes) on sm_100a, as used for per-thread constants.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm ldtm.cu && ./ldtm
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(192, 5) k(unsigned *out, long long *cyc, int iters) {
  __shared__ unsigned base_s;
  const int tau = threadIdx.x;
  if (tau < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&base_s)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tm = base_s + ((unsigned(tau >> 5) & 3u) << 21) + (unsigned(tau) >> 7) * 32u;
  for (int c = 0; c < 32; ++c) asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tm + c), "r"(unsigned(tau * 32 + c)) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  unsigned acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {   // x8 + x2, wait; x16 + x4 + x2, wait  (the k_fdr variant)
      unsigned r[32];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]) : "r"(tm));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[8]),"=r"(r[9]) : "r"(tm + 8));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 10; ++i) acc += r[i];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]) : "r"(tm + 10));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]) : "r"(tm + 26));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[20]),"=r"(r[21]) : "r"(tm + 30));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 22; ++i) acc += r[i];
    } else if (MODE == 1) {   // x16, wait, x16, wait
      unsigned r[16];
      for (int h = 0; h < 2; ++h) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]) : "r"(tm + 16 * h));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) acc += r[i];
      }
    } else if (MODE == 2) {   // one x1, wait
      unsigned r;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(tm + (it & 31)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r;
    } else if (MODE == 3) {   // x16 + x16, one wait
      unsigned r[32];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]) : "r"(tm));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]) : "r"(tm + 16));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 32; ++i) acc += r[i];
    } else if (MODE == 4) {   // the same 32 words from shared memory (LDS.128 x 8) for comparison
      extern __shared__ uint4 sm[];
      for (int h = 0; h < 8; ++h) { uint4 v = sm[(h * 192 + tau)]; acc += v.x + v.y + v.z + v.w; }
      asm volatile("" ::: "memory");
    } else if (MODE == 5) {   // 16x256b.x1: 4 registers per thread, one instruction
      unsigned r[4];
      asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]) : "r"(tm));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r[0] + r[1] + r[2] + r[3];
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * 192 + tau] = acc;
  if (tau == 0) cyc[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (tau < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base_s), "r"(64u) : "memory");
}
template <int MODE> void run(const char *name, int ctas_per_sm) {
  unsigned *out; long long *cyc;
  const int grid = 148 * ctas_per_sm, iters = 2000;
  cudaMalloc(&out, grid * 192 * 4); cudaMalloc(&cyc, grid * 8);
  const size_t sm = MODE == 4 ? 8 * 192 * 16 : 0;
  k<MODE><<<grid, 192, sm>>>(out, cyc, iters);
  k<MODE><<<grid, 192, sm>>>(out, cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148 * 5]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < grid; ++i) s += double(h[i]);
  printf("%-44s %d CTA/SM: %8.1f cycles per iteration per CTA (%s)\n", name, ctas_per_sm, s / grid / iters, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int c : {1, 5}) {
    run<0>("x8+x2 wait x16+x4+x2 wait (32 words)", c);
    run<1>("x16 wait x16 wait (32 words)", c);
    run<3>("x16 x16 wait (32 words)", c);
    run<2>("x1 wait (1 word)", c);
    run<5>("16x256b.x1 wait (4 words)", c);
    run<4>("8 x LDS.128 (32 words)", c);
  }
  return 0;
}
