// fdr_stage.cuh -- single-pass frequency-domain rational resampler ("fdr").
//
// For a plan whose stages are all rational (cr.c:264-496: DFT stage L1/M1 followed by the
// vpoly0 stage L2/M2, poly-fir0.h:24-44) the cascade is ONE zero-phase linear operator
//
//     out[n] = sum_k c[Mt n - Lt k] in[k],   c = (g1 upsampled by L2) * g2,  Lt = L1 L2, Mt = M1 M2
//
// (SURVEY.md 8a-R0).  With lr/mr = Lt/Mt reduced and block sizes NIN = mr a, NOUT = lr a, a block
// of the operator is evaluated as  IFFT_NOUT( H .* FFT_NIN(in)[remapped bins] ):  one forward
// transform, one REAL spectrum product (c is symmetric, the block grid keeps its centre on a
// sample of both rates) and one inverse transform -- instead of the reference's FFT + spectrum
// replication + larger inverse FFT (cr.c:87-146) followed by a 20-tap polyphase dot product per
// output (poly-fir0.h:24-28).  The terms it drops are the aliases a decimating cascade folds back
// from its own stop band (below -att dB); it is used for whole-clip float32 arithmetic with
// HQ/VHQ-class filters (att >= 132 dB: 2.8e-7 / 1.7e-7 RMS from the reference measured, budget
// 1e-6) and for single 1/2 or 2/1 DFT stages at every quality, where the reference computes the
// same thing (fdr_prepare in fdr_stage.cu has the rule).  Blocks overlap by a margin (OVL/2 grid
// steps of mr input / lr output samples on each side) that covers the composite filter's
// half-width; the first outputs of a clip, where the reference's intermediate signal does not
// exist before index 0 (cr.c:487-490), come from the head kernels (fdr_stage.cu).  One FdrCfg per
// ratio family (160:147, 147:160, 2:1, 1:2, 3:1, 1:3, 3:2, 2:3, 441:160, 160:441, 6:1, 1:6, 4:1, 1:4, 441:320, 320:441,
// 147:80, 80:147, 320:147, 147:320).
//
// FFTs: N = RA*RB*RC, three register-radix passes (decimation in time forward, the mirrored
// decimation in frequency backward) with shared-memory exchanges whose strides are chosen so
// that every access is bank-conflict free (scratch/fdr_bank_sim.py); radices 12/14/20 are
// prime-factor compositions of 3/4/5/7 (fft_regs.cuh) -- PFFFT's radix-3/5 passes are compiled
// out in libsoxr (pffft.c:320,424), its block sizes being powers of two.
#pragma once
#include <type_traits>

#include "fft_regs.cuh"
#include "fft_stage.cuh"

namespace b2 {

enum { kP2JFast = 0, kP2Row16 = 1 };

// Twiddle set w[m] = w^m, m = 1 .. R-1: the powers of two come from the table (row m at tw[m * stride],
// correctly rounded), every other one is ONE product  w[m] = w[hi] * w[m - hi]  (hi = top bit of m) of
// earlier entries, so its error is a sum of popcount(m) roundings and not m times that of w^1.
// ceil(log2 R) loads on the L1 data pipe + the rest on the FMA pipe replace R-1 loads: the L1 pipe is
// the busier of the two in k_fdr (profiles/r1j: 68 % vs 40 %).
__host__ __device__ constexpr int fdr_top_bit(int m) { return m >= 16 ? 16 : m >= 8 ? 8 : m >= 4 ? 4 : m >= 2 ? 2 : 1; }   // m < 32
#ifndef B2_FDR_TW_SPARSE
#define B2_FDR_TW_SPARSE 0
#endif
__host__ __device__ constexpr int fdr_np2(int R) { return R > 16 ? 5 : R > 8 ? 4 : R > 4 ? 3 : R > 2 ? 2 : 1; }   // powers of two in [1, R)
// the members of the set that are not powers of two, from the ones that are
template <int R, typename C>
__host__ __device__ __forceinline__ void twiddle_products(C (&w)[R]) {
#pragma unroll
  for (int m = 3; m < R; ++m)
    if (fdr_top_bit(m) != m) w[m] = cmul(w[fdr_top_bit(m)], w[m - fdr_top_bit(m)]);
}
template <int R, typename C>
__host__ __device__ __forceinline__ void twiddle_set(const C *__restrict__ tw, int stride, C (&w)[R]) {
  static_assert(R <= 32, "fdr_top_bit covers m < 32");
#if B2_FDR_TW_SPARSE
  // only w^1, w^4, w^16 come from the table; w^2 = w^1 * w^1 and w^8 = w^4 * w^4 are one product of table entries
  // like every other member of the set (the L1 data pipe is the busiest unit of k_fdr: 75 %)
#pragma unroll
  for (int m = 1; m < R; m *= 4) w[m] = tw[m * stride];
#pragma unroll
  for (int m = 2; m < R; m *= 4) w[m] = cmul(w[m / 2], w[m / 2]);
#else
#pragma unroll
  for (int m = 1; m < R; ++m)
    if (fdr_top_bit(m) == m) w[m] = tw[m * stride];
#endif
#pragma unroll
  for (int m = 3; m < R; ++m)
    if (fdr_top_bit(m) != m) w[m] = cmul(w[fdr_top_bit(m)], w[m - fdr_top_bit(m)]);
}

// Three-pass FFT of N = RA*RB*RC points on a thread group.
//   E1[S1*j + g]      g = RC*m' + m in [0, RB*RC), j in [0, RA)      between passes 1 and 2
//   E2[S2*m + t2]     t2 = j + RA*q2 in [0, RA*RB), m in [0, RC)      between passes 2 and 3
// Pass-2 threads: MODE kP2JFast: tau = j + RA*m;  kP2Row16: j = tau >> 4, m = tau & 15 (RC <= 16).
template <int RA_, int RB_, int RC_, int S1_, int S2_, int MODE_>
struct Fft3 {
  static constexpr int RA = RA_, RB = RB_, RC = RC_, S1 = S1_, S2 = S2_, MODE = MODE_;
  static constexpr int N = RA * RB * RC, G1 = RB * RC, T3 = RA * RB;
  static constexpr int T2 = MODE == kP2Row16 ? 16 * RA : RA * RC;
  static constexpr int E1_ELEMS = S1 * RA, E2_ELEMS = S2 * RC;
  static constexpr int TW2_ELEMS = RB * RA, TW3_ELEMS = RC * T3;

  static __host__ __device__ __forceinline__ bool p2_map(int tau, int &j, int &m) {
    if (MODE == kP2Row16) {
      j = tau >> 4;
      m = tau & 15;
      return j < RA && m < RC;
    }
    j = tau % RA;
    m = tau / RA;
    return m < RC;
  }

  // ---- forward (DIT): x natural -> X natural ---------------------------------------------
  // pass 1: Z[j; g] = sum_n w_RA^(n j) x[g + G1 n];  ld(g, n) returns x[g + G1 n] (n is a compile-time
  // constant after unrolling, so callers can fold G1 * n into an immediate offset)
  // The loads and the butterflies are separate steps so that a kernel can issue the loads of its NEXT block
  // early (software pipelining across blocks) and keep them in registers.
  template <int TP> struct P1Regs { static constexpr int NU = (G1 + TP - 1) / TP; };
  template <int TP, typename C, typename LD>
  static __host__ __device__ __forceinline__ void fwd_p1_load(int tau, C (&a)[(G1 + TP - 1) / TP][RA], LD ld) {
    constexpr int NU = (G1 + TP - 1) / TP;
    if (tau >= TP) return;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int g = tau + TP * u;
      if (NU * TP == G1 || g < G1) {
#pragma unroll
        for (int n = 0; n < RA; ++n) a[u][n] = ld(g, n);
      }
    }
  }
  template <int TP, typename C>
  static __host__ __device__ __forceinline__ void fwd_p1_run(int tau, C *__restrict__ E1, C (&a)[(G1 + TP - 1) / TP][RA]) {
    constexpr int NU = (G1 + TP - 1) / TP;
    if (tau >= TP) return;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int g = tau + TP * u;
      if (NU * TP == G1 || g < G1) {
        Dft<RA>::template run<false>(a[u]);
#pragma unroll
        for (int j = 0; j < RA; ++j) E1[S1 * j + g] = a[u][j];
      }
    }
  }
  template <int TP, typename C, typename LD>
  static __host__ __device__ __forceinline__ void fwd_p1(int tau, C *__restrict__ E1, LD ld) {
    C a[(G1 + TP - 1) / TP][RA];
    fwd_p1_load<TP>(tau, a, ld);
    fwd_p1_run<TP>(tau, E1, a);
  }
  // pass 2: Y_m[j + RA q2] = sum_m' w_RB^(m' q2) w_(RA RB)^(j m') Z[j; RC m' + m]
  // Twiddles are fetched into registers BEFORE the barrier that precedes the pass that uses them (the
  // registers of the pass just finished are free there), so their latency overlaps the barrier wait.
  template <typename C>
  static __host__ __device__ __forceinline__ void ld_tw2(int tau, const C *__restrict__ tw2, C (&w)[RB]) {
    int j, m;
    if (!p2_map(tau, j, m)) j = 0;   // idle threads load (and ignore) valid entries: no control flow around w
    twiddle_set<RB>(tw2 + j, RA, w);
  }
  template <typename C>
  static __host__ __device__ __forceinline__ void ld_tw3(int tau, const C *__restrict__ tw3, C (&w)[RC]) {
    const int t = tau < T3 ? tau : T3 - 1;
    twiddle_set<RC>(tw3 + t, T3, w);
  }
  template <typename R>
  static __host__ __device__ __forceinline__ void ld_h(int tau, const R *__restrict__ H, R (&h)[RC]) {
    const int t = tau < T3 ? tau : T3 - 1;
#pragma unroll
    for (int q = 0; q < RC; ++q) h[q] = H[t + T3 * q];
  }

  // The same per-thread constants as 32-bit words (the form they are kept in tensor memory in, see tmem_* below):
  // the NP2C power-of-two twiddles of the third pass, and the thread's RC response bins.
  static constexpr int NP2C = fdr_np2(RC);
  static __device__ __forceinline__ void tw3_words(int tau, const float2 *__restrict__ tw3, unsigned (&r)[2 * NP2C]) {
    const int t = tau < T3 ? tau : T3 - 1;
#pragma unroll
    for (int i = 0; i < NP2C; ++i) {
      const float2 v = tw3[t + T3 * (1 << i)];
      r[2 * i] = __float_as_uint(v.x);
      r[2 * i + 1] = __float_as_uint(v.y);
    }
  }
  static __device__ __forceinline__ void tw3_from_words(const unsigned *r, float2 (&w)[RC]) {
#pragma unroll
    for (int i = 0; i < NP2C; ++i) w[1 << i] = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
    twiddle_products<RC>(w);
  }
  static __device__ __forceinline__ void h_words(int tau, const float *__restrict__ H, unsigned (&r)[RC]) {
    const int t = tau < T3 ? tau : T3 - 1;
#pragma unroll
    for (int q = 0; q < RC; ++q) r[q] = __float_as_uint(H[t + T3 * q]);
  }

  template <typename C>
  static __host__ __device__ __forceinline__ void fwd_p2(int tau, const C *__restrict__ E1, C *__restrict__ E2,
                                                         const C (&w)[RB]) {
    int j, m;
    if (!p2_map(tau, j, m)) return;
    C v[RB];
    const C *__restrict__ src = E1 + S1 * j + m;
#pragma unroll
    for (int mp = 0; mp < RB; ++mp) v[mp] = src[RC * mp];
#pragma unroll
    for (int mp = 1; mp < RB; ++mp) v[mp] = cmul(v[mp], w[mp]);
    Dft<RB>::template run<false>(v);
    C *__restrict__ dst = E2 + S2 * m + j;
#pragma unroll
    for (int q2 = 0; q2 < RB; ++q2) dst[RA * q2] = v[q2];
  }
  // pass 3: X[t + T3 q] = sum_m w_RC^(m q) w_N^(t m) Y_m[t]
  template <typename C>
  static __host__ __device__ __forceinline__ void fwd_p3(int tau, const C *__restrict__ E2, C *__restrict__ SP,
                                                         const C (&w)[RC]) {
    if (tau >= T3) return;
    C v[RC];
#pragma unroll
    for (int m = 0; m < RC; ++m) v[m] = E2[S2 * m + tau];
#pragma unroll
    for (int m = 1; m < RC; ++m) v[m] = cmul(v[m], w[m]);
    Dft<RC>::template run<false>(v);
#pragma unroll
    for (int q = 0; q < RC; ++q) SP[tau + T3 * q] = v[q];
  }

  // ---- inverse (mirrored DIF, unnormalised): Y natural -> y natural ----------------------------
  // pass 1: reads the forward spectrum of an NIN-point block at the signed frequency of bin F,
  // times the real filter response H[F].
  template <int NIN, typename C, typename R>
  static __host__ __device__ __forceinline__ void inv_p1(int tau, const C *__restrict__ SP, const R (&h)[RC],
                                                         C *__restrict__ E2, const C (&w)[RC]) {
    if (tau >= T3) return;
    C v[RC];
#pragma unroll
    for (int q = 0; q < RC; ++q) {
      const int F = tau + T3 * q;
      const bool pos = ((N / 2) % T3 == 0) ? (q < (N / 2) / T3) : (F < (N + 1) / 2);
      // NIN < N (upsampling): the map reads the periodic images of the input spectrum, exactly
      int k = pos ? F : F + (NIN - N);
      if (N > 2 * NIN) k = pos ? F % NIN : NIN - 1 - (N - 1 - F) % NIN;
      v[q] = scale(h[q], SP[k]);
    }
    Dft<RC>::template run<true>(v);
#pragma unroll
    for (int m = 1; m < RC; ++m) v[m] = cmulc(v[m], w[m]);
#pragma unroll
    for (int m = 0; m < RC; ++m) E2[S2 * m + tau] = v[m];
  }
  template <typename C>
  static __host__ __device__ __forceinline__ void inv_p2(int tau, const C *__restrict__ E2, C *__restrict__ E1,
                                                         const C (&w)[RB]) {
    int j, m;
    if (!p2_map(tau, j, m)) return;
    C v[RB];
    const C *__restrict__ src = E2 + S2 * m + j;
#pragma unroll
    for (int q2 = 0; q2 < RB; ++q2) v[q2] = src[RA * q2];
    Dft<RB>::template run<true>(v);
#pragma unroll
    for (int mp = 1; mp < RB; ++mp) v[mp] = cmulc(v[mp], w[mp]);
    C *__restrict__ dst = E1 + S1 * j + m;
#pragma unroll
    for (int mp = 0; mp < RB; ++mp) dst[RC * mp] = v[mp];
  }
  template <int TP, typename C, typename ST>
  static __host__ __device__ __forceinline__ void inv_p3(int tau, const C *__restrict__ E1, ST st) {
    constexpr int NU = (G1 + TP - 1) / TP;
    if (tau >= TP) return;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int g = tau + TP * u;
      if (NU * TP == G1 || g < G1) {
        C a[RA];
#pragma unroll
        for (int j = 0; j < RA; ++j) a[j] = E1[S1 * j + g];
        Dft<RA>::template run<true>(a);
#pragma unroll
        for (int n = 0; n < RA; ++n) st(g, n, a[n]);
      }
    }
  }
};


#ifdef __CUDACC__
// ---- tensor memory (TMEM) as a per-thread constant store ---------------------------------------
// The per-thread constants of a block (the power-of-two twiddles of the two third passes and the
// thread's RC filter-response bins) are the same for every block a CTA walks.  Registers cannot
// hold them (64 per thread at 5 CTAs/SM), and re-reading them from L1 costs 18 % of the wavefronts
// of the L1 data pipe, the busiest unit of k_fdr (profiles/r2o_ncu_summary.txt: 75 %).  Blackwell's
// tensor memory (256 kB per SM, 128 lanes x 512 columns x 32 bit, read with tcgen05.ld over its own
// datapath) holds them instead: lane = 32 (warp % 4) + lane id, 32-column slot (warp / 4), i.e.
// a private array of 32 words per thread, written once per CTA (tcgen05.st) and read back per
// block (tcgen05.ld.32x32b: one word per lane and column).  No tensor-core instruction is involved.
#ifndef B2_FDR_TMEM
#define B2_FDR_TMEM 1
#endif
__device__ __forceinline__ void tmem_alloc(unsigned *dst_smem, unsigned ncols) {   // one whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols) {     // one whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <int N> struct TmemIo;
template <> struct TmemIo<1> {
  static __device__ __forceinline__ void ld(unsigned a, unsigned *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(a));
  }
  static __device__ __forceinline__ void st(unsigned a, const unsigned *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(a), "r"(r[0]) : "memory");
  }
};
template <> struct TmemIo<2> {
  static __device__ __forceinline__ void ld(unsigned a, unsigned *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
  }
  static __device__ __forceinline__ void st(unsigned a, const unsigned *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(a), "r"(r[0]), "r"(r[1]) : "memory");
  }
};
template <> struct TmemIo<4> {
  static __device__ __forceinline__ void ld(unsigned a, unsigned *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(a));
  }
  static __device__ __forceinline__ void st(unsigned a, const unsigned *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3])
                 : "memory");
  }
};
template <> struct TmemIo<8> {
  static __device__ __forceinline__ void ld(unsigned a, unsigned *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(a));
  }
  static __device__ __forceinline__ void st(unsigned a, const unsigned *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
  }
};
template <> struct TmemIo<16> {
  static __device__ __forceinline__ void ld(unsigned a, unsigned *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(a));
  }
  static __device__ __forceinline__ void st(unsigned a, const unsigned *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(a),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
  }
};
// N words starting at column address a, in power-of-two pieces (largest first)
template <int N>
__device__ __forceinline__ void tmem_ld(unsigned a, unsigned (&r)[N]) {
  static_assert(N < 32, "pieces of 16, 8, 4, 2, 1 words");
  int done = 0;
#pragma unroll
  for (int p = 16; p >= 1; p >>= 1)
    if ((N - done) >= p) {   // N < 32: every piece size occurs at most once
      if (p == 16) TmemIo<16>::ld(a + done, r + done);
      if (p == 8) TmemIo<8>::ld(a + done, r + done);
      if (p == 4) TmemIo<4>::ld(a + done, r + done);
      if (p == 2) TmemIo<2>::ld(a + done, r + done);
      if (p == 1) TmemIo<1>::ld(a + done, r + done);
      done += p;
    }
}
template <int N>
__device__ __forceinline__ void tmem_st(unsigned a, const unsigned (&r)[N]) {
  static_assert(N < 32, "pieces of 16, 8, 4, 2, 1 words");
  int done = 0;
#pragma unroll
  for (int p = 16; p >= 1; p >>= 1)
    if ((N - done) >= p) {
      if (p == 16) TmemIo<16>::st(a + done, r + done);
      if (p == 8) TmemIo<8>::st(a + done, r + done);
      if (p == 4) TmemIo<4>::st(a + done, r + done);
      if (p == 2) TmemIo<2>::st(a + done, r + done);
      if (p == 1) TmemIo<1>::st(a + done, r + done);
      done += p;
    }
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
#endif  // __CUDACC__

constexpr int fdr_cmax(int a, int b) { return a > b ? a : b; }
// threads (a multiple of 32) taking part in a pass over G items done in ceil(G / threads) rounds
#ifndef B2_FDR_TP_GREEDY
#define B2_FDR_TP_GREEDY 0
#endif
constexpr int fdr_tp(int G, int threads) {
  if (B2_FDR_TP_GREEDY && G > threads) return threads;   // every warp in the first round, the remainder in a short second one
  const int rounds = (G + threads - 1) / threads;
  const int per = (G + rounds - 1) / rounds;
  return (per + 31) / 32 * 32;
}

// One block geometry: NIN = RIN*A in, NOUT = ROUT*A out, OVL grid steps of overlap (half on each side).
template <class FW_, class IV_, int RIN_, int ROUT_, int OVL_, int THREADS_, int MIN_CTAS_>
struct FdrCfg {
  typedef FW_ FW;
  typedef IV_ IV;
  static constexpr int RIN = RIN_, ROUT = ROUT_, OVL = OVL_, THREADS = THREADS_, MIN_CTAS = MIN_CTAS_;
  static constexpr int NIN = FW::N, NOUT = IV::N, A = NIN / RIN;
  static_assert(NIN == RIN * A && NOUT == ROUT * A && OVL % 2 == 0, "block sizes must share the grid step");
  static_assert(THREADS % 32 == 0 && THREADS >= FW::T2 && THREADS >= FW::T3 && THREADS >= IV::T2 && THREADS >= IV::T3,
                "every pass must fit the CTA");
  static constexpr int M_IN = RIN * OVL / 2, M_OUT = ROUT * OVL / 2;
  static constexpr int ADV_IN = RIN * (A - OVL), ADV_OUT = ROUT * (A - OVL);
  static constexpr int TP_F1 = fdr_tp(FW::G1, THREADS), TP_I3 = fdr_tp(IV::G1, THREADS);
  // the two exchange buffers swap roles from one block to the next (no barrier between a block's output pass
  // and the next block's first pass), so both hold the larger of the two layouts
  static constexpr int BUF_A = fdr_cmax(fdr_cmax(fdr_cmax(FW::E1_ELEMS, NIN), IV::E1_ELEMS), fdr_cmax(FW::E2_ELEMS, IV::E2_ELEMS));
  static constexpr int BUF_B = BUF_A;
  // tensor-memory slot of a thread: [forward pass-3 twiddles | response bins | inverse pass-3 twiddles]
  // (the inverse twiddles stay in global memory where the three do not fit the 32 words)
  static constexpr int TM_F3 = 2 * FW::NP2C, TM_H = IV::RC;
  static constexpr int TM_I3 = TM_F3 + TM_H + 2 * IV::NP2C <= 32 ? 2 * IV::NP2C : 0;
  static constexpr int TM_WORDS = TM_F3 + TM_H + TM_I3;
  static constexpr int TM_SLOTS = (THREADS / 32 + 3) / 4;                    // warps w and w + 4 share lanes: one 32-column slot each
  static constexpr int TM_COLS = TM_SLOTS <= 1 ? 32 : TM_SLOTS <= 2 ? 64 : TM_SLOTS <= 4 ? 128 : 256;   // allocations are powers of two
  static constexpr bool TM_OK = TM_WORDS <= 32 && TM_COLS * MIN_CTAS <= 512;
};

// 48000 -> 44100 family (io ratio 160/147): 2560 -> 2352 points
// Tunables (scripts/build_variant.sh overrides them for A/B runs; B200 results in profiles/r1m_ab.log):
// 192 threads x 5 CTAs per SM (64 registers) and an L2 prefetch of the pair's next block measured best.
#ifndef B2_FDR_PREFETCH
#define B2_FDR_PREFETCH 1
#endif
#ifndef B2_FDR_TMA
#define B2_FDR_TMA 1
#endif
#ifndef B2_FDR_THREADS
#define B2_FDR_THREADS 192
#endif
#ifndef B2_FDR_MIN_CTAS
#define B2_FDR_MIN_CTAS 5
#endif
typedef Fft3<8, 16, 20, 322, 136, kP2JFast> Fft2560;
typedef Fft3<12, 14, 14, 196, 169, kP2Row16> Fft2352;
typedef Fft3<8, 16, 10, 162, 136, kP2JFast> Fft1280;
typedef Fft3<8, 10, 8, 82, 88, kP2JFast> Fft640;
typedef Fft3<16, 15, 16, 240, 241, kP2Row16> Fft3840;   // all three passes have 240 tasks
typedef Fft3<12, 21, 14, 294, 253, kP2Row16> Fft3528;
typedef Fft3<12, 14, 7, 98, 169, kP2Row16> Fft1176;     // 98 / 84 / 168 tasks
// 48000 -> 44100 family (io ratio 160/147): 2560 -> 2352 points, one grid step (160 in / 147 out) of margin per side
typedef FdrCfg<Fft2560, Fft2352, 160, 147, 2, B2_FDR_THREADS, B2_FDR_MIN_CTAS> FdrCfg160_147;
// 44100 -> 48000 family (147/160): the same two transforms with their roles swapped
// (224 threads: the 12-point first pass then runs one task per thread, 72 registers without spills)
typedef FdrCfg<Fft2352, Fft2560, 147, 160, 2, 224, 4> FdrCfg147_160;
// 2:1 (96000 -> 48000, ...): 2560 -> 1280 points, 208 input samples of margin per side.  The reference's own
// 1/2 DFT stage decimates in the frequency domain the same way (spectrum truncation, cr.c:163-170).
// (HQ-class filters, 389 taps); 288 samples for VHQ-class ones (533 taps).
typedef FdrCfg<Fft2560, Fft1280, 2, 1, 208, B2_FDR_THREADS, B2_FDR_MIN_CTAS> FdrCfg2_1;
typedef FdrCfg<Fft2560, Fft1280, 2, 1, 288, B2_FDR_THREADS, B2_FDR_MIN_CTAS> FdrCfg2_1w;
// 1:2 (48000 -> 96000, ...): 1280 -> 2560 points, 144 input samples of margin per side (HQ and VHQ)
typedef FdrCfg<Fft1280, Fft2560, 1, 2, 288, 192, 4> FdrCfg1_2;
// 3:1 (48000 -> 16000) and 1:3: 3840 <-> 1280 points, 432 / 144 input samples of margin per side (HQ and VHQ)
typedef FdrCfg<Fft3840, Fft1280, 3, 1, 288, 256, 3> FdrCfg3_1;
typedef FdrCfg<Fft1280, Fft3840, 1, 3, 288, 256, 3> FdrCfg1_3;
// 3:2 (48000 -> 32000) and 2:3: 3840 <-> 2560 points, 216 / 144 input samples of margin per side (HQ and VHQ)
typedef FdrCfg<Fft3840, Fft2560, 3, 2, 144, 256, 3> FdrCfg3_2;
typedef FdrCfg<Fft2560, Fft3840, 2, 3, 144, 256, 3> FdrCfg2_3;
// 6:1 (48000 -> 8000: half-band + 1/3 DFT stage), 1:6, 4:1 (48000 -> 12000) and 1:4: 640 points on the low-rate side
typedef FdrCfg<Fft3840, Fft640, 6, 1, 288, 256, 3> FdrCfg6_1;
typedef FdrCfg<Fft640, Fft3840, 1, 6, 224, 256, 3> FdrCfg1_6;
typedef FdrCfg<Fft2560, Fft640, 4, 1, 224, 192, 4> FdrCfg4_1;
typedef FdrCfg<Fft640, Fft2560, 1, 4, 208, 192, 4> FdrCfg1_4;
// 44100 -> 16000 (441/160): 3528 -> 1280 points, one grid step (441 in / 160 out) of margin per side
typedef FdrCfg<Fft3528, Fft1280, 441, 160, 2, 320, 3> FdrCfg441_160;   // 294 first-pass tasks in one round: 63 registers, no spills

// 44100 <-> 32000 (441/320): 3528 <-> 2560 points, one grid step of margin per side
typedef FdrCfg<Fft3528, Fft2560, 441, 320, 2, 320, 3> FdrCfg441_320;
typedef FdrCfg<Fft2560, Fft3528, 320, 441, 2, 320, 3> FdrCfg320_441;
// 88200 -> 48000, 44100 -> 24000 (147/80) and back: 2352 <-> 1280 points
typedef FdrCfg<Fft2352, Fft1280, 147, 80, 4, 224, 4> FdrCfg147_80;
typedef FdrCfg<Fft1280, Fft2352, 80, 147, 4, 192, 4> FdrCfg80_147;
// 96000 -> 44100, 48000 -> 22050 (320/147) and back: 2560 <-> 1176 points
typedef FdrCfg<Fft2560, Fft1176, 320, 147, 2, 192, 5> FdrCfg320_147;
typedef FdrCfg<Fft1176, Fft2560, 147, 320, 2, 192, 5> FdrCfg147_320;
// 16000 -> 44100 (160/441): 1280 -> 3528 points
typedef FdrCfg<Fft1280, Fft3528, 160, 441, 2, 320, 3> FdrCfg160_441;

struct FdrInfo {
  int variant = 0;          // 1: FdrCfg160_147, 2: FdrCfg147_160, 3: FdrCfg2_1, 4: FdrCfg1_2, 5: FdrCfg2_1w, 6: FdrCfg3_1, 7: FdrCfg1_3, 8: FdrCfg441_160, 9: FdrCfg3_2, 10: FdrCfg2_3, 11: FdrCfg6_1, 12: FdrCfg1_6, 13: FdrCfg4_1, 14: FdrCfg1_4
  int n_head = 0;           // outputs [0, n_head) are recomputed by the exact path
  int adv_out = 0, adv_in = 0, m_in = 0, m_out = 0, nin = 0, nout = 0;
  size_t smem_bytes = 0;
  int n_u = 0;              // intermediate samples the head kernel needs
  int m_min = 0;            // first (negative) intermediate index with a non-zero run-in value
  bool head_by_fix = false; // long heads: k_fdr computes them like any output, k_fdr_head_fix removes the run-in terms
  DevBuf H, ftw, itw;       // response; forward / inverse twiddles (tw2 then tw3)
  int ftw3_off = 0, itw3_off = 0;
  ~FdrInfo() { H.release(); ftw.release(); itw.release(); }
};

struct FdrKParams {
  SigDesc in, out;
  unsigned n_pairs, group, ppg;
  int bpair;                        // mono signals: blocks 2i and 2i+1 of ONE signal ride in the real / imaginary parts
  int fast;                         // layout class, checked on the host: 1 every pair is interleaved stereo float32 on
                                    // both sides; 2 / 3 block-paired mono, unit-stride float32 / int16 in and unit-stride
                                    // float32 out; 0 anything else
  long b_first;
  int n_blocks;                     // items per pair (block pairs when bpair)
  long o0, o1;
  const void *H, *ftw2, *ftw3, *itw2, *itw3;
  // Dynamic work distribution: CTAs draw chunks of `chunk` consecutive (pair, block) items from a ticket counter
  // (tickets of this launch start at ticket_base; the counter only ever counts up, so it needs no reset between
  // launches).  With a static split the SM's warp arbiter lets some resident CTAs run ahead and finish early, and
  // the others finish the launch with a fraction of the SM's warps (measured: 22 of 30 warps resident on average).
  // counter == nullptr: static contiguous split.
  // Guided: the first n_big tickets are chunks of `chunk` items (long runs keep the TMA staging of the next block
  // going), the rest of the items are handed out in chunks of `chunk_tail` so the launch ends evenly.
  unsigned *counter;
  unsigned ticket_base, n_chunks, chunk, n_big, chunk_tail;
};

// first outputs of a two-stage cascade, direct form (see k_fdr_head)
struct FdrHeadParams {
  SigDesc in, out;
  unsigned n_sig;
  long o0, o1;                      // outputs [o0, o1), o1 <= n_head
  const float *gA, *gB;             // stage taps
  int nA, nB, LA, MA, LB, MB;
  long CA, CB;
  int n_u;                          // intermediate samples u[0 .. n_u) needed
  int m_min;                        // k_fdr_head_fix: intermediate samples u[m_min .. 0) that the reference does not have
};

#ifdef __CUDACC__
// MODE (FdrKParams::fast, checked on the host): 1 = every pair is interleaved stereo float32 on both sides; 2 / 3 =
// block-paired mono signals with unit-stride float32 / int16 input and unit-stride float32 output (mono batches,
// and the float scratch in front of the integer epilogue); 0 = any layout.  Separate instantiations, so the general
// body's address set-up, predicates and registers cost the common shapes nothing.
template <class Cfg, int MODE>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MIN_CTAS) k_fdr(const FdrKParams P) {
  typedef float R;
  typedef float2 C;
  typedef typename Cfg::FW FW;
  typedef typename Cfg::IV IV;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C *bufA = reinterpret_cast<C *>(smem_raw);
  C *bufB = bufA + Cfg::BUF_A;
  const int tau = int(threadIdx.x);
  const R *__restrict__ H = static_cast<const R *>(P.H);
  const C *__restrict__ ftw2 = static_cast<const C *>(P.ftw2), *__restrict__ ftw3 = static_cast<const C *>(P.ftw3);
  const C *__restrict__ itw2 = static_cast<const C *>(P.itw2), *__restrict__ itw3 = static_cast<const C *>(P.itw3);

  // a contiguous range of (pair, block) items per CTA, walked incrementally: the per-pair address
  // set-up (integer divisions) runs only when the pair changes
  const unsigned long long n_items = (unsigned long long)P.n_pairs * unsigned(P.n_blocks);
  unsigned pair = 0;
  int bi = 0, n_mine = 0;   // 32-bit loop state: registers are scarce
  bool fresh = true;
  __shared__ unsigned s_ticket;
  unsigned next_ticket = 0;   // thread 0: the ticket drawn ahead while the current chunk is processed
  // Per-pair state kept in registers across blocks: only what the interleaved-stereo fast path needs.  The other
  // layouts recompute their address set-up (pair_state) where they use it: at 64 registers per thread the full
  // state lived in local memory, and its reloads were 5 % of the warp stall samples (profiles/r1v).
  constexpr bool fast = MODE == 1, mfast = MODE >= 2;
  const bool bpair = mfast || (MODE == 0 && P.bpair != 0);
  typedef typename std::conditional<MODE == 3, short, float>::type MIn;   // element type of a mono fast input
  const MIn *src_m = nullptr;       // mono fast paths: the signal's logical sample 0 ...
  float *dst_m = nullptr;           // ... and output 0
  struct PairState {
    PairIn<R> pin;
    long long ooff0, ooff1;
    bool has1, ovec;
  };
  auto pair_state = [&](unsigned pr) {
    PairState ps;
    const unsigned clip = pr / P.ppg, cp = pr % P.ppg;
    const unsigned s0 = bpair ? pr : clip * P.group + 2 * cp, s1 = s0 + 1;
    ps.has1 = !bpair && 2 * cp + 1 < P.group;
    ps.pin = make_pair_in<R>(P.in, s0, ps.has1);
    ps.ooff0 = (long long)(s0 / P.out.channels) * P.out.clip_stride + (long long)(s0 % P.out.channels) * P.out.chan_stride -
               P.out.base * P.out.frame_stride;
    ps.ooff1 = ps.has1 ? (long long)(s1 / P.out.channels) * P.out.clip_stride +
                             (long long)(s1 % P.out.channels) * P.out.chan_stride - P.out.base * P.out.frame_stride
                       : ps.ooff0;
    ps.ovec = P.out.dtype == 0 && ps.has1 && ps.ooff1 == ps.ooff0 + 1 && !(ps.ooff0 & 1) && !(P.out.frame_stride & 1) &&
              !(reinterpret_cast<unsigned long long>(P.out.p) & 7);
    return ps;
  };
  const float2 *src_pair = nullptr;
  float2 *dst_pair = nullptr;

#if B2_FDR_TMA
  // The next block's input is staged into the exchange buffer that is idle during the last inverse pass by
  // TMA bulk copies (one per row of the first pass's [RA][G1] view, at the E1 row stride), so the first pass
  // of the next block runs in place out of shared memory instead of waiting for global loads.
  __shared__ __align__(8) unsigned long long in_bar;
  unsigned in_phase = 0;
  bool staged = false;
  if (tau == 0) {
    mbar_init(&in_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#endif
  // tensor-memory slot of this thread (see tmem_* above): allocated by warp 0, filled once per CTA
  constexpr bool TM = Cfg::TM_OK && B2_FDR_TMEM != 0;
  __shared__ unsigned tm_base_s;
  unsigned tm = 0;
  if (TM) {
    if (tau < 32) tmem_alloc(&tm_base_s, Cfg::TM_COLS);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  if (tau == 0) s_ticket = P.counter ? atomicAdd(P.counter, 1u) - P.ticket_base : blockIdx.x;
  __syncthreads();
  if (TM) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // address = lane << 16 | column: lanes 32 (warp % 4) .., columns 32 (warp / 4) ..
    tm = tm_base_s + ((unsigned(tau >> 5) & 3u) << 21) + (unsigned(tau) >> 7) * 32u;
    {
      unsigned r[Cfg::TM_F3];
      FW::tw3_words(tau, ftw3, r);
      tmem_st(tm, r);
    }
    {
      unsigned r[Cfg::TM_H];
      IV::h_words(tau, H, r);
      tmem_st(tm + Cfg::TM_F3, r);
    }
    if (Cfg::TM_I3) {
      unsigned r[2 * IV::NP2C];
      IV::tw3_words(tau, itw3, r);
      tmem_st(tm + Cfg::TM_F3 + Cfg::TM_H, r);
    }
    tmem_wait_st();
  }
#ifdef B2_FDR_TIMING
  // debug build (scripts/build_variant.sh -DB2_FDR_TIMING=1): per-CTA timeline printed by thread 0
  unsigned long long dbg_t0 = 0;
  long long dbg_edge = 0;
  if (tau == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
  int dbg_n_edge = 0, dbg_items = 0;
#endif
  for (;;) {
  {
    const unsigned ticket = s_ticket;
    unsigned long long it0;
    if (P.counter) {
      if (ticket >= P.n_chunks) break;
      if (tau == 0) next_ticket = atomicAdd(P.counter, 1u) - P.ticket_base;
      const bool big = ticket < P.n_big;
      it0 = big ? (unsigned long long)ticket * P.chunk
                : (unsigned long long)P.n_big * P.chunk + (unsigned long long)(ticket - P.n_big) * P.chunk_tail;
      n_mine = int(min((unsigned long long)(big ? P.chunk : P.chunk_tail), n_items - it0));
    } else {
      if (ticket >= gridDim.x) break;
      it0 = n_items * ticket / gridDim.x;
      n_mine = int(n_items * (ticket + 1ull) / gridDim.x - it0);
      next_ticket = gridDim.x;
    }
    pair = unsigned(it0 / unsigned(P.n_blocks));
    bi = int(it0 % unsigned(P.n_blocks));
    fresh = true;
#if B2_FDR_TMA
    staged = false;
#endif
  }
  for (int it = 0; it < n_mine; ++it) {
#ifdef B2_FDR_TIMING
    const long long dbg_b0 = tau == 0 ? clock64() : 0;
#endif
    if (fresh) {
      const PairState ps = pair_state(pair);
      src_pair = reinterpret_cast<const float2 *>(static_cast<const float *>(ps.pin.p) + ps.pin.off0);
      dst_pair = reinterpret_cast<float2 *>(static_cast<float *>(P.out.p) + ps.ooff0);
      if (mfast) {
        src_m = static_cast<const MIn *>(ps.pin.p) + ps.pin.off0;
        dst_m = static_cast<float *>(P.out.p) + ps.ooff0;
      }
      fresh = false;
    }
    const long b = P.b_first + (bpair ? 2 * bi : bi);
    const long kb = b * long(Cfg::ADV_IN) - Cfg::M_IN;      // input index of block point 0
    const long nb = b * long(Cfg::ADV_OUT) - Cfg::M_OUT;    // output index of block point 0
    const bool interior = kb >= P.in.vlo && kb + Cfg::NIN <= P.in.vhi;

#if B2_FDR_TMA
    if (staged) {
      mbar_wait(&in_bar, in_phase);
      in_phase ^= 1u;
      const C *__restrict__ stg = bufA;
      FW::template fwd_p1<Cfg::TP_F1>(tau, bufA, [&](int g, int n) { return stg[FW::S1 * n + g]; });
    } else
#endif
    if (mfast) {
      // block-paired mono signal with unit stride: blocks b and b + 1 straight from the signal's row
      const long vlo = P.in.vlo, vhi = P.in.vhi;
      if (interior && kb + Cfg::ADV_IN + Cfg::NIN <= vhi) {
        const MIn *__restrict__ src = src_m + kb;
#if B2_FDR_PREFETCH
        {
          // pull the new part of the signal's next two blocks (2 ADV_IN samples behind this item's window) towards L2
          // while this item is transformed: one 128-byte line per thread
          constexpr int kPerLine = 128 / int(sizeof(MIn)), kLines = (2 * Cfg::ADV_IN + kPerLine - 1) / kPerLine;
          static_assert(kLines <= 2 * Cfg::THREADS, "two lines per thread cover the next item");
          const long nxt = kb + Cfg::ADV_IN + Cfg::NIN;
          if (bi + 1 < P.n_blocks && nxt + 2 * Cfg::ADV_IN <= vhi) {
            if (tau < kLines) asm volatile("prefetch.global.L2 [%0];" ::"l"(src_m + nxt + tau * kPerLine));
            if (kLines > Cfg::THREADS && tau + Cfg::THREADS < kLines)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(src_m + nxt + (tau + Cfg::THREADS) * kPerLine));
          }
        }
#endif
        FW::template fwd_p1<Cfg::TP_F1>(tau, bufA, [&](int g, int n) {
          return make_float2(float(src[g + FW::G1 * n]), float(src[g + FW::G1 * n + Cfg::ADV_IN]));
        });
      } else {
        const MIn *__restrict__ src = src_m;
        FW::template fwd_p1<Cfg::TP_F1>(tau, bufA, [&](int g, int n) {
          const long k0 = kb + g + FW::G1 * n, k1 = k0 + Cfg::ADV_IN;
          return make_float2((k0 >= vlo && k0 < vhi) ? float(src[k0]) : 0.f, (k1 >= vlo && k1 < vhi) ? float(src[k1]) : 0.f);
        });
      }
    } else if (bpair) {
      // one signal, two consecutive blocks: block b in the real part, block b + 1 in the imaginary part
      const PairIn<R> pin = pair_state(pair).pin;
      const bool both_in = interior && kb + Cfg::ADV_IN + Cfg::NIN <= pin.vhi && pin.fs == 1;
      if (both_in && pin.dtype == 0) {
        const float *__restrict__ src = static_cast<const float *>(pin.p) + pin.off0 + kb;
        FW::template fwd_p1<Cfg::TP_F1>(tau, bufA, [&](int g, int n) {
          return make_float2(src[g + FW::G1 * n], src[g + FW::G1 * n + Cfg::ADV_IN]);
        });
      } else if (both_in && pin.dtype == 3) {
        const short *__restrict__ src = static_cast<const short *>(pin.p) + pin.off0 + kb;
        FW::template fwd_p1<Cfg::TP_F1>(tau, bufA, [&](int g, int n) {
          return make_float2(float(src[g + FW::G1 * n]), float(src[g + FW::G1 * n + Cfg::ADV_IN]));
        });
      } else {
        FW::template fwd_p1<Cfg::TP_F1>(tau, bufA, [&](int g, int n) {
          const long k0 = kb + g + FW::G1 * n, k1 = k0 + Cfg::ADV_IN;
          C c;
          c.x = (k0 >= pin.vlo && k0 < pin.vhi) ? ld_as<R>(pin.p, pin.dtype, pin.off0 + k0 * pin.fs) : R(0);
          c.y = (k1 >= pin.vlo && k1 < pin.vhi) ? ld_as<R>(pin.p, pin.dtype, pin.off0 + k1 * pin.fs) : R(0);
          return c;
        });
      }
    } else if (fast && interior) {
      const float2 *__restrict__ src = src_pair + kb;
#if B2_FDR_PREFETCH
      // pull the next block of this pair towards L2 while this one is transformed (one 128-byte line per thread)
      if (bi + 1 < P.n_blocks && tau * 16 < Cfg::ADV_IN && kb + Cfg::NIN + Cfg::ADV_IN <= P.in.vhi)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(src + Cfg::NIN + tau * 16));
#endif
      FW::template fwd_p1<Cfg::TP_F1>(tau, bufA, [&](int g, int n) { return src[g + FW::G1 * n]; });
    } else if (fast) {
      // blocks that stick out of the clip (its first and last ones): the same float2 loads, zero outside [vlo, vhi)
      const float2 *__restrict__ src = src_pair;
      const long vlo = P.in.vlo, vhi = P.in.vhi;
      FW::template fwd_p1<Cfg::TP_F1>(tau, bufA, [&](int g, int n) {
        const long k = kb + g + FW::G1 * n;
        return (k >= vlo && k < vhi) ? src[k] : make_float2(0.f, 0.f);
      });
    } else {
      const PairIn<R> pin = pair_state(pair).pin;
      if (pin.vec && interior) {
        const float2 *__restrict__ src = src_pair + kb * (pin.fs >> 1);
        const long long st2 = pin.fs >> 1;
        FW::template fwd_p1<Cfg::TP_F1>(tau, bufA, [&](int g, int n) { return src[(g + FW::G1 * n) * st2]; });
      } else if (pin.vec16 && interior) {
        // interleaved int16 neighbours (stereo PCM): one 32-bit load per frame, converted on the ALU / FMA pipes
        const unsigned *__restrict__ src =
            reinterpret_cast<const unsigned *>(static_cast<const short *>(pin.p) + pin.off0) + kb * (pin.fs >> 1);
        const long long st2 = pin.fs >> 1;
#if B2_FDR_PREFETCH
        // pull the next block of this pair towards L2 while this one is transformed (one 128-byte line per thread)
        if (st2 == 1 && bi + 1 < P.n_blocks && tau * 32 < Cfg::ADV_IN && kb + Cfg::NIN + Cfg::ADV_IN <= P.in.vhi)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(src + Cfg::NIN + tau * 32));
#endif
        FW::template fwd_p1<Cfg::TP_F1>(tau, bufA, [&](int g, int n) {
          const unsigned t = src[(g + FW::G1 * n) * st2];
          return make_float2(s16_bits_to_f32(t), s16_bits_to_f32(t >> 16));
        });
      } else {
        FW::template fwd_p1<Cfg::TP_F1>(tau, bufA, [&](int g, int n) { return pin.load(kb + g + FW::G1 * n); });
      }
    }
    {
      C w[FW::RB];
      FW::ld_tw2(tau, ftw2, w);
      __syncthreads();
      FW::fwd_p2(tau, bufA, bufB, w);
    }
    {
      C w[FW::RC];
      if (TM) {
        unsigned r[Cfg::TM_F3];
        tmem_ld(tm, r);
        tmem_wait_ld();
        FW::tw3_from_words(r, w);
      } else {
        FW::ld_tw3(tau, ftw3, w);
      }
      __syncthreads();
      FW::fwd_p3(tau, bufB, bufA, w);
    }
    {
      R h[IV::RC];
      C w[IV::RC];
      if (TM) {
        unsigned r[Cfg::TM_H + Cfg::TM_I3];
        tmem_ld(tm + Cfg::TM_F3, r);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < IV::RC; ++q) h[q] = __uint_as_float(r[q]);
        if (Cfg::TM_I3)
          IV::tw3_from_words(r + Cfg::TM_H, w);
        else
          IV::ld_tw3(tau, itw3, w);
      } else {
        IV::ld_h(tau, H, h);
        IV::ld_tw3(tau, itw3, w);
      }
      __syncthreads();
      IV::template inv_p1<Cfg::NIN>(tau, bufA, h, bufB, w);
    }
    {
      C w[IV::RB];
      IV::ld_tw2(tau, itw2, w);
      __syncthreads();
      IV::inv_p2(tau, bufB, bufA, w);
    }
    __syncthreads();
#if B2_FDR_TMA
    {
      // bufB's last readers (the second inverse pass) are all past the barrier above; it becomes the next
      // block's bufA.  Same pair, whole block inside the clip, 16-byte aligned rows: else the next block loads itself.
      const long kn = kb + Cfg::ADV_IN;
      staged = fast && !bpair && it + 1 < n_mine && bi + 1 < P.n_blocks && kn >= P.in.vlo && kn + Cfg::NIN <= P.in.vhi &&
               !(reinterpret_cast<unsigned long long>(src_pair + kn) & 15);
      if (staged && tau == 0) {
        static_assert(FW::S1 % 2 == 0 && FW::G1 % 2 == 0 && Cfg::BUF_A % 2 == 0, "16-byte rows for the bulk copies");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic accesses to bufB before, async writes after
        mbar_expect_tx(&in_bar, unsigned(Cfg::NIN * sizeof(C)));
#pragma unroll
        for (int n = 0; n < FW::RA; ++n)
          tma_bulk_g2s(bufB + FW::S1 * n, src_pair + kn + FW::G1 * n, unsigned(FW::G1 * sizeof(C)), &in_bar);
      }
    }
#endif
    {
      // owned outputs: block-local o in [M_OUT, M_OUT + ADV_OUT), clipped to the requested range
      const int lo = int(max(long(Cfg::M_OUT), P.o0 - nb)), hi = int(min(long(Cfg::M_OUT + Cfg::ADV_OUT), P.o1 - nb));
      if (mfast) {
        const long nb1 = nb + Cfg::ADV_OUT;                  // block b + 1 (imaginary part)
        const int lo1 = int(max(long(Cfg::M_OUT), P.o0 - nb1)), hi1 = int(min(long(Cfg::M_OUT + Cfg::ADV_OUT), P.o1 - nb1));
        float *__restrict__ dst = dst_m + nb;
        if (lo == Cfg::M_OUT && hi == Cfg::M_OUT + Cfg::ADV_OUT && lo1 == lo && hi1 == hi) {
          // both blocks wholly owned: rows inside [M_OUT, M_OUT + ADV_OUT) need no predicate
          IV::template inv_p3<Cfg::TP_I3>(tau, bufA, [&](int g, int n, C v) {
            const bool row_in = IV::G1 * n >= Cfg::M_OUT && IV::G1 * (n + 1) <= Cfg::M_OUT + Cfg::ADV_OUT;
            const int o = g + IV::G1 * n;
            if (row_in || (o >= Cfg::M_OUT && o < Cfg::M_OUT + Cfg::ADV_OUT)) {
              dst[o] = v.x;
              dst[o + Cfg::ADV_OUT] = v.y;
            }
          });
        } else {
          IV::template inv_p3<Cfg::TP_I3>(tau, bufA, [&](int g, int n, C v) {
            const int o = g + IV::G1 * n;
            if (o >= lo && o < hi) dst[o] = v.x;
            if (o >= lo1 && o < hi1) dst[o + Cfg::ADV_OUT] = v.y;
          });
        }
      } else if (bpair) {
        const long long ooff0 = pair_state(pair).ooff0;
        const long nb1 = nb + Cfg::ADV_OUT;                  // block b + 1 (imaginary part)
        const int lo1 = int(max(long(Cfg::M_OUT), P.o0 - nb1)), hi1 = int(min(long(Cfg::M_OUT + Cfg::ADV_OUT), P.o1 - nb1));
        if (P.out.dtype == 0 && P.out.frame_stride == 1) {
          float *__restrict__ dst = static_cast<float *>(P.out.p) + ooff0 + nb;
          IV::template inv_p3<Cfg::TP_I3>(tau, bufA, [&](int g, int n, C v) {
            const int o = g + IV::G1 * n;
            if (o >= lo && o < hi) dst[o] = v.x;
            if (o >= lo1 && o < hi1) dst[o + Cfg::ADV_OUT] = v.y;
          });
        } else {
          IV::template inv_p3<Cfg::TP_I3>(tau, bufA, [&](int g, int n, C v) {
            const int o = g + IV::G1 * n;
            if (o >= lo && o < hi) st_as<R>(P.out.p, P.out.dtype, ooff0 + (nb + o) * P.out.frame_stride, v.x);
            if (o >= lo1 && o < hi1) st_as<R>(P.out.p, P.out.dtype, ooff0 + (nb1 + o) * P.out.frame_stride, v.y);
          });
        }
      } else if (fast) {
        float2 *__restrict__ dst = dst_pair + nb;
        if (lo == Cfg::M_OUT && hi == Cfg::M_OUT + Cfg::ADV_OUT) {
          // whole block owned: rows fully inside [M_OUT, M_OUT + ADV_OUT) need no predicate
          IV::template inv_p3<Cfg::TP_I3>(tau, bufA, [&](int g, int n, C v) {
            const bool row_in = IV::G1 * n >= Cfg::M_OUT && IV::G1 * (n + 1) <= Cfg::M_OUT + Cfg::ADV_OUT;
            const int o = g + IV::G1 * n;
            if (row_in || (o >= Cfg::M_OUT && o < Cfg::M_OUT + Cfg::ADV_OUT)) dst[o] = v;
          });
        } else {
          IV::template inv_p3<Cfg::TP_I3>(tau, bufA, [&](int g, int n, C v) {
            const int o = g + IV::G1 * n;
            if (o >= lo && o < hi) dst[o] = v;
          });
        }
      } else {
        const PairState ps = pair_state(pair);
        if (ps.ovec) {
          float2 *__restrict__ dst = dst_pair + nb * (P.out.frame_stride >> 1);
          const long long st2 = P.out.frame_stride >> 1;
          IV::template inv_p3<Cfg::TP_I3>(tau, bufA, [&](int g, int n, C v) {
            const int o = g + IV::G1 * n;
            if (o >= lo && o < hi) dst[o * st2] = v;
          });
        } else if (P.out.dtype == 0 && P.out.frame_stride == 1 && ps.has1) {
          // two unit-stride float32 rows (the planar scratch in front of the integer epilogue)
          float *__restrict__ d0 = static_cast<float *>(P.out.p) + ps.ooff0 + nb;
          float *__restrict__ d1 = static_cast<float *>(P.out.p) + ps.ooff1 + nb;
          IV::template inv_p3<Cfg::TP_I3>(tau, bufA, [&](int g, int n, C v) {
            const int o = g + IV::G1 * n;
            if (o >= lo && o < hi) {
              d0[o] = v.x;
              d1[o] = v.y;
            }
          });
        } else {
          IV::template inv_p3<Cfg::TP_I3>(tau, bufA, [&](int g, int n, C v) {
            const int o = g + IV::G1 * n;
            if (o >= lo && o < hi) store_pair<R>(P.out, ps.ooff0, ps.ooff1, ps.has1, false, nb + o, v);
          });
        }
      }
    }
    // No barrier here: the next block's first pass writes into what is bufB now, whose last readers (the second
    // inverse pass) all passed the barrier above; bufA, still being read by slower warps, is not written again
    // before the barrier that follows that first pass.
    {
      C *const t = bufA;
      bufA = bufB;
      bufB = t;
    }
#ifdef B2_FDR_TIMING
    if (!interior && tau == 0) {
      dbg_edge += clock64() - dbg_b0;
      ++dbg_n_edge;
    }
#endif
    if (++bi == P.n_blocks) {
      bi = 0;
      ++pair;
      fresh = true;
    }
  }
#ifdef B2_FDR_TIMING
  dbg_items += n_mine;
#endif
  // next chunk: its ticket was drawn at the start of this one (the atomic's latency is long gone)
  if (tau == 0) s_ticket = next_ticket;
  __syncthreads();
  }
  if (TM) {
    // every thread left the loop behind the same barrier; no tcgen05.ld is in flight (each was waited for)
    if (tau < 32) tmem_dealloc(tm_base_s, Cfg::TM_COLS);
  }
#ifdef B2_FDR_TIMING
  if (tau == 0) {
    unsigned long long dbg_t1;
    unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t1));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    printf("FDRT %u %u %llu %llu %d %d %lld\n", blockIdx.x, smid, dbg_t0, dbg_t1, dbg_items, dbg_n_edge, dbg_edge);
  }
#endif
}
#endif  // __CUDACC__

// ---- host side ---------------------------------------------------------------
// Real zero-phase response of the plan's cascade on the NOUT-bin grid, including every
// normalisation (kernel output = unnormalised inverse transform of H .* X).  Empty if the plan
// is not a cascade of rational stages reducing to rin:rout or the composite filter does not fit
// in the block margins.
std::vector<double> fdr_response(const Plan &plan, int rin, int rout, int nin, int nout, int m_in, int *n_head);

const char *fdr_prepare(DevicePlan *dp);
// final outputs [o0, o1) of all signals: k_fdr_head for [o0, n_head), k_fdr for the rest
const char *fdr_run(const DevicePlan &dp, const FdrInfo &fi, const SigDesc &in, const SigDesc &out, long o0, long o1,
                    unsigned n_sig, unsigned group, cudaStream_t st, Workspace *ws = nullptr);

}  // namespace b2
