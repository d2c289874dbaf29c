// fdr2_stage.cuh -- two-pass (single-exchange) form of the 160:147 block of fdr_stage.cuh: 2560 = 40 x 64 points in,
// 2352 = 49 x 48 points out, every transform held in one thread's registers.
//
// Same operator, same block grid, same response table and head kernels as k_fdr<FdrCfg160_147> (fdr_stage.cuh);
// what changes is the factorisation: two exchanges through shared memory per block instead of four, and
// register-resident 40 / 64 / 49 / 48-point transforms (prime-factor 8x5 and 16x3, Cooley-Tukey 8x8 and 7x7 with
// compile-time twiddles) instead of 8/16/20 and 12/14/14.  Probe: scratch/fdr2/probe.cu, profiles/r2w_two_pass_probe.log.
//
//   forward  n = g + 64 m, k = j + 40 q:   pass A (thread g): Z[j] = DFT40_m x[g + 64 m], times w_2560^(g j)
//                                          pass B (thread j): X[j + 40 q] = DFT64_g Z[j; g]
//   inverse  F = t + 48 q, o = m + 49 r:   pass 1 (thread t): U[m] = IDFT49_q (H X)[t + 48 q], times conj w_2352^(t m)
//                                          pass 2 (thread m): y[m + 49 r] = IDFT48_t U[t; m]
// A CTA of 128 threads works on TWO blocks at a time and re-packs the tasks of every pass over its warps
// (128 / 80 / 96 / 98 tasks), because 40, 48 and 49 tasks alone would leave most of a second warp idle.
#pragma once
#include <utility>

#include "../../python-soxr_b200/csrc/fdr_stage.cuh"
#include "unit_roots.h"

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

// w^m, m = 1 .. R-1, from the six powers of two b[i] = w^(2^i): one product each (error = popcount(m) roundings)
__host__ __device__ constexpr int fdr2_top_bit(int m) { return m >= 32 ? 32 : m >= 16 ? 16 : m >= 8 ? 8 : m >= 4 ? 4 : m >= 2 ? 2 : 1; }
template <int R, typename C>
__host__ __device__ __forceinline__ void twiddle_chain(const C (&b)[6], C (&w)[R]) {
#pragma unroll
  for (int i = 0; i < 6; ++i)
    if ((1 << i) < R) w[1 << i] = b[i];
#pragma unroll
  for (int m = 3; m < R; ++m)
    if (fdr2_top_bit(m) != m) w[m] = cmul(w[fdr2_top_bit(m)], w[m - fdr2_top_bit(m)]);
}

struct Fdr2 {
  typedef float R;
  typedef float2 C;
  static constexpr int NIN = 2560, NOUT = 2352, RA = 40, RB = 64, RC = 49, RD = 48;
  static constexpr int SE = 65;      // exchange 1: E[SE j + g], j < 40 rows
  static constexpr int SE2 = 49;     // exchange 2: E2[SE2 m + t], m < 49 rows (aliases E)
  static constexpr int E_ELEMS = SE * RA > SE2 * RC ? SE * RA : SE2 * RC;   // 2600
  // ONE buffer per block: pass B works in place on its row (thread j owns row j: spectrum bin j + 40 q at
  // E[SE j + q]), and pass 1 of the inverse reads all its bins into registers before anything is overwritten
  static constexpr int SLOT_ELEMS = E_ELEMS + 8;                            // slots 8 (mod 16) elements apart

  // pass A, task g in [0, 64): ld(n) returns x[n]
  template <typename LD>
  static __host__ __device__ __forceinline__ void fwd_a(int g, C *__restrict__ E, LD ld, const C (&wb)[6]) {
    C v[RA];
#pragma unroll
    for (int m = 0; m < RA; ++m) v[m] = ld(g + RB * m);
    Dft<RA>::template run<false>(v);
    C w[RA];
    twiddle_chain<RA>(wb, w);
#pragma unroll
    for (int j = 1; j < RA; ++j) v[j] = cmul(v[j], w[j]);
#pragma unroll
    for (int j = 0; j < RA; ++j) E[SE * j + g] = v[j];
  }
  // pass B, task j in [0, 40), in place: row j of E becomes X[j + 40 q] at E[SE j + q]
  static __host__ __device__ __forceinline__ void fwd_b(int j, C *E) {
    C v[RB];
#pragma unroll
    for (int g = 0; g < RB; ++g) v[g] = E[SE * j + g];
    Dft<RB>::template run<false>(v);
#pragma unroll
    for (int q = 0; q < RB; ++q) E[SE * j + q] = v[q];
  }
  static __host__ __device__ __forceinline__ int sp_at(int k) { return SE * (k % RA) + k / RA; }   // where bin k lives
  // inverse pass 1, task t in [0, 48): h(q) returns H[t + 48 q].  Two halves around a barrier: every bin is read
  // (and transformed) before any thread writes the second exchange layout over the same buffer.
  // bin of task t, term q: index into E of the forward bin feeding inverse bin F = t + 48 q
  static __host__ __device__ __forceinline__ int inv_1_src(int t, int q) {
    const int F = t + RD * q;
    const bool pos = q < 24 || (q == 24 && t < 24);        // F < 1176
    return sp_at(pos ? F : F + (NIN - NOUT));
  }
  static __host__ __device__ __forceinline__ void inv_1_finish(const C (&wb)[6], C (&v)[RC]) {
    Dft<RC>::template run<true>(v);
    C w[RC];
    twiddle_chain<RC>(wb, w);
#pragma unroll
    for (int m = 1; m < RC; ++m) v[m] = cmulc(v[m], w[m]);
  }
  template <typename HF>
  static __host__ __device__ __forceinline__ void inv_1_compute(int t, const C *__restrict__ E, HF h, const C (&wb)[6], C (&v)[RC]) {
#pragma unroll
    for (int q = 0; q < RC; ++q) v[q] = scale(h(q), E[inv_1_src(t, q)]);
    inv_1_finish(wb, v);
  }
  static __host__ __device__ __forceinline__ void inv_1_store(int t, C *__restrict__ E2, const C (&v)[RC]) {
#pragma unroll
    for (int m = 0; m < RC; ++m) E2[SE2 * m + t] = v[m];
  }
  // inverse pass 2, task m in [0, 49): st(o, value) for o = m + 49 r
  template <typename ST>
  static __host__ __device__ __forceinline__ void inv_2(int m, const C *__restrict__ E2, ST st) {
    C v[RD];
#pragma unroll
    for (int t = 0; t < RD; ++t) v[t] = E2[SE2 * m + t];
    Dft<RD>::template run<true>(v);
#pragma unroll
    for (int r = 0; r < RD; ++r) st(m + RC * r, v[r]);
  }
};

#ifdef __CUDACC__
// Per-slot description of the block a CTA half works on, written by thread 0 (two parities: the next pair of blocks
// is described while slow threads still read the current one).
struct Fdr2Slot {
  const float2 *src;     // the pair's logical sample 0 (input) ...
  float2 *dst;           // ... and output 0
  long kb, nb;           // input / output index of block point 0
  int lo, hi;            // owned block-local outputs [lo, hi)
  int active, interior;
};

// tensor-memory words of a thread: [0,12) forward twiddle bases of g = tau % 64, [12,24) inverse bases of
// t = tau % 48, [24,73) the response bins H[t + 48 q]
constexpr int kF2TmFwd = 0, kF2TmInv = 12, kF2TmH = 24, kF2TmCols = 128;

__global__ void __launch_bounds__(128, 3) k_fdr2(const FdrKParams P) {
  typedef Fdr2 F;
  typedef float2 C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C *base = reinterpret_cast<C *>(smem_raw);
  __shared__ Fdr2Slot slots[2][2];
  __shared__ unsigned s_ticket[2];
  __shared__ unsigned tm_base_s;
  const int tau = int(threadIdx.x);
  const float *__restrict__ H = static_cast<const float *>(P.H);
  const C *__restrict__ ftwb = static_cast<const C *>(P.f2tab), *__restrict__ itwb = ftwb + 6 * F::RB;

  if (tau < 32) tmem_alloc(&tm_base_s, kF2TmCols);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tm = tm_base_s + (unsigned(tau >> 5) << 21);
  {
    const int g = tau & 63, t = tau % F::RD;
    unsigned r[12];
#pragma unroll
    for (int i = 0; i < 6; ++i) { const C w = ftwb[g + F::RB * i]; r[2 * i] = __float_as_uint(w.x); r[2 * i + 1] = __float_as_uint(w.y); }
    tmem_st(tm + kF2TmFwd, r);
#pragma unroll
    for (int i = 0; i < 6; ++i) { const C w = itwb[t + F::RD * i]; r[2 * i] = __float_as_uint(w.x); r[2 * i + 1] = __float_as_uint(w.y); }
    tmem_st(tm + kF2TmInv, r);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      unsigned h[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) h[i] = __float_as_uint(H[t + F::RD * (16 * c + i)]);
      tmem_st(tm + kF2TmH + 16 * c, h);
    }
    unsigned h1[1] = {__float_as_uint(H[t + F::RD * 48])};
    tmem_st(tm + kF2TmH + 48, h1);
    tmem_wait_st();
  }

  const unsigned long long n_items = (unsigned long long)P.n_pairs * unsigned(P.n_blocks);
  // The last thread has no task in passes B, 1 and 2: it draws the next ticket and describes the next two blocks
  // while the others work (parity-buffered), so no pass waits for the atomic or the address set-up.
  auto describe = [&](unsigned par) {
    const unsigned ticket = atomicAdd(P.counter, 1u) - P.ticket_base;
    s_ticket[par] = ticket;
    if (ticket >= P.n_chunks) return;
#pragma unroll 1
    for (int sl = 0; sl < 2; ++sl) {
      Fdr2Slot d;
      const unsigned long long it = 2ull * ticket + unsigned(sl);
      d.active = it < n_items;
      d.src = nullptr; d.dst = nullptr; d.kb = d.nb = 0; d.lo = d.hi = 0; d.interior = 0;
      if (d.active) {
        const unsigned pair = unsigned(it / unsigned(P.n_blocks));
        const int bi = int(it % unsigned(P.n_blocks));
        const unsigned clip = pair / P.ppg, cp = pair % P.ppg, s0 = clip * P.group + 2 * cp;
        const long long ioff = (long long)(s0 / P.in.channels) * P.in.clip_stride + (long long)(s0 % P.in.channels) * P.in.chan_stride -
                               P.in.base * P.in.frame_stride;
        const long long ooff = (long long)(s0 / P.out.channels) * P.out.clip_stride + (long long)(s0 % P.out.channels) * P.out.chan_stride -
                               P.out.base * P.out.frame_stride;
        d.src = reinterpret_cast<const float2 *>(static_cast<const float *>(P.in.p) + ioff);
        d.dst = reinterpret_cast<float2 *>(static_cast<float *>(P.out.p) + ooff);
        const long b = P.b_first + bi;
        d.kb = b * long(FdrCfg160_147::ADV_IN) - FdrCfg160_147::M_IN;
        d.nb = b * long(FdrCfg160_147::ADV_OUT) - FdrCfg160_147::M_OUT;
        d.interior = d.kb >= P.in.vlo && d.kb + F::NIN <= P.in.vhi;
        d.lo = int(max(long(FdrCfg160_147::M_OUT), P.o0 - d.nb));
        d.hi = int(min(long(FdrCfg160_147::M_OUT + FdrCfg160_147::ADV_OUT), P.o1 - d.nb));
      }
      slots[par][sl] = d;
    }
  };
  if (tau == 127) describe(0);
  for (unsigned par = 0;; par ^= 1u) {
    __syncthreads();                     // the description of this pair of blocks is there; the previous pass 2 is done with E
    if (s_ticket[par] >= P.n_chunks) break;

    {   // ---- forward pass A: 128 tasks (slot, g) ----
      const int sl = tau >> 6, g = tau & 63;
      const Fdr2Slot &d = slots[par][sl];
      C wb[6];
      {   // (tcgen05.ld is warp-collective: issued by every thread, outside the slot conditionals)
        unsigned r[12];
        tmem_ld(tm + kF2TmFwd, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 6; ++i) wb[i] = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
      }
      if (d.active) {
        C *E = base + sl * F::SLOT_ELEMS;
        if (d.interior) {
          const float2 *__restrict__ src = d.src + d.kb;
          F::fwd_a(g, E, [&](int n) { return src[n]; }, wb);
        } else {
          const float2 *__restrict__ src = d.src;
          const long kb = d.kb, vlo = P.in.vlo, vhi = P.in.vhi;
          F::fwd_a(g, E, [&](int n) { const long k = kb + n; return (k >= vlo && k < vhi) ? src[k] : make_float2(0.f, 0.f); }, wb);
        }
      }
    }
    __syncthreads();
    if (tau < 2 * F::RA) {   // ---- forward pass B: 80 tasks (slot, j), in place ----
      const int sl = tau / F::RA, j = tau % F::RA;
      if (slots[par][sl].active) F::fwd_b(j, base + sl * F::SLOT_ELEMS);
    } else if (tau == 127) {
      describe(par ^ 1u);
    }
    __syncthreads();
    {   // ---- inverse pass 1: 96 tasks (slot, t) on three whole warps ----
      // (a slot without a block -- the odd last one -- is still transformed and only its store is skipped: tcgen05.ld
      // is warp-collective, and the middle warp serves both slots)
      C v[F::RC];
      const int sl1 = tau / F::RD, t1 = tau % F::RD;
      C *E = base + (sl1 & 1) * F::SLOT_ELEMS;
      if (tau < 2 * F::RD) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          unsigned h[16];
          tmem_ld(tm + kF2TmH + 16 * c, h);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[16 * c + i] = scale(__uint_as_float(h[i]), E[F::inv_1_src(t1, 16 * c + i)]);
        }
        C wb[6];
        {
          unsigned h[1], r[12];
          tmem_ld(tm + kF2TmH + 48, h);
          tmem_ld(tm + kF2TmInv, r);
          tmem_wait_ld();
          v[48] = scale(__uint_as_float(h[0]), E[F::inv_1_src(t1, 48)]);
#pragma unroll
          for (int i = 0; i < 6; ++i) wb[i] = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
        }
        F::inv_1_finish(wb, v);
      }
      __syncthreads();                   // every bin has been read: the second exchange layout may overwrite them
      if (tau < 2 * F::RD && slots[par][sl1].active) F::inv_1_store(t1, E, v);
    }
    __syncthreads();
    if (tau < 2 * F::RC) {   // ---- inverse pass 2: 98 tasks (slot, m), results straight to the caller's buffer ----
      const int sl = tau / F::RC, m = tau % F::RC;
      const Fdr2Slot &d = slots[par][sl];
      if (d.active) {
        const C *E = base + sl * F::SLOT_ELEMS;
        float2 *__restrict__ dst = d.dst + d.nb;
        const int lo = d.lo, hi = d.hi;
        F::inv_2(m, E, [&](int o, C v) { if (o >= lo && o < hi) dst[o] = v; });
      }
    }
  }
  if (tau < 32) tmem_dealloc(tm_base_s, kF2TmCols);
}
#endif  // __CUDACC__

}  // namespace b2
