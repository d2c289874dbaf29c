// fft_stage.cuh -- fused shared-memory overlap-save FFT stage (+ rational polyphase).
//
// One CTA computes one block of one PAIR of signals (signal 2q in the real part,
// 2q+1 in the imaginary part of a complex stream: the filters are real, so the
// two never mix mathematically).  It replaces, in one launch,
//
//   K1  deinterleave + convert-in             data-io.c:13-57
//   K2  dft_stage_fn  (overlap-save FIR, L/M)  cr.c:68-174      (PFFFT: pffft.c)
//   K3  vpoly0        (rational polyphase)     poly-fir0.h:24-44
//   K7  interleave + convert-out (float)       data-io.c:161-170
//
// so the L-times oversampled intermediate never leaves shared memory.
//
// Maths (plan.h closed form  out[j] = sum_k g[C + M j - L k] in[k]):
// with k = k0 + M i + s and j = j0 + L i' + r the stage is L*M ordinary
// convolutions at the base rate,
//     out_r[i'] = sum_s sum_m g_rs[m] in_s[i' - m],   g_rs[m] = g[D + M r - L s + L M m],
// D = C + M j0 - L k0, evaluated per block as circular convolutions of length n:
// M forward FFTs, L*M spectrum products, L inverse FFTs.  The block grid is the
// reference's own (k0 = -preload + b M V, V = n - span valid outputs per phase),
// so whenever the reference emits a block's outputs the block's input is
// complete and results do not depend on how the stream was chunked.
//
// FFT: n = R1*16*16 points, T = n/16 threads, 16 points per thread, three
// register-radix passes (R1,16,16) with two shared-memory exchanges; the forward
// transform is DIT and the inverse the mirrored DIF, so the spectrum stays in
// registers between them.  Replaces PFFFT's radix-2/4 passes.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <vector>

#include "fft_regs.cuh"
#include "plan.h"
#include "runtime.h"

namespace b2 {

struct FftStageInfo {
  int stage_index = -1;
  bool fused_poly = false;
  int log2n = 0, n = 0, L = 1, M = 1;
  long k0 = 0;                   // input index of point 0 (phase 0) of block 0
  int span = 0, v_base = 0;      // valid outputs per block and phase = n - span
  int delta = 0;                 // halo (base-rate samples) kept in front of a block for the poly stage
  int pL = 1, pM = 1, p_at0 = 0, p_pre = 0, nc = 0, ncp = 0, nrow = 0, pv = 0, pE = 0;
  int u_elems = 0;
  size_t smem_bytes = 0;
  int threads = 0;
  int max_ctas_per_sm = 1;
};

struct FftKParams {
  SigDesc in, out;
  unsigned n_sig, n_pairs;
  unsigned group, ppg;           // signals per clip; pairs per clip = (group + 1) / 2 (pairs never span clips)
  long b_first;
  int n_blocks;
  long o0, o1;
  long k0;
  int v_base, delta;
  int u_elems;                   // complex elements of U incl. zeroed slack; U[(delta + i) * L + r]
  const void *spec, *tw2, *tw3, *edge, *ptab;
  int pL, pM, p_at0, p_pre, nc, ncp, nrow;   // ncp: taps looped (multiple of 4); nrow: odd row stride
  int pv, pE;                    // sliding-window variant id (0: one output per thread) and its lead zeros
  unsigned ptab_bytes;
};

#ifdef __CUDACC__

template <typename R> __device__ __forceinline__ R ld_as(const void *p, int dtype, long long idx) {
  switch (dtype) {
    case 0: return R(static_cast<const float *>(p)[idx]);
    case 1: return R(static_cast<const double *>(p)[idx]);
    case 2: return R(static_cast<const int *>(p)[idx]);
    default: return R(static_cast<const short *>(p)[idx]);
  }
}
template <typename R> __device__ __forceinline__ void st_as(void *p, int dtype, long long idx, R v) {
  if (dtype == 0) static_cast<float *>(p)[idx] = float(v);
  else static_cast<double *>(p)[idx] = double(v);
}

__host__ __device__ inline long long ceil_div_ll(long long a, long long b) {   // b > 0
  return a >= 0 ? (a + b - 1) / b : -((-a) / b);
}

// int16 -> float32 without a conversion instruction (those run on the 16-lane XU pipe): the biased sample
// written into the low mantissa bits of 1.5 * 2^23 is the float 12582912 + 32768 + v, exactly.
__device__ __forceinline__ float s16_bits_to_f32(unsigned v16) {   // sample in the low 16 bits, anything above
  return __uint_as_float(0x4B400000u | ((v16 ^ 0x8000u) & 0xFFFFu)) - 12615680.0f;
}

// Per-item view of one signal pair.
template <typename R>
struct PairIn {
  const void *p;
  int dtype;
  long long off0, off1, fs;
  long vlo, vhi;
  bool has1, vec, vec16;   // vec / vec16: the pair is one aligned float2 / short2 per frame (interleaved neighbours)
  typedef typename V2<R>::type C;
  __device__ __forceinline__ C load(long k) const {
    C c;
    c.x = R(0);
    c.y = R(0);
    if (k >= vlo && k < vhi) {
      if (vec) {
        const float2 t = *reinterpret_cast<const float2 *>(static_cast<const float *>(p) + off0 + k * fs);
        c.x = R(t.x);
        c.y = R(t.y);
      } else if (vec16) {
        const unsigned t = *reinterpret_cast<const unsigned *>(static_cast<const short *>(p) + off0 + k * fs);
        c.x = R(s16_bits_to_f32(t));
        c.y = R(s16_bits_to_f32(t >> 16));
      } else {
        c.x = ld_as<R>(p, dtype, off0 + k * fs);
        if (has1) c.y = ld_as<R>(p, dtype, off1 + k * fs);
      }
    }
    return c;
  }
};

template <typename R>
__device__ __forceinline__ PairIn<R> make_pair_in(const SigDesc &d, unsigned s0, bool has1) {
  PairIn<R> pi;
  const unsigned s1 = s0 + 1;
  pi.p = d.p;
  pi.dtype = d.dtype;
  pi.fs = d.frame_stride;
  pi.off0 = (long long)(s0 / d.channels) * d.clip_stride + (long long)(s0 % d.channels) * d.chan_stride - d.base * d.frame_stride;
  pi.has1 = has1;
  pi.off1 = pi.has1 ? (long long)(s1 / d.channels) * d.clip_stride + (long long)(s1 % d.channels) * d.chan_stride - d.base * d.frame_stride
                    : pi.off0;
  pi.vlo = d.vlo;
  pi.vhi = d.vhi;
  const bool neighbours = pi.has1 && pi.off1 == pi.off0 + 1 && !(pi.off0 & 1) && !(pi.fs & 1);
  pi.vec = d.dtype == 0 && neighbours && !(reinterpret_cast<unsigned long long>(d.p) & 7);
  pi.vec16 = d.dtype == 3 && neighbours && !(reinterpret_cast<unsigned long long>(d.p) & 3);
  return pi;
}

template <typename R>
__device__ __forceinline__ void store_pair(const SigDesc &d, long long off0, long long off1, bool has1, bool vec, long k,
                                           typename V2<R>::type v) {
  if (vec) {
    *reinterpret_cast<float2 *>(static_cast<float *>(d.p) + off0 + k * d.frame_stride) = make_float2(float(v.x), float(v.y));
  } else {
    st_as<R>(d.p, d.dtype, off0 + k * d.frame_stride, v.x);
    if (has1) st_as<R>(d.p, d.dtype, off1 + k * d.frame_stride, v.y);
  }
}

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
      "r"(phase)
      : "memory");
}
// TMA bulk copy global -> shared (1-D), completion signalled on the mbarrier.
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst_smem)),
               "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

// Padded shared-memory index.  phys is a sum of shifts, so for x, y occupying
// disjoint bit ranges phys(x + y) == phys(x) + phys(y): every access below is
// "per-thread base (computed once per CTA) + compile-time constant".
template <int BYTES> __host__ __device__ constexpr int physc(int p) {
  return BYTES == 8 ? p + (p >> 4) + (p >> 8) : p + (p >> 4) + (p >> 7);
}

// ---------------------------------------------------------------------------
// K3 out of shared memory.  U holds the block's FFT-stage output in output-rate
// order (element j - ubase).  Table rows: [pE zeros][nc taps][zeros], stride nrow.
// Results go to `ostage` (shared) when `staged`, else straight to the caller.
// ---------------------------------------------------------------------------
template <typename R>
struct OutPair {
  const SigDesc *d;
  long long off0, off1;
  bool has1, vec;
  __device__ __forceinline__ void put(long n, typename V2<R>::type v) const {
    store_pair<R>(*d, off0, off1, has1, vec, n, v);
  }
};

// one output per thread per trip; taps in chunks of 4 (rows zero-padded to ncp)
template <typename R, int T>
__device__ __forceinline__ void poly_emit(const FftKParams &P, const typename V2<R>::type *__restrict__ U,
                                          const R *__restrict__ tab, typename V2<R>::type *__restrict__ ostage,
                                          bool staged, long n_lo, long n_hi, long ubase, int t, const OutPair<R> &op) {
  typedef typename V2<R>::type C;
  if (n_lo + t >= n_hi) return;
  const long long at = (long long)P.p_at0 + (long long)(n_lo + t) * P.pM;
  long div = long(at / P.pL);
  int rem = int(at % P.pL);
  const int dstep = int((long long)T * P.pM / P.pL), rstep = int((long long)T * P.pM % P.pL);
  const int nq = P.ncp >> 2;
  for (long nn = n_lo + t; nn < n_hi; nn += T) {
    const C *__restrict__ A = U + int(div - P.p_pre - ubase);
    const R *__restrict__ c = tab + rem * P.nrow + P.pE;
    C o;
    o.x = R(0);
    o.y = R(0);
#pragma unroll 2
    for (int q = 0; q < nq; ++q) {
      const R c0 = c[4 * q], c1 = c[4 * q + 1], c2 = c[4 * q + 2], c3 = c[4 * q + 3];
      const C u0 = A[4 * q], u1 = A[4 * q + 1], u2 = A[4 * q + 2], u3 = A[4 * q + 3];
      o = axpy(c0, u0, o);
      o = axpy(c1, u1, o);
      o = axpy(c2, u2, o);
      o = axpy(c3, u3, o);
    }
    if (staged) ostage[nn - n_lo] = o;
    else op.put(nn, o);
    div += dstep;
    rem += rstep;
    if (rem >= P.pL) { rem -= P.pL; ++div; }
  }
}

// Sliding window: a thread computes RW consecutive outputs from one register-resident run of
// (RW-1)*B + NT samples.  Output i's window starts at w_0 + i*B + e_i with 0 <= e_i <= pE; the
// e_i shift is applied to the COEFFICIENT address (rows carry pE leading zeros), so sample
// registers are indexed statically.  Cuts shared-memory bytes per output ~2x and makes the lane
// strides of both access patterns odd (scratch/poly_bank_sim.py).
template <typename R, int T, int RW, int B, int NT>
__device__ __forceinline__ void poly_emit_slide(const FftKParams &P, const typename V2<R>::type *__restrict__ U,
                                                const R *__restrict__ tab, typename V2<R>::type *__restrict__ ostage,
                                                bool staged, long n_lo, long n_hi, long ubase, int t,
                                                const OutPair<R> &op) {
  typedef typename V2<R>::type C;
  const long ngroups = (n_hi - n_lo + RW - 1) / RW;
  if (t >= ngroups) return;
  const long long at = (long long)P.p_at0 + (long long)(n_lo + (long)t * RW) * P.pM;
  long div = long(at / P.pL);
  int rem = int(at % P.pL);
  const long long gs = (long long)T * RW * P.pM;
  const int dstep = int(gs / P.pL), rstep = int(gs % P.pL);
  const int fr = P.pM - B * P.pL;   // 0 <= fr < pL
  for (long g = t; g < ngroups; g += T) {
    const C *__restrict__ A = U + int(div - P.p_pre - ubase);
    C s[(RW - 1) * B + NT];
#pragma unroll
    for (int m = 0; m < (RW - 1) * B + NT; ++m) s[m] = A[m];
    const long n0 = n_lo + g * RW;
    int ri = rem, sh = P.pE;
#pragma unroll
    for (int i = 0; i < RW; ++i) {
      const R *__restrict__ c = tab + ri * P.nrow + sh;
      C o;
      o.x = R(0);
      o.y = R(0);
#pragma unroll
      for (int k = 0; k < NT; ++k) o = axpy(c[k], s[i * B + k], o);
      if (n0 + i < n_hi) {
        if (staged) ostage[n0 + i - n_lo] = o;
        else op.put(n0 + i, o);
      }
      ri += fr;
      if (ri >= P.pL) { ri -= P.pL; --sh; }
    }
    div += dstep;
    rem += rstep;
    if (rem >= P.pL) { rem -= P.pL; ++div; }
  }
}

// ---------------------------------------------------------------------------
// The kernel.  R: engine real type.  n = 2^LOG2N.  L, M: stage ratio.  POLY: a
// rational polyphase stage (vpoly0) consumes the block from shared memory.
// ---------------------------------------------------------------------------
// Threads per CTA: G independent groups of T = n/16 threads (own exchange buffer and U, named
// barriers) share one coefficient table, so that more warps fit next to it on an SM.
template <typename R, int LOG2N> struct FftCta {
  static constexpr int T = (1 << LOG2N) / 16;
  static constexpr int G = sizeof(R) == 4 ? (T >= 256 ? 1 : 256 / T) : 1;
  static constexpr int THREADS = T * G;
  static constexpr int MIN_CTAS = sizeof(R) == 4 ? 2 : 1;   // f32: <= 128 registers per thread
};

template <int G, int T> __device__ __forceinline__ void group_sync(int grp) {
  if (G == 1) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(T) : "memory");
}

template <typename R, int LOG2N, int L, int M, bool POLY>
__global__ void __launch_bounds__(FftCta<R, LOG2N>::THREADS, FftCta<R, LOG2N>::MIN_CTAS) k_fft_stage(const FftKParams P) {
  typedef typename V2<R>::type C;
  constexpr int N = 1 << LOG2N, R1 = N / 256, T = N / 16, NB1 = 16 / R1, G = FftCta<R, LOG2N>::G;
  constexpr int EB = int(sizeof(C));
  constexpr int BUF_ELEMS = N + (N >> 4) + (EB == 8 ? (N >> 8) : (N >> 7)) + 8;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int grp = G == 1 ? 0 : int(threadIdx.x) / T;
  const int t = G == 1 ? int(threadIdx.x) : int(threadIdx.x) % T;
  C *buf = reinterpret_cast<C *>(smem_raw) + size_t(grp) * (BUF_ELEMS + P.u_elems);
  C *U = buf + BUF_ELEMS;
  R *tab = reinterpret_cast<R *>(reinterpret_cast<C *>(smem_raw) + size_t(G) * (BUF_ELEMS + P.u_elems));
  __shared__ __align__(8) unsigned long long tab_bar;

  const C *__restrict__ tw2 = static_cast<const C *>(P.tw2);
  const C *__restrict__ tw3 = static_cast<const C *>(P.tw3) + t;
  const C *__restrict__ spec = static_cast<const C *>(P.spec) + t;

  if (POLY) {   // K8's table: staged once per CTA with a TMA bulk copy, waited for before the first use
    if (threadIdx.x == 0) {
      mbar_init(&tab_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&tab_bar, P.ptab_bytes);
      tma_bulk_g2s(tab, P.ptab, P.ptab_bytes, &tab_bar);
    }
  }
  bool tab_ready = !POLY;
  if (POLY) {   // slack behind the block: over-read (with zero coefficients) by the polyphase stage
    for (int i = L * (P.delta + P.v_base) + t; i < P.u_elems; i += T) {
      C z;
      z.x = R(0);
      z.y = R(0);
      U[i] = z;
    }
  }

  // per-thread shared-memory bases of the three exchange patterns (see physc)
  const int j2 = t % R1, G2 = t / R1;
  C *const b2p = buf + physc<EB>(G2 * T + j2);     // pass 2: + physc(R1 * m)
  C *const b3p = buf + physc<EB>(t);               // pass 3: + physc(T * m)
  C *b1p[NB1];                                     // pass 1: + q1
#pragma unroll
  for (int u = 0; u < NB1; ++u) {
    const int g = t + T * u;
    b1p[u] = buf + physc<EB>(R1 * (g >> 4) + T * (g & 15));
  }
  tw2 += j2;

  const long LV = long(L) * P.v_base;
  const long long n_items = (long long)P.n_pairs * P.n_blocks;

  for (long long item = (long long)blockIdx.x * G + grp; item < n_items; item += (long long)gridDim.x * G) {
    const unsigned pair = unsigned(item / P.n_blocks);
    const long b = P.b_first + long(item % P.n_blocks);
    // signal 2cp of clip `clip` in the real part, 2cp+1 (if the clip has it) in the imaginary part
    const unsigned clip = pair / P.ppg, cp = pair % P.ppg;
    const unsigned s0 = clip * P.group + 2 * cp, s1 = s0 + 1;
    const bool has1 = 2 * cp + 1 < P.group;
    const PairIn<R> pin = make_pair_in<R>(P.in, s0, has1);
    const long k0 = P.k0 + b * long(M) * P.v_base;   // input index of point 0, phase 0
    // the whole block (and its halo-correction reads) inside the valid input: no per-sample checks
    const bool interior = pin.vec && k0 - long(M) * P.delta >= pin.vlo && k0 + long(M) * N <= pin.vhi;

    // ---- halo correction terms (threads < L*delta), see file header -------------
    C corr;
    corr.x = R(0);
    corr.y = R(0);
    if (POLY && b > 0 && t < L * P.delta) {
      const int r = t / P.delta, ii = t % P.delta - P.delta;   // i' in [-delta, 0)
      const R *__restrict__ edge = static_cast<const R *>(P.edge);
      for (int s = 0; s < M; ++s)
        for (int e = 0; e < -ii; ++e) {       // m = -e, i' - m = ii + e < 0
          const R gv = edge[(r * M + s) * P.delta + e];
          const C a = pin.load(k0 + long(M) * (ii + e) + s);
          const C w = pin.load(k0 + long(M) * (N + ii + e) + s);
          corr.x += gv * (a.x - w.x);
          corr.y += gv * (a.y - w.y);
        }
    }

    // ---- forward transforms: Z[s][q] = spectrum bin t + T q of input phase s -----
    C Z[M][16];
#pragma unroll
    for (int s = 0; s < M; ++s) {
      C v[16];
      if (interior) {
        const float2 *__restrict__ src =
            reinterpret_cast<const float2 *>(static_cast<const float *>(pin.p) + pin.off0 + (k0 + s) * pin.fs) ;
        const long long st2 = (long long)M * (pin.fs >> 1);   // float2 elements per block point
#pragma unroll
        for (int u = 0; u < NB1; ++u)
#pragma unroll
          for (int m1 = 0; m1 < R1; ++m1) {
            const float2 x = src[(long long)(t + T * u + 256 * m1) * st2];
            v[u * R1 + m1].x = R(x.x);
            v[u * R1 + m1].y = R(x.y);
          }
      } else {
#pragma unroll
        for (int u = 0; u < NB1; ++u)
#pragma unroll
          for (int m1 = 0; m1 < R1; ++m1) v[u * R1 + m1] = pin.load(k0 + long(M) * (t + T * u + 256 * m1) + s);
      }
      group_sync<G, T>(grp);   // buf (and U) free: previous item / phase finished reading
#pragma unroll
      for (int u = 0; u < NB1; ++u) {
        C a[R1];
#pragma unroll
        for (int m1 = 0; m1 < R1; ++m1) a[m1] = v[u * R1 + m1];
        Dft<R1>::template run<false>(a);
#pragma unroll
        for (int q1 = 0; q1 < R1; ++q1) b1p[u][q1] = a[q1];
      }
      group_sync<G, T>(grp);
#pragma unroll
      for (int m = 0; m < 16; ++m) v[m] = b2p[physc<EB>(R1 * m)];
#pragma unroll
      for (int m = 1; m < 16; ++m) v[m] = cmul(v[m], tw2[m * R1]);
      Dft<16>::template run<false>(v);
#pragma unroll
      for (int q = 0; q < 16; ++q) b2p[physc<EB>(R1 * q)] = v[q];
      group_sync<G, T>(grp);
#pragma unroll
      for (int m = 0; m < 16; ++m) v[m] = b3p[physc<EB>(T * m)];
#pragma unroll
      for (int m = 1; m < 16; ++m) v[m] = cmul(v[m], tw3[m * T]);
      Dft<16>::template run<false>(v);
#pragma unroll
      for (int q = 0; q < 16; ++q) Z[s][q] = v[q];
    }

    // ---- per output phase: spectrum product, inverse transform into U ------------
#pragma unroll
    for (int r = 0; r < L; ++r) {
      C v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        C acc = cmul(Z[0][q], spec[((r * M + 0) * 16 + q) * T]);
#pragma unroll
        for (int s = 1; s < M; ++s) {
          acc = cadd(acc, cmul(Z[s][q], spec[((r * M + s) * 16 + q) * T]));
        }
        v[q] = acc;
      }
      Dft<16>::template run<true>(v);
#pragma unroll
      for (int m = 1; m < 16; ++m) v[m] = cmulc(v[m], tw3[m * T]);
      group_sync<G, T>(grp);   // previous readers of buf are done
#pragma unroll
      for (int m = 0; m < 16; ++m) b3p[physc<EB>(T * m)] = v[m];
      group_sync<G, T>(grp);
#pragma unroll
      for (int q = 0; q < 16; ++q) v[q] = b2p[physc<EB>(R1 * q)];
      Dft<16>::template run<true>(v);
#pragma unroll
      for (int m = 1; m < 16; ++m) v[m] = cmulc(v[m], tw2[m * R1]);
#pragma unroll
      for (int m = 0; m < 16; ++m) b2p[physc<EB>(R1 * m)] = v[m];
      group_sync<G, T>(grp);
      C *const Ur = U + P.delta * L + r;
#pragma unroll
      for (int u = 0; u < NB1; ++u) {
        C a[R1];
        const int g = t + T * u;
#pragma unroll
        for (int q1 = 0; q1 < R1; ++q1) a[q1] = b1p[u][q1];
        Dft<R1>::template run<true>(a);
#pragma unroll
        for (int m1 = 0; m1 < R1; ++m1) {
          const int i = g + 256 * m1;            // block-local output index i'
          if (i < P.v_base) Ur[i * L] = a[m1];
          else if (POLY && i >= N - P.delta) Ur[(i - N) * L] = a[m1];   // circular tail = halo (uncorrected)
        }
      }
    }
    group_sync<G, T>(grp);

    const long J0 = b * LV;   // first valid output of this block (output-rate index of the FFT stage)
    const long long ooff0 = (long long)(s0 / P.out.channels) * P.out.clip_stride +
                            (long long)(s0 % P.out.channels) * P.out.chan_stride - P.out.base * P.out.frame_stride;
    const long long ooff1 = has1 ? (long long)(s1 / P.out.channels) * P.out.clip_stride +
                                       (long long)(s1 % P.out.channels) * P.out.chan_stride - P.out.base * P.out.frame_stride
                                 : ooff0;
    const bool ovec = P.out.dtype == 0 && has1 && ooff1 == ooff0 + 1 && !(ooff0 & 1) && !(P.out.frame_stride & 1) &&
                      !(reinterpret_cast<unsigned long long>(P.out.p) & 7);
    if (POLY) {
      // halo: add the wrap-around correction; the intermediate does not exist before index 0 (cr.c:487-490)
      if (t < L * P.delta) {
        const int r = t / P.delta, hi = t % P.delta;
        C h = U[hi * L + r];
        h.x += corr.x;
        h.y += corr.y;
        if (b == 0) { h.x = R(0); h.y = R(0); }
        U[hi * L + r] = h;
      }
      if (!tab_ready) {
        mbar_wait(&tab_bar, 0);
        tab_ready = true;
      }
      group_sync<G, T>(grp);
      const long long Q0 = (long long)J0 - P.nc + 1 + P.p_pre, Q1 = Q0 + LV;
      long n_lo = b == 0 ? 0 : long(max(0LL, ceil_div_ll(Q0 * P.pL - P.p_at0, P.pM)));
      long n_hi = long(max(0LL, ceil_div_ll(Q1 * P.pL - P.p_at0, P.pM)));
      n_lo = max(n_lo, P.o0);
      n_hi = min(n_hi, P.o1);
      const OutPair<R> op{&P.out, ooff0, ooff1, has1, ovec};
      const long ubase = J0 - long(L) * P.delta;
      const bool staged = n_hi - n_lo <= BUF_ELEMS;   // buf is free now: stage results for coalesced stores
      bool done = false;
      if (sizeof(R) == 4 && M == 1 && L <= 2) {
        done = true;
        switch (P.pv) {
          case 1: poly_emit_slide<R, T, 7, 2, 22>(P, U, tab, buf, staged, n_lo, n_hi, ubase, t, op); break;
          case 2: poly_emit_slide<R, T, 7, 2, 16>(P, U, tab, buf, staged, n_lo, n_hi, ubase, t, op); break;
          case 3: poly_emit_slide<R, T, 5, 1, 24>(P, U, tab, buf, staged, n_lo, n_hi, ubase, t, op); break;
          case 4: poly_emit_slide<R, T, 6, 1, 16>(P, U, tab, buf, staged, n_lo, n_hi, ubase, t, op); break;
          case 5: poly_emit_slide<R, T, 3, 2, 14>(P, U, tab, buf, staged, n_lo, n_hi, ubase, t, op); break;
          default: done = false;
        }
      }
      if (!done) poly_emit<R, T>(P, U, tab, buf, staged, n_lo, n_hi, ubase, t, op);
      if (staged) {
        group_sync<G, T>(grp);
        for (long nn = n_lo + t; nn < n_hi; nn += T) op.put(nn, buf[nn - n_lo]);
      }
    } else {
      // ---- emit the block: out[J0 + L i' + r], interleaving the L output phases ----
      const long lo = max(J0, P.o0), hi = min(J0 + LV, P.o1);
      for (long j = lo + t; j < hi; j += T) {
        const int idx = int(j - J0);
        store_pair<R>(P.out, ooff0, ooff1, has1, ovec, j, U[idx]);
      }
    }
  }
  if (POLY && !tab_ready) mbar_wait(&tab_bar, 0);   // never exit with a bulk copy in flight
}

#endif  // __CUDACC__

// ---- host side ---------------------------------------------------------------
const char *fft_stage_prepare(DevicePlan *dp);

struct RangeLH { long lo, hi; };
// Input range (of the stage feeding the fused launch) needed for final outputs [lo, hi) of the launch.
RangeLH fft_stage_input_range(const FftStageInfo &fi, long lo, long hi);

// `group`: signals per clip; two signals share a complex FFT only inside one clip.
const char *fft_stage_run(const DevicePlan &dp, const FftStageInfo &fi, const SigDesc &in, const SigDesc &out,
                          long o0, long o1, unsigned n_sig, unsigned group, cudaStream_t st);

}  // namespace b2
