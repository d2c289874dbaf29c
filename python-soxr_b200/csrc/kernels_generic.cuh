// kernels_generic.cuh -- position-keyed stage kernels that work for ANY plan.
//
// These are the complete-coverage path: every stage kind the planner can emit
// runs here as a direct-form kernel over planar signals in HBM.  The hot shapes
// (DFT stage [+ rational polyphase]) have a fused shared-memory FFT kernel in
// fft_stage.cuh; both compute the same closed forms (plan.h).
//
// Signals are planar: signal s (= clip*channels + channel) occupies
// buf[s*stride + (k - base)] for logical sample index k.  Reads outside
// [vlo, vhi) return 0 (zero history before the stream start / flush zeros).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b2 {

// element(s, k) = p[(s / channels) * clip_stride + (s % channels) * chan_stride + (k - base) * frame_stride]; planar
// scratch is channels == 1, clip_stride == row stride, frame_stride == 1.  k_rational needs the planar form
// (`stride`); k_poly_interp and k_cubic take any layout of engine-precision samples, so the first / last stage of a
// plan reads / writes the caller's float buffers directly (no K1 / K7 copy pass).
template <typename T>
struct SigIn {
  const T *p;
  size_t stride;                                  // == clip_stride (planar form)
  long base, vlo, vhi;
  long long chan_stride = 0, frame_stride = 1;
  unsigned channels = 1;
  __device__ __forceinline__ const T *sig(unsigned s) const {   // pointer to logical sample 0 of signal s
    return p + (long long)(s / channels) * (long long)stride + (long long)(s % channels) * chan_stride - (long long)base * frame_stride;
  }
  __device__ __forceinline__ T at(unsigned s, long k) const {
    return (k >= vlo && k < vhi) ? sig(s)[(long long)k * frame_stride] : T(0);
  }
};
template <typename T>
struct SigOut {
  T *p;
  size_t stride;
  long base;
  long long chan_stride = 0, frame_stride = 1;
  unsigned channels = 1;
  __device__ __forceinline__ T *sig(unsigned s) const {
    return p + (long long)(s / channels) * (long long)stride + (long long)(s % channels) * chan_stride - (long long)base * frame_stride;
  }
  __device__ __forceinline__ T &at(unsigned s, long n) const { return sig(s)[(long long)n * frame_stride]; }
};

// ---- K1: deinterleave + convert-in (role of data-io.c:13-57: plain C cast, no scaling)
template <typename Tin, typename T>
__global__ void k_convert_in(const Tin *__restrict__ in, long long clip_stride, long long frame_stride,
                             long long chan_stride, unsigned channels, size_t n_frames, unsigned n_sig,
                             T *__restrict__ X, size_t x_stride, size_t x_off) {
  const size_t k = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n_frames) return;
  for (unsigned s = blockIdx.y; s < n_sig; s += gridDim.y) {
    const unsigned b = s / channels, c = s % channels;
    X[size_t(s) * x_stride + x_off + k] = T(in[b * clip_stride + (long long)k * frame_stride + c * chan_stride]);
  }
}

// ---- rational stage, direct form:  out[n] = sum_k g[C + M n - L k] in[k]
// (half-band half-fir.h:38-50, DFT stage cr.c:68-174, vpoly0 poly-fir0.h:24-44)
// A thread produces output n of SPB signals: each tap is loaded once and used SPB times (per signal the order of
// the accumulation is unchanged, so results do not depend on the grouping).
template <typename T, int SPB>
__global__ void k_rational(SigIn<T> in, const T *__restrict__ g, int n_g, long C, int L, int M,
                           SigOut<T> out, long o0, long n_out, unsigned n_sig) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const long n = o0 + i, top = C + long(M) * n;
  long k_hi = top >= 0 ? top / L : -1;
  const long t = top - (n_g - 1);
  long k_lo = t <= 0 ? 0 : (t + L - 1) / L;
  if (k_lo < in.vlo) k_lo = in.vlo;
  if (k_hi > in.vhi - 1) k_hi = in.vhi - 1;
  const T *gp = g + top;
  for (unsigned s0 = blockIdx.y * unsigned(SPB); s0 < n_sig; s0 += gridDim.y * unsigned(SPB)) {
    const T *src = in.p + size_t(s0) * in.stride - in.base;
    double acc[SPB];
#pragma unroll
    for (int u = 0; u < SPB; ++u) acc[u] = 0;
    if (s0 + SPB <= n_sig) {
      for (long k = k_lo; k <= k_hi; ++k) {
        const double gk = double(gp[-long(L) * k]);
#pragma unroll
        for (int u = 0; u < SPB; ++u) acc[u] = fma(gk, double(src[size_t(u) * in.stride + k]), acc[u]);
      }
    } else {
      for (long k = k_lo; k <= k_hi; ++k) {
        const double gk = double(gp[-long(L) * k]);
#pragma unroll
        for (int u = 0; u < SPB; ++u)
          if (s0 + u < n_sig) acc[u] = fma(gk, double(src[size_t(u) * in.stride + k]), acc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < SPB; ++u)
      if (s0 + u < n_sig) out.at(s0 + u, n) = T(acc[u]);
  }
}

// 32.32 fixed-point stage clock at absolute output index n: at0 + step * n evaluated in 96 bits (64-bit integer
// part, 32-bit fraction), so a stream can run past 2^32 stage-input samples.  The reference carries the same
// clock incrementally and rebases its integer part every call (poly-fir.h:74-86); the values agree exactly.
__device__ __forceinline__ void clock_at(unsigned long long at0, unsigned long long step, unsigned long long n,
                                         long long *whole, unsigned *frac) {
  const unsigned long long lo = step * n, hi = __umul64hi(step, n);
  const unsigned long long sum = lo + at0;
  const unsigned long long carry = sum < lo ? 1ull : 0ull;
  *whole = (long long)(((hi + carry) << 32) | (sum >> 32));
  *frac = unsigned(sum);
}

// ---- interpolated-coefficient polyphase stage (vpoly1..3, poly-fir.h:89-101):
// 32.32 fixed-point clock; phase = top `phase_bits` of the fraction, x = the rest.
// table: [phase][tap][order+1], highest order first.
// All signals share one clock, so the interpolated coefficients of output n (n_taps Horner evaluations fed by
// n_taps * (order + 1) scattered table reads, up to 164 kB of table) are computed ONCE per thread, parked in the
// thread's column of shared memory, and applied to `spb` signals (blockIdx.y selects the group).  The per-output
// arithmetic is the same for every signal whatever the grouping (Horner and accumulation in engine precision like
// the reference), so a stream's results do not depend on how it was chunked or batched.
template <typename T>
__global__ void k_poly_interp(SigIn<T> in, const T *__restrict__ table, int n_taps, int order,
                              int phase_bits, unsigned long long at0, unsigned long long step,
                              int preload, SigOut<T> out, long o0, long n_out, unsigned n_sig, unsigned spb) {
  extern __shared__ __align__(16) unsigned char poly_smem[];
  T *vs = reinterpret_cast<T *>(poly_smem) + threadIdx.x;   // v_j at vs[j * blockDim.x]
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const long n = o0 + i;
  long long whole;
  unsigned frac;
  clock_at(at0, step, (unsigned long long)n, &whole, &frac);
  const long pos = long(whole) - preload;
  const int phase = int(frac >> (32 - phase_bits));
  const T x = T(double(unsigned(frac << phase_bits)) * (1. / 4294967296.));
  const T *c = table + size_t(phase) * n_taps * (order + 1);
  for (int j = 0; j < n_taps; ++j) {
    const T *cj = c + j * (order + 1);
    T v = cj[0];
    for (int o = 1; o <= order; ++o) v = v * x + cj[o];   // Horner in engine precision, as the reference
    vs[j * blockDim.x] = v;
  }
  const bool inside = pos >= in.vlo && pos + n_taps <= in.vhi;   // the whole window is real input: no per-tap checks
  const unsigned s_end = min(n_sig, (blockIdx.y + 1u) * spb);
  // Accumulation in engine precision with fused multiply-adds, two interleaved partial sums (the reference's SIMD
  // engines keep four): float32 for the f32 engine -- no float->double conversion per tap on the 16-lane XU pipe,
  // which was half of this kernel's time (profiles/r2s) -- double for the f64 engine.  Four signals share each
  // coefficient read.
  unsigned s = blockIdx.y * spb;
  const int bd = int(blockDim.x);
  if (inside) {
    const long long fs = in.frame_stride;
    for (; s + 4 <= s_end; s += 4) {
      const T *__restrict__ p0 = in.sig(s) + (long long)pos * fs, *__restrict__ p1 = in.sig(s + 1) + (long long)pos * fs;
      const T *__restrict__ p2 = in.sig(s + 2) + (long long)pos * fs, *__restrict__ p3 = in.sig(s + 3) + (long long)pos * fs;
      T e0 = 0, o0 = 0, e1 = 0, o1 = 0, e2 = 0, o2 = 0, e3 = 0, o3 = 0;
      int j = 0;
      for (; j + 1 < n_taps; j += 2) {
        const T v = vs[j * bd], w = vs[(j + 1) * bd];
        e0 = fma(v, p0[j * fs], e0); o0 = fma(w, p0[(j + 1) * fs], o0);
        e1 = fma(v, p1[j * fs], e1); o1 = fma(w, p1[(j + 1) * fs], o1);
        e2 = fma(v, p2[j * fs], e2); o2 = fma(w, p2[(j + 1) * fs], o2);
        e3 = fma(v, p3[j * fs], e3); o3 = fma(w, p3[(j + 1) * fs], o3);
      }
      if (j < n_taps) {
        const T v = vs[j * bd];
        e0 = fma(v, p0[j * fs], e0); e1 = fma(v, p1[j * fs], e1); e2 = fma(v, p2[j * fs], e2); e3 = fma(v, p3[j * fs], e3);
      }
      out.at(s, n) = e0 + o0;
      out.at(s + 1, n) = e1 + o1;
      out.at(s + 2, n) = e2 + o2;
      out.at(s + 3, n) = e3 + o3;
    }
  }
  for (; s < s_end; ++s) {     // (the same order of operations per signal as above)
    T e = 0, o = 0;
    int j = 0;
    if (inside) {
      const T *__restrict__ src = in.sig(s) + (long long)pos * in.frame_stride;
      const long long fs = in.frame_stride;
      for (; j + 1 < n_taps; j += 2) { e = fma(vs[j * bd], src[j * fs], e); o = fma(vs[(j + 1) * bd], src[(j + 1) * fs], o); }
      if (j < n_taps) e = fma(vs[j * bd], src[j * fs], e);
    } else {
      for (; j + 1 < n_taps; j += 2) { e = fma(vs[j * bd], in.at(s, pos + j), e); o = fma(vs[(j + 1) * bd], in.at(s, pos + j + 1), o); }
      if (j < n_taps) e = fma(vs[j * bd], in.at(s, pos + j), e);
    }
    out.at(s, n) = e + o;
  }
}

// ---- `quick' cubic stage (QQ), cr-core.c:49-67
template <typename T>
__global__ void k_cubic(SigIn<T> in, unsigned long long step, double mult, SigOut<T> out, long o0,
                        long n_out, unsigned n_sig) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const long n = o0 + i;
  long long whole;
  unsigned frac;
  clock_at(0ull, step, (unsigned long long)n, &whole, &frac);
  const long p = long(whole);
  const double x = double(frac) * (1. / 4294967296.);
  for (unsigned s = blockIdx.y; s < n_sig; s += gridDim.y) {
    const double s0 = in.at(s, p), s1 = in.at(s, p + 1), s2 = in.at(s, p + 2), sm = in.at(s, p - 1);   // (any layout)
    const double b = .5 * (s1 + sm) - s0, a = (1 / 6.) * (s2 - s1 + sm - s0 - 4 * b), c = s1 - s0 - a - b;
    out.at(s, n) = T(mult * (((a * x + b) * x + c) * x + s0));
  }
}

// ---- K7 (float outputs): interleave + cast (data-io.c:161-170)
template <typename T, typename Tout>
__global__ void k_convert_out_float(const T *__restrict__ Y, size_t y_stride, size_t n_frames,
                                    unsigned channels, unsigned n_sig, Tout *__restrict__ out,
                                    long long clip_stride, long long frame_stride, long long chan_stride,
                                    size_t frame_off) {
  const size_t k = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n_frames) return;
  for (unsigned s = blockIdx.y; s < n_sig; s += gridDim.y) {
    const unsigned b = s / channels, c = s % channels;
    out[b * clip_stride + (long long)(frame_off + k) * frame_stride + c * chan_stride] = Tout(Y[size_t(s) * y_stride + k]);
  }
}

// ---------------------------------------------------------------------------
// K7 (integer outputs): round-half-even, saturate + clip count, TPDF dither.
// Behaviour restated from rint-clip.h as compiled on x86-64 (x87 `fistp`):
// see SURVEY.md Appendix C.
// ---------------------------------------------------------------------------
#define B2_LCG_A 0xd1342543de82ef95ULL
#define B2_LCG_C 0x6a09e667f3bcc909ULL

__host__ __device__ inline unsigned long long lcg_next(unsigned long long s) { return B2_LCG_A * s + B2_LCG_C; }
// seed after k LCG steps, O(log k)
// k LCG steps as one affine map: seed -> ra * seed + rc (the map depends on k alone, so a thread that owns the same
// group index in many signals builds it once)
__host__ __device__ inline void lcg_jump_map(unsigned long long k, unsigned long long *ra_, unsigned long long *rc_) {
  unsigned long long a = B2_LCG_A, c = B2_LCG_C, ra = 1, rc = 0;
  while (k) {
    if (k & 1) { ra = ra * a; rc = rc * a + c; }
    c = c * a + c;
    a = a * a;
    k >>= 1;
  }
  *ra_ = ra;
  *rc_ = rc;
}
__host__ __device__ inline unsigned long long lcg_jump(unsigned long long s, unsigned long long k) {
  unsigned long long ra, rc;
  lcg_jump_map(k, &ra, &rc);
  return ra * s + rc;
}

template <typename T> __device__ inline unsigned bits32(T v);
template <> __device__ inline unsigned bits32<float>(float v) { return __float_as_uint(v); }
template <> __device__ inline unsigned bits32<double>(double v) { return unsigned(__double_as_longlong(v)); }  // low word

// One block of `n` frames produced by one reference soxr_process call.
struct DitherBlock { long long start; long long n; };
// the block list of a call: a few blocks travel by value in the kernel's parameters (no copy, no pinned staging
// write on the stream: a stream chunk or a whole-clip batch has 1-4), longer lists through device memory
struct DitherBlocksPtr {
  const DitherBlock *p;
  __device__ __forceinline__ DitherBlock operator[](unsigned i) const { return p[i]; }
};
struct DitherBlocks8 {   // (a __grid_constant__ parameter: indexed in place, not copied to the thread's stack)
  DitherBlock b[8];
  __device__ __forceinline__ DitherBlock operator[](unsigned i) const { return b[i]; }
};

// Sequential seed pre-pass: one thread per clip walks blocks x channels.
// seeds[(clip*n_blocks + blk)*channels + ch] = {group-start seed, tail seed (salted, before its 2 steps)}
template <typename T, typename BL>
__global__ void k_dither_seeds(const T *__restrict__ Y, size_t y_stride, unsigned channels, unsigned n_clips,
                               const __grid_constant__ BL blocks, unsigned n_blocks,
                               unsigned long long seed0, const unsigned long long *seed_in, bool split,
                               ulonglong2 *__restrict__ seeds, unsigned long long *seed_end) {
  const unsigned clip = blockIdx.x * blockDim.x + threadIdx.x;
  if (clip >= n_clips) return;
  unsigned long long seed = seed_in ? seed_in[clip] : seed0;   // (seed_in may alias seed_end: read before written)
  for (unsigned b = 0; b < n_blocks; ++b) {
    const long long st = blocks[b].start, n = blocks[b].n;
    if (!split)  // LSX_RINT_CLIP_2 salts once per call with channel 0 (rint-clip.h:104)
      seed = seed ^ (unsigned long long)n ^ (n ? bits32<T>(Y[size_t(clip * channels) * y_stride + st + n / 2]) : 0u);
    for (unsigned c = 0; c < channels; ++c) {
      const T *src = Y + size_t(clip * channels + c) * y_stride + st;
      if (split)  // one LSX_RINT_CLIP call per channel (soxr.c:664-677)
        seed = seed ^ (unsigned long long)n ^ (n ? bits32<T>(src[n / 2]) : 0u);
      ulonglong2 e;
      e.x = seed;
      seed = lcg_jump(seed, 2ull * (unsigned long long)(n / 16));
      seed = seed ^ (unsigned long long)n ^ (n ? bits32<T>(src[n / 2]) : 0u);   // tail RINT_CLIP re-salt
      e.y = seed;
      seed = lcg_next(lcg_next(seed));
      seeds[(size_t(clip) * n_blocks + b) * channels + c] = e;
    }
  }
  if (seed_end) seed_end[clip] = seed;
}

template <int BITS> struct IntOf;
template <> struct IntOf<16> { typedef short type; static constexpr long long vmax = 32767; };
template <> struct IntOf<32> { typedef int type; static constexpr long long vmax = 2147483647LL; };

// fistp: round to nearest even; out of range / NaN => "invalid" => saturate.
// No conversion instructions (they run on the 16-lane XU pipe, which was the busiest unit of k_rint_clip: 43 %,
// profiles/r2s): the largest value of a two's-complement type is odd and the smallest even, so with ties to even
// rint(d) is in range exactly for lo - .5 <= d < hi + .5; inside the range d + 1.5 * 2^52 carries rint(d) in the
// low word of its mantissa (two's complement), rounded by the adder itself.
// the common case of rint_sat without its saturated value: a group with an out-of-range sample is redone as a whole
// (rint-clip.h:70-73), so the first attempt only needs to know THAT one was out of range
template <int BITS>
__device__ __forceinline__ bool rint_try(double d, typename IntOf<BITS>::type *dst) {
  const double hi = double(IntOf<BITS>::vmax) + .5, lo = -double(IntOf<BITS>::vmax) - 1.5;
  *dst = typename IntOf<BITS>::type(__double2loint(d + 6755399441055744.0));
  return !(d >= lo && d < hi);
}
template <int BITS>
__device__ __forceinline__ bool rint_sat(double d, typename IntOf<BITS>::type *dst) {
  const double hi = double(IntOf<BITS>::vmax) + .5, lo = -double(IntOf<BITS>::vmax) - 1.5;
  if (!(d >= lo && d < hi)) {
    *dst = typename IntOf<BITS>::type(d > 0 ? IntOf<BITS>::vmax : -IntOf<BITS>::vmax - 1);
    return true;
  }
  *dst = typename IntOf<BITS>::type(__double2loint(d + 6755399441055744.0));
  return false;
}
// (a - b) / 32 for two 5-bit fields, exactly, without an integer-to-double conversion: 2^47 has an ulp of 2^-5, so
// the integer a - b + 1024 written into the low mantissa bits of 2^47 is the double 2^47 + 32 + (a - b) / 32.
__device__ __forceinline__ double tpdf_term(unsigned long long ran1, unsigned long long ran2) {
  const int k = int(unsigned(ran1) & 31u) - int(unsigned(ran2) & 31u);
  return __longlong_as_double(0x42E0000000000000LL + (long long)(k + 1024)) - 140737488355360.0;
}

// One 16-sample dither group of one signal (LSX_RINT_CLIP's DO_16 body, rint-clip.h:66-74): v[k] for samples
// i0 .. i0 + 15 of `src`, from the group's pre-draw seed.  Returns the clips it counted.
template <typename T, int BITS>
__device__ __forceinline__ unsigned k7_group(const T *__restrict__ src, long long i0, bool dither, unsigned long long seed1,
                                             typename IntOf<BITS>::type (&v)[16]) {
  typedef typename IntOf<BITS>::type I;
  // the group's 16 samples: 16-byte vector loads when the layout allows (float scratch rows are 64-sample aligned,
  // so this is the common case for whole-clip batches)
  T x[16];
  const T *sp = src + i0;
  if (sizeof(T) == 4 && !(reinterpret_cast<unsigned long long>(sp) & 15)) {
    const float4 *s4 = reinterpret_cast<const float4 *>(sp);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 t = s4[k];
      x[4 * k] = T(t.x); x[4 * k + 1] = T(t.y); x[4 * k + 2] = T(t.z); x[4 * k + 3] = T(t.w);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = sp[k];
  }
  // First attempt: values only.  A sample can leave the range only from within 1 (the dither is below 1) of its
  // ends, so one comparison in the source precision, with a generous margin, says whether the exact range check
  // (two double comparisons per sample) is needed at all; NaNs fail it too.
  const T near_end = T(double(IntOf<BITS>::vmax) - 256.0);
  bool maybe = false;
  {
    unsigned long long r1 = 0, r2 = 0;
    if (dither) { r1 = lcg_next(seed1); r2 = lcg_next(r1) >> 3; r1 >>= 3; }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      maybe |= !(fabs(x[k]) < near_end);
      double d = double(x[k]);
      if (dither) {
        r1 >>= 3; r2 >>= 3;
        d = d + tpdf_term(r1, r2);
      }
      v[k] = I(__double2loint(d + 6755399441055744.0));
    }
  }
  if (!maybe) return 0;
  bool invalid = false;
  {
    unsigned long long r1 = 0, r2 = 0;
    if (dither) { r1 = lcg_next(seed1); r2 = lcg_next(r1) >> 3; r1 >>= 3; }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      double d = double(x[k]);
      if (dither) {
        r1 >>= 3; r2 >>= 3;
        d = d + tpdf_term(r1, r2);
      }
      invalid |= rint_try<BITS>(d, &v[k]);
    }
  }
  if (!invalid) return 0;
  // group redone by RINT_CLIP from the pre-group seed, re-salted (rint-clip.h:37-54,70-73)
  unsigned n_clips = 0;
  const long long ne = i0 + 16;
  unsigned long long ran1 = 0, ran2 = 0;
  if (dither) {
    unsigned long long t = seed1 ^ (unsigned long long)ne ^ bits32<T>(src[ne / 2]);
    t = lcg_next(t); ran1 = t >> 3;
    t = lcg_next(t); ran2 = t >> 3;
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    double d = double(x[k]);
    if (dither) {
      ran1 >>= 3; ran2 >>= 3;
      d = d + tpdf_term(ran1, ran2);
    }
    if (rint_sat<BITS>(d, &v[k])) ++n_clips;
  }
  return n_clips;
}

// the samples after the last whole group: one RINT_CLIP pass with one ran pair (rint-clip.h:75)
template <typename T, int BITS>
__device__ __forceinline__ unsigned k7_tail(const T *__restrict__ src, long long i_from, long long n, bool dither,
                                            unsigned long long seed_tail, typename IntOf<BITS>::type *dst,
                                            long long frame_stride) {
  typedef typename IntOf<BITS>::type I;
  unsigned long long ran1 = 0, ran2 = 0;
  if (dither) {
    unsigned long long t = lcg_next(seed_tail);
    ran1 = t >> 3;
    t = lcg_next(t);
    ran2 = t >> 3;
  }
  unsigned n_clips = 0;
  for (long long i = i_from; i < n; ++i) {
    double d = double(src[i]);
    if (dither) {
      ran1 >>= 3; ran2 >>= 3;
      d = d + tpdf_term(ran1, ran2);
    }
    I v;
    if (rint_sat<BITS>(d, &v)) ++n_clips;
    dst[i * frame_stride] = v;
  }
  return n_clips;
}

template <int BITS> struct IntPairOf;
template <> struct IntPairOf<16> { typedef short2 type; };
template <> struct IntPairOf<32> { typedef int2 type; };

// One thread per (unit, block, 16-sample group); group index == n/16 is the tail.  A unit is one signal, or -- PAIR,
// for interleaved outputs with an even channel count -- the two neighbouring channels (c, c + 1) of a clip, whose
// results leave as one 4- / 8-byte element per frame (16-byte vectors for plain stereo) instead of strided 2- / 4-byte
// stores: every channel keeps its own dither sequence, only the stores are shared.
template <typename T, int BITS, typename BL, bool PAIR>
__global__ void __launch_bounds__(128, 6) k_rint_clip(const T *__restrict__ Y, size_t y_stride, unsigned channels, unsigned n_sig,
                            const __grid_constant__ BL blocks, unsigned n_blocks, unsigned max_groups,
                            const ulonglong2 *__restrict__ seeds /* null: no dither */,
                            typename IntOf<BITS>::type *__restrict__ out, long long clip_stride,
                            long long frame_stride, long long chan_stride, long long frame_off,
                            unsigned long long *__restrict__ clips) {
  typedef typename IntOf<BITS>::type I;
  const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned b = blockIdx.y;
  if (g >= max_groups) return;
  const long long st = blocks[b].start, n = blocks[b].n;
  const long long full = n / 16;
  if (g > full) return;
  unsigned my_clips = 0;
  const bool dither = seeds != nullptr;
  // this thread's jump from a block's first seed to its group's: the same affine map in every signal it visits
  unsigned long long jump_a = 1, jump_c = 0;
  if (dither && g < full) lcg_jump_map(2ull * g, &jump_a, &jump_c);
  const long long i0 = (long long)g * 16;
  constexpr unsigned SPU = PAIR ? 2u : 1u;   // signals per unit
  for (unsigned s = blockIdx.z * SPU; s < n_sig; s += gridDim.z * SPU) {
    const unsigned clip = s / channels, c = s % channels;
    const T *src = Y + size_t(s) * y_stride + st;
    I *dst = out + clip * clip_stride + (frame_off + st) * frame_stride + c * chan_stride;
    const ulonglong2 *sdp = seeds + (size_t(clip) * n_blocks + b) * channels + c;
    if (g < full) {
      I v[16];
      my_clips += k7_group<T, BITS>(src, i0, dither, dither ? jump_a * sdp[0].x + jump_c : 0ull, v);
      I *dp = dst + i0 * frame_stride;
      if (PAIR) {
        typedef typename IntPairOf<BITS>::type I2;
        I w[16];
        my_clips += k7_group<T, BITS>(src + y_stride, i0, dither, dither ? jump_a * sdp[1].x + jump_c : 0ull, w);
        if (frame_stride == 2 && !(reinterpret_cast<unsigned long long>(dp) & 15)) {
          constexpr int NV = int(32 * sizeof(I) / 16);
          uint4 pk[NV];
          I *pv = reinterpret_cast<I *>(pk);
#pragma unroll
          for (int k = 0; k < 16; ++k) { pv[2 * k] = v[k]; pv[2 * k + 1] = w[k]; }
#pragma unroll
          for (int k = 0; k < NV; ++k) reinterpret_cast<uint4 *>(dp)[k] = pk[k];
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            I2 e;
            e.x = v[k]; e.y = w[k];
            *reinterpret_cast<I2 *>(dp + k * frame_stride) = e;
          }
        }
      } else if (frame_stride == 1 && !(reinterpret_cast<unsigned long long>(dp) & 15)) {
        constexpr int NV = int(16 * sizeof(I) / 16);
        uint4 pk[NV];
        I *pv = reinterpret_cast<I *>(pk);
#pragma unroll
        for (int k = 0; k < 16; ++k) pv[k] = v[k];
#pragma unroll
        for (int k = 0; k < NV; ++k) reinterpret_cast<uint4 *>(dp)[k] = pk[k];
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) dp[k * frame_stride] = v[k];
      }
    } else {
      my_clips += k7_tail<T, BITS>(src, full * 16, n, dither, dither ? sdp[0].y : 0ull, dst, frame_stride);
      if (PAIR) my_clips += k7_tail<T, BITS>(src + y_stride, full * 16, n, dither, dither ? sdp[1].y : 0ull, dst + chan_stride, frame_stride);
    }
  }
  if (my_clips) atomicAdd(clips, (unsigned long long)my_clips);
}

}  // namespace b2
