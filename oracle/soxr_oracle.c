/* oracle/soxr_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the sample-level arithmetic of libsoxr's constant-rate
 * path, in double precision and direct (time-domain) form.  The stage *plan*
 * and the filters are produced by oracle/soxr_oracle.py; this file only holds
 * the inner loops, each citing the reference code it restates.  PFFFT
 * overlap-save (cr.c:68-174) computes exactly the same linear operator as
 * orc_rational_stage(); only the rounding differs.
 *
 * Pinned against the real reference (oracle/_ref/libsoxr_ref.so) by
 * tests/test_oracle_vs_ref.py and against tests/golden/ fixtures.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stddef.h>

/* out[n] = sum_k g[C + M*n - L*k] * in[k],  in[k] = 0 outside [0, n_in),
 * for n in [n0, n0 + n_out).
 * Closed form of one zero-phase L/M stage:
 *   - DFT stage        cr.c:68-174 (block convolution), cr.c:205-210 (filter placement),
 *                      cr.c:240-241 (preload/at => C = taps-1-post_peak)
 *   - rational vpoly0  poly-fir0.h:24-44, table layout cr.c:31-59
 *                      (C = (nc-1-preload)*L - 1 + at0, preload = (nc-1)>>1, cr.c:428,447)
 *   - half-band        half-fir.h:38-50 expressed as a symmetric 4n-1.. tap filter. */
void orc_rational_stage(const double *in, long n_in, const double *g, long n_g,
                        long C, long L, long M, double *out, long n0, long n_out)
{
  for (long i = 0; i < n_out; ++i) {
    long n = n0 + i;
    long top = C + M * n;              /* g index at k = 0 */
    /* need 0 <= top - L*k < n_g  =>  k in ((top - n_g)/L, top/L] */
    long k_hi = top >= 0 ? top / L : -1;
    long t = top - (n_g - 1);
    long k_lo = t <= 0 ? 0 : (t + L - 1) / L;
    if (k_hi > n_in - 1) k_hi = n_in - 1;
    long double acc = 0;
    for (long k = k_lo; k <= k_hi; ++k)
      acc += (long double)g[top - L * k] * in[k];
    out[i] = (double)acc;
  }
}

/* Interpolated-coefficient polyphase stage, standard-precision 32.32 clock.
 * Restates stdPrecCore of poly-fir.h:89-101 with the scalar coefficient layout
 * coef() of cr.h:23-27: T[phase][tap][order - interp_num].
 *   at.whole starts at at0 (64-bit fixed point) and advances by `step` per
 *   output; input index = at >> 32 (into the stage's FIFO stream, i.e. the
 *   stage input shifted by `preload` zeros, cr.c:447,487-490);
 *   phase = frac >> (32 - phase_bits);  x = (uint32_t)(frac << phase_bits) * 2^-32.
 * in_s is the FIFO stream (preload zeros already prepended, zero-extended by caller).
 * Produces outputs for output indices [n0, n0+n_out). */
void orc_interp_poly_stage(const double *in_s, long n_s, const double *T, int n_taps,
                           int order, int phase_bits, uint64_t at0, uint64_t step,
                           double *out, long n0, long n_out)
{
  for (long i = 0; i < n_out; ++i) {
    uint64_t at = at0 + step * (uint64_t)(n0 + i);
    long pos = (long)(at >> 32);
    uint32_t frac = (uint32_t)at;
    int phase = (int)(frac >> (32 - phase_bits));
    double x = (double)(uint32_t)(frac << phase_bits) * (1. / 4294967296.);
    const double *c = T + (size_t)phase * n_taps * (order + 1);
    long double acc = 0;
    for (int j = 0; j < n_taps; ++j) {
      const double *cj = c + (size_t)j * (order + 1);
      /* cj[0] is the highest-order coef (interp_num = order), cj[order] = a. */
      double v = cj[0];
      for (int o = 1; o <= order; ++o) v = v * x + cj[o];
      double s = (pos + j >= 0 && pos + j < n_s) ? in_s[pos + j] : 0.;
      acc += (long double)v * s;
    }
    out[i] = (double)acc;
  }
}

/* `Quick' cubic stage (quality QQ).  Restates cubic_stage_fn, cr-core.c:49-67:
 *   s = input + at.integer;  x = at.fraction / 2^32
 *   b = .5*(s[1]+s[-1]) - s[0];  a = (1/6)*(s[2]-s[1]+s[-1]-s[0]-4b);  c = s[1]-s[0]-a-b
 *   out = mult * (((a x + b) x + c) x + s[0])
 * in_s: FIFO stream with pre = preload = 1 zero prepended; s points at in_s+1. */
void orc_cubic_stage(const double *in_s, long n_s, uint64_t step, double mult,
                     double *out, long n0, long n_out)
{
#define S(k) (((k) >= 0 && (k) < n_s) ? in_s[(k)] : 0.)
  for (long i = 0; i < n_out; ++i) {
    uint64_t at = step * (uint64_t)(n0 + i);
    long p = (long)(at >> 32) + 1;
    double x = (double)(uint32_t)at * (1. / 4294967296.);
    double s0 = S(p), s1 = S(p + 1), s2 = S(p + 2), sm = S(p - 1);
    double b = .5 * (s1 + sm) - s0, a = (1 / 6.) * (s2 - s1 + sm - s0 - 4 * b);
    double c = s1 - s0 - a - b;
    out[i] = mult * (((a * x + b) * x + c) * x + s0);
  }
#undef S
}

/* ---- output conversion: round-half-even, saturate, clip count, TPDF dither ----
 * Restates rint-clip.h as compiled on x86-64 gcc (FPU_RINT path: `fistp` =
 * round-to-nearest-even; invalid => saturate), SURVEY.md Appendix C.
 *   LCG:    seed = A*seed + C (rint-clip.h:7), ran = seed >> 3, then per sample
 *           ran1 >>= 3, ran2 >>= 3, d = src + (1/32)*(int)((ran1&31)-(ran2&31))
 *   salt:   seed = *seed0 ^ n ^ bits32(src[n/2])   (rint-clip.h:11-13)
 *   groups: a fresh ran1/ran2 pair per 16 samples (rint-clip.h:64-75); a group
 *           with any overflow is redone by RINT_CLIP which re-salts from the
 *           pre-group seed with (n=i, src[i/2]) (rint-clip.h:37-54,70-73);
 *           the tail (n%16) goes through RINT_CLIP too (one pair for the tail).
 * `bits32` of a double source is the LOW word (union pun, rint-clip.h:12-13).
 */
#define LCG_A 0xd1342543de82ef95ULL
#define LCG_C 0x6a09e667f3bcc909ULL

static uint32_t bits32_of(const void *src, int is_double, size_t idx)
{
  uint32_t u;
  if (is_double) memcpy(&u, (const double *)src + idx, 4);   /* low word, little-endian */
  else memcpy(&u, (const float *)src + idx, 4);
  return u;
}

static double src_at(const void *src, int is_double, size_t i)
{
  return is_double ? ((const double *)src)[i] : (double)((const float *)src)[i];
}

/* fistp semantics: nearest-even in the current (default) rounding mode;
 * out-of-range or NaN raises invalid. Returns 1 on overflow. */
static int rint_sat(double d, long vmax, long *res)
{
  double r = nearbyint(d);
  if (!(r >= -(double)vmax - 1 && r <= (double)vmax)) {
    *res = d > 0 ? vmax : -vmax - 1;
    return 1;
  }
  *res = (long)r;
  return 0;
}

static void put(void *dest, int bits, size_t idx, long v)
{
  if (bits == 16) ((int16_t *)dest)[idx] = (int16_t)v;
  else ((int32_t *)dest)[idx] = (int32_t)v;
}

/* RINT_CLIP, rint-clip.h:37-54 */
static void rint_clip(void *dest, int bits, const void *src, int is_double,
                      unsigned stride, size_t i, size_t n, size_t *clips,
                      uint64_t *seed0 /* NULL: no dither */)
{
  long vmax = bits == 16 ? 32767 : 2147483647L;
  uint64_t seed = 0, ran1 = 0, ran2 = 0;
  if (seed0) {
    seed = *seed0 ^ n ^ (n ? bits32_of(src, is_double, n / 2) : 0);
    ran1 = (seed = LCG_A * seed + LCG_C) >> 3;
    ran2 = (seed = LCG_A * seed + LCG_C) >> 3;
  }
  for (; i < n; ++i) {
    double d = src_at(src, is_double, i);
    if (seed0) {
      ran1 >>= 3; ran2 >>= 3;
      d = d + (1. / 32) * (int)((ran1 & 31) - (ran2 & 31));
    } else if (!is_double) {
      /* FLOATD = FLOATX: no widening needed; value identical */
    }
    long v;
    if (rint_sat(d, vmax, &v)) ++*clips;
    put(dest, bits, stride * i, v);
  }
  if (seed0) *seed0 = seed;
}

/* LSX_RINT_CLIP_2 (and LSX_RINT_CLIP for ch==1), rint-clip.h:58-140.
 * srcs: array of `ch` planar channel pointers (float or double), n frames each.
 * dest: interleaved int16/int32 [n][ch].  seed0 NULL => no dither.
 * Returns clip count. */
size_t orc_rint_clip_interleave(void *dest, int bits, const void *const *srcs, int is_double,
                                unsigned ch, size_t n, uint64_t *seed0)
{
  size_t clips = 0;
  long vmax = bits == 16 ? 32767 : 2147483647L;
  uint64_t seed = 0;
  if (seed0) seed = *seed0 ^ n ^ (n ? bits32_of(srcs[0], is_double, n / 2) : 0);
  for (unsigned j = 0; j < ch; ++j) {
    const void *src = srcs[j];
    void *d0 = bits == 16 ? (void *)((int16_t *)dest + j) : (void *)((int32_t *)dest + j);
    size_t i = 0;
    while (i < (n & ~(size_t)15)) {
      uint64_t seed1 = seed, ran1 = 0, ran2 = 0;
      int invalid = 0;
      if (seed0) {
        ran1 = (seed = LCG_A * seed + LCG_C) >> 3;
        ran2 = (seed = LCG_A * seed + LCG_C) >> 3;
      }
      for (int k = 0; k < 16; ++k, ++i) {
        double d = src_at(src, is_double, i);
        if (seed0) {
          ran1 >>= 3; ran2 >>= 3;
          d = d + (1. / 32) * (int)((ran1 & 31) - (ran2 & 31));
        }
        long v;
        invalid |= rint_sat(d, vmax, &v);
        put(d0, bits, (size_t)ch * i, v);
      }
      if (invalid)
        rint_clip(d0, bits, src, is_double, ch, i - 16, i, &clips, seed0 ? &seed1 : NULL);
    }
    rint_clip(d0, bits, src, is_double, ch, i, n, &clips, seed0 ? &seed : NULL);
  }
  if (seed0) *seed0 = seed;
  return clips;
}
