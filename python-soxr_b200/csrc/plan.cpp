// plan.cpp -- quality presets, stage planner, filter/table construction and the
// integer-only emulation of the reference's call schedule.
//
// Behavioural contract (must match the reference plan-for-plan; checked by
// tests/test_plan.py against SOXR_TRACE dumps of the real libsoxr):
//   presets          libsoxr/src/soxr.c:106-145
//   ratio decomposition + attenuation budget   libsoxr/src/cr.c:277-363
//   DFT stage parameters                       libsoxr/src/cr.c:178-248
//   polyphase stage parameters + tables        libsoxr/src/cr.c:31-59,398-476
//   half-band sets                             libsoxr/src/cr-core.c:137-149, half-coefs.h
//   pull scheduler                             libsoxr/src/cr.c:498-581
#include "plan.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "design.h"

namespace b2 {

namespace {
constexpr double kTwo32 = 4294967296.0;
inline double to_db(double x) { return std::log10(x) * 20; }
inline bool is_pow2(int x) { return x >= 2 && !(x & (x - 1)); }
inline float f32(double v) { return float(v); }

constexpr unsigned long kRolloffMedium = 1, kRolloffNone = 2;
constexpr unsigned long kResetOnClear = 1ul << 31;
}  // namespace

QSpec quality_spec(unsigned long recipe, unsigned long flags) {
  QSpec s;
  const unsigned q = unsigned(recipe & 0xf);
  if (q > 7) {
    // python-soxr only ever passes QQ..VHQ (src/soxr/__init__.py:29-35); the libsamplerate-style recipes (8..10) and
    // SOXR_PRECISIONQ are not provided
    s.error = (q > 11) ? "invalid quality type" : "quality recipe not supported by b200soxr";
    return s;
  }
  s.flags = flags | kResetOnClear;
  static const double kPhase[4] = {50, 25, 100, 0};   // "\\62\\31\\144"[(recipe & 0x30) >> 4], soxr.c:118 (index 3 is the NUL)
  s.phase_response = kPhase[(recipe & 0x30) >> 4];
  s.stopband_begin = 1;
  s.precision = q == 0 ? 0 : q <= 3 ? 16 : 4 + q * 4;
  const double rej = s.precision * to_db(2.);
  s.passband_end = q == 1 ? 1385 / 2048. : 1 - .05 / to_3db(rej);
  if (q <= 2) s.flags = (s.flags & ~kRolloffNone) | kRolloffMedium;
  if (recipe & 0x40) s.passband_end = 1 - .01 / to_3db(rej);   // SOXR_STEEP_FILTER (soxr.c:142-143)
  return s;
}

// ---------------------------------------------------------------------------
// tables of the reference that are data, not code
// ---------------------------------------------------------------------------
namespace {

struct HalfSet { int n; float att; const double *c; };
const double kHalf8[] = {3.1154652365332069e-01, -8.7344917685739543e-02, 3.6814458353637280e-02,
                         -1.5189204581464479e-02, 5.4540855610738801e-03, -1.5643862626630416e-03,
                         3.1816575906323303e-04, -3.4799449225005688e-05};
const double kHalf9[] = {3.1227034755311189e-01, -8.9221517147969526e-02, 3.9139704015071934e-02,
                         -1.7250558515852023e-02, 6.8589440230476112e-03, -2.3045049636430419e-03,
                         6.0963740543348963e-04, -1.1323803957431231e-04, 1.1197769991000046e-05};
const double kHalf10[] = {3.1285456012000523e-01, -9.0756740799292787e-02, 4.1096398104193160e-02,
                          -1.9066319572525220e-02, 8.1840569787684902e-03, -3.0766876176359834e-03,
                          9.6396524429277980e-04, -2.3585679989922018e-04, 4.0252189026627833e-05,
                          -3.6298196342497932e-06};
const double kHalf11[] = {3.1333588822574199e-01, -9.2035898673019811e-02, 4.2765169698406408e-02,
                          -2.0673580894964429e-02, 9.4225426824512421e-03, -3.8563379950013192e-03,
                          1.3634742159642453e-03, -3.9874150714431009e-04, 9.0586723632664806e-05,
                          -1.4285617244076783e-05, 1.1834642946400529e-06};
const double kHalf12[] = {3.1373928463345568e-01, -9.3118180335301962e-02, 4.4205005881659098e-02,
                          -2.2103860986973051e-02, 1.0574689371162864e-02, -4.6276428065385065e-03,
                          1.7936153397572132e-03, -5.9617527051353237e-04, 1.6314517495669067e-04,
                          -3.4555126770115446e-05, 5.0617615610782593e-06, -3.8768958592971409e-07};
const double kHalf13[] = {3.1408224847888910e-01, -9.4045836332667387e-02, 4.5459878763259978e-02,
                          -2.3383369012219993e-02, 1.1644273044890753e-02, -5.3806714579057013e-03,
                          2.2429072878264022e-03, -8.2204347506606424e-04, 2.5724946477840893e-04,
                          -6.6072709864248668e-05, 1.3099163296288644e-05, -1.7907147069136000e-06,
                          1.2750825595240592e-07};
// The SIMD engines the reference always selects on x86-64 have no 7-coef set
// (cr-core.c:138-140); the f32 engine stops at 9 coefs.
const HalfSet kHalfSets[] = {{8, 136.51f, kHalf8},   {9, 152.32f, kHalf9},   {10, 168.08f, kHalf10},
                             {11, 183.79f, kHalf11}, {12, 199.46f, kHalf12}, {13, 215.12f, kHalf13}};

// Per (direction, mode): phase-bit budgets and interpolation orders of the
// candidate polyphase kernels, in the order the reference tries them
// (cr-core.c:240-267, SIMD-poly build).  order < 0: candidate absent.
struct PolyCand { float scalar; int fn_order; };
struct PolyRow { float beta; PolyCand c[3]; };
const PolyRow kPolyRows[] = {
    {-1, {{0, 0}, {7.2f, 1}, {5.0f, 2}}},   {-1, {{0, 0}, {9.4f, 1}, {6.7f, 2}}},
    {-1, {{0, 0}, {12.4f, 1}, {7.8f, 2}}},  {-1, {{0, 0}, {13.6f, 1}, {9.3f, 2}}},
    {-1, {{0, 0}, {10.5f, 2}, {8.4f, 3}}},  {-1, {{0, 0}, {11.85f, 2}, {9.0f, 3}}},
    {-1, {{0, 0}, {8.0f, 1}, {5.3f, 2}}},   {-1, {{0, 0}, {8.6f, 1}, {5.7f, 2}}},
    {-1, {{0, 0}, {10.6f, 1}, {6.75f, 2}}}, {-1, {{0, 0}, {12.6f, 1}, {8.6f, 2}}},
    {-1, {{0, 0}, {9.6f, 2}, {7.6f, 3}}},   {-1, {{0, 0}, {11.4f, 2}, {8.65f, 3}}},
    {10.62f, {{0, 0}, {0, -1}, {0, -1}}},   {-1, {{0, 0}, {8, 1}, {6, 2}}},
    {-1, {{0, 0}, {9, 1}, {6, 2}}},         {-1, {{0, 0}, {11, 1}, {7, 2}}},
    {-1, {{0, 0}, {13, 1}, {8, 2}}},        {-1, {{0, 0}, {10, 2}, {8, 3}}},
    {-1, {{0, 0}, {12, 2}, {9, 3}}},
};

int pick_dft_length(int num_taps, int lo, int large) {   // cr.c:178-182
  const double d = std::log(double(num_taps)) / std::log(2.);
  const int e = std::min(std::max(int(d + 2.77), lo), std::max(int(d + 1.77), large));
  return 1 << e;
}

void tracef(std::string *t, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
void tracef(std::string *t, const char *fmt, ...) {
  char buf[256];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  *t += buf;
  *t += '\n';
}

// Overlap-save ("DFT") stage: L-fold zero-stuffing, Kaiser LPF, M-fold decimation.
Stage make_dft_stage(double Fp, double Fs, double Fn, double att, double phase_response, int L, int M, double gain,
                     const RSpec &r, std::string *trace) {
  Stage s;
  s.kind = kDft;
  s.L = L;
  s.M = M;
  s.f_domain_m = std::abs(3 - M) == 1 && Fs <= 1;
  s.fd_shift = s.f_domain_m ? M / 2 : 0;
  const int k = (phase_response == 50 && is_pow2(L) && Fn == L) ? 2 * L : 4;     // tap-count granularity, cr.c:196
  LpfDesign d = design_lpf(Fp, Fs, Fn, att, 0, -k, -1., false);
  int post_peak = d.num_taps / 2;                         // linear phase, cr.c:201
  if (phase_response != 50) {                             // cr.c:199-200
    fir_to_phase(&d.h, &post_peak, phase_response);
    d.num_taps = int(d.h.size());
  }
  s.num_taps = d.num_taps;
  s.dft_length = pick_dft_length(d.num_taps, int(r.log2_min_dft_size), int(r.log2_large_dft_size));
  s.g = std::move(d.h);
  for (double &v : s.g) v *= L * gain;                    // cr.c:206 (FFT normalisation cancels)
  const int preload = post_peak / L, at = post_peak % L;
  const int at_eff = is_pow2(L) ? 0 : at;                 // spectrum replication ignores `at`, cr.c:87-124
  s.C = d.num_taps - 1 - (long(L) * preload + at_eff);
  s.emu_preload = s.dev_preload = preload;
  s.emu_at0 = at;
  s.emu_pre_post = 0;
  s.emu_input_size = (s.dft_length - at + L - 1) / L;
  tracef(trace, "fir_len=%i dft_length=%i Fp=%g Fs=%g Fn=%g att=%g %i/%i", s.num_taps, s.dft_length,
         Fp, Fs, Fn, att, L, M);
  return s;
}

// Coefficient table in tap-major order per phase; for order > 0 each tap holds
// the polynomial coefficients highest order first.  Numerically identical to the
// reference's table (cr.c:31-59) but un-padded and without SIMD interleave.
std::vector<double> build_poly_table(const std::vector<double> &g, int nc, int phases, int order) {
  std::vector<double> T(size_t(nc) * phases * (order + 1), 0.);
  double ahead = g[0], f1 = 0, f2 = 0;     // g is symmetric: g[taps-1] == g[0]
  for (int i = nc - 1; i >= 0; --i)
    for (int j = phases - 1; j >= 0; --j) {
      const double f0 = ahead;
      const int pos = i * phases + j - 1;
      ahead = pos > 0 ? g[pos - 1] : 0;
      double b = 0, c = 0, d = 0;
      if (order == 1) b = f1 - f0;
      else if (order == 2) { c = .5 * (f2 + f0) - f1; b = f1 - c - f0; }
      else if (order == 3) { c = .5 * (f1 + ahead) - f0; d = (1 / 6.) * (f2 - f1 + ahead - f0 - 4 * c); b = f1 - f0 - d - c; }
      double *e = &T[(size_t(j) * nc + (nc - 1 - i)) * (order + 1)];
      if (order > 2) e[order - 3] = d;
      if (order > 1) e[order - 2] = c;
      if (order > 0) e[order - 1] = b;
      e[order] = f0;
      f2 = f1;
      f1 = f0;
    }
  return T;
}

}  // namespace

const char *make_plan(double io_ratio, const QSpec &q, const RSpec &r, double multiplier, Plan *out) {
  Plan &P = *out;
  P = Plan();
  const double tol = 1 + 1e-5;
  double bits = q.precision;
  const unsigned long rolloff = q.flags & 3;
  const int interpolator = int(r.flags & 3) - 1;
  const double Fp0 = q.passband_end, Fs0 = q.stopband_begin, tbw0 = Fs0 - Fp0;
  if (q.error) return q.error;
  if (q.flags & 32) return "variable-rate (SOXR_VR) is not supported by b200soxr";
  const double phase_response = q.phase_response;
  if (0 > phase_response || phase_response > 100) return "phase response not in [0, 100]";   // cr.c:308-309
  if (io_ratio < 1 && Fs0 - 1 > 1 - Fp0 / tol) return "imaging greater than rolloff";
  if (.002 / tol > tbw0 || tbw0 > .5 * tol) return "transition bandwidth not in [0.2,50] % of nyquist";
  if (.5 / tol > Fp0 || Fs0 > 1.5 * tol) return "transition band not within [50,150] % of nyquist";
  if (bits != 0 && (15 > bits || bits > 33)) return "precision not in [15,33] bits";
  if (!(io_ratio > 0)) return "resampling factor not positive";

  P.io_ratio = io_ratio;
  P.precision_bits = bits;
  P.is_double = !(bits <= 20 && !(q.flags & 16));           // soxr.c:447
  P.engine = P.is_double ? "b200-cr64" : "b200-cr32";
  const int sizeof_real = P.is_double ? 8 : 4;

  // ---- 1. decompose the ratio (cr.c:277-346) ----------------------------------
  double arbM = io_ratio;
  const double bits1 = std::min(bits, 33.);
  double att = (bits1 + 1) * to_db(2.), attArb = att;
  int preL = 1, preM = 1, shr = 0, arbL = 1, postL = 1, postM = 1;
  bool upsample = false, rational = false;
  const bool iOpt = !(r.flags & 8);
  const double lq_bw0 = 1385 / 2048.;
  const bool lq_bits = (q.flags & 64) ? bits <= 16 : bits == 16;
  const bool lq_Fp0 = (q.flags & 64) ? Fp0 <= lq_bw0 : Fp0 == lq_bw0;
  int mode = (lq_bits && rolloff == kRolloffMedium)
                 ? int(io_ratio > 1 || phase_response != 50 || !lq_Fp0 || Fs0 != 1)
                 : (int(std::ceil(bits1)) - 6) / 4;
  bool again = bits != 0;
  bool first_pass_flag = true;   // mirrors the reference's `n` (1 unless a re-plan was requested)
  while (again) {
    again = false;
    first_pass_flag = true;
    const int maxL = interpolator > 0 ? 1 : mode ? 2048
                     : int(std::ceil(r.coef_size_kbytes * 1000. / (42 * sizeof_real)));
    upsample = arbM < 1;
    shr = 0;
    for (int i = int(.5 * arbM); i >>= 1; arbM *= .5, ++shr) {}
    preM = upsample || (arbM > 1.5 && arbM < 2);
    postM = 1 + (arbM > 1 && preM);
    arbM /= postM;
    preL = 1 + (!preM && arbM < 2) + (upsample && mode);
    arbM *= preL;
    const double frac = arbM - int(arbM);
    double epsilon = 0;
    if (frac != 0) epsilon = std::fabs(std::floor(frac * kTwo32 + .5) / (frac * kTwo32) - 1);
    rational = frac == 0;
    for (int i = 1; i <= maxL && !rational; ++i) {
      const double d = frac * i;
      const int cand = int(d + .5);
      rational = std::fabs(cand / d - 1) <= epsilon;
      if (rational) {
        if (cand == i) {
          arbM = std::ceil(arbM);
          const int x = arbM > 3;
          shr += x;
          arbM /= 1 + x;
        } else {
          arbM = i * int(arbM) + cand;
          arbL = i;
        }
      }
    }
    int L = preL * arbL, M = int(arbM * postM);
    const int odd = (L | M) & 1;
    L >>= !odd;
    M >>= !odd;
    double d = 0;
    if (iOpt && postL == 1 && (d = preL * arbL / arbM) > 4 && d != 5) {
      int i = int(d / 16);
      for (postL = 4; (i >>= 1) && postL < 256; postL <<= 1) {}
      arbM = arbM * postL / arbL / preL;
      arbL = 1;
      first_pass_flag = false;
      again = true;
    } else if (rational && (std::max(L, M) < 3 + 2 * iOpt || L * M < 6 * iOpt)) {
      preL = L; preM = M; arbM = arbL = postM = 1;
    }
    if (!mode && (!rational || !first_pass_flag)) {
      ++mode;
      again = true;
    }
  }

  const bool have_pre = preM * preL != 1;
  bool have_arb = arbM * arbL != 1;
  const bool have_post = postM * postL != 1;
  int num_stages = shr + have_pre + have_arb + have_post;
  if (!num_stages && multiplier != 1) {   // pure gain: the reference uses the cubic stage, cr.c:349-352
    bits = 0;
    arbL = 0;
    ++num_stages;
    have_arb = true;
  }
  P.mode = mode; P.shr = shr; P.preL = preL; P.preM = preM; P.postL = postL; P.postM = postM;
  P.rational = rational; P.upsample = upsample;

  const double alpha = postM / (io_ratio * postL);
  int n = num_stages;
  if (n > 1) {                                               // attenuation budget, cr.c:357-363
    if (have_arb) { att += to_db(2.); attArb = att; --n; }
    att += to_db(double(n));
  }

  // ---- 2. half-band chain (cr.c:365-372, half-fir.h:38-50) ----------------------
  {
    const int n_sets = P.is_double ? 6 : 2;
    int idx = 0;
    while (idx + 1 < n_sets && att > kHalfSets[idx].att) ++idx;
    for (int i = 0; i < shr; ++i) {
      const HalfSet &h = kHalfSets[idx];
      Stage s;
      s.kind = kHalf;
      s.L = 1; s.M = 2;
      s.num_coefs = h.n;
      s.g.assign(size_t(4 * h.n - 1), 0.);
      const int ctr = 2 * h.n - 1;
      s.g[ctr] = .5;
      for (int j = 0; j < h.n; ++j) {
        const double c = P.is_double ? h.c[j] : double(f32(h.c[j]));
        s.g[ctr - (2 * j + 1)] = c;
        s.g[ctr + (2 * j + 1)] = c;
      }
      s.C = ctr;
      s.emu_pre_post = 4 * h.n;
      s.emu_preload = 2 * h.n;
      s.dev_preload = 0;
      P.stages.push_back(std::move(s));
    }
  }

  // ---- 3. pre DFT stage (cr.c:374-387) ------------------------------------------
  double Fp1 = Fp0, Fs1 = Fs0;
  if (have_pre) {
    const double Fn1 = preM ? std::max(preL, preM) : arbM / arbL;
    P.stages.push_back(make_dft_stage(Fp1, Fs1, Fn1, att, phase_response, preL, std::max(preM, 1), multiplier, r, &P.trace));
    multiplier = 1;
    Fp1 /= Fn1;
    Fs1 /= Fn1;
  }

  // ---- 4. arbitrary-ratio stage (cr.c:389-476) ----------------------------------
  if (bits == 0 && have_arb) {
    Stage s;
    s.kind = kCubic;
    s.cubic_mult = multiplier;
    multiplier = 1;
    s.step_fix = uint64_t(int64_t(arbM * kTwo32 + .5));
    s.at0_fix = 0;
    s.emu_pre_post = std::max(3, int(s.step_fix >> 32));
    s.emu_preload = s.dev_preload = 1;
    P.stages.push_back(std::move(s));
  } else if (have_arb) {
    static const float rolloffs[] = {-.01f, -.3f, 0, -.103f};
    const PolyRow &row = kPolyRows[6 * (int(upsample) + !!preM) + mode - !upsample];
    int num_coefs = int(row.c[0].scalar), phase_bits = 0, phases = 0, order = 0;
    double Fp = Fp1, Fs, Fn;
    const double mult = upsample ? 1 : arbM / arbL;
    if (!upsample && preM) { Fn = 2 * mult; Fs = 3 + std::fabs(Fs1 - 1); }
    else { Fn = 1; Fs = 2 - (mode ? Fp1 + (Fs1 - Fp1) * .7 : Fs1); }
    if (mode) Fp = Fs - (Fs - Fp) / (1 - inv_f_resp(rolloffs[rolloff], attArb));
    int i = (interpolator < 0 ? !rational : std::max(interpolator, int(!rational))) - 1;
    size_t coefs_size = 0;
    double at = 0;
    int nc4 = 0;
    do {
      const PolyCand &cand = row.c[++i];
      assert(cand.fn_order >= 0);
      if (i) { arbM /= arbL; arbL = 1; rational = false; }
      phase_bits = int(std::ceil(cand.scalar - std::log(mult) / std::log(2.)));
      phases = !rational ? (1 << phase_bits) : arbL;
      if (row.c[0].scalar == 0) {
        const int phases0 = std::max(phases, 19);
        const LpfDesign dry = design_lpf(Fp, Fs, Fn, attArb, 0, phases0, row.beta, true);
        num_coefs = dry.num_taps / phases0 + 1;
        num_coefs += num_coefs & !preM;
      }
      if ((num_coefs & 1) && rational && (arbL & 1)) { phases <<= 1; arbL <<= 1; arbM *= 2; }
      at = arbL * (.5 * (num_coefs & 1));
      order = i + (i && mode > 4);
      nc4 = (num_coefs + 3) & ~3;                 // the reference's SIMD padding (affects size heuristics only)
      coefs_size = size_t(nc4) * phases * (order + 1) * sizeof_real;
    } while (interpolator < 0 && i < 2 && row.c[i + 1].fn_order >= 0 &&
             coefs_size / 1000 > r.coef_size_kbytes);

    LpfDesign d = design_lpf(Fp, Fs, Fn, attArb, num_coefs * phases - 1, phases, row.beta, false);
    for (double &v : d.h) v *= multiplier;
    multiplier = 1;
    tracef(&P.trace, "fir_len=%i phases=%i coef_interp=%i size=%.3gk", num_coefs, phases, order,
           double(coefs_size) / 1000.);
    Stage s;
    s.num_coefs = num_coefs;
    s.phases = phases;
    s.order = order;
    s.phase_bits = phase_bits;
    s.dev_preload = (num_coefs - 1) >> 1;
    s.emu_preload = s.dev_preload + (nc4 - num_coefs);
    s.emu_pre_post = nc4 - 1;
    s.table = build_poly_table(d.h, num_coefs, phases, order);
    if (order == 0) {
      s.kind = kPoly;
      s.L = arbL;
      s.M = int(arbM);
      s.at0_int = s.emu_at0 = int(at);
      s.C = long(num_coefs - 1 - s.dev_preload) * s.L - 1 + s.at0_int;
      s.g = std::move(d.h);
    } else {
      s.kind = kPolyInterp;
      s.at0_fix = uint64_t(int64_t(at * kTwo32 + .5));
      s.step_fix = uint64_t(int64_t(arbM * kTwo32 + .5));
    }
    P.stages.push_back(std::move(s));
  }
  P.arbL = arbL;
  P.arbM = arbM;

  // ---- 5. post DFT stage (cr.c:478-482) -----------------------------------------
  if (have_post)
    P.stages.push_back(make_dft_stage(Fp0 / (upsample ? alpha : 1), upsample ? std::max(2 - Fs0 / alpha, 1.) : Fs0,
                                      double(std::max(postL, postM)), att, phase_response, postL, postM, multiplier, r, &P.trace));

  tracef(&P.trace, "%g: >>%i %i/%i %i/%g %i/%i (%x)", 1 / io_ratio, shr, preL, preM, arbL, arbM, postL,
         postM, P.is_double ? 0xf : 0xe);
  return nullptr;
}

// ---------------------------------------------------------------------------

// integer part of the 32.32 stage clock at0 + step * n, in 96 bits (see clock_at in kernels_generic.cuh)
static long clock_whole(uint64_t at0, uint64_t step, long n) {
  const unsigned __int128 at = (unsigned __int128)step * (unsigned __int128)(uint64_t)n + at0;
  return long((unsigned long long)(at >> 32));
}

long stage_input_first(const Stage &s, long n) {
  switch (s.kind) {
    case kHalf: case kDft: case kPoly: {
      // smallest k with C + M n - L k <= taps-1
      // smallest k with C + M n - L k <= taps-1  =>  k >= ceil(t / L)
      const long top = s.C + long(s.M) * n, t = top - (long(s.g.size()) - 1);
      return t >= 0 ? (t + s.L - 1) / s.L : -((-t) / s.L);
    }
    case kPolyInterp: return clock_whole(s.at0_fix, s.step_fix, n) - s.dev_preload;
    case kCubic: return clock_whole(0, s.step_fix, n) - 1;
  }
  return 0;
}

long stage_input_needed(const Stage &s, long n_out) {
  if (n_out <= 0) return 0;
  switch (s.kind) {
    case kHalf: case kDft: case kPoly: {
      const long top = s.C + long(s.M) * (n_out - 1);
      return top < 0 ? 0 : top / s.L + 1;
    }
    case kPolyInterp:
      return std::max(0L, clock_whole(s.at0_fix, s.step_fix, n_out - 1) + s.num_coefs - s.dev_preload);
    case kCubic: return clock_whole(0, s.step_fix, n_out - 1) + 3;
  }
  return 0;
}

// ---------------------------------------------------------------------------
// CallEmu
// ---------------------------------------------------------------------------

CallEmu::CallEmu(const Plan &plan) : plan_(plan) { reset(); }

void CallEmu::reset() {
  s_.clear();
  for (const Stage &st : plan_.stages) {
    S e;
    e.st = &st;
    e.occ = st.emu_preload;
    e.input_size = st.emu_input_size;
    e.at = st.emu_at0;
    e.at_fix = st.at0_fix;
    s_.push_back(e);
  }
  out_occ_ = 0;
  samples_in_ = samples_out_ = total_in_ = total_out_ = 0;
  flushing_ = false;
}

void CallEmu::run_stage(int i) {
  S &e = s_[i];
  const Stage &st = *e.st;
  long &next_occ = (i + 1 < int(s_.size())) ? s_[i + 1].occ : out_occ_;
  switch (st.kind) {
    case kDft: {
      const int N = st.dft_length, overlap = st.num_taps - 1, L = st.L;
      const long num_in = std::max(0L, e.occ);
      if (e.at + L * num_in >= N) {
        const int span = N - overlap - e.at + L - 1;
        e.occ -= span / L;
        if (!is_pow2(L) && L != 1) e.at = L - 1 - span % L;
        if (!st.f_domain_m) {
          if (st.M == 1) next_occ += N - overlap;
          else {
            int j = 0, k = e.remM;
            for (; k < N - overlap; ++j, k += st.M) {}
            e.remM = k - (N - overlap);
            next_occ += j;
          }
        } else {
          const int m = st.fd_shift;
          next_occ += N - ((((1 << m) - 1) * N + overlap) >> m);   // reserve N, trim (cr.c:84,169)
        }
      }
      e.input_size = (N - e.at + L - 1) / L;
      break;
    }
    case kPoly: {
      const long num_in = std::min(std::max(0L, e.occ - st.emu_pre_post), long(e.input_size));
      if (num_in) {
        long at = e.at;
        const long num_out = (num_in * st.L - at + st.M - 1) / st.M;
        at += num_out * st.M;
        e.occ -= at / st.L;
        e.at = int(at % st.L);
        next_occ += num_out;
      }
      break;
    }
    case kHalf: {
      const long num_in = std::min(std::max(0L, e.occ - st.emu_pre_post), long(e.input_size));
      const long num_out = (num_in + 1) >> 1;
      next_occ += num_out;
      e.occ -= 2 * num_out;
      break;
    }
    case kPolyInterp: case kCubic: {
      const long num_in = std::min(std::max(0L, e.occ - st.emu_pre_post), long(e.input_size));
      uint64_t at = e.at_fix;
      long produced = 0;
      while (long(int32_t(at >> 32)) < num_in) { ++produced; at += st.step_fix; }
      e.occ -= long(at >> 32);
      e.at_fix = at & 0xffffffffull;
      next_occ += produced;
      break;
    }
  }
}

bool CallEmu::stage_process(int i) {
  S &e = s_[i];
  bool done = false;
  long want;
  while (!done && (want = e.input_size - e.occ) > 0) {
    if (i == 0) {
      if (flushing_) e.occ += want;   // zeros, cr.c:505-506
      else done = true;
    } else done = stage_process(i - 1);
  }
  run_stage(i);
  return done && e.occ < e.input_size;
}

size_t CallEmu::output(size_t olen) {
  if (s_.empty()) {   // no stages: samples pass straight through the single fifo
    long n = flushing_ ? std::min<long>(-samples_out_, long(olen)) : long(olen);
    n = std::min(n, out_occ_);
    samples_out_ += n; total_out_ += n; out_occ_ -= n;
    return size_t(n);
  }
  const long n = flushing_ ? std::min<long>(-samples_out_, long(olen)) : long(olen);
  bool done = false;
  while (!done && out_occ_ < n) done = stage_process(int(s_.size()) - 1);
  long take = flushing_ ? std::min<long>(-samples_out_, long(olen)) : long(olen);
  take = std::min(take, out_occ_);
  samples_out_ += take;
  total_out_ += take;
  out_occ_ -= take;
  return size_t(take);
}

void CallEmu::add_input(size_t ilen) {
  if (flushing_ || !ilen) return;
  samples_in_ += int64_t(ilen);
  total_in_ += int64_t(ilen);
  if (s_.empty()) out_occ_ += long(ilen); else s_[0].occ += long(ilen);
}

size_t CallEmu::pull(size_t olen, bool flush) {
  if (flush && !flushing_) {          // cr.c:541-547
    samples_out_ -= int64_t(double(samples_in_) / plan_.io_ratio + .5);
    samples_in_ = 0;
    flushing_ = true;
  }
  return output(olen);
}

double CallEmu::delay() const { return double(samples_in_) / plan_.io_ratio - double(samples_out_); }

}  // namespace b2
