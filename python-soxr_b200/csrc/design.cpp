// design.cpp -- Kaiser low-pass filter design (host, double precision).
//
// The numbers produced here must agree with libsoxr's filter.c to ~1e-16 so that
// the GPU runs *the same filters* as the reference (HQ vs VHQ taps differ by
// 7.7e-3 RMS -- a wrong table is not subtle).  The empirical constants (Kaiser
// beta fit, roll-off response fit) are the reference's published fits and are
// restated as data; the code around them is ours.
#include "design.h"

#include <algorithm>
#include <cassert>
#include <cmath>

namespace b2 {

// I0 via its power series  sum_k ((x/2)^k / k!)^2.  The reference evaluates a
// piecewise polynomial (dbesi0.c:133-148, |rel err| ~ 1e-16); the series is
// monotone, has no cancellation for x >= 0 and reaches the same accuracy for
// the window parameters met here (beta < ~30).
double bessel_i0(double x) {
  const double q = 0.25 * x * x;
  double term = 1, acc = 1;
  for (int k = 1; k < 500; ++k) {
    term *= q / (double(k) * k);
    acc += term;
    if (term < 1e-19 * acc) break;
  }
  return acc;
}

namespace {
// cubic-in-att fits of the Kaiser beta, one per octave of transition width
// (filter.c:65-76)
const double kBetaFit[10][4] = {
    {-6.784957e-10, 1.02856e-05, 0.1087556, -0.8988365 + .001},
    {-6.897885e-10, 1.027433e-05, 0.10876, -0.8994658 + .002},
    {-1.000683e-09, 1.030092e-05, 0.1087677, -0.9007898 + .003},
    {-3.654474e-10, 1.040631e-05, 0.1087085, -0.8977766 + .006},
    {8.106988e-09, 6.983091e-06, 0.1091387, -0.9172048 + .015},
    {9.519571e-09, 7.272678e-06, 0.1090068, -0.9140768 + .025},
    {-5.626821e-09, 1.342186e-05, 0.1083999, -0.9065452 + .05},
    {-9.965946e-08, 5.073548e-05, 0.1040967, -0.7672778 + .085},
    {1.604808e-07, -5.856462e-05, 0.1185998, -1.34824 + .1},
    {-1.511964e-07, 6.363034e-05, 0.1064627, -0.9876665 + .18},
};
inline double cubic(const double *c, double x) { return ((c[0] * x + c[1]) * x + c[2]) * x + c[3]; }
inline int clampi(int v, int lo, int hi) { return std::min(std::max(v, lo), hi); }
}  // namespace

double kaiser_beta(double att, double tr_bw) {
  if (att >= 60) {
    const double octave = std::log(tr_bw / .0005) / std::log(2.);
    const int i0 = clampi(int(octave), 0, 9), i1 = clampi(1 + int(octave), 0, 9);
    const double b0 = cubic(kBetaFit[i0], att), b1 = cubic(kBetaFit[i1], att);
    return b0 + (b1 - b0) * (octave - int(octave));
  }
  if (att > 50) return .1102 * (att - 8.7);
  if (att > 20.96) return .58417 * std::pow(att - 20.96, .4) + .07886 * (att - 20.96);
  return 0;
}

void kaiser_params(double att, double fc, double tr_bw, double *beta, int *num_taps) {
  if (*beta < 0) *beta = kaiser_beta(att, tr_bw * .5 / fc);
  const double b = *beta;
  const double width = att < 60 ? (att - 7.95) / (2.285 * M_PI * 2)
                                : ((.0007528358 - 1.577737e-05 * b) * b + .6248022) * b + .06186902;
  if (!*num_taps) *num_taps = int(std::ceil(width / tr_bw + 1));
}

std::vector<double> make_lpf(int num_taps, double fc, double beta, double rho, double scale) {
  assert(fc >= 0 && fc <= 1);
  const int m = num_taps - 1;
  std::vector<double> h(num_taps);
  const double gain = scale / bessel_i0(beta), inv_half = 1 / (.5 * m + rho);
  for (int i = 0; i <= m / 2; ++i) {
    const double z = i - .5 * m, x = z * M_PI, y = z * inv_half;
    double v = x != 0 ? std::sin(fc * x) / x : fc;
    v *= bessel_i0(beta * std::sqrt(1 - y * y)) * gain;
    h[i] = v;
    h[m - i] = v;
  }
  return h;
}

LpfDesign design_lpf(double fp, double fs, double fn, double att, int num_taps, int k, double beta,
                     bool dry_run) {
  LpfDesign d;
  const int requested = num_taps;
  const int phases = std::max(k, 1), modulo = std::max(-k, 1);
  const double rho = phases == 1 ? .5 : (att < 120 ? .63 : .75);
  fp /= std::fabs(fn);
  fs /= std::fabs(fn);
  double tr_bw = .5 * (fs - fp);        // 6 dB point to stop-band edge
  tr_bw /= phases;
  fs /= phases;
  tr_bw = std::min(tr_bw, .5 * fs);
  d.fc = fs - tr_bw;
  assert(d.fc - tr_bw >= 0);
  d.beta = beta;
  d.num_taps = num_taps;
  kaiser_params(att, d.fc, tr_bw, &d.beta, &d.num_taps);
  if (!requested)
    d.num_taps = phases > 1 ? d.num_taps / phases * phases + phases - 1
                            : (d.num_taps + modulo - 2) / modulo * modulo + 1;
  if (!dry_run) d.h = make_lpf(d.num_taps, d.fc, d.beta, rho, double(phases));
  return d;
}

namespace {
// empirical fits of the Kaiser LPF roll-off shape (filter.c:248-251)
inline double sine_phi(double x) { return ((2.0517e-07 * x - 1.1303e-04) * x + .023154) * x + .55924; }
inline double sine_psi(double x) { return ((9.0667e-08 * x - 5.6114e-05) * x + .013658) * x + 1.0977; }
inline double sine_pow(double x) { return std::log(.5) / std::log(std::sin(x * .5)); }
inline double lin_to_db(double x) { return std::log10(x) * 20; }
inline double db_to_lin(double x) { return std::exp(x * (M_LN10 * 0.05)); }
}  // namespace

double f_resp(double t, double a) {
  if (t > (a <= 160 ? .8 : .82)) {
    const double a1 = a + 15, p = .00035 * a + .375;
    const double w = 1 / (1 - .597) * std::asin(std::pow((a1 - 10.6) / a1, 1 / p));
    const double c = 1 + std::asin(std::pow(1 - a / a1, 1 / p)) / w;
    return a1 * (std::pow(std::sin((c - t) * w), p) - 1);
  }
  double x;
  if (t > .5) {
    x = sine_psi(a);
    x = std::pow(std::sin((1 - t) * x), sine_pow(x));
  } else {
    x = sine_phi(a);
    x = 1 - std::pow(std::sin(t * x), sine_pow(x));
  }
  return lin_to_db(x);
}

double inv_f_resp(double drop, double a) {
  double x = sine_phi(a);
  drop = db_to_lin(drop);
  const double s = drop > .5 ? 1 - drop : drop;
  x = std::asin(std::pow(s, 1 / sine_pow(x))) / x;
  return drop > .5 ? x : 1 - x;
}

// ---------------------------------------------------------------------------
// fir_to_phase (filter.c:154-245)
// ---------------------------------------------------------------------------
namespace {
// in-place radix-2 complex FFT, sign = -1 forward (e^{-i...}), +1 backward (unnormalised)
void cfft(std::vector<double> &re, std::vector<double> &im, int sign) {
  const size_t n = re.size();
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
  }
  const double pi = 3.14159265358979323846264338327950288;
  for (size_t len = 2; len <= n; len <<= 1) {
    const double ang = sign * 2 * pi / double(len);
    for (size_t i = 0; i < n; i += len)
      for (size_t k = 0; k < len / 2; ++k) {
        const double c = std::cos(ang * double(k)), s = std::sin(ang * double(k));
        const size_t a = i + k, b = a + len / 2;
        const double tr = re[b] * c - im[b] * s, ti = re[b] * s + im[b] * c;
        re[b] = re[a] - tr; im[b] = im[a] - ti;
        re[a] += tr; im[a] += ti;
      }
  }
}
// The reference's real transform pair in its own (Ooura) convention on the unpacked array
//   w[2k] = R_k = sum_j a_j cos(2 pi j k / n),  w[2k+1] = I_k = sum_j a_j sin(2 pi j k / n),  k = 0 .. n/2
// (I_k is MINUS the imaginary part of the usual DFT); the backward transform followed by the 2/n scaling of
// filter.c:190,215 is the exact inverse.
void real_forward(std::vector<double> &w, int n) {       // w: n reals in, n + 2 values (R, I pairs) out
  std::vector<double> re(w.begin(), w.begin() + n), im(size_t(n), 0.);
  cfft(re, im, -1);
  for (int k = 0; k <= n / 2; ++k) { w[size_t(2 * k)] = re[size_t(k)]; w[size_t(2 * k + 1)] = -im[size_t(k)]; }
}
void real_backward_scaled(std::vector<double> &w, int n) {   // n + 2 values (R, I pairs) in, n reals out
  std::vector<double> re(size_t(n), 0.), im(size_t(n), 0.);
  for (int k = 0; k <= n / 2; ++k) {
    re[size_t(k)] = w[size_t(2 * k)];
    im[size_t(k)] = (k == 0 || k == n / 2) ? 0. : -w[size_t(2 * k + 1)];
    if (k && k < n / 2) { re[size_t(n - k)] = re[size_t(k)]; im[size_t(n - k)] = -im[size_t(k)]; }
  }
  cfft(re, im, +1);
  for (int j = 0; j < n; ++j) w[size_t(j)] = re[size_t(j)] / double(n);
}
double safe_log(double x) { return x != 0 ? std::log(x) : -26; }
}  // namespace

void fir_to_phase(std::vector<double> *h, int *post_len, double phase) {
  const double pi = 3.14159265358979323846264338327950288;
  const double phase1 = (phase > 50 ? 100 - phase : phase) / 50;
  int len = int(h->size());
  int work_len = 2 * 2 * 8;
  for (int i = len; i > 1; work_len <<= 1, i >>= 1) {}
  std::vector<double> work(size_t(work_len) + 2, 0.), pi_wraps(size_t(work_len) / 2 + 1, 0.);
  for (int i = 0; i < len; ++i) work[size_t(i)] = (*h)[size_t(i)];
  real_forward(work, work_len);                                   // cepstral
  double prev_angle2 = 0, cum_2pi = 0, prev_angle1 = 0, cum_1pi = 0;
  for (int i = 0; i <= work_len; i += 2) {
    double angle = std::atan2(work[size_t(i) + 1], work[size_t(i)]);
    double detect = 2 * pi;
    double delta = angle - prev_angle2;
    double adjust = detect * (double(delta < -detect * .7) - double(delta > detect * .7));
    prev_angle2 = angle;
    cum_2pi += adjust;
    angle += cum_2pi;
    detect = pi;
    delta = angle - prev_angle1;
    adjust = detect * (double(delta < -detect * .7) - double(delta > detect * .7));
    prev_angle1 = angle;
    cum_1pi += std::fabs(adjust);                                 // fabs for when 2pi and 1pi have combined
    pi_wraps[size_t(i >> 1)] = cum_1pi;
    work[size_t(i)] = safe_log(std::sqrt(work[size_t(i)] * work[size_t(i)] + work[size_t(i) + 1] * work[size_t(i) + 1]));
    work[size_t(i) + 1] = 0;
  }
  real_backward_scaled(work, work_len);
  for (int i = 1; i < work_len / 2; ++i) {                        // window to reject acausal components
    work[size_t(i)] *= 2;
    work[size_t(i + work_len / 2)] = 0;
  }
  work[size_t(work_len)] = work[size_t(work_len) + 1] = 0;
  real_forward(work, work_len);
  // the reference keeps R_{n/2} packed in slot 1 and leaves slots n, n+1 alone in this second half: mirror that
  const double nyq_log = work[size_t(work_len)];
  for (int i = 2; i < work_len; i += 2)                           // interpolate between linear & min phase
    work[size_t(i) + 1] = phase1 * i / work_len * pi_wraps[size_t(work_len >> 1)] +
                          (1 - phase1) * (work[size_t(i) + 1] + pi_wraps[size_t(i >> 1)]) - pi_wraps[size_t(i >> 1)];
  work[0] = std::exp(work[0]);
  const double nyq = std::exp(nyq_log);
  for (int i = 2; i < work_len; i += 2) {
    const double x = std::exp(work[size_t(i)]);
    const double a = work[size_t(i) + 1];
    work[size_t(i)] = x * std::cos(a);
    work[size_t(i) + 1] = x * std::sin(a);
  }
  work[1] = 0;
  work[size_t(work_len)] = nyq;
  work[size_t(work_len) + 1] = 0;
  real_backward_scaled(work, work_len);

  int peak = 0, imp_peak = 0;
  double imp_sum = 0, peak_imp_sum = 0;
  for (int i = 0; i <= int(pi_wraps[size_t(work_len >> 1)] / pi + .5); ++i) {   // find peak pos.
    imp_sum += work[size_t(i)];
    if (std::fabs(imp_sum) > std::fabs(peak_imp_sum)) { peak_imp_sum = imp_sum; peak = i; }
    if (work[size_t(i)] > work[size_t(imp_peak)]) imp_peak = i;
  }
  while (peak && std::fabs(work[size_t(peak - 1)]) > std::fabs(work[size_t(peak)]) && work[size_t(peak - 1)] * work[size_t(peak)] > 0) --peak;
  int begin, end;
  if (phase1 == 0) begin = 0;
  else if (phase1 == 1) begin = peak - len / 2;
  else {
    begin = int((.997 - (2 - phase1) * .22) * len + .5);
    end = int((.997 + (0 - phase1) * .22) * len + .5);
    begin = peak - (begin & ~3);
    end = peak + 1 + ((end + 3) & ~3);
    len = end - begin;
  }
  std::vector<double> out(static_cast<size_t>(len));
  for (int i = 0; i < len; ++i)
    out[size_t(i)] = work[size_t((begin + (phase > 50 ? len - 1 - i : i) + work_len) & (work_len - 1))];
  *post_len = phase > 50 ? peak - begin : begin + len - (peak + 1);
  h->swap(out);
}

}  // namespace b2
