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

}  // namespace b2
