// design.h -- double-precision Kaiser low-pass design used by the host planner.
// Restates (does not copy) the maths of libsoxr's filter.c / dbesi0.c; each
// function cites the reference lines it must agree with numerically.
#pragma once
#include <vector>

namespace b2 {

double bessel_i0(double x);                                   // role of dbesi0.c:16-149
double kaiser_beta(double att_db, double tr_bw);              // filter.c:62-87
// Returns the number of taps and (in/out) beta; filter.c:109-115
void kaiser_params(double att_db, double fc, double tr_bw, double *beta, int *num_taps);
// Symmetric Kaiser-windowed sinc, filter.c:89-107
std::vector<double> make_lpf(int num_taps, double fc, double beta, double rho, double scale);

struct LpfDesign {
  std::vector<double> h;   // empty for a dry run
  int num_taps = 0;
  double beta = 0, fc = 0;
};
// filter.c:117-143.  k > 0: number of phases; k < 0: num_taps == 1 (mod -k).
// dry_run == true corresponds to the reference's `Fn < 0` call.
LpfDesign design_lpf(double fp, double fs, double fn, double att_db, int num_taps, int k,
                     double beta, bool dry_run);

// Non-linear phase responses (quality recipes SOXR_MINIMUM_PHASE / INTERMEDIATE_PHASE / MAXIMUM_PHASE or an explicit
// quality_spec.phase_response): turns the symmetric prototype `h` into the filter with the same magnitude response and
// `phase` per cent of the way from minimum (0) over linear (50) to maximum (100) phase, by the cepstral method of
// filter.c:154-245 (restated with a plain radix-2 complex FFT instead of Ooura's packed real FFT).  *post_len
// receives the number of taps after the impulse peak (the stage's delay bookkeeping, cr.c:199-201).
void fir_to_phase(std::vector<double> *h, int *post_len, double phase);

double f_resp(double t, double a);                            // filter.c:253-268
double inv_f_resp(double drop_db, double a);                  // filter.c:270-277
inline double to_3db(double a) { return 1 - inv_f_resp(-3., a); }   // filter.h:42

}  // namespace b2
