// Host emulation of the fdr kernel's passes (same templates, double2, threads run sequentially between
// "barriers") against the direct closed form of the plan's cascade.
#include <cstdio>
#include <cmath>
#include <vector>
#include <random>
#include "../../python-soxr_b200/csrc/fdr_stage.cuh"
using namespace b2;
#ifndef CFG
#define CFG FdrCfg160_147
#define IN_RATE 48000.0
#define OUT_RATE 44100.0
#endif
typedef CFG Cfg;
typedef Cfg::FW FW; typedef Cfg::IV IV;
template <class F3> void twid(std::vector<double2> &tw2, std::vector<double2> &tw3) {
  tw2.resize(F3::TW2_ELEMS); tw3.resize(F3::TW3_ELEMS);
  for (int mp = 0; mp < F3::RB; ++mp) for (int j = 0; j < F3::RA; ++j) { double a = -2*M_PI*j*mp/double(F3::RA*F3::RB); tw2[mp*F3::RA+j] = make_double2(cos(a), sin(a)); }
  for (int m = 0; m < F3::RC; ++m) for (int t = 0; t < F3::T3; ++t) { double a = -2*M_PI*double(t)*m/double(F3::N); tw3[m*F3::T3+t] = make_double2(cos(a), sin(a)); }
}
int main(int argc, char **argv) {
  const unsigned long q = argc > 1 ? atoi(argv[1]) : 6;
  Plan plan;
  const char *e = make_plan(IN_RATE/OUT_RATE, quality_spec(q, 0), RSpec(), 1.0, &plan);
  if (e) { printf("plan error %s\n", e); return 1; }
  int n_head = -1;
  std::vector<double> H = fdr_response(plan, Cfg::RIN, Cfg::ROUT, Cfg::NIN, Cfg::NOUT, Cfg::M_IN, &n_head);
  printf("stages %zu, H size %zu, n_head %d, THREADS %d TP_F1 %d TP_I3 %d BUF_A %d BUF_B %d\n", plan.stages.size(), H.size(), n_head, Cfg::THREADS, Cfg::TP_F1, Cfg::TP_I3, Cfg::BUF_A, Cfg::BUF_B);
  if (H.empty()) return 1;
  const long n_in = 9000, n_out = long(output_frames(plan.io_ratio, n_in));
  std::mt19937 rng(1); std::normal_distribution<double> nd;
  std::vector<double2> x(n_in); for (auto &v : x) v = make_double2(nd(rng), nd(rng));
  // reference: direct closed form, stage by stage; intermediates are zero before index 0 and extend as far as needed
  std::vector<double2> cur = x;
  long cur_len_needed = n_out;
  std::vector<long> need(plan.stages.size() + 1);
  need[plan.stages.size()] = n_out;
  for (size_t s = plan.stages.size(); s-- > 0;) need[s] = stage_input_needed(plan.stages[s], need[s + 1]);
  for (size_t s = 0; s < plan.stages.size(); ++s) {
    const Stage &st = plan.stages[s];
    std::vector<double2> nxt(need[s + 1]);
    for (long n = 0; n < need[s + 1]; ++n) {
      double ax = 0, ay = 0;
      const long top = st.C + long(st.M) * n;
      for (long k = std::max(0L, stage_input_first(st, n)); k < (long)cur.size(); ++k) {
        const long gi = top - st.L * k;
        if (gi < 0) break;
        if (gi < (long)st.g.size()) { ax += st.g[gi] * cur[k].x; ay += st.g[gi] * cur[k].y; }
      }
      nxt[n] = make_double2(ax, ay);
    }
    cur.swap(nxt);
  }
  (void)cur_len_needed;
  // fdr emulation
  std::vector<double2> ftw2, ftw3, itw2, itw3;
  twid<FW>(ftw2, ftw3); twid<IV>(itw2, itw3);
  std::vector<double2> bufA(Cfg::BUF_A), bufB(Cfg::BUF_B), y(n_out, make_double2(0, 0));
  const long n_blocks = (n_out + Cfg::ADV_OUT - 1) / Cfg::ADV_OUT;
  for (long b = 0; b < n_blocks; ++b) {
    const long kb = b * Cfg::ADV_IN - Cfg::M_IN, nb = b * Cfg::ADV_OUT - Cfg::M_OUT;
    for (int tau = 0; tau < Cfg::THREADS; ++tau)
      FW::fwd_p1<Cfg::TP_F1>(tau, bufA.data(), [&](int g, int n) { long k = kb + g + FW::G1 * n; return (k >= 0 && k < n_in) ? x[k] : make_double2(0, 0); });
    for (int tau = 0; tau < Cfg::THREADS; ++tau) { double2 w[FW::RB]; FW::ld_tw2(tau, ftw2.data(), w); FW::fwd_p2(tau, bufA.data(), bufB.data(), w); }
    std::vector<double2> sp(Cfg::BUF_A);
    for (int tau = 0; tau < Cfg::THREADS; ++tau) { double2 w[FW::RC]; FW::ld_tw3(tau, ftw3.data(), w); FW::fwd_p3(tau, bufB.data(), sp.data(), w); }
    bufA = sp;
    if (b == 1) {   // check the forward transform against the definition at a few bins
      double worst = 0;
      for (int k : {0, 1, 5, 127, 128, 1000, 1175, Cfg::NIN - 3, Cfg::NIN - 1}) {
        double re = 0, im = 0;
        for (int n = 0; n < Cfg::NIN; ++n) { long kk = kb + n; if (kk < 0 || kk >= n_in) continue; double a = -2*M_PI*double((long long)n*k % Cfg::NIN)/Cfg::NIN; re += x[kk].x*cos(a) - x[kk].y*sin(a); im += x[kk].x*sin(a) + x[kk].y*cos(a); }
        worst = fmax(worst, fmax(fabs(re - bufA[k].x), fabs(im - bufA[k].y)));
      }
      printf("forward FFT check: max err %.3e\n", worst);
    }
    for (int tau = 0; tau < Cfg::THREADS; ++tau) { double h[IV::RC]; IV::ld_h(tau, H.data(), h); double2 w[IV::RC]; IV::ld_tw3(tau, itw3.data(), w); IV::inv_p1<Cfg::NIN>(tau, bufA.data(), h, bufB.data(), w); }
    std::vector<double2> e1(Cfg::BUF_A);
    for (int tau = 0; tau < Cfg::THREADS; ++tau) { double2 w[IV::RB]; IV::ld_tw2(tau, itw2.data(), w); IV::inv_p2(tau, bufB.data(), e1.data(), w); }
    for (int tau = 0; tau < Cfg::THREADS; ++tau)
      IV::inv_p3<Cfg::TP_I3>(tau, e1.data(), [&](int g, int n, double2 v) { const int o = g + IV::G1 * n;
        long nn = nb + o;
        if (o >= Cfg::M_OUT && o < Cfg::M_OUT + Cfg::ADV_OUT && nn >= 0 && nn < n_out) y[nn] = v;
      });
  }
  double se = 0, mx = 0; long cnt = 0;
  for (long n = 0; n < n_out; ++n) {
    double d = fmax(fabs(y[n].x - cur[n].x), fabs(y[n].y - cur[n].y));
    if (n < n_head) { printf("  head n=%ld err %.3e\n", n, d); continue; }
    se += d * d; mx = fmax(mx, d); ++cnt;
  }
  printf("n_out %ld: rms(max-abs per frame) %.3e  max %.3e  (outputs >= n_head)\n", n_out, sqrt(se / cnt), mx);
  return 0;
}
