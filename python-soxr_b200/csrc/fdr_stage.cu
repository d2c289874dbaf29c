// fdr_stage.cu -- host side of the single-pass frequency-domain resampler: eligibility, the
// cascade's zero-phase response (double precision, from the same prototype filters the planner
// designs: cr.c:184-248 / cr.c:398-443 via plan.cpp), twiddles, launcher.
#include "fdr_stage.cuh"
#include "fdr_launch.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <numeric>
#include <vector>

namespace b2 {

namespace {
const double kPi = 3.14159265358979323846264338327950288;

template <typename R>
const char *upload_vec(const std::vector<R> &v, DevBuf *dst) {
  if (const char *e = dst->reserve(std::max<size_t>(16, v.size() * sizeof(R)))) return e;
  B2_CUDA(cudaMemcpy(dst->p, v.data(), v.size() * sizeof(R), cudaMemcpyHostToDevice));
  return nullptr;
}

// exp(-2 pi i a / n) with exact argument reduction
void unit_root(long long a, long long n, double *re, double *im) {
  a %= n;
  if (a < 0) a += n;
  const double ang = -2 * kPi * double(a) / double(n);
  *re = std::cos(ang);
  *im = std::sin(ang);
}

template <class F3>
std::vector<float> make_twiddles(int *tw3_off) {
  std::vector<float> tw(size_t(2) * (F3::TW2_ELEMS + F3::TW3_ELEMS));
  for (int mp = 0; mp < F3::RB; ++mp)
    for (int j = 0; j < F3::RA; ++j) {
      double re, im;
      unit_root((long long)j * mp, F3::RA * F3::RB, &re, &im);
      tw[2 * size_t(mp * F3::RA + j)] = float(re);
      tw[2 * size_t(mp * F3::RA + j) + 1] = float(im);
    }
  *tw3_off = F3::TW2_ELEMS;
  for (int m = 0; m < F3::RC; ++m)
    for (int t = 0; t < F3::T3; ++t) {
      double re, im;
      unit_root((long long)t * m, F3::N, &re, &im);
      const size_t o = size_t(F3::TW2_ELEMS) + size_t(m) * F3::T3 + t;
      tw[2 * o] = float(re);
      tw[2 * o + 1] = float(im);
    }
  return tw;
}

bool serial_head() {
  static const bool v = [] { const char *e = std::getenv("B200SOXR_SERIAL_HEAD"); return e && e[0] == '1'; }();
  return v;
}

bool fdr_disabled() {
  const char *env = std::getenv("B200SOXR_NO_FDR");
  if (env && env[0] == '1') return true;
  env = std::getenv("B200SOXR_EXACT");
  return env && env[0] == '1';
}
}  // namespace

std::vector<double> fdr_response(const Plan &plan, int rin, int rout, int nin, int nout, int m_in, int *n_head) {
  std::vector<double> none;
  const size_t S = plan.stages.size();
  if (S < 1 || S > 2) return none;   // k_fdr_head restates a cascade of at most two stages
  long long Lt = 1, Mt = 1;
  for (const Stage &s : plan.stages) {
    if (!(s.kind == kDft || s.kind == kPoly || s.kind == kHalf) || s.g.empty()) return none;
    Lt *= s.L;
    Mt *= s.M;
    if (Lt > (1 << 20) || Mt > (1 << 20)) return none;
  }
  const long long g = std::gcd(Lt, Mt);
  if (Lt / g != rout || Mt / g != rin) return none;
  // global tick = 1 / Lt input samples; stage i's fine grid (its input rate * L_i) has
  // tick_i = prod_{s<i} M_s * prod_{s>i} L_s ticks per sample
  std::vector<long long> tick(S);
  double half_support = 0;   // in ticks
  for (size_t i = 0; i < S; ++i) {
    long long t = 1;
    for (size_t s = 0; s < i; ++s) t *= plan.stages[s].M;
    for (size_t s = i + 1; s < S; ++s) t *= plan.stages[s].L;
    tick[i] = t;
    const Stage &st = plan.stages[i];
    if (2 * st.C != long(st.g.size()) - 1) return none;   // every stage zero-phase about a tap: real response
    half_support += double(st.C) * double(t);
  }
  if (half_support > double(m_in) * double(Lt)) return none;
  const long long Pp = Lt * nin;   // block period in ticks
  std::vector<double> H(size_t(nout), 0.0);
  for (int F = 0; F <= (nout - 1) / 2; ++F) {
    double prod = 1;
    for (size_t i = 0; i < S; ++i) {
      const Stage &st = plan.stages[i];
      const long c = st.C;
      double acc = st.g[size_t(c)];
      for (long k = 1; k <= c; ++k) {
        // cos(2 pi F k tick / Pp), argument reduced in integers
        long long a = ((long long)F * k) % Pp;
        a = (a * (tick[i] % Pp)) % Pp;
        acc += (st.g[size_t(c - k)] + st.g[size_t(c + k)]) * std::cos(2 * kPi * double(a) / double(Pp));
      }
      prod *= acc;
    }
    const double h = prod / double(Pp);
    H[size_t(F)] = h;
    if (F) H[size_t(nout - F)] = h;
  }
  // outputs whose chain of windows reaches an intermediate signal before its index 0 (which the
  // reference treats as zero, cr.c:487-490, not as the filter's run-in)
  int head = 0;
  for (int n = 0; n < 65536; ++n) {
    long r = n;
    bool ok = true;
    for (size_t s = S; s-- > 1;) {
      r = stage_input_first(plan.stages[s], r);
      if (r < 0) { ok = false; break; }
    }
    if (ok) { head = n; break; }
    head = n + 1;
  }
  if (n_head) *n_head = head;
  return H;
}

namespace {
// One warp per signal: u[m] = sum_k gA[CA + MA m - LA k] x[k] (lane m), then
// out[n] = sum_{m >= 0} gB[CB + MB n - LB m] u[m] (lane n - o0) -- the reference's cascade with its
// intermediate signal starting at index 0 (cr.c:487-490), in double accumulation like k_rational.
__global__ void __launch_bounds__(128) k_fdr_head(const FdrHeadParams P) {
  extern __shared__ float u_all[];   // [4 warps][n_u]
  const int w = int(threadIdx.x) >> 5, lane = int(threadIdx.x) & 31;
  float *u_s = u_all + size_t(w) * size_t(P.n_u);
  const unsigned s = blockIdx.x * 4u + unsigned(w);
  if (s < P.n_sig) {
    const long long off = (long long)(s / P.in.channels) * P.in.clip_stride + (long long)(s % P.in.channels) * P.in.chan_stride -
                          P.in.base * P.in.frame_stride;
    for (int m = lane; m < P.n_u; m += 32) {
      const long top = P.CA + long(P.MA) * m;
      long k_hi = top / P.LA;
      const long t = top - (P.nA - 1);
      long k_lo = t <= 0 ? 0 : (t + P.LA - 1) / P.LA;
      if (k_lo < P.in.vlo) k_lo = P.in.vlo;
      if (k_hi > P.in.vhi - 1) k_hi = P.in.vhi - 1;
      double acc = 0;
      for (long k = k_lo; k <= k_hi; ++k)
        acc = fma(double(P.gA[top - P.LA * k]), ld_as<double>(P.in.p, P.in.dtype, off + k * P.in.frame_stride), acc);
      u_s[m] = float(acc);
    }
  }
  __syncwarp();
  if (s < P.n_sig) {
    for (long n = P.o0 + lane; n < P.o1; n += 32) {
      const long top = P.CB + long(P.MB) * n;
      long m_hi = top / P.LB;
      const long t = top - (P.nB - 1);
      const long m_lo = t <= 0 ? 0 : (t + P.LB - 1) / P.LB;
      if (m_hi > P.n_u - 1) m_hi = P.n_u - 1;
      double acc = 0;
      for (long m = m_lo; m <= m_hi; ++m) acc = fma(double(P.gB[top - P.LB * m]), double(u_s[m]), acc);
      const long long oo = (long long)(s / P.out.channels) * P.out.clip_stride + (long long)(s % P.out.channels) * P.out.chan_stride +
                           (n - P.out.base) * P.out.frame_stride;
      st_as<double>(P.out.p, P.out.dtype, oo, acc);
    }
  }
}

// The same head for SHORT heads (a few outputs per signal, e.g. 5 at 48k -> 44.1k, whose 19 intermediate samples
// are 300-tap dot products): one 128-thread block per signal, every dot product split over S lanes (a power of
// two chosen from its length) and reduced with shuffles, so the long first-stage filter does not sit on one
// thread's critical path (20 instead of 30 us for 2048 signals).  Long heads (102 outputs at 48k -> 8k) keep all
// lanes busy already and are faster with the warp-per-signal kernel above.
__device__ __forceinline__ int head_split(int taps) {   // lanes per dot product: about 8 taps per lane, at most a warp
  int s = 1;
  while (s < 32 && taps > 8 * s) s *= 2;
  return s;
}
__global__ void __launch_bounds__(128) k_fdr_head_split(const FdrHeadParams P) {
  extern __shared__ float u_blk[];   // [n_u]
  const unsigned s = blockIdx.x;
  const int tid = int(threadIdx.x);
  const long long off = (long long)(s / P.in.channels) * P.in.clip_stride + (long long)(s % P.in.channels) * P.in.chan_stride -
                        P.in.base * P.in.frame_stride;
  {
    const int S = head_split((P.nA + P.LA - 1) / P.LA), part = tid & (S - 1);
    for (int m0 = 0; m0 < P.n_u; m0 += 128 / S) {   // the same trip count for every thread: whole shuffle groups
      const int m = m0 + tid / S;
      double acc = 0;
      if (m < P.n_u) {
        const long top = P.CA + long(P.MA) * m;
        long k_hi = top / P.LA;
        const long t = top - (P.nA - 1);
        long k_lo = t <= 0 ? 0 : (t + P.LA - 1) / P.LA;
        if (k_lo < P.in.vlo) k_lo = P.in.vlo;
        if (k_hi > P.in.vhi - 1) k_hi = P.in.vhi - 1;
        for (long k = k_lo + part; k <= k_hi; k += S)
          acc = fma(double(P.gA[top - P.LA * k]), ld_as<double>(P.in.p, P.in.dtype, off + k * P.in.frame_stride), acc);
      }
      for (int d = S >> 1; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
      if (m < P.n_u && part == 0) u_blk[m] = float(acc);
    }
  }
  __syncthreads();
  {
    const int S = head_split((P.nB + P.LB - 1) / P.LB), part = tid & (S - 1);
    for (long n0 = P.o0; n0 < P.o1; n0 += 128 / S) {
      const long n = n0 + tid / S;
      double acc = 0;
      if (n < P.o1) {
        const long top = P.CB + long(P.MB) * n;
        long m_hi = top / P.LB;
        const long t = top - (P.nB - 1);
        const long m_lo = t <= 0 ? 0 : (t + P.LB - 1) / P.LB;
        if (m_hi > P.n_u - 1) m_hi = P.n_u - 1;
        for (long m = m_lo + part; m <= m_hi; m += S) acc = fma(double(P.gB[top - P.LB * m]), double(u_blk[m]), acc);
      }
      for (int d = S >> 1; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
      if (n < P.o1 && part == 0) {
        const long long oo = (long long)(s / P.out.channels) * P.out.clip_stride +
                             (long long)(s % P.out.channels) * P.out.chan_stride + (n - P.out.base) * P.out.frame_stride;
        st_as<double>(P.out.p, P.out.dtype, oo, acc);
      }
    }
  }
}

// Long heads (e.g. 102 outputs at 48k -> 8k, each a 609-tap dot product): cheaper the other way round.  k_fdr
// computes the head outputs like any other (the ideal operator, whose intermediate signal u has a run-in
// u[m], m_min <= m < 0, from the first stage's filter reaching back before the stream), and this kernel
// subtracts what the reference does not have (its u starts at index 0, cr.c:487-490):
//     out[n] -= sum_{m_min <= m < 0} gB[CB + MB n - LB m] u[m],   u[m] = sum_{k >= 0} gA[CA + MA m - LA k] x[k].
// One warp per signal; u[m_min .. 0) in shared memory.
__global__ void __launch_bounds__(128) k_fdr_head_fix(const FdrHeadParams P) {
  extern __shared__ float u_neg[];   // [4 warps][-m_min]
  const int w = int(threadIdx.x) >> 5, lane = int(threadIdx.x) & 31;
  const int n_neg = -P.m_min;
  float *u_s = u_neg + size_t(w) * size_t(n_neg);
  const unsigned s = blockIdx.x * 4u + unsigned(w);
  if (s < P.n_sig) {
    const long long off = (long long)(s / P.in.channels) * P.in.clip_stride + (long long)(s % P.in.channels) * P.in.chan_stride -
                          P.in.base * P.in.frame_stride;
    for (int i = lane; i < n_neg; i += 32) {
      const long top = P.CA + long(P.MA) * (P.m_min + i);   // >= 0 by the choice of m_min
      long k_hi = top / P.LA;
      const long t = top - (P.nA - 1);
      long k_lo = t <= 0 ? 0 : (t + P.LA - 1) / P.LA;
      if (k_lo < P.in.vlo) k_lo = P.in.vlo;
      if (k_hi > P.in.vhi - 1) k_hi = P.in.vhi - 1;
      double acc = 0;
      for (long k = k_lo; k <= k_hi; ++k)
        acc = fma(double(P.gA[top - P.LA * k]), ld_as<double>(P.in.p, P.in.dtype, off + k * P.in.frame_stride), acc);
      u_s[i] = float(acc);
    }
  }
  __syncwarp();
  if (s < P.n_sig) {
    for (long n = P.o0 + lane; n < P.o1; n += 32) {
      const long top = P.CB + long(P.MB) * n;
      long m_hi = top / P.LB;
      if (m_hi > -1) m_hi = -1;
      long m_lo = long(ceil_div_ll(top - (P.nB - 1), P.LB));
      if (m_lo < P.m_min) m_lo = P.m_min;
      double acc = 0;
      for (long m = m_lo; m <= m_hi; ++m) acc = fma(double(P.gB[top - P.LB * m]), double(u_s[m - P.m_min]), acc);
      if (m_lo <= m_hi) {
        const long long oo = (long long)(s / P.out.channels) * P.out.clip_stride + (long long)(s % P.out.channels) * P.out.chan_stride +
                             (n - P.out.base) * P.out.frame_stride;
        st_as<double>(P.out.p, P.out.dtype, oo, ld_as<double>(P.out.p, P.out.dtype, oo) - acc);
      }
    }
  }
}

template <class Cfg>
const char *prepare_cfg(DevicePlan *dp, int variant) {
  int n_head = 0;
  const std::vector<double> Hd = fdr_response(dp->plan, Cfg::RIN, Cfg::ROUT, Cfg::NIN, Cfg::NOUT, Cfg::M_IN, &n_head);
  if (Hd.empty()) return nullptr;
  auto fi = std::make_shared<FdrInfo>();
  fi->variant = variant;
  fi->n_head = n_head;
  if (n_head > 4096) return nullptr;
  if (dp->plan.stages.size() == 2) {
    const Stage &sb = dp->plan.stages[1];
    fi->n_u = int((sb.C + long(sb.M) * std::max(0, n_head - 1)) / sb.L) + 1;
    // long heads: is it cheaper to let k_fdr compute them and subtract the first stage's run-in (k_fdr_head_fix)?
    const Stage &sa = dp->plan.stages[0];
    fi->m_min = -int(sa.C / sa.M);
    const double tapsA = double(sa.g.size()) / sa.L, tapsB = double(sb.g.size()) / sb.L, n_neg = double(-fi->m_min);
    const double cost_direct = fi->n_u * tapsA + n_head * tapsB;
    const double cost_fix = n_neg * tapsA * .5 + n_head * std::min(tapsB, n_neg);
    fi->head_by_fix = n_head > 32 && fi->m_min < 0 && 2 * cost_fix < cost_direct && -fi->m_min <= 3072 && n_head < Cfg::ADV_OUT;
    if (!fi->head_by_fix && fi->n_u > 3072) return nullptr;   // k_fdr_head keeps u[0 .. n_u) of 4 signals in 48 kB of shared memory
  }
  fi->adv_in = Cfg::ADV_IN; fi->adv_out = Cfg::ADV_OUT; fi->m_in = Cfg::M_IN; fi->m_out = Cfg::M_OUT;
  fi->nin = Cfg::NIN; fi->nout = Cfg::NOUT;
  fi->smem_bytes = size_t(Cfg::BUF_A + Cfg::BUF_B) * sizeof(float2);
  std::vector<float> Hf(Hd.size());
  for (size_t i = 0; i < Hd.size(); ++i) Hf[i] = float(Hd[i]);
  if (const char *e = upload_vec(Hf, &fi->H)) return e;
  if (const char *e = upload_vec(make_twiddles<typename Cfg::FW>(&fi->ftw3_off), &fi->ftw)) return e;
  if (const char *e = upload_vec(make_twiddles<typename Cfg::IV>(&fi->itw3_off), &fi->itw)) return e;
  dp->fdr = fi;
  return nullptr;
}

}  // namespace

const char *fdr_prepare(DevicePlan *dp) {
  if (dp->is_double || fdr_disabled()) return nullptr;
  const Plan &pl = dp->plan;
  // What the single-pass form drops are the aliases a decimating cascade folds back from its own stop band
  // (<= -att dB, att = (bits + 1) * 6.02 + 6 dB, cr.c:288,359-363).  It is used
  //  * for HQ/VHQ-class filters (>= 20 bits, att >= 132 dB = 2.5e-7): measured on B200 against the reference,
  //    float32 HQ 48k<->44.1k: 2.8e-7 RMS with it, 2.3e-7 without (float32 rounding dominates both;
  //    scratch/gpu_hq_fdr.py), VHQ 1.7e-7 either way; north_star's budget is 1e-6;
  //  * for plans that are ONE DFT stage with L/M = 1/2 or 2/1 at every quality: there the reference itself
  //    decimates by truncating the spectrum (cr.c:163-170) or interpolates by replicating it (cr.c:87-124).
  // B200SOXR_FDR_ALL=1 lifts the precision condition, B200SOXR_FDR_VHQ_ONLY=1 restores the round-1i rule.
  const char *env = std::getenv("B200SOXR_FDR_ALL");
  const bool any_precision = env && env[0] == '1';
  env = std::getenv("B200SOXR_FDR_VHQ_ONLY");
  const bool vhq_only = env && env[0] == '1';
  const bool single_dft = pl.stages.size() == 1 && pl.stages[0].kind == kDft &&
                          ((pl.stages[0].L == 1 && pl.stages[0].M == 2) || (pl.stages[0].L == 2 && pl.stages[0].M == 1));
  const bool fine_enough = vhq_only ? pl.is_double : pl.precision_bits >= 20;
  if (!any_precision && !fine_enough && !single_dft) return nullptr;
  if (const char *e = prepare_cfg<FdrCfg160_147>(dp, 1)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg147_160>(dp, 2)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg2_1>(dp, 3)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg1_2>(dp, 4)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg2_1w>(dp, 5)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg3_1>(dp, 6)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg1_3>(dp, 7)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg441_160>(dp, 8)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg3_2>(dp, 9)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg2_3>(dp, 10)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg6_1>(dp, 11)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg1_6>(dp, 12)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg4_1>(dp, 13)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg1_4>(dp, 14)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg441_320>(dp, 15)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg320_441>(dp, 16)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg147_80>(dp, 17)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg80_147>(dp, 18)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg320_147>(dp, 19)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg147_320>(dp, 20)) return e;
  if (!dp->fdr) if (const char *e = prepare_cfg<FdrCfg160_441>(dp, 21)) return e;
  return nullptr;
}

static const char *launch_variant(const FdrInfo &fi, const FdrKParams &P, long long n_items, cudaStream_t st, Workspace *ws);

const char *fdr_run(const DevicePlan &dp, const FdrInfo &fi, const SigDesc &in, const SigDesc &out, long o0, long o1,
                    unsigned n_sig, unsigned group, cudaStream_t st, Workspace *ws) {
  if (o1 <= o0 || !n_sig) return nullptr;
  FdrHeadParams fix, head_params;
  bool fix_after = false, head_after = false;
  cudaStream_t head_stream = st;
  if (o0 < fi.n_head && fi.head_by_fix && dp.plan.stages.size() == 2) {
    const Stage &sa = dp.plan.stages[0], &sb = dp.plan.stages[1];
    fix.in = in; fix.out = out; fix.n_sig = n_sig;
    fix.o0 = o0; fix.o1 = std::min<long>(o1, fi.n_head);
    fix.gA = dp.stages[0].g.as<float>(); fix.gB = dp.stages[1].g.as<float>();
    fix.nA = int(sa.g.size()); fix.nB = int(sb.g.size());
    fix.LA = sa.L; fix.MA = sa.M; fix.LB = sb.L; fix.MB = sb.M;
    fix.CA = sa.C; fix.CB = sb.C;
    fix.n_u = 0; fix.m_min = fi.m_min;
    fix_after = true;     // k_fdr below computes [o0, o1) including the head; the run-in terms are removed after it
  } else if (o0 < fi.n_head) {
    // the first outputs, where the reference's intermediate signal starts at index 0 (cr.c:487-490)
    if (dp.plan.stages.size() != 2) return "internal: fdr head needs a two-stage plan";
    const Stage &sa = dp.plan.stages[0], &sb = dp.plan.stages[1];
    FdrHeadParams hp;
    hp.in = in;
    hp.out = out;
    hp.n_sig = n_sig;
    hp.o0 = o0;
    hp.o1 = std::min<long>(o1, fi.n_head);
    hp.gA = dp.stages[0].g.as<float>();
    hp.gB = dp.stages[1].g.as<float>();
    hp.nA = int(sa.g.size()); hp.nB = int(sb.g.size());
    hp.LA = sa.L; hp.MA = sa.M; hp.LB = sb.L; hp.MB = sb.M;
    hp.CA = sa.C; hp.CB = sb.C;
    hp.n_u = fi.n_u;
    hp.m_min = 0;
    // The head kernel and k_fdr touch disjoint outputs.  When k_fdr follows, the head kernel goes to a side stream
    // behind it: its blocks fill the SMs that k_fdr's persistent CTAs free at the end instead of adding their own
    // pass (and tail) in front.
    head_stream = st;
    if (ws && o1 > fi.n_head && !serial_head()) {
      if (const char *e = ws->need_aux()) return e;
      B2_CUDA(cudaEventRecord(ws->aux_ev[0], st));
      B2_CUDA(cudaStreamWaitEvent(ws->aux_stream, ws->aux_ev[0], 0));
      head_stream = ws->aux_stream;
      head_after = true;
      head_params = hp;
    } else {
      if (fi.n_head <= 32)
        k_fdr_head_split<<<n_sig, 128, size_t(std::max(1, fi.n_u)) * sizeof(float), st>>>(hp);
      else
        k_fdr_head<<<(n_sig + 3) / 4, 128, size_t(4) * size_t(std::max(1, fi.n_u)) * sizeof(float), st>>>(hp);
      B2_COUNT_LAUNCH();
      B2_CUDA(cudaGetLastError());
    }
    o0 = fi.n_head;
    if (o1 <= o0) return nullptr;
  }
  FdrKParams P;
  P.in = in;
  P.out = out;
  if (!group || n_sig % group) group = n_sig;
  P.group = group;
  P.b_first = o0 / fi.adv_out;
  const int n_blk = int((o1 - 1) / fi.adv_out - P.b_first + 1);
  // mono signals: two consecutive blocks of one signal share a transform (real / imaginary parts), so a
  // signal's result still depends on nothing but the signal itself
  P.bpair = group == 1;
  if (P.bpair) {
    P.ppg = 1;
    P.n_pairs = n_sig;
    P.n_blocks = (n_blk + 1) / 2;
  } else {
    P.ppg = (group + 1) / 2;
    P.n_pairs = (n_sig / group) * P.ppg;
    P.n_blocks = n_blk;
  }
  // interleaved stereo float32 on both sides, every pair 8-byte aligned: one float2 per frame, unit stride
  auto stereo_f32 = [&](const SigDesc &d) {
    return d.dtype == 0 && d.channels == 2 && group == 2 && d.chan_stride == 1 && d.frame_stride == 2 &&
           !(d.clip_stride & 1) && !((d.base * d.frame_stride) & 1) && !(reinterpret_cast<unsigned long long>(d.p) & 7);
  };
  // block-paired mono signals, unit stride: float32 or int16 in, float32 out (the last also covers the planar float
  // scratch in front of the integer epilogue)
  auto mono_unit = [&](const SigDesc &d, int dtype) { return d.dtype == dtype && d.frame_stride == 1; };
  static const bool no_mono_fast = [] { const char *v = std::getenv("B200SOXR_FDR_NO_MONO_FAST"); return v && v[0] == '1'; }();
  P.fast = 0;
  if (!P.bpair && stereo_f32(in) && stereo_f32(out)) P.fast = 1;
  else if (P.bpair && !no_mono_fast && mono_unit(out, 0) && mono_unit(in, 0)) P.fast = 2;
  else if (P.bpair && !no_mono_fast && mono_unit(out, 0) && mono_unit(in, 3)) P.fast = 3;
  P.o0 = o0;
  P.o1 = o1;
  P.H = fi.H.p;
  P.ftw2 = fi.ftw.p;
  P.ftw3 = static_cast<const char *>(fi.ftw.p) + size_t(fi.ftw3_off) * sizeof(float2);
  P.itw2 = fi.itw.p;
  P.itw3 = static_cast<const char *>(fi.itw.p) + size_t(fi.itw3_off) * sizeof(float2);
  const long long n_items = (long long)P.n_pairs * P.n_blocks;
  const char *err = launch_variant(fi, P, n_items, st, ws);
  if (err) return err;
  if (head_after) {
    if (fi.n_head <= 32)
      k_fdr_head_split<<<n_sig, 128, size_t(std::max(1, fi.n_u)) * sizeof(float), head_stream>>>(head_params);
    else
      k_fdr_head<<<(n_sig + 3) / 4, 128, size_t(4) * size_t(std::max(1, fi.n_u)) * sizeof(float), head_stream>>>(head_params);
    B2_COUNT_LAUNCH();
    B2_CUDA(cudaGetLastError());
    B2_CUDA(cudaEventRecord(ws->aux_ev[1], head_stream));
    B2_CUDA(cudaStreamWaitEvent(st, ws->aux_ev[1], 0));
  }
  if (!fix_after) return nullptr;
  k_fdr_head_fix<<<(n_sig + 3) / 4, 128, size_t(4) * size_t(-fi.m_min) * sizeof(float), st>>>(fix);
  B2_COUNT_LAUNCH();
  B2_CUDA(cudaGetLastError());
  return nullptr;
}

static const char *launch_variant(const FdrInfo &fi, const FdrKParams &P, long long n_items, cudaStream_t st, Workspace *ws) {
  return P.fast >= 2 ? fdr_launch_mono(fi, P, n_items, st, ws) : fdr_launch_modes<0, 1>(fi, P, n_items, st, ws);
}

}  // namespace b2
