// fft_stage.cu -- host side of the fused FFT stage: eligibility, spectra / twiddle /
// table construction (double precision, like the reference's one-time filter
// set-up in dft_stage_init cr.c:184-248 and prepare_poly_fir_coefs cr.c:31-59),
// and the launcher.
#include "fft_stage.cuh"

#include <algorithm>
#include <complex>
#include <cstring>
#include <mutex>
#include <vector>

namespace b2 {

namespace {

typedef std::complex<double> cd;
const double kPi = 3.14159265358979323846264338327950288;

// in-place iterative radix-2 FFT (double), forward
void fft_double(std::vector<cd> &a) {
  const size_t n = a.size();
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    for (size_t i = 0; i < n; i += len)
      for (size_t k = 0; k < len / 2; ++k) {
        const double ang = -2 * kPi * double(k) / double(len);
        const cd w(std::cos(ang), std::sin(ang));
        const cd u = a[i + k], v = a[i + k + len / 2] * w;
        a[i + k] = u + v;
        a[i + k + len / 2] = u - v;
      }
  }
}

template <typename R>
const char *upload_vec(const std::vector<R> &v, DevBuf *dst) {
  if (const char *e = dst->reserve(std::max<size_t>(16, v.size() * sizeof(R)))) return e;
  B2_CUDA(cudaMemcpy(dst->p, v.data(), v.size() * sizeof(R), cudaMemcpyHostToDevice));
  return nullptr;
}

bool is_pow2i(int x) { return x > 0 && !(x & (x - 1)); }

size_t tab_bytes(const FftStageInfo &fi, size_t real_size);
size_t smem_for(const FftStageInfo &fi, size_t real_size, int groups);
int groups_for(const FftStageInfo &fi, bool is_double) {
  const int T = fi.n / 16;
  return is_double ? 1 : (T >= 256 ? 1 : 256 / T);   // == FftCta<R, LOG2N>::G
}

template <typename R>
const char *prepare_t(DevicePlan *dp, FftStageInfo &fi) {
  const Stage &st = dp->plan.stages[fi.stage_index];
  DevStage &ds = dp->stages[fi.stage_index];
  const int n = fi.n, L = fi.L, M = fi.M, T = n / 16, R1 = n / 256;
  const long D = st.C - long(L) * fi.k0;   // j0 == 0
  const long taps = long(st.g.size());
  // spectra: [r][s][q][t] = FFT_n(hc_rs)[t + T q] / n, hc_rs[m mod n] = g_rs[m], m in [-span, 0]
  std::vector<R> spec(size_t(L) * M * n * 2);
  std::vector<R> edge(size_t(L) * M * std::max(fi.delta, 1), R(0));
  for (int r = 0; r < L; ++r)
    for (int s = 0; s < M; ++s) {
      std::vector<cd> h(n, cd(0, 0));
      for (int e = 0; e <= fi.span; ++e) {
        const long idx = D + long(M) * r - long(L) * s - long(L) * M * e;
        if (idx >= 0 && idx < taps) {
          h[(n - e) % n] += st.g[idx];
          if (e < fi.delta) edge[(size_t(r) * M + s) * fi.delta + e] = R(st.g[idx]);
        }
      }
      fft_double(h);
      for (int q = 0; q < 16; ++q)
        for (int t = 0; t < T; ++t) {
          const cd v = h[t + T * q] / double(n);
          const size_t o = ((size_t(r) * M + s) * 16 + q) * T + t;
          spec[2 * o] = R(v.real());
          spec[2 * o + 1] = R(v.imag());
        }
    }
  std::vector<R> tw2(size_t(16) * R1 * 2), tw3(size_t(16) * T * 2);
  for (int m = 0; m < 16; ++m) {
    for (int j = 0; j < R1; ++j) {
      const double a = -2 * kPi * double(j) * m / double(16 * R1);
      tw2[2 * (size_t(m) * R1 + j)] = R(std::cos(a));
      tw2[2 * (size_t(m) * R1 + j) + 1] = R(std::sin(a));
    }
    for (int t = 0; t < T; ++t) {
      const double a = -2 * kPi * double(t) * m / double(n);
      tw3[2 * (size_t(m) * T + t)] = R(std::cos(a));
      tw3[2 * (size_t(m) * T + t) + 1] = R(std::sin(a));
    }
  }
  const char *e = upload_vec(spec, &ds.spec);
  if (!e) e = upload_vec(edge, &ds.edge);
  if (!e) {
    // both twiddle tables in one buffer: tw2 then tw3
    std::vector<R> tw(tw2);
    tw.insert(tw.end(), tw3.begin(), tw3.end());
    e = upload_vec(tw, &ds.twid);
  }
  if (e) return e;
  if (fi.fused_poly) {
    const Stage &ps = dp->plan.stages[fi.stage_index + 1];
    DevStage &pd = dp->stages[fi.stage_index + 1];
    // rows: [pE zeros][nc taps][zeros up to nrow]; a trailing all-zero row absorbs shifted over-reads
    std::vector<R> tab(tab_bytes(fi, sizeof(R)) / sizeof(R), R(0));
    for (int ph = 0; ph < fi.pL; ++ph)
      for (int k = 0; k < fi.nc; ++k) tab[size_t(ph) * fi.nrow + fi.pE + k] = R(ps.table[size_t(ph) * fi.nc + k]);
    if ((e = upload_vec(tab, &pd.ptab))) return e;
  }
  return nullptr;
}

template <typename R, int LOG2N, int L, int M, bool POLY>
const char *launch(const FftStageInfo &fi, const FftKParams &P, long long n_items, cudaStream_t st) {
  auto kern = k_fft_stage<R, LOG2N, L, M, POLY>;
  int dev = 0;
  cudaGetDevice(&dev);
  constexpr int G = FftCta<R, LOG2N>::G, THREADS = FftCta<R, LOG2N>::THREADS;
  const size_t smem = smem_for(fi, sizeof(R), G);
  // per instantiation and device, under a mutex: many host threads launch concurrently (the reference's
  // thread-pool tests).  The dynamic shared-memory opt-in only ever grows; occupancy is cached per size.
  struct DevCfg { size_t attr_smem = 0, occ_smem = ~size_t(0); int per_sm = 1, sm_count = 0; };
  static std::mutex mu;
  static std::vector<DevCfg> cfgs;
  int sms, per_sm;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (cfgs.size() <= size_t(dev)) cfgs.resize(size_t(dev) + 1);
    DevCfg &c = cfgs[size_t(dev)];
    if (c.attr_smem < smem) {   // opt in to > 48 kB of dynamic shared memory
      B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
      c.attr_smem = smem;
    }
    if (!c.sm_count && (cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || c.sm_count < 1))
      c.sm_count = 148;
    if (c.occ_smem != smem) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.per_sm, kern, THREADS, smem) != cudaSuccess || c.per_sm < 1) {
        cudaGetLastError();
        c.per_sm = 1;
      }
      c.occ_smem = smem;
    }
    sms = c.sm_count;
    per_sm = c.per_sm;
  }
  const long long grid = std::min<long long>((n_items + G - 1) / G, (long long)sms * per_sm);
  kern<<<unsigned(grid), THREADS, smem, st>>>(P);
  B2_COUNT_LAUNCH();
  B2_CUDA(cudaGetLastError());
  return nullptr;
}

template <typename R, int LOG2N, bool POLY>
const char *launch_lm(const FftStageInfo &fi, const FftKParams &P, long long n_items, cudaStream_t st) {
  if (fi.L == 1 && fi.M == 1) return launch<R, LOG2N, 1, 1, POLY>(fi, P, n_items, st);
  if (fi.L == 2 && fi.M == 1) return launch<R, LOG2N, 2, 1, POLY>(fi, P, n_items, st);
  if (fi.L == 1 && fi.M == 2) return launch<R, LOG2N, 1, 2, POLY>(fi, P, n_items, st);
  if (fi.L == 4 && fi.M == 1) return launch<R, LOG2N, 4, 1, POLY>(fi, P, n_items, st);
  return "internal: fused FFT stage built for L/M in {1/1, 2/1, 1/2, 4/1} only";
}

template <typename R, bool POLY>
const char *launch_n(const FftStageInfo &fi, const FftKParams &P, long long n_items, cudaStream_t st) {
  switch (fi.log2n) {
    case 9: return launch_lm<R, 9, POLY>(fi, P, n_items, st);
    case 10: return launch_lm<R, 10, POLY>(fi, P, n_items, st);
    case 11: return launch_lm<R, 11, POLY>(fi, P, n_items, st);
    case 12:
      if (sizeof(R) == 4) return launch_lm<float, 12, POLY>(fi, P, n_items, st);
  }
  return "internal: fused FFT stage built for n in [512, 4096] only";
}

size_t tab_bytes(const FftStageInfo &fi, size_t real_size) {
  return ((size_t(fi.pL) + 1) * fi.nrow * real_size + 64 * real_size + 15) / 16 * 16;
}

size_t smem_for(const FftStageInfo &fi, size_t real_size, int groups) {
  const size_t csz = 2 * real_size;
  const size_t buf_elems = size_t(fi.n) + (fi.n >> 4) + (csz == 8 ? (fi.n >> 8) : (fi.n >> 7)) + 8;
  size_t bytes = size_t(groups) * (buf_elems + size_t(fi.u_elems)) * csz;
  if (fi.fused_poly) bytes += tab_bytes(fi, real_size);
  return bytes;
}

// Sliding-window variants compiled into the kernel (fft_stage.cuh): {id, RW, B, NT}.
struct SlideVariant { int id, rw, b, nt; };
const SlideVariant kSlide[] = {{1, 7, 2, 22}, {2, 7, 2, 16}, {3, 5, 1, 24}, {4, 6, 1, 16}, {5, 3, 2, 14}};

// Largest coefficient shift e_i = floor((rem0 + i*pM)/pL) - i*B over all start phases, i < rw.
int max_shift(int pL, int pM, int B, int rw) {
  int mx = 0;
  for (int rem0 = 0; rem0 < pL; ++rem0) {
    const long long v = ((long long)rem0 + (long long)(rw - 1) * pM) / pL - (long long)(rw - 1) * B;
    mx = std::max<long long>(mx, v);
  }
  return mx;
}

void choose_poly_variant(FftStageInfo *fi, bool is_double) {
  fi->pv = 0;
  fi->pE = 0;
  fi->ncp = (fi->nc + 3) & ~3;          // one output per thread: taps in chunks of 4
  fi->nrow = fi->ncp | 1;               // odd row stride: phases spread over all banks
  if (is_double || fi->M != 1 || fi->L > 2) return;
  if (const char *env = std::getenv("B200SOXR_NO_SLIDE"))
    if (env[0] == '1') return;
  const int B = fi->pM / fi->pL;
  int best_waste = 1 << 30;
  for (const SlideVariant &v : kSlide) {
    if (v.b != B) continue;
    const int E = v.nt - fi->nc;
    if (E < 0 || E < max_shift(fi->pL, fi->pM, B, v.rw)) continue;
    const int waste = E * 8 - v.rw;     // fewer padded taps first, then the longer run
    if (waste < best_waste) {
      best_waste = waste;
      fi->pv = v.id;
      fi->pE = E;
      fi->nrow = (v.nt + 1) | 1;
    }
  }
}

}  // namespace

const char *fft_stage_prepare(DevicePlan *dp) {
  if (const char *env = std::getenv("B200SOXR_NO_FUSED"))
    if (env[0] == '1') return nullptr;
  const Plan &plan = dp->plan;
  for (size_t i = 0; i < plan.stages.size(); ++i) {
    const Stage &st = plan.stages[i];
    if (st.kind != kDft) continue;
    const int L = st.L, M = st.M;
    if (!((L == 1 && M == 1) || (L == 2 && M == 1) || (L == 1 && M == 2) || (L == 4 && M == 1))) continue;
    if (st.dft_length % (L * M)) continue;
    const int n = st.dft_length / (L * M);
    if (!is_pow2i(n) || n < 512 || n > (dp->is_double ? 2048 : 4096)) continue;
    auto fi = std::make_shared<FftStageInfo>();
    fi->stage_index = int(i);
    fi->n = n;
    fi->log2n = 0;
    while ((1 << fi->log2n) < n) ++fi->log2n;
    fi->L = L;
    fi->M = M;
    fi->k0 = -long(st.emu_preload);
    // taps of g_rs[m] = g[D + M r - L s + L M m] lie in m in [m_lo, m_hi]; the block grid needs m_hi == 0
    const long D = st.C - long(L) * fi->k0, taps = long(st.g.size());
    long m_lo = 0, m_hi = -(1L << 40);
    for (int r = 0; r < L; ++r)
      for (int s = 0; s < M; ++s) {
        const long c = D + long(M) * r - long(L) * s, lm = long(L) * M;
        const long hi = (taps - 1 - c) >= 0 ? (taps - 1 - c) / lm : -((c - (taps - 1) + lm - 1) / lm);
        const long lo = -c >= 0 ? (-c + lm - 1) / lm : -(c / lm);
        m_hi = std::max(m_hi, hi);
        m_lo = std::min(m_lo, lo);
      }
    if (m_hi != 0 || -m_lo >= n / 2) continue;
    fi->span = int(-m_lo);
    fi->v_base = n - fi->span;
    // the reference's block must advance by exactly M * v_base inputs (cr.c:79-82) for stream alignment
    if ((st.dft_length - (st.num_taps - 1)) != fi->v_base * L * M) continue;
    fi->threads = n / 16;
    // fuse a following rational polyphase stage if its table fits next to the block
    if (i + 1 < plan.stages.size() && plan.stages[i + 1].kind == kPoly) {
      const Stage &ps = plan.stages[i + 1];
      fi->pL = ps.L;
      fi->pM = ps.M;
      fi->p_at0 = ps.at0_int;
      fi->p_pre = ps.dev_preload;
      fi->nc = ps.num_coefs;
      fi->delta = (ps.num_coefs - 1 + L - 1) / L;
      fi->fused_poly = true;
      choose_poly_variant(fi.get(), dp->is_double);
      fi->u_elems = (L * (fi->delta + fi->v_base) + 64 + 1) & ~1;
      const size_t need = smem_for(*fi, dp->is_double ? 8 : 4, groups_for(*fi, dp->is_double));
      if (need > 220 * 1024 || fi->delta >= fi->span || L * fi->delta > fi->threads) {
        fi->fused_poly = false;
        fi->delta = 0;
      }
    }
    if (!fi->fused_poly) fi->u_elems = (L * fi->v_base + 1) & ~1;
    fi->smem_bytes = smem_for(*fi, dp->is_double ? 8 : 4, groups_for(*fi, dp->is_double));
    if (fi->smem_bytes > 220 * 1024) continue;
    const char *e = dp->is_double ? prepare_t<double>(dp, *fi) : prepare_t<float>(dp, *fi);
    if (e) return e;
    dp->fft = fi;
    break;   // one fused launch per plan
  }
  return nullptr;
}

namespace {
// block that owns final output index o
long block_of(const FftStageInfo &fi, long o) {
  const long LV = long(fi.L) * fi.v_base;
  if (fi.fused_poly) {
    const long long at = (long long)fi.p_at0 + (long long)o * fi.pM;
    const long we = long(at / fi.pL) - fi.p_pre + fi.nc - 1;
    return we < 0 ? 0 : we / LV;
  }
  return o < 0 ? 0 : o / LV;
}
}  // namespace

RangeLH fft_stage_input_range(const FftStageInfo &fi, long lo, long hi) {
  if (hi <= lo) return {0, 0};
  const long b0 = block_of(fi, lo), b1 = block_of(fi, hi - 1);
  RangeLH r;
  r.lo = fi.k0 + b0 * long(fi.M) * fi.v_base - long(fi.M) * fi.delta;
  r.hi = fi.k0 + b1 * long(fi.M) * fi.v_base + long(fi.M) * fi.n;
  return r;
}

const char *fft_stage_run(const DevicePlan &dp, const FftStageInfo &fi, const SigDesc &in, const SigDesc &out,
                          long o0, long o1, unsigned n_sig, unsigned group, cudaStream_t st) {
  if (o1 <= o0 || !n_sig) return nullptr;
  const DevStage &ds = dp.stages[fi.stage_index];
  FftKParams P;
  P.in = in;
  P.out = out;
  P.n_sig = n_sig;
  if (!group || n_sig % group) group = n_sig;
  P.group = group;
  P.ppg = (group + 1) / 2;
  P.n_pairs = (n_sig / group) * P.ppg;
  P.b_first = block_of(fi, o0);
  P.n_blocks = int(block_of(fi, o1 - 1) - P.b_first + 1);
  P.o0 = o0;
  P.o1 = o1;
  P.k0 = fi.k0;
  P.v_base = fi.v_base;
  P.delta = fi.delta;
  P.u_elems = fi.u_elems;
  const size_t csz = dp.is_double ? 16 : 8;
  P.spec = ds.spec.p;
  P.tw2 = ds.twid.p;
  P.tw3 = static_cast<const char *>(ds.twid.p) + size_t(16) * (fi.n / 256) * csz;
  P.edge = ds.edge.p;
  P.ptab = nullptr;
  P.pL = fi.pL; P.pM = fi.pM; P.p_at0 = fi.p_at0; P.p_pre = fi.p_pre; P.nc = fi.nc; P.ncp = fi.ncp; P.nrow = fi.nrow; P.pv = fi.pv; P.pE = fi.pE;
  P.ptab_bytes = 0;
  if (fi.fused_poly) {
    P.ptab = dp.stages[fi.stage_index + 1].ptab.p;
    P.ptab_bytes = unsigned(tab_bytes(fi, dp.is_double ? 8 : 4));
  }
  const long long n_items = (long long)P.n_pairs * P.n_blocks;
  if (dp.is_double)
    return fi.fused_poly ? launch_n<double, true>(fi, P, n_items, st) : launch_n<double, false>(fi, P, n_items, st);
  return fi.fused_poly ? launch_n<float, true>(fi, P, n_items, st) : launch_n<float, false>(fi, P, n_items, st);
}

}  // namespace b2
