// runtime.cu -- plan upload/cache, workspace pool, range executor, dtype conversion.
#include "runtime.h"

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <cstdio>
#include <map>
#include <thread>
#include <cstdint>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#endif

#include "fdr_stage.cuh"
#include "fft_stage.cuh"
#include "kernels_generic.cuh"

namespace b2 {

std::atomic<unsigned long long> g_launches{0};

const char *cuda_err(cudaError_t e) {
  // cudaGetErrorString returns a static string
  return cudaGetErrorString(e);
}

const char *DevBuf::reserve(size_t bytes) {
  if (bytes <= cap) return nullptr;
  release();
  size_t want = std::max(bytes, size_t(1) << 16);
  want = (want + (want >> 2) + 255) & ~size_t(255);   // 25 % head-room
  B2_CUDA(cudaMalloc(&p, want));
  cap = want;
  return nullptr;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}
const char *PinnedBuf::reserve(size_t bytes) {
  if (bytes <= cap) return nullptr;
  release();
  size_t want = std::max(bytes, size_t(1) << 16);
  want = (want + (want >> 2) + 255) & ~size_t(255);
  B2_CUDA(cudaMallocHost(&p, want));
  cap = want;
  return nullptr;
}
void PinnedBuf::release() {
  if (p) cudaFreeHost(p);
  p = nullptr;
  cap = 0;
}

DevicePlan::~DevicePlan() {
  for (DevStage &s : stages) { s.g.release(); s.table.release(); s.spec.release(); s.twid.release(); s.edge.release(); s.ptab.release(); }
}

Workspace::~Workspace() {
  for (DevBuf *b : {&in_stage, &out_stage, &x, &y, &t0, &t1, &seeds, &blocks, &misc, &fdr_tickets}) b->release();
  for (PinnedBuf *b : {&pin_in, &pin_out, &pin_misc}) b->release();
  for (PinnedBuf &b : pin_ring_in) b.release();
  for (PinnedBuf &b : pin_ring_out) b.release();
  if (stream) cudaStreamDestroy(stream);
  if (h2d_stream) cudaStreamDestroy(h2d_stream);
  if (d2h_stream) cudaStreamDestroy(d2h_stream);
  for (cudaEvent_t e : events) cudaEventDestroy(e);
  if (busy_ev) cudaEventDestroy(busy_ev);
  if (aux_stream) cudaStreamDestroy(aux_stream);
  for (cudaEvent_t e : aux_ev) if (e) cudaEventDestroy(e);
}

const char *Workspace::need_aux() {
  if (!aux_stream) B2_CUDA(cudaStreamCreateWithFlags(&aux_stream, cudaStreamNonBlocking));
  for (cudaEvent_t &e : aux_ev)
    if (!e) B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return nullptr;
}

const char *Workspace::mark_busy(cudaStream_t st) {
  if (!busy_ev) B2_CUDA(cudaEventCreateWithFlags(&busy_ev, cudaEventDisableTiming));
  B2_CUDA(cudaEventRecord(busy_ev, st));
  busy = true;
  return nullptr;
}

const char *Workspace::need_events(size_t n) {
  if (!h2d_stream) B2_CUDA(cudaStreamCreateWithFlags(&h2d_stream, cudaStreamNonBlocking));
  if (!d2h_stream) B2_CUDA(cudaStreamCreateWithFlags(&d2h_stream, cudaStreamNonBlocking));
  while (events.size() < n) {
    cudaEvent_t e;
    B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    events.push_back(e);
  }
  return nullptr;
}

// ---- host copy pool ---------------------------------------------------------------------------------------
namespace {
// Large host-to-host copy with non-temporal stores (SSE2, baseline x86-64): the destination lines are not read for
// ownership first, which takes a third off the memory traffic of a copy -- the staging of pageable batches is bound
// by host memory bandwidth (754 MB in, 754 MB out per headline batch), not by the link.
#if defined(__x86_64__) || defined(_M_X64)
inline void copy_stream(char *dst, const char *src, size_t n) {
  if (n < (size_t(256) << 10)) { std::memcpy(dst, src, n); return; }
  const size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
  if (head) { std::memcpy(dst, src, head); dst += head; src += head; n -= head; }
  size_t i = 0;
  for (; i + 64 <= n; i += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 32));
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 48));
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), a);
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 48), d);
  }
  _mm_sfence();
  if (i < n) std::memcpy(dst + i, src + i, n - i);
}
#else
inline void copy_stream(char *dst, const char *src, size_t n) { std::memcpy(dst, src, n); }
#endif
const bool kStreamCopies = [] { const char *v = std::getenv("B200SOXR_NT_COPY"); return !(v && v[0] == '0'); }();
inline void copy_part(char *dst, const char *src, size_t n) {
  if (kStreamCopies) copy_stream(dst, src, n); else std::memcpy(dst, src, n);
}

class CopyPool {
 public:
  static CopyPool &get() {
    static CopyPool *pool = new CopyPool();   // never destroyed: worker threads must not outlive their state at exit
    return *pool;
  }
  void run(char *dst, const char *src, size_t bytes) {
    const size_t kMin = size_t(256) << 10;   // a worker's wake-up costs about as much as copying this
    const unsigned parts = unsigned(std::min<size_t>(workers_.size() + 1, std::max<size_t>(1, bytes / kMin)));
    if (parts <= 1) { std::memcpy(dst, src, bytes); return; }
    const size_t per = ((bytes / parts) + 4095) & ~size_t(4095);
    std::unique_lock<std::mutex> call_lock(call_mu_);        // one parallel copy at a time
    {
      std::lock_guard<std::mutex> l(mu_);
      dst_ = dst; src_ = src; bytes_ = bytes; per_ = per; parts_ = parts; next_ = 1; pending_ = parts - 1;
      ++generation_;
    }
    cv_.notify_all();
    copy_part(dst, src, std::min(per, bytes));              // the caller takes part 0
    std::unique_lock<std::mutex> l(mu_);
    done_cv_.wait(l, [&] { return pending_ == 0; });
  }

 private:
  CopyPool() {
    unsigned n = std::thread::hardware_concurrency();
    if (const char *e = std::getenv("B200SOXR_COPY_THREADS")) n = unsigned(std::max(1L, std::atol(e)));
    else n = std::max(1u, std::min(16u, n));   // measured on a 16-vCPU B200 host: 2 / 4 / 8 / 16 threads -> 46.5 / 23.4 / 15.1 / 13.2 ms per 754 MB batch
    for (unsigned i = 1; i < n; ++i) workers_.emplace_back([this] { loop(); });
    for (std::thread &t : workers_) t.detach();
  }
  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      unsigned part;
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return generation_ != seen && next_ < parts_; });
        part = next_++;
        if (next_ >= parts_) seen = generation_;
      }
      const size_t off = size_t(part) * per_;
      if (off < bytes_) copy_part(dst_ + off, src_ + off, std::min(per_, bytes_ - off));
      {
        std::lock_guard<std::mutex> l(mu_);
        if (--pending_ == 0) done_cv_.notify_all();
      }
    }
  }
  std::mutex call_mu_, mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> workers_;
  char *dst_ = nullptr;
  const char *src_ = nullptr;
  size_t bytes_ = 0, per_ = 0;
  unsigned parts_ = 0, next_ = 0, pending_ = 0;
  unsigned long long generation_ = 0;
};
}  // namespace

void parallel_memcpy(void *dst, const void *src, size_t bytes) {
  if (!bytes) return;
  CopyPool::get().run(static_cast<char *>(dst), static_cast<const char *>(src), bytes);
}

namespace {
std::mutex g_ws_mu;
std::vector<Workspace *> g_ws_free;

template <typename T>
const char *upload(const std::vector<double> &v, DevBuf *dst) {
  if (v.empty()) return nullptr;
  std::vector<T> tmp(v.size());
  for (size_t i = 0; i < v.size(); ++i) tmp[i] = T(v[i]);
  if (const char *e = dst->reserve(tmp.size() * sizeof(T))) return e;
  B2_CUDA(cudaMemcpy(dst->p, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice));
  return nullptr;
}

struct PlanKey {
  unsigned long long ratio_bits, gain_bits;
  double precision, passband_end, stopband_begin, phase_response;
  unsigned long qflags, rflags;
  unsigned min_dft, large_dft, coef_kb;
  int device, arith_double;
  bool operator<(const PlanKey &o) const { return std::memcmp(this, &o, sizeof *this) < 0; }
};
// Bounded (least recently used entry dropped): a long-running process that sweeps ratios must not grow device
// memory without limit.  Handles keep their plan alive through the shared_ptr.
struct PlanEntry { std::shared_ptr<DevicePlan> dp; unsigned long long last_use = 0; };
constexpr size_t kMaxCachedPlans = 64;
std::mutex g_plan_mu;
std::map<PlanKey, PlanEntry> g_plans;
unsigned long long g_plan_tick = 0;
}  // namespace

Workspace *acquire_workspace(bool host_wait) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
  {
    Workspace *w = nullptr;
    {
      std::lock_guard<std::mutex> l(g_ws_mu);
      for (size_t i = g_ws_free.size(); i-- > 0;)
        if (g_ws_free[i]->device == dev) {   // streams and buffers belong to one device
          w = g_ws_free[i];
          g_ws_free.erase(g_ws_free.begin() + long(i));
          break;
        }
    }
    if (w) {
      if (w->busy && host_wait) {
        cudaEventSynchronize(w->busy_ev);
        w->busy = false;
      }
      return w;
    }
  }
  Workspace *w = new Workspace();
  w->device = dev;
  if (cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess) w->stream = nullptr;
  return w;
}
void release_workspace(Workspace *w) {
  std::lock_guard<std::mutex> l(g_ws_mu);
  g_ws_free.push_back(w);
}

const char *get_device_plan(double io_ratio, const QSpec &q, const RSpec &r, double multiplier,
                            bool arith_double, std::shared_ptr<DevicePlan> *out) {
  PlanKey key;
  std::memset(&key, 0, sizeof key);
  std::memcpy(&key.ratio_bits, &io_ratio, 8);
  std::memcpy(&key.gain_bits, &multiplier, 8);
  key.precision = q.precision;
  key.passband_end = q.passband_end;
  key.stopband_begin = q.stopband_begin;
  key.phase_response = q.phase_response;
  key.qflags = q.flags;
  key.rflags = r.flags & 0xff;
  key.min_dft = r.log2_min_dft_size;
  key.large_dft = r.log2_large_dft_size;
  key.coef_kb = r.coef_size_kbytes;
  int dev = 0;
  B2_CUDA(cudaGetDevice(&dev));
  key.device = dev;
  key.arith_double = arith_double ? 1 : 0;
  std::lock_guard<std::mutex> l(g_plan_mu);
  auto it = g_plans.find(key);
  if (it != g_plans.end()) { it->second.last_use = ++g_plan_tick; *out = it->second.dp; return nullptr; }
  auto dp = std::make_shared<DevicePlan>();
  if (const char *e = make_plan(io_ratio, q, r, multiplier, &dp->plan)) return e;
  dp->is_double = arith_double;
  dp->stages.resize(dp->plan.stages.size());
  for (size_t i = 0; i < dp->plan.stages.size(); ++i) {
    const Stage &s = dp->plan.stages[i];
    DevStage &d = dp->stages[i];
    const char *e = nullptr;
    if (dp->is_double) {
      e = upload<double>(s.g, &d.g);
      if (!e && s.kind == kPolyInterp) e = upload<double>(s.table, &d.table);
    } else {
      e = upload<float>(s.g, &d.g);
      if (!e && s.kind == kPolyInterp) e = upload<float>(s.table, &d.table);
    }
    if (e) return e;
  }
  if (const char *e = fft_stage_prepare(dp.get())) return e;
  if (const char *e = fdr_prepare(dp.get())) return e;
  {
    // one line saying which kernels serve the plan (no silent fall-through to the slow paths)
    std::string exact;
    const size_t S = dp->plan.stages.size();
    for (size_t i = 0; i < S; ++i) {
      const Stage &st = dp->plan.stages[i];
      const bool fused = dp->fft && dp->fft->stage_index == int(i);
      char buf[128];
      const char *kind = st.kind == kHalf ? "half-band" : st.kind == kDft ? "dft" : st.kind == kPoly ? "vpoly0"
                       : st.kind == kPolyInterp ? "vpoly-interp" : "cubic";
      if (fused) {
        std::snprintf(buf, sizeof buf, "k_fft_stage[dft N=%d %d/%d%s]", st.dft_length, st.L, st.M,
                      dp->fft->fused_poly ? " + vpoly0 fused" : "");
        if (dp->fft->fused_poly) ++i;
      } else {
        std::snprintf(buf, sizeof buf, "%s[%s %d/%d direct form]",
                      st.kind == kPolyInterp ? "k_poly_interp" : st.kind == kCubic ? "k_cubic" : "k_rational", kind, st.L, st.M);
      }
      exact += (exact.empty() ? "" : " -> ");
      exact += buf;
    }
    if (exact.empty()) exact = "copy (unity ratio)";
    std::string batch = "same as exact grid";
    if (dp->fdr) {
      char buf[96];
      std::snprintf(buf, sizeof buf, "k_fdr[%d -> %d points, variant %d]", dp->fdr->nin, dp->fdr->nout, dp->fdr->variant);
      batch = buf;
    }
    dp->path = std::string(dp->is_double ? "b200-cr64" : "b200-cr32") + " | exact grid (resample, streams): " + exact +
               " | resample_batch default: " + batch;
  }
  if (g_plans.size() >= kMaxCachedPlans) {
    auto victim = g_plans.begin();
    for (auto j = g_plans.begin(); j != g_plans.end(); ++j)
      if (j->second.last_use < victim->second.last_use) victim = j;
    g_plans.erase(victim);
  }
  g_plans[key] = PlanEntry{dp, ++g_plan_tick};
  *out = dp;
  return nullptr;
}

// ---------------------------------------------------------------------------

namespace {
inline dim3 grid_for(size_t n, unsigned n_sig, unsigned threads = 256) {
  return dim3(unsigned((n + threads - 1) / threads), std::min(n_sig, 65535u), 1);
}

template <typename Tin, typename T>
const char *convert_in_t(const IoLayout &lay, const void *in, size_t n_frames, T *X, size_t xs, size_t xo,
                         cudaStream_t st) {
  const unsigned n_sig = lay.n_clips * lay.channels;
  if (!n_frames || !n_sig) return nullptr;
  k_convert_in<Tin, T><<<grid_for(n_frames, n_sig), 256, 0, st>>>(
      static_cast<const Tin *>(in), lay.clip_stride, lay.frame_stride, lay.chan_stride, lay.channels, n_frames,
      n_sig, X, xs, xo);
  B2_COUNT_LAUNCH();
  B2_CUDA(cudaGetLastError());
  return nullptr;
}
template <typename T>
const char *convert_in_d(const IoLayout &lay, const void *in, size_t n, void *X, size_t xs, size_t xo, cudaStream_t st) {
  switch (lay.dtype) {
    case 0: return convert_in_t<float, T>(lay, in, n, static_cast<T *>(X), xs, xo, st);
    case 1: return convert_in_t<double, T>(lay, in, n, static_cast<T *>(X), xs, xo, st);
    case 2: return convert_in_t<int, T>(lay, in, n, static_cast<T *>(X), xs, xo, st);
    case 3: return convert_in_t<short, T>(lay, in, n, static_cast<T *>(X), xs, xo, st);
  }
  return "invalid io datatype(s)";
}
}  // namespace

const char *convert_in(const IoLayout &lay, const void *dev_in, size_t n_frames, bool is_double, void *X,
                       size_t x_stride, size_t x_off, cudaStream_t stream) {
  return is_double ? convert_in_d<double>(lay, dev_in, n_frames, X, x_stride, x_off, stream)
                   : convert_in_d<float>(lay, dev_in, n_frames, X, x_stride, x_off, stream);
}

namespace {
template <typename T, int BITS>
const char *convert_out_int(const IoLayout &lay, void *out, size_t frame_off, size_t n_frames, const T *Y,
                            size_t ys, const DitherCfg &dc, unsigned long long *clips_dev, Workspace *ws,
                            cudaStream_t st) {
  const unsigned n_sig = lay.n_clips * lay.channels;
  // block partition (one block == one reference soxr_process call)
  std::vector<DitherBlock> blocks;
  if (dc.enabled && !dc.call_lens.empty()) {
    long long pos = 0;
    for (long long n : dc.call_lens) { blocks.push_back({pos, n}); pos += n; }
  } else {
    blocks.push_back({0, (long long)n_frames});
  }
  const unsigned nb = unsigned(blocks.size());
  long long max_n = 0;
  for (auto &b : blocks) max_n = std::max(max_n, b.n);
  const bool by_value = nb <= 8;
  DitherBlocks8 b8{};
  DitherBlocksPtr bp{nullptr};
  if (by_value) {
    for (unsigned i = 0; i < nb; ++i) b8.b[i] = blocks[i];
  } else {
    if (const char *e = ws->blocks.reserve(nb * sizeof(DitherBlock))) return e;
    if (const char *e = ws->pin_misc.reserve(nb * sizeof(DitherBlock))) return e;
    std::memcpy(ws->pin_misc.p, blocks.data(), nb * sizeof(DitherBlock));
    B2_CUDA(cudaMemcpyAsync(ws->blocks.p, ws->pin_misc.p, nb * sizeof(DitherBlock), cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaStreamSynchronize(st));   // the pinned list is rewritten by the next call (callers may return without waiting)
    bp.p = ws->blocks.as<DitherBlock>();
  }
  ulonglong2 *seeds = nullptr;
  if (dc.enabled) {
    if (const char *e = ws->seeds.reserve(size_t(lay.n_clips) * nb * lay.channels * sizeof(ulonglong2))) return e;
    seeds = ws->seeds.as<ulonglong2>();
    if (by_value)
      k_dither_seeds<T, DitherBlocks8><<<(lay.n_clips + 63) / 64, 64, 0, st>>>(
          Y, ys, lay.channels, lay.n_clips, b8, nb, dc.seed0, dc.seed_in_dev, dc.split, seeds, dc.seed_end_dev);
    else
      k_dither_seeds<T, DitherBlocksPtr><<<(lay.n_clips + 63) / 64, 64, 0, st>>>(
          Y, ys, lay.channels, lay.n_clips, bp, nb, dc.seed0, dc.seed_in_dev, dc.split, seeds, dc.seed_end_dev);
    B2_COUNT_LAUNCH();
    B2_CUDA(cudaGetLastError());
  }
  if (n_sig && nb) {
    const unsigned max_groups = unsigned(max_n / 16 + 1);
    typedef typename IntOf<BITS>::type I;
    // interleaved output with an even channel count: a thread takes the channel pair (c, c + 1) of its group and
    // stores one 2-sample element per frame (k_rint_clip<PAIR>)
    const bool pair = lay.channels % 2 == 0 && lay.chan_stride == 1 && lay.frame_stride % 2 == 0 && lay.clip_stride % 2 == 0 &&
                      !(reinterpret_cast<unsigned long long>(out) % (2 * sizeof(I)));
    const unsigned n_units = pair ? n_sig / 2 : n_sig;
    // A thread keeps its group index and walks gridDim.z-strided units (its LCG jump to that group is built once):
    // about 24 CTAs of work per SM in all, the rest of the batch goes into the per-thread loop.
    unsigned long long live_ctas = 0;   // per unit: CTAs that own at least one group
    for (auto &b : blocks) live_ctas += (unsigned long long)(b.n / 16 + 1 + 127) / 128;
    const unsigned long long want_z = (148ull * 24 + live_ctas - 1) / std::max<unsigned long long>(1, live_ctas);
    dim3 grid((max_groups + 127) / 128, nb, unsigned(std::min<unsigned long long>(std::max<unsigned long long>(1, want_z), std::min(n_units, 65535u))));
    if (grid.y > 65535) return "too many dither blocks";
    I *o = static_cast<I *>(out);
#define B2_K7_LAUNCH(BL, PAIR, blk)                                                                                       \
    k_rint_clip<T, BITS, BL, PAIR><<<grid, 128, 0, st>>>(Y, ys, lay.channels, n_sig, blk, nb, max_groups, seeds, o,        \
                                                        lay.clip_stride, lay.frame_stride, lay.chan_stride,               \
                                                        (long long)frame_off, clips_dev)
    if (by_value) {
      if (pair) B2_K7_LAUNCH(DitherBlocks8, true, b8); else B2_K7_LAUNCH(DitherBlocks8, false, b8);
    } else {
      if (pair) B2_K7_LAUNCH(DitherBlocksPtr, true, bp); else B2_K7_LAUNCH(DitherBlocksPtr, false, bp);
    }
#undef B2_K7_LAUNCH
    B2_COUNT_LAUNCH();
    B2_CUDA(cudaGetLastError());
  }
  return nullptr;
}

template <typename T>
const char *convert_out_t(const IoLayout &lay, void *out, size_t frame_off, size_t n_frames, const T *Y,
                          size_t ys, const DitherCfg &dc, unsigned long long *clips_dev, Workspace *ws,
                          cudaStream_t st) {
  const unsigned n_sig = lay.n_clips * lay.channels;
  switch (lay.dtype) {
    case 0:
    case 1:
      if (!n_frames || !n_sig) return nullptr;
      if (lay.dtype == 0)
        k_convert_out_float<T, float><<<grid_for(n_frames, n_sig), 256, 0, st>>>(
            Y, ys, n_frames, lay.channels, n_sig, static_cast<float *>(out), lay.clip_stride, lay.frame_stride,
            lay.chan_stride, frame_off);
      else
        k_convert_out_float<T, double><<<grid_for(n_frames, n_sig), 256, 0, st>>>(
            Y, ys, n_frames, lay.channels, n_sig, static_cast<double *>(out), lay.clip_stride, lay.frame_stride,
            lay.chan_stride, frame_off);
      B2_COUNT_LAUNCH();
      B2_CUDA(cudaGetLastError());
      return nullptr;
    case 2: {
      DitherCfg nd;   // int32 is never dithered (data-io.c:207-209)
      return convert_out_int<T, 32>(lay, out, frame_off, n_frames, Y, ys, nd, clips_dev, ws, st);
    }
    case 3: return convert_out_int<T, 16>(lay, out, frame_off, n_frames, Y, ys, dc, clips_dev, ws, st);
  }
  return "invalid io datatype(s)";
}
}  // namespace

const char *convert_out(const IoLayout &lay, void *dev_out, size_t frame_off, size_t n_frames, bool is_double,
                        const void *Y, size_t y_stride, const DitherCfg &dc, unsigned long long *clips_dev,
                        Workspace *ws, cudaStream_t stream) {
  return is_double ? convert_out_t<double>(lay, dev_out, frame_off, n_frames, static_cast<const double *>(Y),
                                           y_stride, dc, clips_dev, ws, stream)
                   : convert_out_t<float>(lay, dev_out, frame_off, n_frames, static_cast<const float *>(Y),
                                          y_stride, dc, clips_dev, ws, stream);
}

// ---------------------------------------------------------------------------
// Range executor
// ---------------------------------------------------------------------------

namespace {

struct Range { long lo, hi; };

// Output range of stage s-1 needed to produce outputs [r.lo, r.hi) of stage `last`,
// where stages [s, last] run as one launch.
Range input_range(const DevicePlan &dp, size_t s, size_t last, Range r, bool generic = false) {
  if (r.hi <= r.lo) return {0, 0};
  Range need;
  if (!generic && dp.fft && dp.fft->stage_index == int(s)) {
    const RangeLH t = fft_stage_input_range(*dp.fft, r.lo, r.hi);
    need = {t.lo, t.hi};
  } else {
    const Stage &st = dp.plan.stages[s];
    need = {stage_input_first(st, r.lo), stage_input_needed(st, r.hi)};
  }
  (void)last;
  need.lo = std::max(need.lo, 0L);
  need.hi = std::max(need.hi, need.lo);
  return need;
}

// element-wise copy between two descriptors (K1 / float K7 when no fused launch touches the user buffers)
template <typename T>
__global__ void k_sig_copy(SigDesc in, SigDesc out, long o0, long n, unsigned n_sig) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long k = o0 + i;
  for (unsigned s = blockIdx.y; s < n_sig; s += gridDim.y) {
    T v = T(0);
    if (k >= in.vlo && k < in.vhi)
      v = ld_as<T>(in.p, in.dtype, (long long)(s / in.channels) * in.clip_stride + (long long)(s % in.channels) * in.chan_stride +
                                       (long long)(k - in.base) * in.frame_stride);
    st_as<T>(out.p, out.dtype, (long long)(s / out.channels) * out.clip_stride + (long long)(s % out.channels) * out.chan_stride +
                                   (long long)(k - out.base) * out.frame_stride, v);
  }
}

template <typename T>
const char *sig_copy(const SigDesc &in, const SigDesc &out, long o0, long o1, unsigned n_sig, cudaStream_t st) {
  if (o1 <= o0 || !n_sig) return nullptr;
  k_sig_copy<T><<<grid_for(size_t(o1 - o0), n_sig), 256, 0, st>>>(in, out, o0, o1 - o0, n_sig);
  B2_COUNT_LAUNCH();
  B2_CUDA(cudaGetLastError());
  return nullptr;
}

template <typename T> bool is_planar(const SigDesc &d) {
  return d.dtype == (sizeof(T) == 8 ? 1 : 0) && d.frame_stride == 1 && d.channels == 1;
}
// engine-precision samples in any layout: what k_poly_interp / k_cubic read and write directly
template <typename T> bool is_engine_type(const SigDesc &d) { return d.dtype == (sizeof(T) == 8 ? 1 : 0); }
inline bool takes_any_layout(const Stage &s) { return s.kind == kPolyInterp || s.kind == kCubic; }
template <typename T> SigIn<T> sig_in_of(const SigDesc &d) {
  SigIn<T> g{static_cast<const T *>(d.p), size_t(d.clip_stride), d.base, d.vlo, d.vhi};
  g.chan_stride = d.chan_stride; g.frame_stride = d.frame_stride; g.channels = d.channels;
  return g;
}
template <typename T> SigOut<T> sig_out_of(const SigDesc &d) {
  SigOut<T> g{static_cast<T *>(d.p), size_t(d.clip_stride), d.base};
  g.chan_stride = d.chan_stride; g.frame_stride = d.frame_stride; g.channels = d.channels;
  return g;
}

template <typename T>
const char *run_stage_generic(const DevicePlan &dp, size_t si, SigIn<T> in, SigOut<T> out, Range r,
                              unsigned n_sig, cudaStream_t st) {
  const Stage &s = dp.plan.stages[si];
  const DevStage &d = dp.stages[si];
  const long n_out = r.hi - r.lo;
  if (n_out <= 0 || !n_sig) return nullptr;
  const dim3 grid = grid_for(size_t(n_out), n_sig, 128);
  switch (s.kind) {
    case kHalf: case kDft: case kPoly: {
      // 8 signals per thread when that still leaves enough blocks to fill the GPU, else one
      if (size_t(n_sig / 8) * grid.x >= 2368) {
        const dim3 g8(grid.x, std::min((n_sig + 7) / 8, 65535u), 1);
        k_rational<T, 8><<<g8, 128, 0, st>>>(in, d.g.as<T>(), int(s.g.size()), s.C, s.L, s.M, out, r.lo, n_out, n_sig);
      } else {
        k_rational<T, 1><<<grid, 128, 0, st>>>(in, d.g.as<T>(), int(s.g.size()), s.C, s.L, s.M, out, r.lo, n_out, n_sig);
      }
      break;
    }
    case kPolyInterp: {
      // signals per block: enough blocks to fill the GPU (148 SMs x 16 resident blocks), at most 16 signals each
      const unsigned bx = grid.x;
      const unsigned spb = unsigned(std::max<size_t>((n_sig + 65534u) / 65535u, std::min<size_t>(16, size_t(n_sig) * bx / 2368)));
      const dim3 g2(bx, (n_sig + spb - 1) / spb, 1);
      const size_t smem = size_t(s.num_coefs) * 128 * sizeof(T);
      if (g2.y > 65535u || smem > 200 * 1024) return "polyphase stage too large for the interpolating kernel";
      if (smem > 48 * 1024)   // rare: more than 96 (f32) / 48 (f64) taps per phase
        B2_CUDA(cudaFuncSetAttribute(k_poly_interp<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
      k_poly_interp<T><<<g2, 128, smem, st>>>(in, d.table.as<T>(), s.num_coefs, s.order, s.phase_bits, s.at0_fix,
                                               s.step_fix, s.dev_preload, out, r.lo, n_out, n_sig, spb);
      break;
    }
    case kCubic:
      k_cubic<T><<<grid, 128, 0, st>>>(in, s.step_fix, s.cubic_mult, out, r.lo, n_out, n_sig);
      break;
  }
  B2_COUNT_LAUNCH();
  B2_CUDA(cudaGetLastError());
  return nullptr;
}

template <typename T>
const char *run_range_t(const DevicePlan &dp, const SigDesc &in0, long n_in_total, unsigned n_sig, unsigned group, long o0, long o1,
                        const SigDesc &out0, Workspace *ws, cudaStream_t st, unsigned flags) {
  const size_t S = dp.plan.stages.size();
  if (o1 <= o0 || !n_sig) return nullptr;
  SigDesc cur = in0;
  if (n_in_total >= 0) cur.vhi = std::min(cur.vhi, n_in_total);
  if ((flags & kRunAllowFdr) && !(flags & kRunGeneric) && dp.fdr && sizeof(T) == 4 && n_in_total >= 0 &&
      out0.dtype <= 1 && in0.base == 0) {
    // whole signals: single-pass frequency-domain kernel (the first few outputs, where the reference's
    // intermediate signal is cut at index 0, cr.c:487-490, come from its direct-form head kernel)
    return fdr_run(dp, *dp.fdr, cur, out0, o0, o1, n_sig, group, st, ws);
  }
  if (S == 0)   // unity ratio: samples pass through (cr.c:347; zero-extended by flush)
    return sig_copy<T>(cur, out0, o0, o1, n_sig, st);

  // launches: [first, last] stage index ranges, and the output range of each, back to front
  struct Launch { size_t first, last; bool fft; Range r; };
  std::vector<Launch> launches;
  for (size_t s = 0; s < S; ++s) {
    const bool use_fft = !(flags & kRunGeneric) && dp.fft && dp.fft->stage_index == int(s);
    const size_t last = (use_fft && dp.fft->fused_poly) ? s + 1 : s;
    launches.push_back({s, last, use_fft, {0, 0}});
    s = last;
  }
  launches.back().r = {o0, o1};
  for (size_t i = launches.size() - 1; i > 0; --i)
    launches[i - 1].r = input_range(dp, launches[i].first, launches[i].last, launches[i].r, (flags & kRunGeneric) != 0);
  const Range xin = input_range(dp, launches[0].first, launches[0].last, launches[0].r, (flags & kRunGeneric) != 0);
  if (xin.hi > xin.lo && xin.lo < in0.base) return "internal: input history trimmed too far";
  if (n_in_total < 0 && xin.hi > in0.vhi) return "internal: output range needs input that has not arrived";

  DevBuf *cur_buf = nullptr;   // which scratch buffer `cur` lives in
  for (size_t i = 0; i < launches.size(); ++i) {
    const Launch &ln = launches[i];
    const bool is_last = i + 1 == launches.size();
    const size_t len = size_t(std::max(0L, ln.r.hi - ln.r.lo));
    // generic kernels read planar engine-precision input only (K1 as a separate pass)
    const bool any_layout = !ln.fft && takes_any_layout(dp.plan.stages[ln.first]);
    if (!ln.fft && !is_planar<T>(cur) && !(any_layout && is_engine_type<T>(cur))) {
      const long lo = std::max(xin.lo, cur.vlo), hi = std::max(lo, std::min(xin.hi, cur.vhi));
      const size_t stride = (size_t(hi - lo) + 63) & ~size_t(63);
      if (const char *e = ws->x.reserve(std::max<size_t>(1, stride * n_sig) * sizeof(T))) return e;
      const SigDesc px = planar_desc(ws->x.p, sizeof(T) == 8, stride, lo, lo, hi);
      if (const char *e = sig_copy<T>(cur, px, lo, hi, n_sig, st)) return e;
      cur = px;
    }
    SigDesc dst;
    DevBuf *dst_buf = nullptr;
    const bool direct = is_last && (ln.fft || is_planar<T>(out0) || (any_layout && is_engine_type<T>(out0)));
    if (direct) dst = out0;
    else {
      dst_buf = (cur_buf == &ws->t0) ? &ws->t1 : &ws->t0;
      const size_t stride = (len + 63) & ~size_t(63);
      if (const char *e = dst_buf->reserve(std::max<size_t>(1, stride * n_sig) * sizeof(T))) return e;
      // intermediate signals exist for index >= 0 only (FIFO preloads are zeros, cr.c:487-490)
      dst = planar_desc(dst_buf->p, sizeof(T) == 8, stride, ln.r.lo, std::max(0L, ln.r.lo), ln.r.hi);
    }
    const char *e;
    if (ln.fft) e = fft_stage_run(dp, *dp.fft, cur, dst, ln.r.lo, ln.r.hi, n_sig, group, st);
    else {
      const SigIn<T> gi = sig_in_of<T>(cur);
      const SigOut<T> go = sig_out_of<T>(dst);
      e = run_stage_generic<T>(dp, ln.first, gi, go, ln.r, n_sig, st);
    }
    if (e) return e;
    if (is_last && !direct) return sig_copy<T>(dst, out0, o0, o1, n_sig, st);
    cur = dst;
    cur_buf = dst_buf;
  }
  return nullptr;
}
}  // namespace

const char *run_range(const DevicePlan &dp, const SigDesc &in, long n_in_total, unsigned n_sig, unsigned group,
                      long o0, long o1, const SigDesc &out, Workspace *ws, cudaStream_t stream, unsigned flags) {
  return dp.is_double ? run_range_t<double>(dp, in, n_in_total, n_sig, group, o0, o1, out, ws, stream, flags)
                      : run_range_t<float>(dp, in, n_in_total, n_sig, group, o0, o1, out, ws, stream, flags);
}

long history_needed_from(const DevicePlan &dp, long o, bool generic) {
  Range r{o, o + 1};
  std::vector<std::pair<size_t, size_t>> launches;
  const size_t S = dp.plan.stages.size();
  for (size_t s = 0; s < S; ++s) {
    const bool use_fft = !generic && dp.fft && dp.fft->stage_index == int(s);
    const size_t last = (use_fft && dp.fft->fused_poly) ? s + 1 : s;
    launches.push_back({s, last});
    s = last;
  }
  for (size_t i = launches.size(); i-- > 0;) r = input_range(dp, launches[i].first, launches[i].second, r, generic);
  return r.lo;
}

}  // namespace b2
