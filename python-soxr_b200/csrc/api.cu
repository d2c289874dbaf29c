// api.cu -- the C ABI of libb200soxr.so (include/b200soxr.h).
//
// Host-side mirror of libsoxr's public API for the constant-rate path
// (libsoxr/src/soxr.c) plus the batch entry point that replaces the per-clip
// driver loops of python-soxr's nanobind binding (src/soxr_ext.cpp:210-402).
// There is NO CPU fallback: without a usable CUDA device every compute entry
// point returns an error string.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iterator>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/b200soxr.h"
#include "plan.h"
#include "fft_stage.cuh"
#include "runtime.h"
#include "vr.h"

using namespace b2;

namespace {

const char *kNoDevice = "b200soxr: no usable CUDA device (this library has no CPU fallback)";

bool device_ok() {
  static std::once_flag once;
  static bool ok = false;
  std::call_once(once, [] {
    int n = 0;
    ok = cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
    if (!ok) cudaGetLastError();
  });
  return ok;
}

QSpec to_qspec(const b200soxr_quality_spec_t &q) {
  QSpec s;
  s.precision = q.precision;
  s.phase_response = q.phase_response;
  s.passband_end = q.passband_end;
  s.stopband_begin = q.stopband_begin;
  s.flags = q.flags;
  s.error = static_cast<const char *>(q.e);
  // backwards compatibility with the original API (soxr.c:408-413)
  if (s.passband_end > 2) s.passband_end /= 100;
  if (s.stopband_begin > 2) s.stopband_begin = 2 - s.stopband_begin / 100;
  return s;
}
// libsoxr's run-time overrides from the environment (soxr.c:340-363,426-433): a number inside its range replaces the
// field, a flag value replaces its bit field.  Read at every create like the reference does.
void runtime_num(const char *name, int lo, int hi, unsigned *field) {
  if (const char *e = std::getenv(name)) {
    const int i = std::atoi(e);
    if (i >= lo && i <= hi) *field = unsigned(i);
  }
}
void runtime_flag(const char *name, unsigned n_bits, unsigned n_shift, unsigned long *flags) {
  if (const char *e = std::getenv(name)) {
    const int i = std::atoi(e);
    const unsigned long mask = (1ul << n_bits) - 1;
    if (i >= 0 && i <= int(mask)) *flags = (*flags & ~(mask << n_shift)) | ((unsigned long)i << n_shift);
  }
}
RSpec to_rspec(const b200soxr_runtime_spec_t *r) {
  RSpec s;
  if (r) {
    s.log2_min_dft_size = r->log2_min_dft_size;
    s.log2_large_dft_size = r->log2_large_dft_size;
    s.coef_size_kbytes = r->coef_size_kbytes;
    s.flags = r->flags & 0xff;      // (the bits above are this library's own extensions)
  }
  runtime_num("SOXR_MIN_DFT_SIZE", 8, 15, &s.log2_min_dft_size);
  runtime_num("SOXR_LARGE_DFT_SIZE", 8, 20, &s.log2_large_dft_size);
  runtime_num("SOXR_COEFS_SIZE", 100, 800, &s.coef_size_kbytes);
  runtime_flag("SOXR_COEF_INTERP", 2, 0, &s.flags);
  runtime_flag("SOXR_NOSMALLINTOPT", 1, 3, &s.flags);
  // (SOXR_NUM_THREADS and SOXR_STRICT_BUF steer the reference's OpenMP loop and its buffering check; neither has a
  // counterpart here)
  return s;
}
const double kFullScale[] = {1, 1, 65536. * 32768, 32768};   // soxr.c:377

// Arithmetic type of the device path.  The reference runs every recipe above 20 bits in
// its f64 engine (soxr.c:447) whatever the I/O type.  With float32 in AND out the result is
// rounded to 24 bits anyway, so the f32 path (measured ~2e-7 RMS from the f64 engine on
// unit-variance noise, inside the 1e-6 parity budget) is used unless the caller asked for
// SOXR_DOUBLE_PRECISION or B200SOXR_STRICT_ENGINE=1 is set.  The filters are always the
// reference's for the recipe.
bool arith_double_for(const QSpec &q, int itype, int otype) {
  const bool ref_double = !(q.precision <= 20 && !(q.flags & B200SOXR_DOUBLE_PRECISION));
  if (!ref_double) return false;
  if ((itype & 3) != B200SOXR_FLOAT32 || (otype & 3) != B200SOXR_FLOAT32) return true;
  if (q.flags & B200SOXR_DOUBLE_PRECISION) return true;
  static const bool strict = [] { const char *e = std::getenv("B200SOXR_STRICT_ENGINE"); return e && e[0] == '1'; }();
  return strict;
}

// B200SOXR_TRACE_TIMING=1: per-handle wall-clock breakdown of b200soxr_process, printed to stderr by b200soxr_delete
// (where the time of a stream call goes: history append, queueing device work, waiting for it, copy-out)
const bool kTraceTiming = [] { const char *e = std::getenv("B200SOXR_TRACE_TIMING"); return e && e[0] == '1'; }();
inline double now_us() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return double(ts.tv_sec) * 1e6 + double(ts.tv_nsec) * 1e-3;
}

size_t div_len_of(double in_rate, double out_rate) {          // soxr_ext.cpp:71,239
  return size_t(std::max(1000., 48000 * in_rate / out_rate));
}

}  // namespace

// ---------------------------------------------------------------------------
// stream handle
// ---------------------------------------------------------------------------
struct b200soxr {
  int device = -1;        // the CUDA device this handle's plan, stream and buffers live on
  unsigned channels = 0;
  unsigned n_streams = 1; // batched handle (b200soxr_create_batch): independent streams of `channels` channels in lockstep
  double io_ratio = 0;
  b200soxr_io_spec_t io{};
  QSpec q;
  RSpec r;
  bool device_io = false;
  bool low_latency = false;   // B200SOXR_RT_LOW_LATENCY: direct-form kernels, every output whose support has arrived
  std::shared_ptr<DevicePlan> dp;
  std::unique_ptr<CallEmu> emu;
  std::unique_ptr<VrEngine> vr;   // quality_spec.flags & B200SOXR_VR: the variable-rate engine (vr.h) instead of a plan
  double vr_scale = 1;
  Workspace *ws = nullptr;
  cudaStream_t stream = nullptr;
  // Input history, one of two homes:
  //  * host_hist: mapped pinned HOST memory in the caller's own dtype, [stream][frame][channel].  A call that returns
  //    no output is then a memcpy and nothing else (no launch, no copy engine, no synchronise); a call that returns
  //    output is ONE launch whose first kernel reads the history over PCIe (K1 fused) and whose last kernel writes
  //    the result into mapped pinned memory (K7 fused), plus one synchronise.  This is the home for streams fed in
  //    small chunks (real-time use, BASELINE configs[2]).
  //  * xbuf: planar engine-precision DEVICE memory (large chunks, split layouts, device I/O).
  // A handle starts on the host and moves to the device for good at the first chunk above the zero-copy size.
  bool host_hist = false;
  PinnedBuf hx;
  size_t hx_cap = 0;      // frames per stream in hx
  DevBuf xbuf, clips_dev, seed_dev;
  size_t x_cap = 0;       // elements per signal in xbuf
  long x_base = 0, x_hi = 0, n_in_total = -1, out_done = 0;
  size_t clips = 0;
  unsigned long long seed = 0;   // create-time dither seed (soxr.c:439); the running per-stream seeds live in seed_dev
  bool flushing = false;
  const char *error = nullptr;
  double t_in = 0, t_queue = 0, t_wait = 0, t_out = 0;   // B200SOXR_TRACE_TIMING
  unsigned long n_calls = 0, n_due = 0;

  ~b200soxr() {
    if (kTraceTiming && n_calls)
      std::fprintf(stderr, "b200soxr timing: %lu calls (%lu with output): append %.1f us, queue %.1f us, wait %.1f us, copy-out %.1f us per call\n",
                   n_calls, n_due, t_in / n_calls, t_queue / n_calls, t_wait / n_calls, t_out / n_calls);
    DeviceGuard guard(device);
    if (stream) cudaStreamSynchronize(stream);
    hx.release();
    xbuf.release();
    clips_dev.release();
    seed_dev.release();
    if (ws) release_workspace(ws);
  }
  unsigned n_sig() const { return n_streams * channels; }
  size_t esz() const { return dp->is_double ? 8 : 4; }
  const char *reset_seeds(unsigned long long v) {
    std::vector<unsigned long long> sd(n_streams, v);
    B2_CUDA(cudaMemcpyAsync(seed_dev.p, sd.data(), sd.size() * 8, cudaMemcpyHostToDevice, stream));
    B2_CUDA(cudaStreamSynchronize(stream));   // sd is a stack-owned staging vector
    return nullptr;
  }
  const char *init_state() {
    emu.reset(new CallEmu(dp->plan));
    x_base = x_hi = out_done = 0;
    n_in_total = -1;
    flushing = false;
    host_hist = !device_io && !(io.itype & B200SOXR_SPLIT);
    return nullptr;
  }
  // ---- host-side history ------------------------------------------------------------------------------------
  // make room for `extra` more frames per stream, dropping history no output can need
  const char *grow_host_history(size_t extra) {
    const long keep_from = std::max(x_base, std::min(x_hi, history_needed_from(*dp, out_done, low_latency)));
    const size_t live = size_t(x_hi - keep_from);
    const size_t fsz = size_t(channels) * dtype_size(io.itype & 3);
    if (size_t(x_hi - x_base) + extra <= hx_cap) return nullptr;
    if (live + extra <= hx_cap) {        // compact in place
      for (unsigned b = 0; b < n_streams; ++b) {
        char *row = static_cast<char *>(hx.p) + size_t(b) * hx_cap * fsz;
        std::memmove(row, row + size_t(keep_from - x_base) * fsz, live * fsz);
      }
      x_base = keep_from;
      return nullptr;
    }
    size_t new_cap = std::max<size_t>(4096, (live + extra) * 2);
    PinnedBuf nb;
    if (const char *e = nb.reserve(new_cap * fsz * n_streams)) return e;
    for (unsigned b = 0; b < n_streams; ++b)
      std::memcpy(static_cast<char *>(nb.p) + size_t(b) * new_cap * fsz,
                  static_cast<char *>(hx.p) + (size_t(b) * hx_cap + size_t(keep_from - x_base)) * fsz, live * fsz);
    hx.release();
    hx = nb;
    hx_cap = new_cap;
    x_base = keep_from;
    return nullptr;
  }
  SigDesc host_desc() const {
    SigDesc d;
    d.p = hx.p; d.dtype = io.itype & 3; d.channels = channels;
    d.clip_stride = (long long)(hx_cap * channels); d.frame_stride = channels; d.chan_stride = 1;
    d.base = x_base; d.vlo = std::max(0L, x_base); d.vhi = x_hi;
    return d;
  }
  // move the live history into planar device memory (first large chunk): K1 reads the pinned rows
  const char *migrate_to_device(size_t extra) {
    const long keep_from = std::max(x_base, std::min(x_hi, history_needed_from(*dp, out_done, low_latency)));
    const size_t live = size_t(x_hi - keep_from);
    size_t new_cap = std::max<size_t>(4096, (live + extra) * 2);
    new_cap = (new_cap + 63) & ~size_t(63);
    xbuf.release();
    if (const char *e = xbuf.reserve(new_cap * n_sig() * esz())) return e;
    x_cap = new_cap;
    if (live) {
      IoLayout lay;
      lay.dtype = io.itype & 3; lay.n_clips = n_streams; lay.channels = channels;
      lay.clip_stride = (long long)(hx_cap * channels); lay.frame_stride = channels; lay.chan_stride = 1;
      const char *src = static_cast<const char *>(hx.p) + size_t(keep_from - x_base) * channels * dtype_size(io.itype & 3);
      if (const char *e = convert_in(lay, src, live, dp->is_double, xbuf.p, x_cap, 0, stream)) return e;
      B2_CUDA(cudaStreamSynchronize(stream));
    }
    x_base = keep_from;
    host_hist = false;
    hx.release();
    hx_cap = 0;
    return nullptr;
  }
  // ---- device-side history ----------------------------------------------------------------------------------
  // make room for `extra` more samples per signal, dropping history no output can need
  const char *grow_history(size_t extra) {
    const long keep_from = std::max(x_base, std::min(x_hi, history_needed_from(*dp, out_done, low_latency)));
    const size_t live = size_t(x_hi - keep_from);
    const size_t es = esz();
    const unsigned ns = n_sig();
    if (live + extra <= x_cap) {
      if (keep_from > x_base && live + extra + size_t(keep_from - x_base) > x_cap) {
        // compact in place (per signal; ranges may overlap -> go through scratch)
        if (const char *e = ws->misc.reserve(std::max<size_t>(1, live * ns) * es)) return e;
        B2_CUDA(cudaMemcpy2DAsync(ws->misc.p, live * es, static_cast<char *>(xbuf.p) + size_t(keep_from - x_base) * es,
                                  x_cap * es, live * es, ns, cudaMemcpyDeviceToDevice, stream));
        B2_CUDA(cudaMemcpy2DAsync(xbuf.p, x_cap * es, ws->misc.p, live * es, live * es, ns,
                                  cudaMemcpyDeviceToDevice, stream));
        x_base = keep_from;
      }
      if (size_t(x_hi - x_base) + extra <= x_cap) return nullptr;
    }
    // reallocate
    size_t new_cap = std::max<size_t>(4096, (live + extra) * 2);
    new_cap = (new_cap + 63) & ~size_t(63);
    DevBuf nb;
    if (const char *e = nb.reserve(new_cap * ns * es)) return e;
    if (live)
      B2_CUDA(cudaMemcpy2DAsync(nb.p, new_cap * es, static_cast<char *>(xbuf.p) + size_t(keep_from - x_base) * es,
                                x_cap * es, live * es, ns, cudaMemcpyDeviceToDevice, stream));
    B2_CUDA(cudaStreamSynchronize(stream));
    xbuf.release();
    xbuf = nb;
    x_cap = new_cap;
    x_base = keep_from;
    return nullptr;
  }
  // low-latency mode: number of outputs [0, n) whose whole filter support has arrived (all of them once flushing)
  long outputs_ready() const {
    const long total = long(output_frames(io_ratio, size_t(n_in_total >= 0 ? n_in_total : x_hi)));
    if (n_in_total >= 0) return total;
    long lo = out_done, hi = total + 1;          // outputs [0, lo) are ready; [0, hi) are not (zero-phase look-ahead)
    if (outputs_need(hi) <= x_hi) return hi;     // (degenerate: filters shorter than one sample)
    while (hi - lo > 1) {
      const long mid = lo + (hi - lo) / 2;
      if (outputs_need(mid) <= x_hi) lo = mid; else hi = mid;
    }
    return lo;
  }
  long outputs_need(long n_out) const {          // input frames [0, need) read by outputs [0, n_out)
    long need = n_out;
    for (size_t s = dp->plan.stages.size(); s-- > 0;) need = stage_input_needed(dp->plan.stages[s], need);
    return need;
  }
};

extern "C" {

b200soxr_quality_spec_t b200soxr_quality_spec(unsigned long recipe, unsigned long flags) {
  b200soxr_quality_spec_t out;
  std::memset(&out, 0, sizeof out);
  const QSpec q = quality_spec(recipe, flags);
  out.precision = q.precision;
  out.phase_response = q.phase_response;
  out.passband_end = q.passband_end;
  out.stopband_begin = q.stopband_begin;
  out.flags = q.flags;
  out.e = const_cast<char *>(q.error);
  return out;
}

b200soxr_io_spec_t b200soxr_io_spec(int itype, int otype) {
  b200soxr_io_spec_t s;
  std::memset(&s, 0, sizeof s);
  if ((itype | otype) >= B200SOXR_SPLIT * 2 || itype < 0 || otype < 0)
    s.e = const_cast<char *>("invalid io datatype(s)");
  else {
    s.itype = itype;
    s.otype = otype;
    s.scale = 1;
  }
  return s;
}

b200soxr_runtime_spec_t b200soxr_runtime_spec(unsigned num_threads) {
  b200soxr_runtime_spec_t s;
  std::memset(&s, 0, sizeof s);
  s.log2_min_dft_size = 10;
  s.log2_large_dft_size = 17;
  s.coef_size_kbytes = 400;
  s.num_threads = num_threads;
  return s;
}

char const *b200soxr_version(void) { return "b200soxr-0.1.0 (libsoxr-0.1.3 compatible)"; }

int b200soxr_device_available(void) { return device_ok() ? 1 : 0; }

unsigned long long b200soxr_launch_count(void) { return g_launches.load(); }

size_t b200soxr_output_frames(double in_rate, double out_rate, size_t ilen) {
  return output_frames(in_rate / out_rate, ilen);
}

size_t b200soxr_plan_dump(double in_rate, double out_rate, unsigned long quality, unsigned long q_flags,
                          char *buf, size_t buf_size) {
  Plan p;
  const QSpec q = quality_spec(quality, q_flags);
  const char *err = make_plan(in_rate / out_rate, q, to_rspec(nullptr), 1.0, &p);
  std::string text = err ? std::string("error: ") + err + "\n" : p.trace;
  if (buf && buf_size) {
    const size_t n = std::min(buf_size - 1, text.size());
    std::memcpy(buf, text.data(), n);
    buf[n] = 0;
  }
  return text.size() + 1;
}

b200soxr_error_t b200soxr_emulate_calls(double in_rate, double out_rate, unsigned long quality,
                                        unsigned long q_flags, size_t const *chunk_lens, size_t n_chunks,
                                        size_t olen_each, size_t *odone_each, double *delay_each) {
  Plan p;
  const QSpec q = quality_spec(quality, q_flags);
  if (const char *err = make_plan(in_rate / out_rate, q, to_rspec(nullptr), 1.0, &p)) return err;
  CallEmu emu(p);
  for (size_t i = 0; i < n_chunks; ++i) {
    const size_t od = chunk_lens[i] ? emu.process(chunk_lens[i], olen_each) : emu.flush(olen_each);
    if (odone_each) odone_each[i] = od;
    if (delay_each) delay_each[i] = emu.delay();
  }
  return nullptr;
}

// Control flow of the variable-rate engine without a device (host-side tests pin it against the live reference):
// step i optionally changes the ratio (io_ratio[i] > 0), appends ilen[i] frames, optionally starts the flush, and
// asks for olen[i] frames; odone[i] is what the reference's soxr_process would return.
b200soxr_error_t b200soxr_vr_emulate(double max_io_ratio, size_t n_steps, double const *io_ratio,
                                     size_t const *slew_len, size_t const *ilen, unsigned char const *flush,
                                     size_t const *olen, size_t *odone) {
  if (!(max_io_ratio > 0)) return "I/O ratio out-of-range";
  VrEngine vr(max_io_ratio, 1.0, 1);
  vr.set_dry_run();
  if (const char *e = vr.init()) return e;
  for (size_t i = 0; i < n_steps; ++i) {
    if (io_ratio && io_ratio[i] > 0) vr.set_io_ratio(io_ratio[i], slew_len ? slew_len[i] : 0);
    if (ilen && ilen[i]) if (const char *e = vr.dry_input(ilen[i])) return e;
    if (flush && flush[i]) vr.flush();
    size_t od = 0;
    vr.plan(olen[i], &od);
    if (const char *e = vr.run(nullptr, 0, nullptr)) return e;
    if (odone) odone[i] = od;
  }
  return nullptr;
}

static b200soxr_t create_handle(double input_rate, double output_rate, unsigned n_streams, unsigned num_channels,
                                b200soxr_error_t *error0, b200soxr_io_spec_t const *io_spec,
                                b200soxr_quality_spec_t const *q_spec, b200soxr_runtime_spec_t const *rt_spec) {
  const char *error = nullptr;
  b200soxr *p = nullptr;
  const double io_ratio = output_rate != 0 ? (input_rate != 0 ? input_rate / output_rate : -1)
                                           : (input_rate != 0 ? -1 : 0);
  if (q_spec && q_spec->e) error = static_cast<const char *>(q_spec->e);
  else if (io_spec && (io_spec->e || (io_spec->itype | io_spec->otype) >= B200SOXR_SPLIT * 2))
    error = "invalid io datatype(s)";
  else if (!num_channels || !n_streams) error = "invalid # of channels";
  else if (!(io_ratio > 0)) error = "I/O ratio out-of-range";
  else if (!device_ok()) error = kNoDevice;
  if (!error && !(p = new (std::nothrow) b200soxr())) error = "malloc failed";
  if (p) {
    if (cudaGetDevice(&p->device) != cudaSuccess) { cudaGetLastError(); p->device = 0; }
    p->channels = num_channels;
    p->n_streams = n_streams;
    p->io_ratio = io_ratio;
    if (io_spec) p->io = *io_spec;
    else { std::memset(&p->io, 0, sizeof p->io); p->io.scale = 1; }
    p->q = q_spec ? to_qspec(*q_spec) : quality_spec(B200SOXR_HQ, 0);
    p->r = to_rspec(rt_spec);
    p->device_io = rt_spec && (rt_spec->flags & B200SOXR_RT_DEVICE_IO);
    p->low_latency = rt_spec && (rt_spec->flags & B200SOXR_RT_LOW_LATENCY);
    if (((p->io.itype ^ p->io.otype) & B200SOXR_SPLIT))
      error = "b200soxr: input and output must both be interleaved or both split";
    else if (n_streams > 1 && (p->io.itype & B200SOXR_SPLIT))
      error = "b200soxr: batched stream handles take interleaved layouts only";
    const double scale = p->io.scale * kFullScale[p->io.otype & 3] / kFullScale[p->io.itype & 3];   // soxr.c:435-436
    // deterministic dither seed, soxr.c:439
    unsigned long long bits;
    std::memcpy(&bits, &io_ratio, 8);
    p->seed = num_channels ^ bits;
    const bool want_vr = (p->q.flags & B200SOXR_VR) != 0;     // soxr.c:443-455: SOXR_VR selects the vr32 engine
    if (!error && want_vr && (p->device_io || p->low_latency))
      error = "b200soxr: variable-rate handles take host buffers and the reference's scheduling only";
    if (!error && !want_vr)
      error = get_device_plan(io_ratio, p->q, p->r, scale, arith_double_for(p->q, p->io.itype, p->io.otype), &p->dp);
    if (!error && want_vr) {
      p->vr_scale = scale;
      p->vr.reset(new (std::nothrow) VrEngine(io_ratio, scale, n_streams * num_channels));
      error = p->vr ? p->vr->init() : "malloc failed";
    }
    if (!error) {
      p->ws = acquire_workspace();
      p->stream = (rt_spec && rt_spec->e) ? static_cast<cudaStream_t>(rt_spec->e) : p->ws->stream;
      error = p->clips_dev.reserve(8);
      if (!error) error = p->seed_dev.reserve(8 * size_t(n_streams));
      if (!error && cudaMemset(p->clips_dev.p, 0, 8) != cudaSuccess) error = "cudaMemset failed";
      if (!error) error = p->reset_seeds(p->seed);
      if (!error && !want_vr) error = p->init_state();
    }
    if (error) { delete p; p = nullptr; }
  }
  if (error0) *error0 = error;
  return p;
}

b200soxr_t b200soxr_create(double input_rate, double output_rate, unsigned num_channels,
                           b200soxr_error_t *error0, b200soxr_io_spec_t const *io_spec,
                           b200soxr_quality_spec_t const *q_spec, b200soxr_runtime_spec_t const *rt_spec) {
  return create_handle(input_rate, output_rate, 1, num_channels, error0, io_spec, q_spec, rt_spec);
}

b200soxr_t b200soxr_create_batch(double input_rate, double output_rate, unsigned num_streams, unsigned num_channels,
                                 b200soxr_error_t *error0, b200soxr_io_spec_t const *io_spec,
                                 b200soxr_quality_spec_t const *q_spec, b200soxr_runtime_spec_t const *rt_spec) {
  return create_handle(input_rate, output_rate, num_streams, num_channels, error0, io_spec, q_spec, rt_spec);
}

double b200soxr_delay(b200soxr_t p) {
  if (p && !p->error && p->vr) return 100;   // vr_delay is a constant in the reference (vr32.c:612-616, "TODO")
  if (!p || p->error || !p->emu) return 0;
  if (!p->low_latency) return p->emu->delay();
  // cr.c:578-581 with this mode's own output count: samples_in / io_ratio - samples_out before the flush,
  // what is left of round(samples_in / io_ratio) after it (cr.c:541-547)
  if (p->n_in_total >= 0) return double(long(output_frames(p->io_ratio, size_t(p->n_in_total))) - p->out_done);
  return double(p->x_hi) / p->io_ratio - double(p->out_done);
}
size_t *b200soxr_num_clips(b200soxr_t p) { return &p->clips; }
char const *b200soxr_engine(b200soxr_t p) { return p->vr ? "b200-vr32" : p->dp->is_double ? "b200-cr64" : "b200-cr32"; }

// soxr_set_io_ratio (soxr.c:579-598) on a live handle: the variable-rate engine takes the new ratio, reached over
// `slew_len` output samples (0: at once); a constant-rate handle accepts only its own ratio.
b200soxr_error_t b200soxr_set_io_ratio(b200soxr_t p, double io_ratio, size_t slew_len) {
  if (!p) return "invalid soxr_t pointer";
  if (p->error) return p->error;
  if (!(io_ratio > 0)) return "I/O ratio out-of-range";
  if (p->vr) { p->vr->set_io_ratio(io_ratio, slew_len); return nullptr; }
  return std::fabs(p->io_ratio - io_ratio) < 1e-15 ? nullptr : "varying O/I ratio is not supported with this quality level";
}

size_t b200soxr_describe_path(double input_rate, double output_rate, unsigned long quality, int dtype,
                              char *buf, size_t buf_size) {
  std::string text;
  if (!(input_rate > 0) || !(output_rate > 0)) text = "error: I/O ratio out-of-range";
  else if (dtype < 0 || dtype > 3) text = "error: invalid io datatype(s)";
  else if (!device_ok()) text = std::string("error: ") + kNoDevice;
  else {
    const QSpec q = quality_spec(quality, 0);
    std::shared_ptr<DevicePlan> dp;
    const char *e = get_device_plan(input_rate / output_rate, q, to_rspec(nullptr), 1.0, arith_double_for(q, dtype, dtype), &dp);
    text = e ? std::string("error: ") + e : dp->path;
  }
  if (buf && buf_size) {
    const size_t n = std::min(buf_size - 1, text.size());
    std::memcpy(buf, text.data(), n);
    buf[n] = 0;
  }
  return text.size() + 1;
}

b200soxr_error_t b200soxr_clear(b200soxr_t p) {
  if (!p) return "invalid soxr_t pointer";
  DeviceGuard guard(p->device);
  // The reference rebuilds the resampler and zeroes everything else, incl. the
  // clip counter and the dither seed (soxr.c:609-628).
  p->clips = 0;
  p->seed = 0;
  p->error = nullptr;
  // the device-side accumulators k_rint_clip / k_dither_seeds add to restart as well
  if (p->clips_dev.p) B2_CUDA(cudaMemsetAsync(p->clips_dev.p, 0, 8, p->stream));
  if (const char *e = p->reset_seeds(0)) return e;
  if (p->vr) {   // RESET_ON_CLEAR: a fresh engine at the create-time (maximum) ratio (soxr.c:609-628)
    p->flushing = false;
    p->vr.reset(new (std::nothrow) VrEngine(p->io_ratio, p->vr_scale, p->n_sig()));
    return p->vr ? p->vr->init() : "malloc failed";
  }
  return p->init_state();
}

void b200soxr_delete(b200soxr_t p) { delete p; }   // ~b200soxr switches to the handle's device itself

// host chunks / results up to this size bypass the copy engines (see b200soxr_process)
static const size_t kZeroCopyBytes = [] {
  const char *v = std::getenv("B200SOXR_ZERO_COPY_KB");
  return size_t(v ? std::max(0L, std::atol(v)) : 256) << 10;
}();

// history windows up to this size are read by the kernels straight from mapped pinned memory; larger ones are
// moved next to them by one DMA first
static const size_t kDirectReadBytes = [] {
  const char *v = std::getenv("B200SOXR_DIRECT_READ_KB");
  return size_t(v ? std::max(0L, std::atol(v)) : 192) << 10;
}();

// One soxr_process call of a variable-rate handle (soxr.c:640-733 over vr32.c): append the input behind stage 0,
// let the engine produce what it can of `olen` frames, convert out (K7, with the reference's dither for int16).
static b200soxr_error_t process_vr(b200soxr_t p, void const *in, size_t ilen, size_t *idone0, void *out, size_t olen,
                                   size_t *odone0) {
  const unsigned ch = p->channels, B = p->n_streams, ns = p->n_sig();
  const bool split = (p->io.itype & B200SOXR_SPLIT) != 0;
  const int it = p->io.itype & 3, ot = p->io.otype & 3;
  const size_t isz = dtype_size(it), osz = dtype_size(ot);
  cudaStream_t st = p->stream;
  Workspace *ws = p->ws;
  const char *e = nullptr;
  if (ilen) {
    IoLayout lay;
    lay.dtype = it; lay.n_clips = B; lay.channels = ch; lay.clip_stride = (long long)(ilen * ch);
    if ((e = ws->in_stage.reserve(ilen * ch * isz * B))) return p->error = e;
    if (split) {
      const void *const *ptrs = static_cast<const void *const *>(in);
      for (unsigned c = 0; c < ch; ++c)
        B2_CUDA(cudaMemcpyAsync(static_cast<char *>(ws->in_stage.p) + size_t(c) * ilen * isz, ptrs[c], ilen * isz,
                                cudaMemcpyHostToDevice, st));
      lay.frame_stride = 1; lay.chan_stride = (long long)ilen;
    } else {
      B2_CUDA(cudaMemcpyAsync(ws->in_stage.p, in, ilen * ch * isz * B, cudaMemcpyHostToDevice, st));
      lay.frame_stride = ch; lay.chan_stride = 1;
    }
    if ((e = p->vr->input(lay, ws->in_stage.p, ilen, st))) return p->error = e;
  }
  if (p->flushing) p->vr->flush();
  size_t odone = 0;
  p->vr->plan(olen, &odone);       // control flow first: the scratch is sized by what the call yields, not by the room offered
  const size_t ys = (std::max<size_t>(odone, 1) + 63) & ~size_t(63);
  if ((e = ws->y.reserve(ys * ns * sizeof(float)))) return p->error = e;
  if ((e = p->vr->run(ws->y.as<float>(), ys, st))) return p->error = e;
  const bool dith = ot == B200SOXR_INT16 && !(p->io.flags & B200SOXR_NO_DITHER);
  if (odone || dith) {
    IoLayout lay;
    lay.dtype = ot; lay.n_clips = B; lay.channels = ch; lay.clip_stride = (long long)(odone * ch);
    if (split) { lay.frame_stride = 1; lay.chan_stride = (long long)odone; }
    else { lay.frame_stride = ch; lay.chan_stride = 1; }
    if ((e = ws->out_stage.reserve(std::max<size_t>(1, odone * ch * B) * osz))) return p->error = e;
    DitherCfg dc;
    dc.enabled = dith;
    dc.split = split;
    dc.seed_in_dev = p->seed_dev.as<unsigned long long>();
    dc.call_lens.push_back((long long)odone);
    dc.seed_end_dev = p->seed_dev.as<unsigned long long>();
    if ((e = convert_out(lay, ws->out_stage.p, 0, odone, false, ws->y.p, ys, dc, p->clips_dev.as<unsigned long long>(), ws, st)))
      return p->error = e;
    if (odone) {
      if (split) {
        void *const *ptrs = static_cast<void *const *>(out);
        for (unsigned c = 0; c < ch; ++c)
          B2_CUDA(cudaMemcpyAsync(ptrs[c], static_cast<char *>(ws->out_stage.p) + size_t(c) * odone * osz, odone * osz,
                                  cudaMemcpyDeviceToHost, st));
      } else {
        B2_CUDA(cudaMemcpyAsync(out, ws->out_stage.p, odone * ch * osz * B, cudaMemcpyDeviceToHost, st));
      }
    }
    if (ot >= B200SOXR_INT32) {
      unsigned long long host = 0;
      B2_CUDA(cudaMemcpyAsync(&host, p->clips_dev.p, 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
      p->clips = size_t(host);
    }
  }
  B2_CUDA(cudaStreamSynchronize(st));
  if (idone0) *idone0 = ilen;
  if (odone0) *odone0 = odone;
  return nullptr;
}

b200soxr_error_t b200soxr_process(b200soxr_t p, void const *in, size_t ilen0, size_t *idone0, void *out,
                                  size_t olen, size_t *odone0) {
  if (!p) return "null pointer";
  if (p->error) return p->error;
  DeviceGuard guard(p->device);
  size_t ilen;
  bool flush_requested = false;
  if (!in) { flush_requested = true; ilen = ilen0 = 0; }
  else {
    if ((ptrdiff_t)ilen0 < 0) { flush_requested = true; ilen0 = ~ilen0; }
    ilen = idone0 ? std::min(size_t(std::ceil(double(olen) * p->io_ratio)), ilen0) : ilen0;   // soxr.c:736-747,775-780
  }
  p->flushing |= (ilen == ilen0 && flush_requested);
  if (!out && olen) return p->error = "null output buffer pointer";
  if (p->vr) return process_vr(p, in, ilen, idone0, out, olen, odone0);

  const unsigned ch = p->channels, B = p->n_streams, ns = p->n_sig();
  const bool split = (p->io.itype & B200SOXR_SPLIT) != 0;
  const int it = p->io.itype & 3, ot = p->io.otype & 3;
  cudaStream_t st = p->stream;
  Workspace *ws = p->ws;
  const char *e = nullptr;
  bool queued = false;   // device work of this call is on the stream (the caller's buffers may still be in use)
  const double t0 = kTraceTiming ? now_us() : 0;

  // ---- input (soxr.c:640-660): append to the history
  if (ilen && !p->emu->flushing()) {
    const size_t isz = dtype_size(it);
    if (p->host_hist && ilen * ch * isz * B > kZeroCopyBytes)
      if ((e = p->migrate_to_device(ilen))) return p->error = e;
    if (p->host_hist) {
      // small chunk (streaming): a memcpy into the mapped pinned history, nothing else
      if ((e = p->grow_host_history(ilen))) return p->error = e;
      const size_t fsz = size_t(ch) * isz;
      for (unsigned b = 0; b < B; ++b)
        std::memcpy(static_cast<char *>(p->hx.p) + (size_t(b) * p->hx_cap + size_t(p->x_hi - p->x_base)) * fsz,
                    static_cast<const char *>(in) + size_t(b) * ilen * fsz, ilen * fsz);
    } else {
      if ((e = p->grow_history(ilen))) return p->error = e;
      const void *dev_in = in;
      IoLayout lay;
      lay.dtype = it;
      lay.n_clips = B;
      lay.channels = ch;
      lay.clip_stride = (long long)(ilen * ch);
      if (p->device_io) {
        if (split) return p->error = "b200soxr: device I/O supports interleaved layouts only";
        lay.frame_stride = ch; lay.chan_stride = 1;
      } else {
        if ((e = ws->in_stage.reserve(ilen * ch * isz * B))) return p->error = e;
        if (split) {
          const void *const *ptrs = static_cast<const void *const *>(in);
          for (unsigned c = 0; c < ch; ++c)
            B2_CUDA(cudaMemcpyAsync(static_cast<char *>(ws->in_stage.p) + size_t(c) * ilen * isz, ptrs[c], ilen * isz,
                                    cudaMemcpyHostToDevice, st));
          lay.frame_stride = 1; lay.chan_stride = (long long)ilen;
        } else {
          // (Staging a large pageable chunk through a pinned slot with the host copy pool was measured and rejected for
          // this path: 256 streams x 32 kB per call take 1.75 ms that way against 1.05 ms with the driver's own staging,
          // profiles/r2v_stream_staging_rejected.log -- the pool pays off from tens of MB, as in resample_batch.)
          B2_CUDA(cudaMemcpyAsync(ws->in_stage.p, in, ilen * ch * isz * B, cudaMemcpyHostToDevice, st));
          lay.frame_stride = ch; lay.chan_stride = 1;
        }
        dev_in = ws->in_stage.p;
      }
      if ((e = convert_in(lay, dev_in, ilen, p->dp->is_double, p->xbuf.p, p->x_cap, size_t(p->x_hi - p->x_base), st)))
        return p->error = e;
      queued = true;
    }
    p->x_hi += long(ilen);
    p->emu->add_input(ilen);
  }
  if (p->flushing && p->n_in_total < 0) p->n_in_total = p->x_hi;
  const double t1 = kTraceTiming ? now_us() : 0;

  // ---- output (soxr.c:664-733): how much the reference would return now
  size_t odone;
  if (p->low_latency) {
    if (p->flushing) p->emu->pull(0, true);   // keep the emulator's flushing state in step (input is refused after it)
    odone = size_t(std::max(0L, std::min(long(olen), p->outputs_ready() - p->out_done)));
  } else {
    odone = p->emu->pull(olen, p->flushing);
  }
  bool out_zero_copy = false;
  const bool int_out = ot >= B200SOXR_INT32;
  const bool dith = ot == B200SOXR_INT16 && !(p->io.flags & B200SOXR_NO_DITHER);
  if (odone || dith) {
    const size_t es = p->esz(), osz = dtype_size(ot);
    IoLayout lay;
    lay.dtype = ot;
    lay.n_clips = B;
    lay.channels = ch;
    lay.clip_stride = (long long)(odone * ch);     // batched handles: [stream][odone][channel], packed
    void *dev_out = out;
    if (p->device_io) { lay.frame_stride = ch; lay.chan_stride = 1; }
    else if (!split && odone * ch * osz * B <= kZeroCopyBytes) {
      // small result: the last kernel writes straight into mapped pinned host memory (K7 fused, no D2H copy)
      if ((e = ws->pin_out.reserve(std::max<size_t>(1, odone * ch * B) * osz))) return p->error = e;
      dev_out = ws->pin_out.p;
      out_zero_copy = true;
      lay.frame_stride = ch; lay.chan_stride = 1;
    } else {
      if ((e = ws->out_stage.reserve(std::max<size_t>(1, odone * ch * B) * osz))) return p->error = e;
      dev_out = ws->out_stage.p;
      if (split) { lay.frame_stride = 1; lay.chan_stride = (long long)odone; }
      else { lay.frame_stride = ch; lay.chan_stride = 1; }
    }
    SigDesc xin;
    if (p->host_hist) {
      // One DMA moves the window of the pinned history these outputs read ([stream][frames][channel] rows) next to
      // the kernels; they take it from there in the caller's dtype and layout (K1 fused).  (Reading the mapped
      // host rows from the kernel instead costs a PCIe round trip per load: 30 us for one 2048-point block.)
      const long lo = std::max(p->x_base, std::min(p->x_hi, history_needed_from(*p->dp, p->out_done, p->low_latency)));
      const size_t frames = size_t(p->x_hi - lo), fsz = size_t(ch) * dtype_size(it);
      if (frames * fsz * B <= kDirectReadBytes) {
        xin = p->host_desc();    // a few blocks: the kernel reads the mapped pinned rows itself (no copy engine in the path)
      } else {
        if ((e = ws->in_stage.reserve(std::max<size_t>(1, frames * fsz * B)))) return p->error = e;
        B2_CUDA(cudaMemcpy2DAsync(ws->in_stage.p, frames * fsz, static_cast<const char *>(p->hx.p) + size_t(lo - p->x_base) * fsz,
                                  p->hx_cap * fsz, frames * fsz, B, cudaMemcpyHostToDevice, st));
        xin.p = ws->in_stage.p; xin.dtype = it; xin.channels = ch;
        xin.clip_stride = (long long)(frames * ch); xin.frame_stride = ch; xin.chan_stride = 1;
        xin.base = lo; xin.vlo = std::max(0L, lo); xin.vhi = p->x_hi;
      }
    } else {
      xin = planar_desc(p->xbuf.p, p->dp->is_double, p->x_cap, p->x_base, std::max(0L, p->x_base), p->x_hi);
    }
    const unsigned run_flags = p->low_latency ? unsigned(kRunGeneric) : 0u;
    if (!int_out) {
      // float output: the last launch writes the caller's layout directly (K7 fused)
      if (odone) {
        SigDesc od;
        od.p = dev_out; od.dtype = ot; od.channels = ch; od.clip_stride = lay.clip_stride;
        od.frame_stride = lay.frame_stride; od.chan_stride = lay.chan_stride; od.base = p->out_done;
        if ((e = run_range(*p->dp, xin, p->n_in_total, ns, ch, p->out_done, p->out_done + long(odone), od, ws, st, run_flags)))
          return p->error = e;
      }
    } else {
      const size_t ys = (std::max<size_t>(odone, 1) + 63) & ~size_t(63);
      if ((e = ws->y.reserve(ys * ns * es))) return p->error = e;
      if (odone) {
        const SigDesc yd = planar_desc(ws->y.p, p->dp->is_double, ys, p->out_done, 0, 0);
        if ((e = run_range(*p->dp, xin, p->n_in_total, ns, ch, p->out_done, p->out_done + long(odone), yd, ws, st, run_flags)))
          return p->error = e;
      }
      DitherCfg dc;
      dc.enabled = dith;
      dc.split = split;
      dc.seed0 = 0;
      dc.seed_in_dev = p->seed_dev.as<unsigned long long>();    // running per-stream seeds stay on the device
      dc.call_lens.push_back((long long)odone);
      dc.seed_end_dev = p->seed_dev.as<unsigned long long>();
      if ((e = convert_out(lay, dev_out, 0, odone, p->dp->is_double, ws->y.p, ys, dc, p->clips_dev.as<unsigned long long>(),
                           ws, st)))
        return p->error = e;
    }
    queued = true;
    if (!p->device_io && odone && !out_zero_copy) {
      if (split) {
        void *const *ptrs = static_cast<void *const *>(out);
        for (unsigned c = 0; c < ch; ++c)
          B2_CUDA(cudaMemcpyAsync(ptrs[c], static_cast<char *>(dev_out) + size_t(c) * odone * osz, odone * osz,
                                  cudaMemcpyDeviceToHost, st));
      } else
        B2_CUDA(cudaMemcpyAsync(out, dev_out, odone * ch * osz * B, cudaMemcpyDeviceToHost, st));
    }
    if (int_out) {
      unsigned long long host = 0;
      B2_CUDA(cudaMemcpyAsync(&host, p->clips_dev.p, 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
      p->clips = size_t(host);
      queued = false;
    }
  }
  // host buffers are the caller's again on return; device I/O stays asynchronous on the handle's stream
  const double t2 = kTraceTiming ? now_us() : 0;
  if (queued && !p->device_io) B2_CUDA(cudaStreamSynchronize(st));
  const double t3 = kTraceTiming ? now_us() : 0;
  if (out_zero_copy && odone) std::memcpy(out, ws->pin_out.p, odone * ch * B * dtype_size(ot));
  if (kTraceTiming) {
    p->t_in += t1 - t0; p->t_queue += t2 - t1; p->t_wait += t3 - t2; p->t_out += now_us() - t3;
    ++p->n_calls;
    p->n_due += odone != 0;
  }
  p->out_done += long(odone);
  if (idone0) *idone0 = ilen;
  if (odone0) *odone0 = odone;
  return nullptr;
}

b200soxr_error_t b200soxr_oneshot(double irate, double orate, unsigned num_channels, void const *in, size_t ilen,
                                  size_t *idone, void *out, size_t olen, size_t *odone,
                                  b200soxr_io_spec_t const *io_spec, b200soxr_quality_spec_t const *q_spec,
                                  b200soxr_runtime_spec_t const *runtime_spec) {
  b200soxr_error_t error = q_spec ? static_cast<const char *>(q_spec->e) : nullptr;
  b200soxr_t h = nullptr;
  if (!error) {
    b200soxr_quality_spec_t q1;
    if (!q_spec) { q1 = b200soxr_quality_spec(B200SOXR_LQ, 0); q_spec = &q1; }   // soxr.c:827-829
    h = b200soxr_create(irate, orate, num_channels, &error, io_spec, q_spec, runtime_spec);
  }
  if (!error) {
    error = b200soxr_process(h, in, ~ilen, idone, out, olen, odone);
    b200soxr_delete(h);
  }
  return error;
}

// ---------------------------------------------------------------------------
// batch entry point
// ---------------------------------------------------------------------------
b200soxr_error_t b200soxr_resample_batch(b200soxr_batch_desc_t const *desc, void const *in, void *out, size_t *odone,
                                         size_t *clips, void *cuda_stream) {
  if (!desc) return "null pointer";
  // One channel has no channel stride: a mono clip handed over in "split-channel" form (python-soxr sends every 1-D
  // array that way, soxr/__init__.py) IS the interleaved layout, and takes its paths (time-split pipeline, vector I/O).
  b200soxr_batch_desc_t norm = *desc;
  if (norm.channels == 1 && norm.in_frame_stride == 1 && norm.out_frame_stride == 1) norm.in_chan_stride = norm.out_chan_stride = 1;
  b200soxr_batch_desc_t const *const d = &norm;
  if (!(d->in_rate > 0) || !(d->out_rate > 0)) return "I/O ratio out-of-range";
  if (d->dtype < 0 || d->dtype > 3) return "invalid io datatype(s)";
  if (!d->channels || !d->n_clips) return "invalid # of channels";
  if (!device_ok()) return kNoDevice;
  int run_device = -1;
  if ((d->flags & B200SOXR_BATCH_DEVICE_PTRS) && in) {
    // run on the device that owns the buffers (one process may drive several GPUs)
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, in) == cudaSuccess && pa.type == cudaMemoryTypeDevice) run_device = pa.device;
    else cudaGetLastError();
  }
  DeviceGuard guard(run_device);   // the caller's current device is restored on return
  const double io_ratio = d->in_rate / d->out_rate;
  const QSpec q = quality_spec(d->quality, (d->flags & B200SOXR_BATCH_DOUBLE) ? B200SOXR_DOUBLE_PRECISION : 0);
  std::shared_ptr<DevicePlan> dp;
  if (const char *e = get_device_plan(io_ratio, q, to_rspec(nullptr), 1.0, arith_double_for(q, d->dtype, d->dtype), &dp)) return e;
  const size_t total = output_frames(io_ratio, d->ilen);
  if (odone) *odone = total;
  if (clips) *clips = 0;
  if (total > d->olen) return "b200soxr: output buffer too small";
  const bool dev_ptrs = d->flags & B200SOXR_BATCH_DEVICE_PTRS;
  const size_t esz = dp->is_double ? 8 : 4, isz = dtype_size(d->dtype);

  Workspace *ws = acquire_workspace(false);
  struct Rel { Workspace *w; ~Rel() { release_workspace(w); } } rel{ws};
  // device pointers: the caller's stream as given (NULL = the legacy default stream, CUDA's own
  // convention) so the call orders with the caller's work; host pointers: NULL = a private stream
  const bool dev_ptrs0 = d->flags & B200SOXR_BATCH_DEVICE_PTRS;
  cudaStream_t st = (cuda_stream || dev_ptrs0) ? static_cast<cudaStream_t>(cuda_stream) : ws->stream;
  const char *e = nullptr;
  if (ws->busy) {   // scratch still in use by an earlier asynchronous call: order this stream after it
    B2_CUDA(cudaStreamWaitEvent(st, ws->busy_ev, 0));
    ws->busy = false;
  }

  IoLayout li{d->dtype, d->n_clips, d->channels, d->in_clip_stride, d->in_frame_stride, d->in_chan_stride};
  IoLayout lo{d->dtype, d->n_clips, d->channels, d->out_clip_stride, d->out_frame_stride, d->out_chan_stride};
  auto span = [](const IoLayout &l, size_t frames) -> size_t {
    if (!frames) return 0;
    return size_t((l.n_clips - 1) * l.clip_stride + (long long)(frames - 1) * l.frame_stride +
                  (l.channels - 1) * l.chan_stride + 1);
  };
  const size_t in_span = span(li, d->ilen), out_span = span(lo, total);

  // dither bookkeeping (identical for every clip: the seed walk restarts per clip)
  DitherCfg dc;
  const bool int_out = d->dtype > B200SOXR_FLOAT64;
  if (int_out && total) {
    dc.enabled = d->dtype == B200SOXR_INT16 && !(d->flags & B200SOXR_BATCH_NO_DITHER);
    dc.split = d->flags & B200SOXR_BATCH_SPLIT_SEMANTICS;
    unsigned long long bits;
    std::memcpy(&bits, &io_ratio, 8);
    dc.seed0 = d->channels ^ bits;                                 // soxr.c:439
    if (dc.enabled) {
      CallEmu emu(dp->plan);
      const size_t cap = size_t(double(d->ilen) * d->out_rate / d->in_rate + 1);   // soxr_ext.cpp:238
      size_t pos = 0;
      if (d->flags & B200SOXR_BATCH_EMULATE_CALLS) {
        const size_t dl = div_len_of(d->in_rate, d->out_rate);
        for (size_t idx = 0; idx < d->ilen; idx += dl) {
          const size_t od = emu.process(std::min(dl, d->ilen - idx), cap - pos);
          dc.call_lens.push_back((long long)od);
          pos += od;
        }
        const size_t od = emu.flush(cap - pos);
        dc.call_lens.push_back((long long)od);
        pos += od;
      } else if (d->flags & B200SOXR_BATCH_ONESHOT_CALLS) {
        emu.add_input(d->ilen);
        const size_t od = emu.pull(cap, true);
        dc.call_lens.push_back((long long)od);
        pos += od;
      } else {
        dc.call_lens.push_back((long long)total);
        pos = total;
      }
      if (pos != total) return "internal: emulated call partition does not cover the output";
    }
    if ((e = ws->misc.reserve(8))) return e;
    B2_CUDA(cudaMemsetAsync(ws->misc.p, 0, 8, st));
  }

  const unsigned run_flags = (d->flags & B200SOXR_BATCH_EXACT_GRID) ? 0u : unsigned(kRunAllowFdr);
  // Compute clips [c0, c0 + nc) from device buffers laid out like the caller's.
  // The first / last launches read and write the caller's layout directly (K1 / float K7 fused).
  auto compute = [&](unsigned c0, unsigned nc, const void *din, void *dout) -> const char * {
    SigDesc xin;
    xin.p = static_cast<char *>(const_cast<void *>(din)) + (long long)c0 * d->in_clip_stride * (long long)isz;
    xin.dtype = d->dtype; xin.channels = d->channels;
    xin.clip_stride = d->in_clip_stride; xin.frame_stride = d->in_frame_stride; xin.chan_stride = d->in_chan_stride;
    xin.base = 0; xin.vlo = 0; xin.vhi = long(d->ilen);
    void *po = static_cast<char *>(dout) + (long long)c0 * d->out_clip_stride * (long long)isz;
    const unsigned ns = nc * d->channels;
    if (!int_out) {
      SigDesc od;
      od.p = po; od.dtype = d->dtype; od.channels = d->channels;
      od.clip_stride = d->out_clip_stride; od.frame_stride = d->out_frame_stride; od.chan_stride = d->out_chan_stride;
      od.base = 0;
      return run_range(*dp, xin, long(d->ilen), ns, d->channels, 0, long(total), od, ws, st, run_flags);
    }
    const size_t ys = (total + 63) & ~size_t(63);
    if (const char *e2 = ws->y.reserve(ys * ns * esz)) return e2;
    const SigDesc yd = planar_desc(ws->y.p, dp->is_double, ys, 0, 0, 0);
    if (const char *e2 = run_range(*dp, xin, long(d->ilen), ns, d->channels, 0, long(total), yd, ws, st, run_flags)) return e2;
    IoLayout lc = lo;
    lc.n_clips = nc;
    return convert_out(lc, po, 0, total, dp->is_double, ws->y.p, ys, dc, ws->misc.as<unsigned long long>(), ws, st);
  };

  if (dev_ptrs) {
    if (total && (e = compute(0, d->n_clips, in, out))) return e;
  } else {
    // Host buffers: clips are cut into chunks that flow  H2D -> compute -> D2H  on three streams, so the
    // two copy directions (one DMA engine each) and the kernels overlap; a chunk is whole clips because
    // clips are the independent unit (soxr_ext.cpp:230: one soxr_t per clip).
    if ((e = ws->in_stage.reserve(std::max<size_t>(1, in_span) * isz))) return e;
    if ((e = ws->out_stage.reserve(std::max<size_t>(1, out_span) * isz))) return e;
    const bool clip_major =
        d->n_clips > 1 && size_t(d->in_clip_stride) >= span(IoLayout{d->dtype, 1, d->channels, 0, d->in_frame_stride, d->in_chan_stride}, d->ilen) &&
        size_t(d->out_clip_stride) >= span(IoLayout{d->dtype, 1, d->channels, 0, d->out_frame_stride, d->out_chan_stride}, total) &&
        d->in_clip_stride > 0 && d->out_clip_stride > 0;
    unsigned per_chunk = d->n_clips;
    if (clip_major && total) {
      static const size_t kChunkBytes = [] {
        const char *v = std::getenv("B200SOXR_CHUNK_MB");
        const long mb = v ? std::atol(v) : 16;
        return size_t(mb > 0 ? mb : 16) << 20;
      }();
      const size_t clip_bytes = std::max<size_t>(1, size_t(d->in_clip_stride) * isz);
      per_chunk = unsigned(std::max<size_t>(1, kChunkBytes / clip_bytes));
      per_chunk = std::max(per_chunk, (d->n_clips + 63) / 64);   // at most 64 chunks
      per_chunk = std::min(per_chunk, d->n_clips);
    }
    const unsigned n_chunks = (d->n_clips + per_chunk - 1) / per_chunk;
    // Pageable caller memory (an ordinary ndarray: the drop-in case) cannot be DMA'd directly; the driver's own
    // staging runs at a fraction of the link.  It goes through pinned staging slots filled / drained by the host
    // copy pool while the copy engines and the kernels work on the other slots (or pieces).
    auto pageable = [](const void *ptr) {
      cudaPointerAttributes pa;
      if (cudaPointerGetAttributes(&pa, ptr) != cudaSuccess) { cudaGetLastError(); return true; }
      return pa.type == cudaMemoryTypeUnregistered;
    };
    static const bool no_staging = [] { const char *v = std::getenv("B200SOXR_NO_HOST_STAGING"); return v && v[0] == '1'; }();
    // Who copies between the caller's pageable memory and the pinned slots of a single-chunk call.  Measured on a
    // 16-vCPU host (profiles/r4h_pool_small.log, r4i, r4j; p50 over 100 calls): from about 2 MB in + out the pool
    // wins for multi-channel clips (5 s stereo float32 0.28 -> 0.21 ms, 10 s stereo int16 0.39 -> 0.27, 2.5 s stereo VHQ
    // 0.21 -> 0.14 when it is allowed at 1.8 MB), but a mono clip of the same size is bound by its transforms (twice
    // the blocks per byte), hides the calling thread's copies behind them and only inherits the pool's run-to-run
    // spread (10 s mono: 0.38 ms on the calling thread, 0.37 - 0.54 with the pool; spinning workers or a polling
    // caller did not change that).  B200SOXR_POOL_MIN_KB overrides the size rule.
    static const long kPoolMinKb = [] { const char *v = std::getenv("B200SOXR_POOL_MIN_KB"); return v ? std::atol(v) : -1L; }();
    const size_t io_bytes = (in_span + out_span) * isz;
    const bool pool_copies = kPoolMinKb >= 0 ? io_bytes >= (size_t(kPoolMinKb) << 10)
                                             : io_bytes >= (size_t(8) << 20) || (d->channels >= 2 && io_bytes >= (size_t(1536) << 10));
    auto hcopy = [&](void *dst, const void *src, size_t n) {
      if (pool_copies) parallel_memcpy(dst, src, n); else std::memcpy(dst, src, n);
    };
    // one chunk (a single clip -- soxr.resample -- or a small batch) of at least this many bytes in + out
    static const size_t kStageOneMin = [] {
      const char *v = std::getenv("B200SOXR_STAGE_ONE_KB");
      return size_t(v ? std::max(0L, std::atol(v)) : 512) << 10;
    }();
    // ONE clip on the reference's block grid with float output (soxr.resample / _resample_oneshot on an ndarray):
    // the clip is cut in TIME and run like a stream fed piece by piece -- piece k + 1 is copied in while piece k's
    // outputs (the frames the reference would return after that much input, CallEmu) are computed and copied back.
    // The block grid is keyed on absolute positions, so the result is bit for bit the unsplit one.
    static const size_t kTimeSplitMin = [] {
      const char *v = std::getenv("B200SOXR_TIME_SPLIT_KB");
      return size_t(v ? std::max(0L, std::atol(v)) : 1024) << 10;
    }();
    const size_t in_bytes = in_span * isz, out_bytes = out_span * isz;
    const bool time_split = d->n_clips == 1 && (d->flags & B200SOXR_BATCH_EXACT_GRID) && !int_out && total &&
                            d->in_chan_stride == 1 && d->in_frame_stride == (long long)d->channels &&
                            d->out_chan_stride == 1 && d->out_frame_stride == (long long)d->channels &&
                            in_bytes >= kTimeSplitMin;
    // Pieces of a single-chunk call go through rings of three pinned slots per direction (a slot is reused once its
    // previous tenant has left), so the page-locked staging memory is bounded by the piece size, not by the call.
    constexpr unsigned kRing = 3;
    constexpr size_t kPieceMax = size_t(8) << 20;
    if (time_split) {
      const unsigned ch = d->channels;
      const size_t fsz = size_t(ch) * isz;
      const bool stage_in = !no_staging && pageable(in), stage_out = !no_staging && pageable(out);
      const size_t piece_b = std::min(kPieceMax, std::max<size_t>(size_t(512) << 10, in_bytes / 8));
      const unsigned P = unsigned(std::max<size_t>(2, in_bytes / piece_b));   // (a piece stays below 2 * kPieceMax)
      if ((e = ws->need_events(2 * size_t(P) + 1))) return e;
      cudaEvent_t ev_start = ws->events[2 * P];
      B2_CUDA(cudaMemsetAsync(ws->out_stage.p, 0, out_bytes, st));
      B2_CUDA(cudaEventRecord(ev_start, st));          // order after earlier work queued on st
      B2_CUDA(cudaStreamWaitEvent(ws->h2d_stream, ev_start, 0));
      B2_CUDA(cudaStreamWaitEvent(ws->d2h_stream, ev_start, 0));
      auto sync_all = [&] { cudaStreamSynchronize(ws->h2d_stream); cudaStreamSynchronize(st); cudaStreamSynchronize(ws->d2h_stream); };
      CallEmu emu(dp->plan);
      struct Piece { size_t off, bytes; bool copied; };
      std::vector<Piece> pieces(P, Piece{0, 0, true});
      auto drain = [&](unsigned k, bool wait) -> bool {   // piece k's frames: pinned slot -> caller memory
        if (pieces[k].copied) return true;
        if (wait) cudaEventSynchronize(ws->events[2 * k + 1]);
        else if (cudaEventQuery(ws->events[2 * k + 1]) != cudaSuccess) { cudaGetLastError(); return false; }
        hcopy(static_cast<char *>(out) + pieces[k].off, ws->pin_ring_out[k % kRing].p, pieces[k].bytes);
        pieces[k].copied = true;
        return true;
      };
      size_t in_done = 0, o_done = 0;
      unsigned next_drain = 0;
      for (unsigned k = 0; k < P; ++k) {
        const bool last = k + 1 == P;
        const size_t hi = last ? d->ilen : d->ilen * (k + 1) / P;
        const size_t ioff = in_done * fsz, ilen_b = (hi - in_done) * fsz;
        const char *src = static_cast<const char *>(in) + ioff;
        if (stage_in && ilen_b) {
          PinnedBuf &slot = ws->pin_ring_in[k % kRing];
          if (k >= kRing) B2_CUDA(cudaEventSynchronize(ws->events[2 * (k - kRing)]));   // its previous tenant is on the device
          if ((e = slot.reserve(ilen_b))) { sync_all(); return e; }
          hcopy(slot.p, src, ilen_b);
          src = static_cast<const char *>(slot.p);
        }
        if (ilen_b) B2_CUDA(cudaMemcpyAsync(static_cast<char *>(ws->in_stage.p) + ioff, src, ilen_b, cudaMemcpyHostToDevice, ws->h2d_stream));
        B2_CUDA(cudaEventRecord(ws->events[2 * k], ws->h2d_stream));
        size_t od = emu.process(hi - in_done, total - o_done);
        if (last) {
          od += emu.flush(total - o_done - od);
          if (o_done + od != total) { sync_all(); return "internal: emulated call partition does not cover the output"; }
        }
        in_done = hi;
        if (od) {
          B2_CUDA(cudaStreamWaitEvent(st, ws->events[2 * k], 0));
          SigDesc xin, odsc;
          xin.p = ws->in_stage.p; xin.dtype = d->dtype; xin.channels = ch;
          xin.clip_stride = d->in_clip_stride; xin.frame_stride = ch; xin.chan_stride = 1;
          xin.base = 0; xin.vlo = 0; xin.vhi = long(hi);
          odsc.p = ws->out_stage.p; odsc.dtype = d->dtype; odsc.channels = ch;
          odsc.clip_stride = d->out_clip_stride; odsc.frame_stride = ch; odsc.chan_stride = 1; odsc.base = 0;
          if ((e = run_range(*dp, xin, last ? long(d->ilen) : -1L, ch, ch, long(o_done), long(o_done + od), odsc, ws, st, 0u))) {
            sync_all();
            return e;
          }
          B2_CUDA(cudaEventRecord(ws->events[2 * k + 1], st));
          B2_CUDA(cudaStreamWaitEvent(ws->d2h_stream, ws->events[2 * k + 1], 0));
          const size_t ooff = o_done * fsz, ob = od * fsz;
          char *dst = static_cast<char *>(out) + ooff;
          if (stage_out) {
            while (next_drain + kRing <= k) drain(next_drain++, true);   // the slot's previous tenant has been copied out
            PinnedBuf &slot = ws->pin_ring_out[k % kRing];
            if ((e = slot.reserve(ob))) { sync_all(); return e; }
            dst = static_cast<char *>(slot.p);
          }
          B2_CUDA(cudaMemcpyAsync(dst, static_cast<char *>(ws->out_stage.p) + ooff, ob, cudaMemcpyDeviceToHost, ws->d2h_stream));
          B2_CUDA(cudaEventRecord(ws->events[2 * k + 1], ws->d2h_stream));   // re-recorded: now marks the end of the D2H
          if (stage_out) pieces[k] = Piece{ooff, ob, false};
          o_done += od;
        }
        // results that have landed meanwhile leave for the caller's array (in order, without waiting)
        while (stage_out && next_drain < k && drain(next_drain, false)) ++next_drain;
      }
      for (; stage_out && next_drain < P; ++next_drain) drain(next_drain, true);
      sync_all();
    } else if (n_chunks <= 1 && !no_staging && (in_span + out_span) * isz >= kStageOneMin && (pageable(in) || pageable(out))) {
      // the one chunk cut into pieces: piece k + 1 is copied into its pinned slot while piece k is on the link; on
      // the way back piece k is copied out while piece k + 1 is on the link
      const size_t ib = in_span * isz, ob = out_span * isz;
      constexpr size_t kPiece = size_t(768) << 10;
      const bool stage_in = ib && pageable(in), stage_out = ob && pageable(out);
      auto n_pieces = [&](size_t bytes) {
        return unsigned(std::max<size_t>(std::min<size_t>(4, std::max<size_t>(1, bytes / kPiece)), (bytes + kPieceMax - 1) / kPieceMax));
      };
      auto cut = [](size_t bytes, unsigned k, unsigned n) { return k >= n ? bytes : (bytes * k / n) & ~size_t(255); };
      const unsigned n_ip = stage_in ? n_pieces(ib) : 1u, n_op = stage_out ? n_pieces(ob) : 1u;
      if ((e = ws->need_events(size_t(n_ip) + n_op))) return e;
      for (unsigned k = 0; k < n_ip && ib; ++k) {
        const size_t lo = cut(ib, k, n_ip), hi = cut(ib, k + 1, n_ip);
        const char *src = static_cast<const char *>(in) + lo;
        if (stage_in) {
          PinnedBuf &slot = ws->pin_ring_in[k % kRing];
          if (k >= kRing) B2_CUDA(cudaEventSynchronize(ws->events[k - kRing]));
          if ((e = slot.reserve(hi - lo))) { cudaStreamSynchronize(st); return e; }
          hcopy(slot.p, src, hi - lo);
          src = static_cast<const char *>(slot.p);
        }
        B2_CUDA(cudaMemcpyAsync(static_cast<char *>(ws->in_stage.p) + lo, src, hi - lo, cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaEventRecord(ws->events[k], st));
      }
      if (ob) B2_CUDA(cudaMemsetAsync(ws->out_stage.p, 0, ob, st));
      if (total && (e = compute(0, d->n_clips, ws->in_stage.p, ws->out_stage.p))) { cudaStreamSynchronize(st); return e; }
      unsigned drained = 0;
      auto drain_out = [&](unsigned k) {
        const size_t lo = cut(ob, k, n_op), hi = cut(ob, k + 1, n_op);
        cudaEventSynchronize(ws->events[n_ip + k]);
        hcopy(static_cast<char *>(out) + lo, ws->pin_ring_out[k % kRing].p, hi - lo);
      };
      for (unsigned k = 0; k < n_op && ob; ++k) {
        const size_t lo = cut(ob, k, n_op), hi = cut(ob, k + 1, n_op);
        char *dst = static_cast<char *>(out) + lo;
        if (stage_out) {
          while (drained + kRing <= k) drain_out(drained++);
          PinnedBuf &slot = ws->pin_ring_out[k % kRing];
          if ((e = slot.reserve(hi - lo))) { cudaStreamSynchronize(st); return e; }
          dst = static_cast<char *>(slot.p);
        }
        B2_CUDA(cudaMemcpyAsync(dst, static_cast<char *>(ws->out_stage.p) + lo, hi - lo, cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaEventRecord(ws->events[n_ip + k], st));
      }
      for (; stage_out && ob && drained < n_op; ++drained) drain_out(drained);
    } else if (n_chunks <= 1) {
      if (in_span) B2_CUDA(cudaMemcpyAsync(ws->in_stage.p, in, in_span * isz, cudaMemcpyHostToDevice, st));
      if (out_span) B2_CUDA(cudaMemsetAsync(ws->out_stage.p, 0, out_span * isz, st));
      if (total && (e = compute(0, d->n_clips, ws->in_stage.p, ws->out_stage.p))) return e;
      if (out_span) B2_CUDA(cudaMemcpyAsync(out, ws->out_stage.p, out_span * isz, cudaMemcpyDeviceToHost, st));
    } else {
      if ((e = ws->need_events(2 * size_t(n_chunks) + 1))) return e;
      cudaEvent_t ev_start = ws->events[2 * n_chunks];
      B2_CUDA(cudaEventRecord(ev_start, st));          // order after earlier work queued on st
      B2_CUDA(cudaStreamWaitEvent(ws->h2d_stream, ev_start, 0));
      B2_CUDA(cudaStreamWaitEvent(ws->d2h_stream, ev_start, 0));
      // pageable chunks go through three pinned staging slots per direction
      const bool stage_in = !no_staging && pageable(in), stage_out = !no_staging && pageable(out);
      constexpr unsigned kSlots = 3;
      struct Pending { size_t ooff, obytes; unsigned slot; };
      std::vector<Pending> pend(n_chunks);
      auto sync_all = [&] { cudaStreamSynchronize(ws->h2d_stream); cudaStreamSynchronize(st); cudaStreamSynchronize(ws->d2h_stream); };
      auto drain = [&](unsigned c) -> const char * {      // chunk c's result: staging slot -> caller memory
        B2_CUDA(cudaEventSynchronize(ws->events[2 * c + 1]));
        if (stage_out) parallel_memcpy(static_cast<char *>(out) + pend[c].ooff, ws->pin_ring_out[pend[c].slot].p, pend[c].obytes);
        return nullptr;
      };
      for (unsigned c = 0; c < n_chunks; ++c) {
        const unsigned c0 = c * per_chunk, nc = std::min(per_chunk, d->n_clips - c0), slot = c % kSlots;
        IoLayout l1 = li, l2 = lo;
        l1.n_clips = l2.n_clips = nc;
        const size_t ispan = span(l1, d->ilen), ospan = span(l2, total);
        const size_t ioff = size_t(c0) * size_t(d->in_clip_stride) * isz, ooff = size_t(c0) * size_t(d->out_clip_stride) * isz;
        pend[c] = Pending{ooff, ospan * isz, slot};
        // the slot's previous tenant (chunk c - kSlots) must have left: its D2H done and copied out
        if ((stage_in || stage_out) && c >= kSlots)
          if ((e = drain(c - kSlots))) { sync_all(); return e; }
        const void *h2d_src = static_cast<const char *>(in) + ioff;
        if (stage_in) {
          if ((e = ws->pin_ring_in[slot].reserve(ispan * isz))) { sync_all(); return e; }
          parallel_memcpy(ws->pin_ring_in[slot].p, h2d_src, ispan * isz);
          h2d_src = ws->pin_ring_in[slot].p;
        }
        B2_CUDA(cudaMemcpyAsync(static_cast<char *>(ws->in_stage.p) + ioff, h2d_src, ispan * isz, cudaMemcpyHostToDevice,
                                ws->h2d_stream));
        B2_CUDA(cudaEventRecord(ws->events[2 * c], ws->h2d_stream));
        B2_CUDA(cudaStreamWaitEvent(st, ws->events[2 * c], 0));
        B2_CUDA(cudaMemsetAsync(static_cast<char *>(ws->out_stage.p) + ooff, 0, ospan * isz, st));
        if ((e = compute(c0, nc, ws->in_stage.p, ws->out_stage.p))) { sync_all(); return e; }
        B2_CUDA(cudaEventRecord(ws->events[2 * c + 1], st));
        B2_CUDA(cudaStreamWaitEvent(ws->d2h_stream, ws->events[2 * c + 1], 0));
        void *d2h_dst = static_cast<char *>(out) + ooff;
        if (stage_out) {
          if ((e = ws->pin_ring_out[slot].reserve(ospan * isz))) { sync_all(); return e; }
          d2h_dst = ws->pin_ring_out[slot].p;
        }
        B2_CUDA(cudaMemcpyAsync(d2h_dst, static_cast<char *>(ws->out_stage.p) + ooff, ospan * isz, cudaMemcpyDeviceToHost,
                                ws->d2h_stream));
        B2_CUDA(cudaEventRecord(ws->events[2 * c + 1], ws->d2h_stream));   // re-recorded: now marks the end of the D2H
      }
      if (stage_in || stage_out)
        for (unsigned c = n_chunks > kSlots ? n_chunks - kSlots : 0; c < n_chunks; ++c)
          if ((e = drain(c))) { sync_all(); return e; }
      B2_CUDA(cudaStreamSynchronize(ws->d2h_stream));
    }
  }
  if (int_out && total && clips) {
    unsigned long long c = 0;
    B2_CUDA(cudaMemcpyAsync(&c, ws->misc.p, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    *clips = size_t(c);
  }
  if (!dev_ptrs) {
    B2_CUDA(cudaStreamSynchronize(st));
  } else {
    // asynchronous return: the work stays queued on the caller's stream; the pooled scratch it may still be using is
    // fenced by Workspace::busy_ev for whoever takes it next.  Integer batches whose caller asked for the clip count
    // (clips != NULL) have waited for it above.
    if (d->dtype <= B200SOXR_FLOAT64 || !clips) {
      if ((e = ws->mark_busy(st))) return e;
    } else B2_CUDA(cudaStreamSynchronize(st));
  }
  return nullptr;
}

// ---- pinned result buffers (include/b200soxr.h) -----------------------------------------------------------------
namespace {
struct ResultPool {
  std::mutex mu;
  std::map<void *, size_t> live;              // handed out: pointer -> capacity
  std::multimap<size_t, void *> cached;       // free: capacity -> pointer
  size_t live_bytes = 0, cached_bytes = 0;
  const size_t budget = [] {
    const char *v = std::getenv("B200SOXR_PINNED_RESULT_MB");
    return size_t(v ? std::max(0L, std::atol(v)) : 1024) << 20;
  }();
  static size_t bucket(size_t n) {            // four sizes per octave from 256 kB: little waste, few distinct sizes
    size_t b = size_t(256) << 10;
    while (b < n) b <<= 1;
    const size_t q = b >> 3;                  // b/2 < n <= b: steps of b/8 above b/2
    return b > (size_t(256) << 10) ? (n + q - 1) / q * q : b;
  }
};
ResultPool &result_pool() {
  static ResultPool *p = new ResultPool();    // never destroyed: arrays may be released at interpreter exit
  return *p;
}
}  // namespace

void *b200soxr_result_alloc(size_t bytes) {
  ResultPool &rp = result_pool();
  if (!bytes || !rp.budget || !device_ok()) return nullptr;
  const size_t cap = ResultPool::bucket(bytes);
  std::lock_guard<std::mutex> l(rp.mu);
  auto it = rp.cached.lower_bound(cap);
  if (it != rp.cached.end() && it->first <= cap + cap / 4) {
    void *p = it->second;
    const size_t c = it->first;
    rp.cached.erase(it);
    rp.cached_bytes -= c;
    rp.live[p] = c;
    rp.live_bytes += c;
    return p;
  }
  while (rp.live_bytes + rp.cached_bytes + cap > rp.budget && !rp.cached.empty()) {   // make room: drop cached buffers
    auto big = std::prev(rp.cached.end());
    cudaFreeHost(big->second);
    rp.cached_bytes -= big->first;
    rp.cached.erase(big);
  }
  if (rp.live_bytes + rp.cached_bytes + cap > rp.budget) return nullptr;
  void *p = nullptr;
  if (cudaHostAlloc(&p, cap, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  rp.live[p] = cap;
  rp.live_bytes += cap;
  return p;
}

void b200soxr_result_free(void *p) {
  if (!p) return;
  ResultPool &rp = result_pool();
  std::lock_guard<std::mutex> l(rp.mu);
  auto it = rp.live.find(p);
  if (it == rp.live.end()) return;
  const size_t c = it->second;
  rp.live.erase(it);
  rp.live_bytes -= c;
  constexpr size_t kBig = size_t(64) << 20;
  if (c >= kBig) {   // at most one large buffer stays cached (enough for `y = f(x)` in a loop: one live, one free)
    for (auto jt = rp.cached.lower_bound(kBig); jt != rp.cached.end();) {
      cudaFreeHost(jt->second);
      rp.cached_bytes -= jt->first;
      jt = rp.cached.erase(jt);
    }
  }
  if (rp.cached_bytes + c <= rp.budget / 2) {   // keep up to half the budget around for the next results
    rp.cached.emplace(c, p);
    rp.cached_bytes += c;
  } else {
    cudaFreeHost(p);
  }
}

b200soxr_error_t b200soxr_quantize(void const *planar, int is_double, unsigned channels, size_t n,
                                   size_t const *call_lens, size_t n_calls, int out_dtype, int dither, int split,
                                   unsigned long long *seed, void *out, size_t *clips) {
  if (!device_ok()) return kNoDevice;
  if (out_dtype != B200SOXR_INT16 && out_dtype != B200SOXR_INT32) return "invalid io datatype(s)";
  Workspace *ws = acquire_workspace();
  struct Rel { Workspace *w; ~Rel() { release_workspace(w); } } rel{ws};
  cudaStream_t st = ws->stream;
  const size_t es = is_double ? 8 : 4, osz = dtype_size(out_dtype);
  const char *e;
  if ((e = ws->y.reserve(std::max<size_t>(1, n * channels) * es))) return e;
  if ((e = ws->out_stage.reserve(std::max<size_t>(1, n * channels) * osz))) return e;
  if ((e = ws->misc.reserve(16))) return e;
  B2_CUDA(cudaMemsetAsync(ws->misc.p, 0, 16, st));
  if (n) B2_CUDA(cudaMemcpyAsync(ws->y.p, planar, n * channels * es, cudaMemcpyHostToDevice, st));
  IoLayout lay{out_dtype, 1, channels, 0, (long long)channels, 1};
  DitherCfg dc;
  dc.enabled = dither && out_dtype == B200SOXR_INT16;
  dc.split = split;
  dc.seed0 = seed ? *seed : 0;
  for (size_t i = 0; i < n_calls; ++i) dc.call_lens.push_back((long long)call_lens[i]);
  dc.seed_end_dev = ws->misc.as<unsigned long long>() + 1;
  if ((e = convert_out(lay, ws->out_stage.p, 0, n, is_double, ws->y.p, n, dc, ws->misc.as<unsigned long long>(), ws, st)))
    return e;
  unsigned long long host[2] = {0, 0};
  if (n) B2_CUDA(cudaMemcpyAsync(out, ws->out_stage.p, n * channels * osz, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaMemcpyAsync(host, ws->misc.p, 16, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  if (clips) *clips = size_t(host[0]);
  if (seed && dc.enabled) *seed = host[1];
  return nullptr;
}

}  // extern "C"
