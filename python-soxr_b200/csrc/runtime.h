// runtime.h -- device-side plan objects, workspace pool and the range executor.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "plan.h"

namespace b2 {

extern std::atomic<unsigned long long> g_launches;
#define B2_COUNT_LAUNCH() (::b2::g_launches.fetch_add(1, std::memory_order_relaxed))

const char *cuda_err(cudaError_t e);   // static string for the error
#define B2_CUDA(expr)                                   \
  do {                                                  \
    cudaError_t e__ = (expr);                           \
    if (e__ != cudaSuccess) return ::b2::cuda_err(e__); \
  } while (0)

// Makes `device` current for the calling thread and restores the previous one on scope exit, so an entry point
// never changes the caller's current device (torch and other libraries rely on it).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int device) {
    if (device >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != device)
      switched = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard &) = delete;
  DeviceGuard &operator=(const DeviceGuard &) = delete;
};

// Growable device buffer (never shrinks; freed with its owner).
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  const char *reserve(size_t bytes);
  void release();
  template <typename T> T *as() const { return static_cast<T *>(p); }
};
struct PinnedBuf {
  void *p = nullptr;
  size_t cap = 0;
  const char *reserve(size_t bytes);
  void release();
};

// Device copy of one planned stage.
struct DevStage {
  DevBuf g;       // taps, engine precision
  DevBuf table;   // interpolated poly table
  // fused FFT stage data (fft_stage.cuh)
  DevBuf spec;    // filter spectra in kernel order
  DevBuf twid;    // twiddles
  DevBuf edge;    // first taps of each polyphase component (halo correction)
  DevBuf ptab;    // rational poly table [phase][tap] for the fused kernel
};

struct FftStageInfo;   // fft_stage.cuh
struct FdrInfo;        // fdr_stage.cuh

struct DevicePlan {
  Plan plan;
  bool is_double = false;   // ARITHMETIC type of the device path (plan.is_double is the reference's engine)
  std::vector<DevStage> stages;
  std::shared_ptr<FftStageInfo> fft;   // non-null if the fused FFT kernel applies
  std::shared_ptr<FdrInfo> fdr;        // non-null if the single-pass frequency-domain kernel applies (batches)
  // which kernels this plan runs through: "b200-cr32 | exact: ... | batch: ..." (b200soxr_engine / b200soxr_describe_path)
  std::string path;
  ~DevicePlan();
};

// Immutable plan cache (keyed by ratio bits, preset, flags, gain, device).
// arith_double: run the device arithmetic in f64 (else f32).  The filter design always follows
// the reference's engine choice (soxr.c:447), only the sample arithmetic type differs.
const char *get_device_plan(double io_ratio, const QSpec &q, const RSpec &r, double multiplier,
                            bool arith_double, std::shared_ptr<DevicePlan> *out);

// Per-call scratch: reused across calls through a pool so concurrent callers
// (the reference's thread-pool tests) never share buffers.
struct Workspace {
  int device = 0;
  cudaStream_t stream = nullptr;
  // host-buffer batches: copy-in / copy-out streams (one DMA engine each way) and a pool of
  // untimed events chaining  H2D(chunk i) -> compute(chunk i) -> D2H(chunk i)
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  std::vector<cudaEvent_t> events;
  const char *need_events(size_t n);
  // set when a call returned with work still queued that uses the scratch buffers below: the next
  // user orders itself after `busy_ev` (on its stream, or on the host)
  cudaEvent_t busy_ev = nullptr;
  bool busy = false;
  const char *mark_busy(cudaStream_t st);
  DevBuf in_stage, out_stage, x, y, t0, t1, seeds, blocks, misc;
  DevBuf fdr_tickets;               // k_fdr's chunk ticket counter (counts up across launches; never reset)
  unsigned fdr_ticket_base = 0;
  PinnedBuf pin_in, pin_out, pin_misc;
  PinnedBuf pin_ring_in[3], pin_ring_out[3];   // staging slots of the pageable-host batch pipeline
  // side stream for small kernels that may overlap the tail of a large one (k_fdr's head kernel), fenced by events
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t aux_ev[2] = {nullptr, nullptr};
  const char *need_aux();
  ~Workspace();
};
// host_wait: block until queued work of the previous user has finished (callers that pass false
// order their own stream after Workspace::busy_ev instead)
Workspace *acquire_workspace(bool host_wait = true);
void release_workspace(Workspace *);

// memcpy split over a small persistent pool of host threads (pageable <-> pinned staging of large batches: one
// core moves ~10 GB/s, the PCIe link ~50 GB/s).  Blocks until the whole range is copied.
void parallel_memcpy(void *dst, const void *src, size_t bytes);

// User-side layout of a batch of clips.
struct IoLayout {
  int dtype = 0;   // B200SOXR_FLOAT32..INT16
  unsigned n_clips = 1, channels = 1;
  long long clip_stride = 0, frame_stride = 1, chan_stride = 1;   // elements
};
inline size_t dtype_size(int dt) { return dt == 0 ? 4 : dt == 1 ? 8 : dt == 2 ? 4 : 2; }

// Signals in user layout or planar scratch:
//   element(s, k) = p[(s / channels) * clip_stride + (s % channels) * chan_stride + (k - base) * frame_stride]
struct SigDesc {
  void *p = nullptr;
  int dtype = 0;                 // B200SOXR_FLOAT32.. (0 f32, 1 f64, 2 i32, 3 i16)
  unsigned channels = 1;
  long long clip_stride = 0, frame_stride = 1, chan_stride = 0;
  long base = 0;                 // logical index of element 0
  long vlo = 0, vhi = 0;         // reads outside [vlo, vhi) are zero (inputs only)
};
// planar scratch of engine-precision samples: signal s at p[s*stride + (k - base)]
inline SigDesc planar_desc(void *p, bool is_double, size_t stride, long base, long vlo, long vhi) {
  SigDesc d;
  d.p = p; d.dtype = is_double ? 1 : 0; d.channels = 1; d.clip_stride = (long long)stride;
  d.frame_stride = 1; d.chan_stride = 0; d.base = base; d.vlo = vlo; d.vhi = vhi;
  return d;
}

// Compute final outputs [o0, o1) of every signal.  `in` holds input samples for
// logical indices [in.vlo, in.vhi) (zero outside; n_in_total < 0 means more input
// may still arrive, so reading beyond in.vhi is an error).  `out` is a float32 /
// float64 destination in any layout (element 0 of a signal == output out.base).
// `group` = channels per clip (n_sig is a multiple of it).
// `flags`: kRunAllowFdr lets whole-signal (batch / one-shot `resample`) calls use the single-pass
// frequency-domain kernel (fdr_stage.cuh) when the plan has one; streams never set it, so a stream's
// result does not depend on how it was chunked.  kRunGeneric forces the direct-form kernels.
enum { kRunAllowFdr = 1, kRunGeneric = 2 };
const char *run_range(const DevicePlan &dp, const SigDesc &in, long n_in_total, unsigned n_sig, unsigned group,
                      long o0, long o1, const SigDesc &out, Workspace *ws, cudaStream_t stream, unsigned flags = 0);

// Lowest input index any output >= o can still depend on (for history trimming).
long history_needed_from(const DevicePlan &dp, long o, bool generic = false);   // generic: the direct-form kernels' ranges

const char *convert_in(const IoLayout &lay, const void *dev_in, size_t n_frames, bool is_double, void *X,
                       size_t x_stride, size_t x_off, cudaStream_t stream);

struct DitherCfg {
  bool enabled = false;     // int16 and !NO_DITHER
  bool split = false;
  unsigned long long seed0 = 0;
  const unsigned long long *seed_in_dev = nullptr;   // optional [n_clips]: per-clip start seeds on the device (streams)
  std::vector<long long> call_lens;   // frames per emulated call (sum == n_frames)
  unsigned long long *seed_end_dev = nullptr;   // optional [n_clips]
};
// Planar engine-precision Y (n_frames per signal) -> user layout at frame offset `frame_off`.
const char *convert_out(const IoLayout &lay, void *dev_out, size_t frame_off, size_t n_frames, bool is_double,
                        const void *Y, size_t y_stride, const DitherCfg &dc, unsigned long long *clips_dev,
                        Workspace *ws, cudaStream_t stream);

}  // namespace b2
