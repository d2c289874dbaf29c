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
  DevBuf ptab;    // rational poly table [phase][tap] for the fused kernel
};

struct FftStageInfo;   // fft_stage.cuh

struct DevicePlan {
  Plan plan;
  bool is_double = false;
  std::vector<DevStage> stages;
  std::shared_ptr<FftStageInfo> fft;   // non-null if the fused FFT kernel applies
  ~DevicePlan();
};

// Immutable plan cache (keyed by ratio bits, preset, flags, gain, device).
const char *get_device_plan(double io_ratio, const QSpec &q, const RSpec &r, double multiplier,
                            std::shared_ptr<DevicePlan> *out);

// Per-call scratch: reused across calls through a pool so concurrent callers
// (the reference's thread-pool tests) never share buffers.
struct Workspace {
  cudaStream_t stream = nullptr;
  DevBuf in_stage, out_stage, x, y, t0, t1, seeds, blocks, misc;
  PinnedBuf pin_in, pin_out, pin_misc;
  ~Workspace();
};
Workspace *acquire_workspace();
void release_workspace(Workspace *);

// User-side layout of a batch of clips.
struct IoLayout {
  int dtype = 0;   // B200SOXR_FLOAT32..INT16
  unsigned n_clips = 1, channels = 1;
  long long clip_stride = 0, frame_stride = 1, chan_stride = 1;   // elements
};
inline size_t dtype_size(int dt) { return dt == 0 ? 4 : dt == 1 ? 8 : dt == 2 ? 4 : 2; }

// Planar engine-precision view of the input history.
struct XView {
  void *p;            // T*
  size_t stride;      // elements per signal
  long base;          // logical index of element 0
  long hi;            // exclusive logical end of valid data
  long n_in_total;    // known total length (flush seen) or -1
};

// Compute final outputs [o0, o1) of every signal into planar Y (stride y_stride,
// element 0 == output o0).  Generic path; the fused path is selected inside.
const char *run_range(const DevicePlan &dp, const XView &x, unsigned n_sig, long o0, long o1, void *Y,
                      size_t y_stride, Workspace *ws, cudaStream_t stream);

// Lowest input index any output >= o can still depend on (for history trimming).
long history_needed_from(const DevicePlan &dp, long o);

const char *convert_in(const IoLayout &lay, const void *dev_in, size_t n_frames, bool is_double, void *X,
                       size_t x_stride, size_t x_off, cudaStream_t stream);

struct DitherCfg {
  bool enabled = false;     // int16 and !NO_DITHER
  bool split = false;
  unsigned long long seed0 = 0;
  std::vector<long long> call_lens;   // frames per emulated call (sum == n_frames)
  unsigned long long *seed_end_dev = nullptr;   // optional [n_clips]
};
// Planar engine-precision Y (n_frames per signal) -> user layout at frame offset `frame_off`.
const char *convert_out(const IoLayout &lay, void *dev_out, size_t frame_off, size_t n_frames, bool is_double,
                        const void *Y, size_t y_stride, const DitherCfg &dc, unsigned long long *clips_dev,
                        Workspace *ws, cudaStream_t stream);

}  // namespace b2
