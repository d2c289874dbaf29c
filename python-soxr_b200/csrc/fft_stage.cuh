// fft_stage.cuh -- fused shared-memory overlap-save FFT stage (+ rational polyphase).
// (stub: filled in by the next milestone; the generic kernels cover every plan.)
#pragma once
#include <cuda_runtime.h>

#include "plan.h"
#include "runtime.h"

namespace b2 {

struct FftStageInfo {
  int stage_index = -1;
  bool fused_poly = false;
};

struct RangeLH { long lo, hi; };

inline const char *fft_stage_prepare(DevicePlan *) { return nullptr; }

template <typename R>
inline R fft_stage_input_range(const FftStageInfo &, const Stage &, R r) { return r; }

template <typename T>
inline const char *fft_stage_run(const DevicePlan &, const FftStageInfo &, const T *, size_t, long, long, long,
                                 T *, size_t, long, long, long, unsigned, cudaStream_t) {
  return "internal: fused FFT stage not built";
}

}  // namespace b2
