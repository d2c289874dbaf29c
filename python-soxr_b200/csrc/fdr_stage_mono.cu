// fdr_stage_mono.cu -- k_fdr's instantiations for block-paired mono signals (MODE 2: float32 in, MODE 3: int16 in;
// unit stride, float32 out), in their own translation unit so that they compile beside the others.
#include "fdr_launch.cuh"

namespace b2 {

const char *fdr_launch_mono(const FdrInfo &fi, const FdrKParams &P, long long n_items, cudaStream_t st, Workspace *ws) {
  return fdr_launch_modes<2, 3>(fi, P, n_items, st, ws);
}

}  // namespace b2
