// vr.h -- variable-rate engine (`ResampleStream(vr=True)`, `set_io_ratio`): SURVEY.md section 8f-4.
//
// The reference's vr32 engine (libsoxr/src/vr32.c) is a per-channel state machine: a ladder of half-band stages
// (x2 interpolator below the input rate, decimators above it), a poly-phase interpolator whose 32.32 fixed-point
// step can slew, cross-fades when the ratio crosses an octave, and a recursive all-pass half-band behind the
// interpolator.  Its CONTROL flow never looks at sample values, only at lengths, clocks and the requested ratio.
// So it is split here the same way the constant-rate path is (CallEmu + position-keyed kernels):
//   * host: `VrEngine` runs the control state machine in integers (FIFO occupancies, clocks, fades, stage
//     switches -- one instance for all channels and streams of a handle, they are in lockstep) and turns one
//     `process` call into a list of data operations with closed-form clocks;
//   * device: `k_vr_ops` (vr.cu), one CTA per channel, executes the list: FIR stages and poly-phase evaluations in
//     parallel over samples, the recursive all-pass sequentially per channel at the end.
// Arithmetic follows the reference operation by operation in float32 without contraction (__fmul_rn / __fadd_rn),
// so results are bit-identical to vr32.c's, not merely within tolerance.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include "runtime.h"

namespace b2 {

struct VrOp {
  int kind;            // VrOpKind
  int flag;            // kind-specific (fast / should_be_fast / is_d | accumulate << 1)
  int n;               // samples to produce
  int aux;             // x_fade_len at the start (cross-fades) / vol index (fades)
  int vstep;           // fade window step per evaluation
  int dstride;         // destination stride (fade of the 1x stream writes even slots only)
  int sa, da;          // stage whose buffer src / dst (arena kinds) is relative to
  long long src, dst;  // element offsets inside that stage's buffer (sources, arena destinations) or inside T / Y
  long long at, step, ss;   // 32.32 clock at the first evaluation, its step, the step's step (slew)
};

enum VrOpKind {
  kVrZero = 0,     // dst[0..n) = 0                                  (arena)
  kVrZeroT,        // T[dst..dst+n) = 0
  kVrDouble,       // x2 interpolator: n outputs from src            (arena -> arena)
  kVrHalf,         // half-band decimator, flag 1: the short one     (arena -> arena)
  kVrXfade,        // cross-fade between the two decimators          (arena -> arena)
  kVrMove,         // dst[0..n) = src[0..n), dst below src             (a FIFO's compaction, fifo.h:56-61)
  kVrPoly,         // poly-phase evaluations                          (arena -> T or Y)
  kVrIir,          // all-pass half-band over pairs of T -> Y         (deferred, sequential)
  kVrPhase,        // all-pass phase equaliser in place on Y          (deferred, sequential)
};

class VrEngine {
 public:
  VrEngine(double max_io_ratio, double scale, unsigned n_sig);
  ~VrEngine();
  const char *init();                                    // device tables and buffers
  void set_io_ratio(double io_ratio, size_t slew_len);   // vr32.c:340-366
  // append n frames per signal (vr_input, vr32.c:572-575): K1 writes them straight behind stage 0's data
  const char *input(const IoLayout &lay, const void *dev_in, size_t n, cudaStream_t st);
  void flush();                                          // vr32.c:592-598
  // vr_process + vr_output (vr32.c:425-590) in two steps, so that the caller can size the result from the count:
  // plan() runs the control flow of one call (up to olen frames) and returns how many frames it yields; run() then
  // executes the planned operations, writing them to planar Y[s * y_stride + i], i < odone
  void plan(size_t olen, size_t *odone);
  const char *run(float *Y, size_t y_stride, cudaStream_t st);
  const char *process(size_t olen, float *Y, size_t y_stride, size_t *odone, cudaStream_t st) {
    plan(olen, odone);
    return run(Y, y_stride, st);
  }
  unsigned n_sig() const { return n_sig_; }
  // control flow only (no device): the frames per call the reference would return, for host-side tests
  void set_dry_run() { dry_ = true; }
  const char *dry_input(size_t n) { IoLayout lay; return input(lay, nullptr, n, nullptr); }

 private:
  // fifo.h's bookkeeping, position for position (allocation grows by the request, data moves to the front only
  // once more than FIFO_MIN bytes were consumed): after two upward stage switches inside one call the reference
  // reads a few samples past a FIFO's end, i.e. whatever that buffer held before -- equal positions keep even
  // those reads equal
  struct Fifo { long alloc = 8192, begin = 0, end = 0; long occ() const { return end - begin; } };
  struct Stage { Fifo f; double step_mult = 0; bool is_fast = true; int x_fade_len = 0, preload = 0; };
  struct Stream {
    int64_t at = 0, step = 0, step_step = 0;
    long input = 0;          // element offset (within its stage's buffer) of the reference's `input` pointer
    int len = 0, stage_num = 0;
    bool is_d = false;
    double step_mult = 0;
  };
  Stage &stg(int i) { return stages_[size_t(i + 1)]; }
  long reserve(int stage, long n);                       // fifo_reserve: offset of the n new elements
  void preload(int stage);
  bool do_input_stage(int stage_num, int sign, int min_stage_num);
  void enter_new_stage(int occupancy0);
  void set_step(Stream *s, double io_ratio) const;
  bool set_step_step(Stream *s, double io_ratio, int slew_len) const;
  int emu_poly_d(Stream *s, int olen, bool fade, int vol, int vstep, long long dst, bool to_y);
  int emu_poly_u(Stream *s, int olen, bool fade, int vol, int vstep, long long dst, bool to_y);
  const char *fit_arena(cudaStream_t st);                 // regions at least as long as the FIFOs' allocations
  const char *run_ops(float *Y, size_t y_stride, long long t_used, cudaStream_t st);
  void push(const VrOp &op) { ops_.push_back(op); }

  bool dry_ = false;
  double max_io_ratio_, scale_;
  unsigned n_sig_;
  int num_stages0_ = 0, num_stages_ = 1, flushing_ = 0;
  int fade_len_ = 0, slew_len_ = 0, xfade_ = 0, stage_inc_ = 0, switch_stage_num_ = 0;
  double new_io_ratio_ = 0, default_io_ratio_ = 0;
  std::vector<Stage> stages_;
  Stream current_, fadeout_;
  std::vector<VrOp> ops_;
  // device side
  long cap_ = 0;                   // elements per stage region; a signal's row holds num_stages + 1 regions
  DevBuf arena_[2];
  int cur_arena_ = 0;
  DevBuf tbuf_, ops_dev_, iir_, tables_;
  PinnedBuf ops_pin_;
  size_t t_stride_ = 0;
  long long t_used_ = 0;           // elements of T the planned call needs
};

}  // namespace b2
