// plan.h -- host-side stage planner and call-schedule emulator.
//
// The planner decides, for an I/O ratio and a quality spec, the same cascade of
// stages libsoxr's constant-rate engine would build (cr.c:264-496), and designs
// the same filters.  Unlike the reference, a stage here is a *stateless closed
// form keyed on absolute sample position* (SURVEY.md 8a-R0):
//
//     rational stage :  out[n] = sum_k g[C + M*n - L*k] * in[k]
//
// so that any range of outputs can be computed by any GPU thread block, in any
// order, with bit-identical results regardless of how the input was chunked.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace b2 {

struct QSpec {  // role of soxr_quality_spec_t, soxr.h:239-246
  double precision = 20, phase_response = 50, passband_end = 0, stopband_begin = 1;
  unsigned long flags = 0;
  const char *error = nullptr;
};
QSpec quality_spec(unsigned long recipe, unsigned long flags);   // soxr.c:106-145

struct RSpec {  // role of soxr_runtime_spec_t, soxr.h:258-265 / soxr.c:163-173
  unsigned log2_min_dft_size = 10, log2_large_dft_size = 17, coef_size_kbytes = 400;
  unsigned long flags = 0;
};

enum StageKind { kHalf, kDft, kPoly, kPolyInterp, kCubic };

struct Stage {
  StageKind kind;
  // ---- closed form of a rational stage (kHalf, kDft, kPoly) -------------------
  std::vector<double> g;   // prototype taps incl. gain
  long C = 0;              // centre offset
  int L = 1, M = 1;
  // ---- kDft (cr.c:184-248) ----------------------------------------------------
  int num_taps = 0, dft_length = 0;
  bool f_domain_m = false;   // decimate by spectrum truncation (cr.c:163-170,192)
  int fd_shift = 0;          // m = log2 of that decimation
  // ---- kPoly / kPolyInterp (cr.c:398-476, poly-fir0.h, poly-fir.h) --------------
  int num_coefs = 0, phases = 0, order = 0, phase_bits = 0;
  int at0_int = 0;                 // vpoly0 initial clock (integer part)
  uint64_t at0_fix = 0, step_fix = 0;   // 32.32 clocks for kPolyInterp / kCubic
  std::vector<double> table;       // kPoly: [phase][tap];  kPolyInterp: [phase][tap][order+1] (highest order first)
  double cubic_mult = 1;
  // ---- what the reference's FIFO bookkeeping uses (for the call emulator) -------
  int emu_preload = 0, emu_pre_post = 0, emu_input_size = 8192, emu_at0 = 0;
  int dev_preload = 0;             // un-padded preload used by the device maths
};

struct Plan {
  double io_ratio = 1;
  bool is_double = false;     // engine precision: soxr.c:447
  double precision_bits = 0;  // the quality spec's precision (stop-band attenuation = (bits + 1) * 6.02 dB, cr.c:288)
  int mode = 0, shr = 0, preL = 1, preM = 1, arbL = 1, postL = 1, postM = 1;
  double arbM = 1;
  bool rational = false, upsample = false;
  std::vector<Stage> stages;
  std::string trace;          // SOXR_TRACE=1-compatible lines
  const char *engine = "b200-cr32";
};

// Returns nullptr or a static error string (same messages as cr.c:299-310).
const char *make_plan(double io_ratio, const QSpec &q, const RSpec &r, double multiplier, Plan *out);

// Total frames produced for n_in input frames (cr.c:544).
inline size_t output_frames(double io_ratio, size_t n_in) { return size_t(double(n_in) / io_ratio + .5); }

// For a stage, the exclusive upper bound of input indices read to produce
// outputs [0, n_out) -- used to size intermediates back to front.
long stage_input_needed(const Stage &s, long n_out);
// Lowest input index read to produce output index n (may be negative => zeros).
long stage_input_first(const Stage &s, long n);

// ---------------------------------------------------------------------------
// Emulation of the reference's pull scheduler (cr.c:498-581) in integers only:
// tells how many frames each soxr_process call returns (bursty), so streaming
// results are partitioned exactly like the reference's, and int16 dither can
// replay the reference's block partition.  No samples are touched.
// ---------------------------------------------------------------------------
class CallEmu {
 public:
  explicit CallEmu(const Plan &plan);
  void reset();
  // soxr_input: append ilen frames (ignored once flushing, cr.c:525-531).
  void add_input(size_t ilen);
  // soxr_output for one call; flush == true first switches to flushing (cr.c:541-547).
  size_t pull(size_t olen, bool flush);
  // One soxr_process(in != NULL, ilen, out, olen) call; returns odone.
  size_t process(size_t ilen, size_t olen) { add_input(ilen); return pull(olen, false); }
  // One soxr_process(NULL, 0, out, olen) call.
  size_t flush(size_t olen) { return pull(olen, true); }
  double delay() const;       // cr.c:578-581
  int64_t samples_in_total() const { return total_in_; }
  int64_t samples_out_total() const { return total_out_; }
  bool flushing() const { return flushing_; }

 private:
  struct S {
    const Stage *st;
    long occ = 0;            // fifo occupancy
    int input_size = 0;
    int at = 0, remM = 0;    // integer clocks
    uint64_t at_fix = 0;
  };
  bool stage_process(int i);
  void run_stage(int i);
  size_t output(size_t olen);
  const Plan &plan_;
  std::vector<S> s_;
  long out_occ_ = 0;          // occupancy of the final fifo
  int64_t samples_in_ = 0, samples_out_ = 0, total_in_ = 0, total_out_ = 0;
  bool flushing_ = false;
};

}  // namespace b2
