// vr.cu -- variable-rate engine: host control emulator + device operation interpreter (see vr.h).
//
// Reference behaviour restated: libsoxr/src/vr32.c (state machine :253-570, filters :35-109, :155-186,
// table preparation :115-153), libsoxr/src/fifo.h (occupancy bookkeeping).
#include "vr.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "vr_tables.h"

namespace b2 {

namespace {
// vr32.c:16-27
constexpr int kFadeLenBits = 9, kPhaseBitsD = 10, kPhaseBitsU = 9;
constexpr int kPhases0D = 12, kPolyLenD = 20, kPhases0U = 6, kPolyLenU = 12;
constexpr int kPhasesD = 1 << kPhaseBitsD, kPhasesU = 1 << kPhaseBitsU;
constexpr int kHfl2 = 120, kHfl4 = 60;              // HALF_FIR_LEN_2 / _4 (vr32.c:34-35)
constexpr int kFadeN = (2 << kFadeLenBits) + 1;     // entries of the cross-fade window (vr32.c:275)
constexpr double kMult32 = 65536. * 65536.;
// device table layout (floats)
constexpr int kOffHalf = 0, kOffFast = kOffHalf + 121, kOffAllpass = kOffFast + 6, kOffFade = kOffAllpass + 15,
              kOffPolyD = kOffFade + kFadeN, kOffPolyU = kOffPolyD + kPolyLenD * kPhasesD * 2,
              kTabFloats = kOffPolyU + kPolyLenU * kPhasesU * 2;
constexpr int kIirState = 17;                       // x[2], y[15] per signal (half_iir_t, vr32.c:86)

const bool kVrTrace = [] { const char *e = std::getenv("B200SOXR_VR_TRACE"); return e && e[0] == '1'; }();   // stage switches, per-call state on stderr

inline int64_t shr(int64_t x, int by) { return by < 0 ? x << (-by) : x >> by; }      // shiftr, vr32.c:251
inline int64_t shl(int64_t x, int by) { return shr(x, -by); }
inline int32_t int_part(int64_t v) { return int32_t(v >> 32); }                      // INT(), little-endian union (:197-205)
inline uint32_t frac_part(int64_t v) { return uint32_t(uint64_t(v)); }

// Interpolated poly-phase table from a prototype (prepare_coefs, vr32.c:115-153): quintic interpolation of the
// prototype from phases0 to phases points per tap in double, then [phase][tap]{slope, value} pairs in float.
void make_poly_table(float *coefs, int n, int phases0, int phases, const float *proto, double multiplier) {
  const int length0 = n * phases0, length = n * phases, K0 = 2;
  auto sym = [](const float *h, int l, int i) -> float { return (i < 0 || i >= l) ? 0.f : h[i > l / 2 ? l - i : i]; };
  std::vector<float> fine(size_t(length / 2 + 1));
  float *p = fine.data();
  double k[6];
  for (int j = 0; j < 6; ++j) k[j] = sym(proto, length0, j - K0);
  int pos = 0;
  for (int i = 0; i < length0 / 2; ++i) {
    const double b = (1 / 24.) * (k[0] + k[4] + 6 * k[2] - 4 * (k[1] + k[3])), d = .5 * (k[1] + k[3]) - k[2] - b;
    const double a = (1 / 120.) * (k[5] - k[2] - 9 * (9 * b + d) + 2.5 * (k[3] - k[1]) - 2 * (k[4] - k[0]));
    const double c = (1 / 12.) * (k[4] - k[0] - 2 * (k[3] - k[1]) - 60 * a), e = .5 * (k[3] - k[1]) - a - c;
    for (; pos / phases == i; pos += phases0) {
      const double x = double(pos % phases) / phases;
      *p++ = float(k[K0] + ((((a * x + b) * x + c) * x + d) * x + e) * x);
    }
    for (int j = 0; j < 5; ++j) k[j] = k[j + 1];
    k[5] = sym(proto, length0, i + 6 / 2 + 1);
  }
  if (!(length & 1)) *p++ = float(k[K0]);
  float f0, f1 = 0;
  const float mult = float(multiplier);
  for (int i = 0; i < n; ++i)
    for (int j = phases - 1; j >= 0; --j, f1 = f0) {
      const int at = (n - 1 - i) * phases + j;
      f0 = sym(fine.data(), length, at) * mult;
      coefs[size_t(n) * 2 * j + 2 * i + 1] = f0;        // value   (coef(.., 0, i), vr32.c:111-113)
      coefs[size_t(n) * 2 * j + 2 * i + 0] = f1 - f0;   // slope   (coef(.., 1, i))
    }
}

// ---- device side ------------------------------------------------------------------------------------------
struct VrKParams {
  float *arena; long long row, cap;      // a signal's row = (stages + 1) regions of `cap` elements, stage i at (i + 1) * cap
  float *T; long long t_stride;
  float *Y; long long y_stride;
  float *iir;
  const VrOp *ops; int n_ops;
  const float *tab;
};

// Every product and sum is rounded to float32 on its own, like the reference compiled for x86-64 (no FMA).
__device__ __forceinline__ float mulr(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float addr(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float subr(float a, float b) { return __fsub_rn(a, b); }

__device__ float d_half_fir(const float *in, const float *c) {           // vr32.c:37-46
  float sum = mulr(in[0], c[0]);
  for (int i = 1; i <= kHfl2; ++i) sum = addr(sum, mulr(addr(in[-i], in[i]), c[i]));
  return sum;
}
__device__ float d_fast_half_fir(const float *in, const float *fc) {     // vr32.c:70-79
  float sum = mulr(in[0], .5f);
#pragma unroll
  for (int i = 0; i < 6; ++i) sum = addr(sum, mulr(addr(in[-(2 * i + 1)], in[2 * i + 1]), fc[i]));
  return sum;
}
__device__ float d_double_fir0(const float *in, const float *c) {        // vr32.c:48-57
  float sum = mulr(in[0], c[0]);
  for (int i = 1; i <= kHfl4; ++i) sum = addr(sum, mulr(addr(in[-i], in[i]), c[2 * i]));
  return mulr(sum, 2.f);
}
__device__ float d_double_fir1(const float *in, const float *c) {        // vr32.c:59-68
  float sum = 0;
  for (int i = 0; i < kHfl4; ++i) sum = addr(sum, mulr(addr(in[-i], in[1 + i]), c[2 * i + 1]));
  return mulr(sum, 2.f);
}
template <int LEN, int PHASE_BITS>
__device__ float d_poly(const float *in, unsigned frac, const float *table) {   // vr32.c:155-186
  const int phase = int(frac >> (32 - PHASE_BITS));
  const float x = mulr(__uint2float_rn(frac << PHASE_BITS), float(1 / kMult32));
  const float *t = table + size_t(LEN) * 2 * phase;
  float sum = 0;
#pragma unroll 4
  for (int i = 0; i < LEN; ++i) sum = addr(sum, mulr(addr(mulr(t[2 * i], x), t[2 * i + 1]), in[i]));
  return sum;
}
// one step of the all-pass half-band (half_iir1, vr32.c:88-98): two paths of first-order all-pass sections
__device__ float d_half_iir1(float *st, const float *c, float in0, float in1) {
  float *x = st, *y = st + 2;
  float tmp0 = x[0], tmp1 = x[1];
  x[0] = in0;
  x[1] = in1;
  int i = 0;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    in1 = addr(mulr(subr(in1, y[i]), c[i]), tmp1); tmp1 = y[i]; y[i] = in1; ++i;
    in0 = addr(mulr(subr(in0, y[i]), c[i]), tmp0); tmp0 = y[i]; y[i] = in0; ++i;
  }
  y[i] = in1 = addr(mulr(subr(in1, y[i]), c[i]), tmp1);
  return addr(in1, in0);
}

__global__ void __launch_bounds__(256) k_vr_ops(const VrKParams P) {
  const unsigned s = blockIdx.x;
  float *A = P.arena + (long long)s * P.row;
  float *T = P.T + (long long)s * P.t_stride;
  float *Y = P.Y + (long long)s * P.y_stride;
  const float *tab = P.tab;
  const int tid = int(threadIdx.x), nt = int(blockDim.x);
  bool deferred = false;
  for (int o = 0; o < P.n_ops; ++o) {
    const VrOp op = P.ops[o];
    switch (op.kind) {
      case kVrZero:
        for (int i = tid; i < op.n; i += nt) A[(op.da + 1) * P.cap + op.dst + i] = 0.f;
        break;
      case kVrMove: {   // towards lower addresses, ranges may overlap: a round reads before it writes
        float *b = A + (op.da + 1) * P.cap;
        for (int base = 0; base < op.n; base += nt) {
          const int i = base + tid;
          const float v = i < op.n ? b[op.src + i] : 0.f;
          __syncthreads();
          if (i < op.n) b[op.dst + i] = v;
          __syncthreads();
        }
        break;
      }
      case kVrZeroT:
        for (int i = tid; i < op.n; i += nt) T[op.dst + i] = 0.f;
        break;
      case kVrDouble:   // do_input_stage, stage < 0 (vr32.c:385-387): pairs of outputs per input position
        for (int i = tid; 2 * i < op.n; i += nt) {
          const float *src = A + (op.sa + 1) * P.cap + op.src + i;
          float *dst = A + (op.da + 1) * P.cap + op.dst;
          dst[2 * i] = d_double_fir0(src, tab + kOffHalf);
          dst[2 * i + 1] = d_double_fir1(src, tab + kOffHalf);
        }
        break;
      case kVrHalf:     // vr32.c:408-410
        for (int i = tid; i < op.n; i += nt) {
          const float *src = A + (op.sa + 1) * P.cap + op.src + 2 * i;
          A[(op.da + 1) * P.cap + op.dst + i] = op.flag ? d_fast_half_fir(src, tab + kOffFast) : d_half_fir(src, tab + kOffHalf);
        }
        break;
      case kVrXfade: {  // vr32.c:396-407: the window runs down for one filter and up for the other
        const float *fade = tab + kOffFade;
        for (int i = tid; i < op.n; i += nt) {
          const float *src = A + (op.sa + 1) * P.cap + op.src + 2 * i;
          const float v1 = fade[(op.aux << 1) - 2 * i], v2 = fade[(((1 << kFadeLenBits) - op.aux) << 1) + 2 * i];
          const float f = d_fast_half_fir(src, tab + kOffFast), h = d_half_fir(src, tab + kOffHalf);
          A[(op.da + 1) * P.cap + op.dst + i] = op.flag ? addr(mulr(v1, f), mulr(v2, h)) : addr(mulr(v2, f), mulr(v1, h));
        }
        break;
      }
      case kVrPoly: {   // poly_fir_d / _u and their fade variants (vr32.c:215-249), clocks in closed form
        const bool is_d = op.flag & 1, fade = (op.flag & 2) != 0, to_y = (op.flag & 4) != 0;
        float *dst = (to_y ? Y : T) + op.dst;
        const float *in = A + (op.sa + 1) * P.cap + op.src;
        for (int i = tid; i < op.n; i += nt) {
          long long at;
          if (is_d) {   // pairs share a step; the step moves by ss after each pair
            const long long m = i >> 1;
            at = op.at + 2 * m * op.step + op.ss * m * (m - 1);
            if (i & 1) at += op.step + m * op.ss;
          } else {
            const long long m = i;
            at = op.at + m * op.step + op.ss * (m * (m - 1) / 2);
          }
          const int whole = int(at >> 32);
          const unsigned frac = unsigned((unsigned long long)at);
          const float v = is_d ? d_poly<kPolyLenD, kPhaseBitsD>(in + whole, frac, tab + kOffPolyD)
                               : d_poly<kPolyLenU, kPhaseBitsU>(in + whole, frac, tab + kOffPolyU);
          float *d = dst + (long long)i * op.dstride;
          if (fade) *d = addr(*d, mulr(tab[kOffFade + op.aux + i * op.vstep], v));
          else *d = v;
        }
        break;
      }
      default:
        deferred = true;
        break;
    }
    __syncthreads();
  }
  // the recursive part: sequential per channel (vr32.c:100-109), in call order
  if (deferred && tid == 0) {
    float st[kIirState];
    float *g = P.iir + (long long)s * kIirState;
    for (int i = 0; i < kIirState; ++i) st[i] = g[i];
    const float *c = tab + kOffAllpass;
    for (int o = 0; o < P.n_ops; ++o) {
      const VrOp op = P.ops[o];
      if (op.kind == kVrIir) {
        for (int i = 0; i < op.n; ++i) Y[op.dst + i] = d_half_iir1(st, c, T[op.src + 2 * i], T[op.src + 2 * i + 1]);
      } else if (op.kind == kVrPhase) {
        for (int i = 0; i < op.n; ++i) Y[op.dst + i] = d_half_iir1(st, c, Y[op.dst + i], 0.f);
        const float small_normal = float(1 / kMult32 / kMult32);   // quashes denormals on path 0 (vr32.c:104-108)
        for (int i = 0; i < 15; i += 2) st[2 + i] = addr(st[2 + i], small_normal);
        for (int i = 0; i < 15; i += 2) st[2 + i] = subr(st[2 + i], small_normal);
      }
    }
    for (int i = 0; i < kIirState; ++i) g[i] = st[i];
  }
}
}  // namespace

// ---- host side --------------------------------------------------------------------------------------------
VrEngine::VrEngine(double max_io_ratio, double scale, unsigned n_sig)
    : max_io_ratio_(max_io_ratio), scale_(scale), n_sig_(n_sig) {
  // vr_create (vr32.c:618-627): one decimating stage per octave of the largest ratio
  int n = 0;
  for (double x = max_io_ratio; x > 1; x *= .5) ++n;
  // vr_init (vr32.c:277-305)
  num_stages0_ = n;
  num_stages_ = std::max(n, 1);
  stages_.resize(size_t(num_stages_) + 1);
  for (int i = -1; i < num_stages_; ++i) {
    Stage &s = stg(i);
    s.step_mult = 2 * kMult32 / double(shl(2, i));
    s.preload = i < 0 ? 0 : i == 0 ? 2 * kHfl2 : 3 * kHfl2 / 2;
    s.is_fast = true;
  }
  default_io_ratio_ = max_io_ratio;
}

VrEngine::~VrEngine() {
  arena_[0].release(); arena_[1].release();
  tbuf_.release(); ops_dev_.release(); iir_.release(); tables_.release(); ops_pin_.release();
}

const char *VrEngine::init() {
  // fifo_create + stage_preload (vr32.c:284-293): every FIFO starts with FIFO_MIN bytes and its preload of zeros
  for (int i = -1; i < num_stages_; ++i) { Stage &st = stg(i); st.f = Fifo(); st.f.end = st.preload; }
  cap_ = 1 << 14;
  if (dry_) return nullptr;
  // tables: the five prototypes, the raised-cosine cross-fade window (vr32.c:298-300) and the two interpolated
  // poly-phase tables (the x2-rate one carries half the gain, vr32.c:301-302)
  std::vector<float> tab(kTabFloats);
  std::memcpy(&tab[kOffHalf], vrt::kHalfFir, sizeof vrt::kHalfFir);
  std::memcpy(&tab[kOffFast], vrt::kFastHalfFir, sizeof vrt::kFastHalfFir);
  std::memcpy(&tab[kOffAllpass], vrt::kAllpass, sizeof vrt::kAllpass);
  for (int i = 0; i < kFadeN; ++i) tab[kOffFade + i] = float(.5 * (1 + std::cos(M_PI * i / (kFadeN - 1))));
  make_poly_table(&tab[kOffPolyU], kPolyLenU, kPhases0U, kPhasesU, vrt::kProtoUp, scale_);
  make_poly_table(&tab[kOffPolyD], kPolyLenD, kPhases0D, kPhasesD, vrt::kProtoDown, scale_ * .5);
  if (const char *e = tables_.reserve(tab.size() * 4)) return e;
  B2_CUDA(cudaMemcpy(tables_.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
  if (const char *e = iir_.reserve(size_t(n_sig_) * kIirState * 4)) return e;
  B2_CUDA(cudaMemset(iir_.p, 0, size_t(n_sig_) * kIirState * 4));
  // stage arena: one region of cap_ elements per stage and signal, zero-filled (a fresh heap block's likeliest content)
  const size_t bytes = size_t(cap_) * size_t(num_stages_ + 1) * n_sig_ * 4;
  if (const char *e = arena_[0].reserve(bytes)) return e;
  B2_CUDA(cudaMemset(arena_[0].p, 0, bytes));
  cur_arena_ = 0;
  return nullptr;
}

long VrEngine::reserve(int stage, long n) {          // fifo_reserve (fifo.h:41-66): same positions as the reference's
  Fifo &f = stg(stage).f;
  if (f.begin == f.end) f.begin = f.end = 0;
  for (;;) {
    if (f.end + n <= f.alloc) {
      const long at = f.end;
      f.end += n;
      return at;
    }
    if (f.begin > 8192) {                            // FIFO_MIN = 0x8000 bytes consumed: data moves to the front
      VrOp op{}; op.kind = kVrMove; op.da = stage; op.n = int(f.occ()); op.src = f.begin; op.dst = 0; push(op);
      f.end -= f.begin;
      f.begin = 0;
      continue;
    }
    f.alloc += n;                                    // realloc: what is there stays where it is
  }
}

void VrEngine::preload(int stage) {                  // stage_preload (vr32.c:255-256)
  const int n = stg(stage).preload;
  const long dst = reserve(stage, n);
  if (n) { VrOp op{}; op.kind = kVrZero; op.da = stage; op.n = n; op.dst = dst; push(op); }
}

void VrEngine::set_step(Stream *s, double io_ratio) const { s->step = int64_t(io_ratio * s->step_mult + .5); }

bool VrEngine::set_step_step(Stream *s, double io_ratio, int slew_len) const {   // vr32.c:326-338
  Stream tmp = *s;
  set_step(&tmp, io_ratio);
  int64_t dif = tmp.step - s->step;
  dif = dif < 0 ? dif - (slew_len >> 1) : dif + (slew_len >> 1);
  s->step_step = dif / slew_len;
  return s->step_step != 0;
}

void VrEngine::enter_new_stage(int occupancy0) {     // vr32.c:307-317
  current_.len = int(shr(occupancy0, current_.stage_num));
  current_.input = stg(current_.stage_num).f.begin + 2 * kHfl2;
  current_.step_mult = stg(current_.stage_num).step_mult;
  current_.is_d = current_.stage_num >= 0;
  if (current_.is_d) current_.step_mult *= .5;
}

void VrEngine::set_io_ratio(double io_ratio, size_t slew_len) {   // vr_set_io_ratio, vr32.c:340-366
  if (slew_len) {
    if (!set_step_step(&current_, io_ratio, slew_len_ = int(slew_len))) {
      slew_len_ = 0; new_io_ratio_ = 0; fadeout_.step_step = 0;
    } else {
      new_io_ratio_ = io_ratio;
      if (fade_len_) set_step_step(&fadeout_, io_ratio, slew_len_);
    }
  } else {
    if (default_io_ratio_ != 0) {   // the first call
      const int octave = int(std::floor(std::log(io_ratio) / M_LN2));
      current_.stage_num = octave < 0 ? -1 : std::min(octave, num_stages0_ - 1);
      enter_new_stage(0);
    } else if (fade_len_) {
      set_step(&fadeout_, io_ratio);
    }
    set_step(&current_, io_ratio);
    if (default_io_ratio_ != 0)
      current_.at = (current_.at & ~int64_t(0xffffffffll)) | int64_t(frac_part(current_.step) >> 1);
    default_io_ratio_ = 0;
  }
}

bool VrEngine::do_input_stage(int stage_num, int sign, int min_stage_num) {   // vr32.c:368-416
  Stage &s = stg(stage_num), &s1 = stg(stage_num - sign);
  long src = s1.f.begin + kHfl2;
  int len = int(shr(s1.f.occ() - kHfl2 * 2, sign));
  const int already_done = int(s.f.occ()) - s.preload;
  if (kVrTrace) std::fprintf(stderr, "vr: input stage %d sign %d: src occ %ld len %d already %d\n", stage_num, sign, s1.f.occ(), len, already_done);
  if ((len -= already_done) <= 0) return false;
  src += long(shl(already_done, sign));
  const long dest = reserve(stage_num, len);
  int i = 0;
  if (stage_num < 0) {
    VrOp op{}; op.kind = kVrDouble; op.sa = stage_num - sign; op.da = stage_num; op.n = len; op.src = src; op.dst = dest; push(op);
  } else {
    const bool should_be_fast = stage_inc_ != 0;
    if (!s.x_fade_len && stage_num == switch_stage_num_) {
      switch_stage_num_ = 0;
      if (s.is_fast != should_be_fast) { s.x_fade_len = 1 << kFadeLenBits; s.is_fast = should_be_fast; ++xfade_; }
    }
    if (s.x_fade_len) {
      const int n = std::min(len, s.x_fade_len);
      VrOp op{}; op.kind = kVrXfade; op.sa = stage_num - sign; op.da = stage_num; op.flag = should_be_fast; op.n = n;
      op.aux = s.x_fade_len; op.src = src; op.dst = dest;
      push(op);
      i = n;
      s.x_fade_len -= n;
      xfade_ -= !s.x_fade_len;
    }
    if (i < len) {
      VrOp op{}; op.kind = kVrHalf; op.sa = stage_num - sign; op.da = stage_num; op.flag = stage_num < min_stage_num;
      op.n = len - i; op.src = src + 2 * i; op.dst = dest + i;
      push(op);
    }
  }
  if (flushing_ > 0) preload(stage_num);
  return true;
}

// poly_fir_d / poly_fir_fade_d (vr32.c:215-249): the clock walk decides how many evaluations fit the data at hand;
// the device repeats them from the closed form
int VrEngine::emu_poly_d(Stream *s, int olen, bool fade, int vol, int vstep, long long dst, bool to_y) {
  VrOp op{}; op.kind = kVrPoly; op.flag = 1 | (fade ? 2 : 0) | (to_y ? 4 : 0);
  op.sa = s->stage_num; op.src = s->input - kPolyLenD / 2 + 1; op.dst = dst; op.dstride = 1;
  op.at = s->at; op.step = s->step; op.ss = s->step_step; op.aux = vol; op.vstep = vstep;
  int i;
  for (i = 0; i < olen && int_part(s->at) < s->len; ++i) {
    s->at += s->step;
    if (!(int_part(s->at) < s->len)) { s->at -= s->step; break; }
    ++i;
    s->at += s->step;
    s->step += s->step_step;
  }
  op.n = i;
  if (i) push(op);
  return i;
}
int VrEngine::emu_poly_u(Stream *s, int olen, bool fade, int vol, int vstep, long long dst, bool to_y) {
  VrOp op{}; op.kind = kVrPoly; op.flag = 0 | (fade ? 2 : 0) | (to_y ? 4 : 0);
  op.sa = s->stage_num; op.src = s->input - kPolyLenU / 2 + 1; op.dst = dst; op.dstride = fade ? 2 : 1;
  op.at = s->at; op.step = s->step; op.ss = s->step_step; op.aux = vol; op.vstep = vstep;
  int i, n = 0;
  for (i = 0; i < olen && int_part(s->at) < s->len; i += fade ? 2 : 1) {   // (the fade writes every other slot)
    s->at += s->step;
    s->step += s->step_step;
    ++n;
  }
  op.n = n;
  if (n) push(op);
  return i;
}

// Regions at least as long as the FIFOs' allocations (a realloc in the reference keeps what the block held: the old
// regions are carried over whole, the new tail reads as zeros).
const char *VrEngine::fit_arena(cudaStream_t st) {
  long need = cap_;
  for (int i = -1; i < num_stages_; ++i) need = std::max(need, stg(i).f.alloc);
  if (need <= cap_ || dry_) return nullptr;
  long new_cap = cap_;
  while (new_cap < need) new_cap *= 2;
  const size_t regions = size_t(num_stages_ + 1);
  DevBuf &dst = arena_[cur_arena_ ^ 1];
  if (const char *e = dst.reserve(size_t(new_cap) * regions * n_sig_ * 4)) return e;
  B2_CUDA(cudaMemsetAsync(dst.p, 0, size_t(new_cap) * regions * n_sig_ * 4, st));
  B2_CUDA(cudaMemcpy2DAsync(dst.p, size_t(new_cap) * 4, arena_[cur_arena_].p, size_t(cap_) * 4, size_t(cap_) * 4,
                            regions * n_sig_, cudaMemcpyDeviceToDevice, st));
  cap_ = new_cap;
  cur_arena_ ^= 1;
  return nullptr;
}

const char *VrEngine::run_ops(float *Y, size_t y_stride, long long t_used, cudaStream_t st) {
  if (ops_.empty() || dry_) { ops_.clear(); return nullptr; }
  if (const char *e = fit_arena(st)) return e;
  t_stride_ = size_t(std::max<long long>(t_used, 1) + 63) & ~size_t(63);
  if (const char *e = tbuf_.reserve(t_stride_ * n_sig_ * 4)) return e;
  const size_t bytes = ops_.size() * sizeof(VrOp);
  if (const char *e = ops_pin_.reserve(bytes)) return e;
  if (const char *e = ops_dev_.reserve(bytes)) return e;
  std::memcpy(ops_pin_.p, ops_.data(), bytes);
  B2_CUDA(cudaMemcpyAsync(ops_dev_.p, ops_pin_.p, bytes, cudaMemcpyHostToDevice, st));
  VrKParams P;
  P.arena = arena_[cur_arena_].as<float>(); P.cap = cap_; P.row = (long long)cap_ * (num_stages_ + 1);
  P.T = tbuf_.as<float>(); P.t_stride = (long long)t_stride_;
  P.Y = Y; P.y_stride = (long long)y_stride;
  P.iir = iir_.as<float>();
  P.ops = ops_dev_.as<VrOp>(); P.n_ops = int(ops_.size());
  P.tab = tables_.as<float>();
  k_vr_ops<<<n_sig_, 256, 0, st>>>(P);
  B2_COUNT_LAUNCH();
  B2_CUDA(cudaGetLastError());
  B2_CUDA(cudaStreamSynchronize(st));   // the pinned operation list is reused by the next call
  ops_.clear();
  return nullptr;
}

const char *VrEngine::input(const IoLayout &lay, const void *dev_in, size_t n, cudaStream_t st) {
  if (!n) return nullptr;
  ops_.clear();
  const long at = reserve(0, long(n));                 // (may ask for a compaction first)
  if (const char *e = run_ops(nullptr, 0, 0, st)) return e;
  if (dry_) return nullptr;
  if (const char *e = fit_arena(st)) return e;
  float *X = arena_[cur_arena_].as<float>() + size_t(cap_);     // stage 0's region
  return convert_in(lay, dev_in, n, false, X, size_t(cap_) * size_t(num_stages_ + 1), size_t(at), st);
}

void VrEngine::flush() {            // vr_flush (vr32.c:592-598); the zeros are written by the next process()
  if (!flushing_) flushing_ = 2;    // 2: preload of stage 0 still owed (becomes 1 in process())
}

void VrEngine::plan(size_t olen_, size_t *odone_out) {
  ops_.clear();
  if (flushing_ == 2) { preload(0); flushing_ = 1; }
  if (default_io_ratio_ != 0) set_io_ratio(default_io_ratio_, 0);
  const int olen0 = int(std::min<size_t>(olen_, size_t(1) << 30));
  long long t_used = 0;
  int odone0 = 0, min_stage_num = current_.stage_num, max_stage_num = min_stage_num;
  if (fade_len_) {
    min_stage_num = std::min(min_stage_num, fadeout_.stage_num);
    max_stage_num = std::max(max_stage_num, fadeout_.stage_num);
  }
  for (int j = std::min(min_stage_num, 0); j <= max_stage_num; ++j)
    if (j && !do_input_stage(j, j < 0 ? -1 : 1, min_stage_num)) break;
  if (flushing_ > 0) flushing_ = -1;

  auto stage_occupancy = [&](int i) { return int(stg(i).f.occ()) - 4 * kHfl2; };
  auto read_p = [&](int i) { return stg(i).f.begin + 2 * kHfl2; };
  const int occupancy0 = int(shl(std::max(0, stage_occupancy(max_stage_num)), max_stage_num));
  current_.len = int(shr(occupancy0, current_.stage_num));
  current_.input = read_p(current_.stage_num);
  if (fade_len_) {
    fadeout_.len = int(shr(occupancy0, fadeout_.stage_num));
    fadeout_.input = read_p(fadeout_.stage_num);
  }

  while (odone0 < olen0) {
    int odone, olen = olen0 - odone0, stage_dif = 0;
    olen = std::min(olen, 64);
    if (slew_len_) olen = std::min(olen, slew_len_);
    else if (new_io_ratio_ != 0) {
      set_step(&current_, new_io_ratio_);
      set_step(&fadeout_, new_io_ratio_);
      fadeout_.step_step = current_.step_step = 0;
      new_io_ratio_ = 0;
    }
    if (!flushing_ && !fade_len_ && !xfade_) {
      const int32_t si = int_part(current_.step);
      const uint32_t sf = frac_part(current_.step);
      if (current_.is_d) {
        if (si && sf) { stage_dif = 1; ++max_stage_num; }
        else if (!si && sf < (1u << 31)) { stage_dif = -1; --min_stage_num; }
      } else if (si > 1 && sf) { stage_dif = 1; ++max_stage_num; }
    }
    if (stage_dif) {
      const int n = current_.stage_num + stage_dif;
      if (n >= num_stages_) --max_stage_num;
      else {
        stage_inc_ = stage_dif > 0;
        fadeout_ = current_;
        current_.stage_num += stage_dif;
        if (!stage_inc_) switch_stage_num_ = current_.stage_num;
        if ((current_.stage_num < 0 && stage_dif < 0) || (current_.stage_num > 0 && stage_dif > 0)) {
          Stage &s = stg(current_.stage_num);
          s.f.begin = s.f.end = 0;
          preload(current_.stage_num);
          s.is_fast = false;
          do_input_stage(current_.stage_num, stage_dif, current_.stage_num);
        }
        if (current_.stage_num > 0 && stage_dif < 0) {
          const int idone = int_part(current_.at);
          Stage &s = stg(current_.stage_num);
          s.f.end = s.f.begin + (2 * kHfl2 + idone + (kPolyLenD >> 1));   // fifo_trim_to
          do_input_stage(current_.stage_num, 1, current_.stage_num);
        }
        enter_new_stage(occupancy0);
        int shift = -stage_dif;
        current_.at = shl(current_.at, shift);
        shift += int(fadeout_.is_d) - int(current_.is_d);
        current_.step = shl(current_.step, shift);
        current_.step_step = shl(current_.step_step, shift);
        fade_len_ = kFadeN - 1;
        if (kVrTrace)
          std::fprintf(stderr, "vr: switch from stage %d to %d, x2 from %d to %d (odone0 %d)\n", fadeout_.stage_num,
                       current_.stage_num, int(fadeout_.is_d), int(current_.is_d), odone0);
      }
    }

    if (fade_len_) {
      const int vol1 = fade_len_, vol2 = kFadeN - 1 - fade_len_;
      const int olen2 = (olen = std::min(olen, fade_len_ >> 1)) << 1;
      { VrOp z{}; z.kind = kVrZeroT; z.n = olen2; z.dst = t_used; if (olen2) push(z); }
      // the x2-rate stream goes first: it is the finer-grained one and decides how many pairs there are
      if (current_.is_d && fadeout_.is_d) {
        odone = emu_poly_d(&current_, olen2, true, vol1, -1, t_used, false);
        emu_poly_d(&fadeout_, odone, true, vol2, 1, t_used, false);
      } else if (current_.is_d) {
        odone = emu_poly_d(&current_, olen2, true, vol1, -1, t_used, false);
        emu_poly_u(&fadeout_, odone, true, vol2, 2, t_used, false);
      } else {
        odone = emu_poly_d(&fadeout_, olen2, true, vol2, 1, t_used, false);
        emu_poly_u(&current_, odone, true, vol1, -2, t_used, false);
      }
      fade_len_ -= odone;
      if (!fade_len_) {
        if (stage_inc_) switch_stage_num_ = min_stage_num++;
        else --max_stage_num;
      }
      odone >>= 1;
      { VrOp q{}; q.kind = kVrIir; q.n = odone; q.src = t_used; q.dst = odone0; if (odone) push(q); }
      t_used += olen2;
    } else if (current_.is_d) {
      odone = emu_poly_d(&current_, olen << 1, false, 0, 0, t_used, false) >> 1;
      { VrOp q{}; q.kind = kVrIir; q.n = odone; q.src = t_used; q.dst = odone0; if (odone) push(q); }
      t_used += olen << 1;
    } else {
      odone = emu_poly_u(&current_, olen, false, 0, 0, odone0, true);
      if (num_stages0_) { VrOp q{}; q.kind = kVrPhase; q.n = odone; q.dst = odone0; push(q); }
    }
    odone0 += odone;
    if (slew_len_) slew_len_ -= odone;
    if (odone != olen) break;   // needs more input
  }
  {
    const int from = std::max(0, max_stage_num), to = std::min(0, min_stage_num);
    int idone = int(shr(int_part(current_.at), from - current_.stage_num));
    current_.at -= int64_t(shl(idone, from - current_.stage_num)) << 32;
    if (fade_len_) fadeout_.at -= int64_t(shl(idone, from - fadeout_.stage_num)) << 32;
    if (kVrTrace) std::fprintf(stderr, "vr: drain from %d to %d idone %d\n", from, to, idone);
    for (int i = from; i >= to; --i, idone <<= 1) {
      Fifo &f = stg(i).f;
      if (idone >= 0 && long(idone) <= f.occ()) f.begin += idone;     // fifo_read drops nothing when asked for too much
    }
  }
  *odone_out = size_t(odone0);
  if (kVrTrace) {
    std::fprintf(stderr, "vr: process(%d) -> %d; stage %d len %d at %d flushing %d fade %d xfade %d slew %d occ:", olen0, odone0,
                 current_.stage_num, current_.len, int_part(current_.at), flushing_, fade_len_, xfade_, slew_len_);
    for (int i = -1; i < num_stages_; ++i) std::fprintf(stderr, " %ld", stg(i).f.occ());
    std::fprintf(stderr, "\n");
  }
  t_used_ = t_used;
}

const char *VrEngine::run(float *Y, size_t y_stride, cudaStream_t st) { return run_ops(Y, y_stride, t_used_, st); }

}  // namespace b2
