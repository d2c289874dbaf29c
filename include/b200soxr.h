/* b200soxr.h -- C ABI of libb200soxr.so, the B200-native (sm_100a) drop-in for
 * the constant-rate resampling path of libsoxr as python-soxr binds it.
 *
 * Every entry point cites the reference interface it replaces
 * (paths relative to /root/reference).  Plain pointers and sizes only: no
 * torch / numpy / C++ types cross this boundary.  Lengths are FRAMES per
 * channel, exactly as in libsoxr (libsoxr/src/soxr.h:53-54).
 *
 * Error convention = libsoxr's (libsoxr/src/soxr.h:76,95-96): functions return
 * NULL on success or a static, NUL-terminated message.
 */
#ifndef B200SOXR_H
#define B200SOXR_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200SOXR_API __attribute__((visibility("default")))
#else
#define B200SOXR_API
#endif

typedef struct b200soxr *b200soxr_t;          /* replaces soxr_t, soxr.h:74 */
typedef char const *b200soxr_error_t;         /* replaces soxr_error_t, soxr.h:76 */

/* datatypes: same numeric values as soxr_datatype_t (soxr.h:207-219) */
enum {
  B200SOXR_FLOAT32 = 0, B200SOXR_FLOAT64 = 1, B200SOXR_INT32 = 2, B200SOXR_INT16 = 3,
  B200SOXR_SPLIT = 4,
  B200SOXR_FLOAT32_I = 0, B200SOXR_FLOAT64_I = 1, B200SOXR_INT32_I = 2, B200SOXR_INT16_I = 3,
  B200SOXR_FLOAT32_S = 4, B200SOXR_FLOAT64_S = 5, B200SOXR_INT32_S = 6, B200SOXR_INT16_S = 7
};

/* quality recipes: same values as soxr.h:284-293 */
#define B200SOXR_QQ  0
#define B200SOXR_LQ  1
#define B200SOXR_MQ  2
#define B200SOXR_HQ  4
#define B200SOXR_VHQ 6

/* io_spec.flags (soxr.h:234-235) */
#define B200SOXR_TPDF       0u
#define B200SOXR_NO_DITHER  8u
/* quality_spec.flags (soxr.h:248-254) */
#define B200SOXR_ROLLOFF_SMALL    0u
#define B200SOXR_ROLLOFF_MEDIUM   1u
#define B200SOXR_ROLLOFF_NONE     2u
#define B200SOXR_HI_PREC_CLOCK    8u
#define B200SOXR_DOUBLE_PRECISION 16u
#define B200SOXR_VR               32u   /* variable-rate engine (vr32.c); ratio changes via b200soxr_set_io_ratio */

/* Layout-compatible with soxr_io_spec_t (soxr.h:226-232). */
typedef struct b200soxr_io_spec {
  int itype, otype;
  double scale;
  void *e;
  unsigned long flags;
} b200soxr_io_spec_t;

/* Layout-compatible with soxr_quality_spec_t (soxr.h:239-246). */
typedef struct b200soxr_quality_spec {
  double precision, phase_response, passband_end, stopband_begin;
  void *e;
  unsigned long flags;
} b200soxr_quality_spec_t;

/* Layout-compatible with soxr_runtime_spec_t (soxr.h:258-265).
 * Extension: `e` may carry a cudaStream_t on which all device work of the
 * handle is enqueued (NULL => a private non-blocking stream). */
typedef struct b200soxr_runtime_spec {
  unsigned log2_min_dft_size, log2_large_dft_size, coef_size_kbytes, num_threads;
  void *e;
  unsigned long flags;
} b200soxr_runtime_spec_t;

/* ---- spec constructors: soxr_quality_spec / soxr_io_spec / soxr_runtime_spec
 *      (libsoxr/src/soxr.c:106-197, declared soxr.h:273-313) ------------------ */
B200SOXR_API b200soxr_quality_spec_t b200soxr_quality_spec(unsigned long recipe, unsigned long flags);
B200SOXR_API b200soxr_io_spec_t b200soxr_io_spec(int itype, int otype);
B200SOXR_API b200soxr_runtime_spec_t b200soxr_runtime_spec(unsigned num_threads);

/* ---- version strings: soxr_version (soxr.c:43-46) and the binding's
 *      libsoxr_version() (src/csoxr_version.cpp) ------------------------------ */
B200SOXR_API char const *b200soxr_version(void);

/* ---- stream handle API: soxr_create / soxr_process / soxr_delay /
 *      soxr_num_clips / soxr_engine / soxr_clear / soxr_delete
 *      (libsoxr/src/soxr.h:101-179; soxr.c:367-486,761-813,519-523,156-159,149-152,
 *       609-628,601-605).  `in`/`out` are HOST pointers unless the handle was
 *      created with B200SOXR_RT_DEVICE_IO in runtime_spec.flags. ---------------- */
#define B200SOXR_RT_DEVICE_IO 0x100ul   /* extension: in/out of process() are device pointers */
/* extension: low-latency streaming.  The reference returns output in bursts of one FFT block (about 1000 samples,
 * libsoxr/README:57-61); with this flag every call returns ALL output whose filter support has arrived (latency =
 * half the filter length, about 3 ms at 48 kHz HQ), computed by the direct-form kernels.  Results stay within the
 * parity budget of the reference but are not bit-identical to the default mode; b200soxr_delay() follows
 * cr.c:578-581 with this mode's own output count. */
#define B200SOXR_RT_LOW_LATENCY 0x200ul

B200SOXR_API b200soxr_t b200soxr_create(double input_rate, double output_rate, unsigned num_channels,
                                        b200soxr_error_t *error, b200soxr_io_spec_t const *io_spec,
                                        b200soxr_quality_spec_t const *quality_spec,
                                        b200soxr_runtime_spec_t const *runtime_spec);
/* extension: ONE handle for `num_streams` independent streams of `num_channels` channels each that are fed in
 * lockstep (same ilen every call) -- what a server running many ResampleStreams would otherwise do with
 * num_streams soxr_t handles and num_streams soxr_process calls (src/soxr_ext.cpp:129-188 per stream).  Every
 * b200soxr_process call on it takes `in` as [stream][ilen][channel] and writes `out` packed as
 * [stream][*odone][channel] (olen is the per-stream capacity; all streams return the same *odone); results,
 * per-call counts, delay and dither are per stream exactly those of num_streams separate handles. */
B200SOXR_API b200soxr_t b200soxr_create_batch(double input_rate, double output_rate, unsigned num_streams,
                                              unsigned num_channels, b200soxr_error_t *error,
                                              b200soxr_io_spec_t const *io_spec,
                                              b200soxr_quality_spec_t const *quality_spec,
                                              b200soxr_runtime_spec_t const *runtime_spec);
/* in == NULL  => end of input; flush (soxr.h:131-133).  For split (*_S) types
 * in/out are arrays of per-channel pointers (soxr.h:81-87). */
B200SOXR_API b200soxr_error_t b200soxr_process(b200soxr_t p, void const *in, size_t ilen, size_t *idone,
                                               void *out, size_t olen, size_t *odone);
B200SOXR_API double b200soxr_delay(b200soxr_t p);
B200SOXR_API size_t *b200soxr_num_clips(b200soxr_t p);
B200SOXR_API char const *b200soxr_engine(b200soxr_t p);
B200SOXR_API b200soxr_error_t b200soxr_clear(b200soxr_t p);

/* Replaces soxr_set_io_ratio (libsoxr/src/soxr.h:170-172, soxr.c:579-598; bound at src/soxr_ext.cpp:200-203).
 * A handle created with quality_spec.flags & B200SOXR_VR (its create-time in/out rates give the MAXIMUM I/O ratio,
 * vr32.c:618-627) takes the new ratio from the next output on, or glides to it over `slew_len` output samples.
 * Any other handle accepts only its own ratio, with the reference's error string otherwise. */
B200SOXR_API b200soxr_error_t b200soxr_set_io_ratio(b200soxr_t p, double io_ratio, size_t slew_len);
B200SOXR_API void b200soxr_delete(b200soxr_t p);

/* ---- one-shot: soxr_oneshot (soxr.h:188-198, soxr.c:816-838) --------------- */
B200SOXR_API b200soxr_error_t b200soxr_oneshot(double input_rate, double output_rate, unsigned num_channels,
                                               void const *in, size_t ilen, size_t *idone,
                                               void *out, size_t olen, size_t *odone,
                                               b200soxr_io_spec_t const *io_spec,
                                               b200soxr_quality_spec_t const *quality_spec,
                                               b200soxr_runtime_spec_t const *runtime_spec);

/* ---- batch entry point (B200 extension).  Replaces the per-clip driver loops of
 *      the nanobind binding -- csoxr_divide_proc / csoxr_split_ch / csoxr_oneshot
 *      (src/soxr_ext.cpp:210-402) -- for `n_clips` independent clips of equal
 *      length in ONE launch sequence.
 *
 *      Clip b, frame k, channel c lives at
 *          in  + (b*in_clip_stride  + k*in_frame_stride  + c*in_chan_stride ) elements
 *          out + (b*out_clip_stride + n*out_frame_stride + c*out_chan_stride) elements
 *      so both `[frame, channel]` (frame_stride=ch, chan_stride=1) and
 *      split-channel / Fortran order (frame_stride=1, chan_stride=len) are covered.
 *      `dtype` is one of B200SOXR_{FLOAT32,FLOAT64,INT32,INT16} for both sides.
 *      *odone receives frames written per clip = floor(ilen*out/in + .5) (cr.c:544).
 *      flags: B200SOXR_BATCH_DEVICE_PTRS  in/out are device pointers (no copies)
 *             B200SOXR_BATCH_NO_DITHER    == io_spec.flags SOXR_NO_DITHER
 *             B200SOXR_BATCH_DOUBLE       force the f64 engine (SOXR_DOUBLE_PRECISION)
 *             B200SOXR_BATCH_EXACT_GRID   never use the single-pass frequency-domain kernel: results are
 *                                         bit-identical to b200soxr_process on the same input
 *             B200SOXR_BATCH_EMULATE_CALLS int16 dither: replay the div_len call
 *                                          partition of csoxr_divide_proc (soxr_ext.cpp:239-257)
 *      `cuda_stream` is a cudaStream_t.  Host pointers: NULL => a private stream of the
 *      library; the call is synchronous (copies included).  Device pointers: the work
 *      is enqueued on `cuda_stream` as given (NULL => CUDA's legacy default stream) and
 *      the call returns without waiting -- except that an integer batch with
 *      `clips` != NULL waits for its clip count (soxr_num_clips' figure, summed over
 *      the clips); pass NULL to keep it asynchronous.                             */
#define B200SOXR_BATCH_DEVICE_PTRS   1u
#define B200SOXR_BATCH_NO_DITHER     2u
#define B200SOXR_BATCH_DOUBLE        4u
#define B200SOXR_BATCH_EMULATE_CALLS 8u
#define B200SOXR_BATCH_ONESHOT_CALLS 16u  /* dither partition of a single soxr_oneshot call */
#define B200SOXR_BATCH_SPLIT_SEMANTICS 32u /* int16 dither salted per channel, as for *_S types (soxr.c:664-677) */
#define B200SOXR_BATCH_EXACT_GRID 64u /* always use the reference's block grid (bit-identical to a stream of the same input) */

typedef struct b200soxr_batch_desc {
  double in_rate, out_rate;
  unsigned long quality;          /* B200SOXR_QQ .. B200SOXR_VHQ */
  int dtype;                      /* B200SOXR_FLOAT32 ... (no SPLIT bit: use strides) */
  unsigned n_clips, channels;
  size_t ilen;                    /* frames per clip */
  long long in_clip_stride, in_frame_stride, in_chan_stride;     /* in elements */
  long long out_clip_stride, out_frame_stride, out_chan_stride;  /* in elements */
  size_t olen;                    /* capacity in frames per clip of `out` */
  unsigned flags;
} b200soxr_batch_desc_t;

B200SOXR_API b200soxr_error_t b200soxr_resample_batch(b200soxr_batch_desc_t const *desc,
                                                      void const *in, void *out, size_t *odone,
                                                      size_t *clips, void *cuda_stream);

/* frames produced for `ilen` input frames: (size_t)(ilen / io_ratio + .5), cr.c:544 */
B200SOXR_API size_t b200soxr_output_frames(double input_rate, double output_rate, size_t ilen);

/* ---- introspection (tests / tracing; equivalent of SOXR_TRACE=1, soxr.c:381-397,
 *      line formats cr.c:232-233,440-441,484-485).  Writes the plan lines for the
 *      given conversion into buf (NUL-terminated, '\n'-separated); returns bytes
 *      needed.  Pure host code: works without a GPU. ---------------------------- */
B200SOXR_API size_t b200soxr_plan_dump(double input_rate, double output_rate, unsigned long quality,
                                       unsigned long q_flags, char *buf, size_t buf_size);

/* Which kernels serve a conversion, in one line -- e.g. "b200-cr32 | exact grid (resample, streams): k_fft_stage[dft
 * N=4096 2/1 + vpoly0 fused] | resample_batch default: k_fdr[2560 -> 2352 points, variant 1]" -- so that a plan that
 * falls to the direct-form kernels is visible (the role of SOXR_TRACE's engine line, soxr.c:381-397).  `dtype` is
 * the I/O type (it decides the arithmetic type).  Needs a CUDA device (the plan is built and cached).  Returns the
 * bytes needed. */
B200SOXR_API size_t b200soxr_describe_path(double input_rate, double output_rate, unsigned long quality, int dtype,
                                           char *buf, size_t buf_size);

/* Pure-host emulation of the variable-rate engine's control flow (vr32.c:340-570): step i optionally sets a new
 * ratio (io_ratio[i] > 0, over slew_len[i] output samples), appends ilen[i] frames, optionally starts the flush
 * (flush[i] != 0) and asks for olen[i] frames; odone[i] = frames the reference's soxr_process call would return.
 * Works without a GPU. */
B200SOXR_API b200soxr_error_t b200soxr_vr_emulate(double max_io_ratio, size_t n_steps, double const *io_ratio,
                                                  size_t const *slew_len, size_t const *ilen,
                                                  unsigned char const *flush, size_t const *olen, size_t *odone);

/* Pure-host emulation of the reference's pull scheduler (cr.c:498-581): number of
 * output frames each soxr_process call of the given chunk partition returns.
 * chunk_lens[i] == 0 means a flush call (in == NULL).  Works without a GPU. */
B200SOXR_API b200soxr_error_t b200soxr_emulate_calls(double input_rate, double output_rate,
                                                     unsigned long quality, unsigned long q_flags,
                                                     size_t const *chunk_lens, size_t n_chunks,
                                                     size_t olen_each, size_t *odone_each, double *delay_each);

/* ---- K7 in isolation: planar engine-precision samples -> interleaved int16/int32 with
 *      libsoxr's rounding, saturation, clip counting and TPDF dither
 *      (libsoxr/src/data-io.c:200-223, rint-clip.h).  `planar` is HOST memory,
 *      [channels][n] float (is_double == 0) or double; `call_lens` partitions n
 *      into the frames each reference soxr_process call produced (dither is
 *      re-salted per call); *seed is read and updated (soxr.c:439).  Used by the
 *      parity tests to pin the device K7 bit-for-bit against the reference. */
B200SOXR_API b200soxr_error_t b200soxr_quantize(void const *planar, int is_double, unsigned channels, size_t n,
                                                size_t const *call_lens, size_t n_calls, int out_dtype,
                                                int dither, int split, unsigned long long *seed,
                                                void *out_interleaved, size_t *clips);

/* ---- result buffers in page-locked host memory, for front-ends that hand the caller an array they
 *      allocated themselves: the counterpart of `new T[olen * channels]` + capsule in the reference
 *      binding (src/soxr_ext.cpp:240,268-272).  The device-to-host copy of a result lands in such a
 *      buffer directly (no staging copy on the host).  Buffers are recycled through a bounded cache;
 *      NULL when the budget (B200SOXR_PINNED_RESULT_MB, default 1024; 0 disables) is spent or no
 *      device is usable -- the caller then uses ordinary memory.                              */
B200SOXR_API void *b200soxr_result_alloc(size_t bytes);
B200SOXR_API void b200soxr_result_free(void *p);

/* Number of CUDA kernels this library has launched in this process (bench.py's
 * `gpu_launches`). */
B200SOXR_API unsigned long long b200soxr_launch_count(void);
/* 1 if a CUDA device is usable, else 0 (never falls back to the CPU: compute
 * entry points return an error string instead). */
B200SOXR_API int b200soxr_device_available(void);

#ifdef __cplusplus
}
#endif
#endif /* B200SOXR_H */
