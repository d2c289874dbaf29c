/* tests/c_abi/test_abi.c -- the C ABI of libb200soxr.so driven from plain C, the way a python-soxr maintainer's
 * re-pointed binding would call it: no Python shim in the path.
 *
 *   oneshot    b200soxr_oneshot                                     <- soxr_oneshot, soxr.h:188-198 / soxr_ext.cpp:383-389
 *   split      b200soxr_process with *_S pointer arrays, div_len     <- csoxr_split_ch, soxr_ext.cpp:299-346
 *              chunks + a flush call
 *   stream     b200soxr_process, interleaved, 480-frame chunks        <- CSoxr::process, soxr_ext.cpp:129-188
 *   device     the same with B200SOXR_RT_DEVICE_IO (device pointers)
 *   cstream    the same on a caller-supplied cudaStream_t (runtime_spec.e)
 *   batch      b200soxr_create_batch: 3 lockstep streams through one handle
 *   vr         quality_spec.flags |= B200SOXR_VR, 1000-frame chunks, b200soxr_set_io_ratio half-way <- soxr_set_io_ratio,
 *              (to 0.6 of the create-time ratio over 500 output samples), flush                       soxr_ext.cpp:200-203
 *
 * usage: test_abi <mode> <in_rate> <out_rate> <channels> <quality> <in.f32> <out.f32>
 * in.f32 holds interleaved float32 frames; the result is written interleaved (for `batch`, in.f32 holds
 * [stream][frame][channel] and the result is [stream][frame][channel]).  Exit code 0 = every call succeeded;
 * tests/test_gpu_c_abi.py compares the files with the golden expectations.
 * Built by __graft_entry__.build():  gcc tests/c_abi/test_abi.c -Iinclude -lb200soxr -lcudart
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "b200soxr.h"

#define CHECK(err) do { b200soxr_error_t e__ = (err); if (e__) { fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, e__); return 2; } } while (0)
#define CUCHECK(x) do { cudaError_t c__ = (x); if (c__ != cudaSuccess) { fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(c__)); return 3; } } while (0)

static float *read_all(const char *path, size_t *n_floats) {
  FILE *f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  float *p = (float *)malloc(bytes > 0 ? (size_t)bytes : 1);
  if (bytes > 0 && fread(p, 1, (size_t)bytes, f) != (size_t)bytes) { fclose(f); free(p); return NULL; }
  fclose(f);
  *n_floats = (size_t)bytes / sizeof(float);
  return p;
}

static int write_all(const char *path, const float *p, size_t n_floats) {
  FILE *f = fopen(path, "wb");
  if (!f) return 1;
  fwrite(p, sizeof(float), n_floats, f);
  fclose(f);
  return 0;
}

int main(int argc, char **argv) {
  if (argc != 8) { fprintf(stderr, "usage: %s mode in_rate out_rate channels quality in.f32 out.f32\n", argv[0]); return 1; }
  const char *mode = argv[1];
  const double in_rate = atof(argv[2]), out_rate = atof(argv[3]);
  const unsigned ch = (unsigned)atoi(argv[4]);
  const unsigned long quality = strtoul(argv[5], NULL, 10);
  size_t n_floats = 0;
  float *x = read_all(argv[6], &n_floats);
  if (!x) { fprintf(stderr, "cannot read %s\n", argv[6]); return 1; }
  const unsigned n_streams = strcmp(mode, "batch") == 0 ? 3u : 1u;
  const size_t ilen = n_floats / ch / n_streams;
  const size_t olen = (size_t)((double)ilen * out_rate / in_rate + 1) + 16;      /* soxr_ext.cpp:238 (+ slack) */
  float *y = (float *)calloc(olen * ch * n_streams, sizeof(float));
  size_t out_pos = 0, odone = 0;
  b200soxr_quality_spec_t q = b200soxr_quality_spec(quality, 0);

  if (strcmp(mode, "oneshot") == 0) {
    b200soxr_io_spec_t io = b200soxr_io_spec(B200SOXR_FLOAT32_I, B200SOXR_FLOAT32_I);
    size_t idone = 0;
    CHECK(b200soxr_oneshot(in_rate, out_rate, ch, x, ilen, &idone, y, olen, &odone, &io, &q, NULL));
    if (idone != ilen) { fprintf(stderr, "oneshot consumed %zu of %zu frames\n", idone, ilen); return 4; }
    out_pos = odone;
  } else if (strcmp(mode, "split") == 0) {
    /* planar copies of the input / output: channel c at xs + c*ilen (what a Fortran-order ndarray looks like) */
    b200soxr_io_spec_t io = b200soxr_io_spec(B200SOXR_FLOAT32_S, B200SOXR_FLOAT32_S);
    b200soxr_error_t err = NULL;
    b200soxr_t h = b200soxr_create(in_rate, out_rate, ch, &err, &io, &q, NULL);
    CHECK(err);
    float *xs = (float *)malloc((ilen ? ilen : 1) * ch * sizeof(float)), *ys = (float *)calloc(olen * ch, sizeof(float));
    for (size_t k = 0; k < ilen; ++k) for (unsigned c = 0; c < ch; ++c) xs[c * ilen + k] = x[k * ch + c];
    const float **ip = (const float **)malloc(ch * sizeof(*ip));
    float **op = (float **)malloc(ch * sizeof(*op));
    double dl = 48000 * in_rate / out_rate;
    const size_t div_len = (size_t)(dl > 1000. ? dl : 1000.);                     /* soxr_ext.cpp:313 */
    for (size_t idx = 0; idx < ilen; idx += div_len) {
      for (unsigned c = 0; c < ch; ++c) { ip[c] = xs + c * ilen + idx; op[c] = ys + c * olen + out_pos; }
      CHECK(b200soxr_process(h, ip, ilen - idx < div_len ? ilen - idx : div_len, NULL, op, olen - out_pos, &odone));
      out_pos += odone;
    }
    for (unsigned c = 0; c < ch; ++c) op[c] = ys + c * olen + out_pos;
    CHECK(b200soxr_process(h, NULL, 0, NULL, op, olen - out_pos, &odone));
    out_pos += odone;
    for (size_t k = 0; k < out_pos; ++k) for (unsigned c = 0; c < ch; ++c) y[k * ch + c] = ys[c * olen + k];
    if (strstr(b200soxr_engine(h), "b200") == NULL) { fprintf(stderr, "unexpected engine %s\n", b200soxr_engine(h)); return 4; }
    b200soxr_delete(h);
    free(xs); free(ys); free(ip); free(op);
  } else if (strcmp(mode, "stream") == 0 || strcmp(mode, "device") == 0 || strcmp(mode, "cstream") == 0 ||
             strcmp(mode, "batch") == 0) {
    const int device_io = strcmp(mode, "device") == 0;
    b200soxr_io_spec_t io = b200soxr_io_spec(B200SOXR_FLOAT32_I, B200SOXR_FLOAT32_I);
    b200soxr_runtime_spec_t rt = b200soxr_runtime_spec(1);
    cudaStream_t user_stream = NULL;
    if (device_io) rt.flags |= B200SOXR_RT_DEVICE_IO;
    if (device_io || strcmp(mode, "cstream") == 0) {
      CUCHECK(cudaStreamCreateWithFlags(&user_stream, cudaStreamNonBlocking));
      rt.e = (void *)user_stream;                                                   /* all device work of the handle goes here */
    }
    b200soxr_error_t err = NULL;
    b200soxr_t h = n_streams > 1 ? b200soxr_create_batch(in_rate, out_rate, n_streams, ch, &err, &io, &q, &rt)
                                 : b200soxr_create(in_rate, out_rate, ch, &err, &io, &q, &rt);
    CHECK(err);
    const size_t chunk = 480;
    float *dx = NULL, *dy = NULL, *piece = (float *)malloc(olen * ch * n_streams * sizeof(float));
    float *cin = (float *)malloc(chunk * ch * n_streams * sizeof(float));
    if (device_io) {
      CUCHECK(cudaMalloc((void **)&dx, (n_floats ? n_floats : 1) * sizeof(float)));
      CUCHECK(cudaMalloc((void **)&dy, olen * ch * sizeof(float)));
      CUCHECK(cudaMemcpyAsync(dx, x, n_floats * sizeof(float), cudaMemcpyHostToDevice, user_stream));
    }
    double delay_seen = 0;
    size_t idx = 0;
    for (;;) {
      const int flush = idx >= ilen;                                 /* after the last (possibly short) chunk: in == NULL */
      const size_t n = flush ? 0 : (ilen - idx < chunk ? ilen - idx : chunk);
      const void *in_ptr = NULL;
      if (!flush) {
        if (device_io) in_ptr = dx + idx * ch;
        else if (n_streams == 1) in_ptr = x + idx * ch;
        else {                                                       /* [stream][n][channel] for this call */
          for (unsigned b = 0; b < n_streams; ++b) memcpy(cin + b * n * ch, x + (b * ilen + idx) * ch, n * ch * sizeof(float));
          in_ptr = cin;
        }
      }
      void *out_ptr = device_io ? (void *)(dy + out_pos * ch) : (void *)(n_streams == 1 ? y + out_pos * ch : piece);
      CHECK(b200soxr_process(h, in_ptr, n, NULL, out_ptr, olen - out_pos, &odone));
      if (n_streams > 1)                                             /* packed [stream][odone][channel] */
        for (unsigned b = 0; b < n_streams; ++b)
          memcpy(y + (b * olen + out_pos) * ch, piece + b * odone * ch, odone * ch * sizeof(float));
      out_pos += odone;
      if (b200soxr_delay(h) > delay_seen) delay_seen = b200soxr_delay(h);
      if (flush && odone == 0) break;                                /* flush until nothing more comes (soxr_ext.cpp:109-127) */
      idx += n;
    }
    if (device_io) {
      CUCHECK(cudaMemcpyAsync(y, dy, out_pos * ch * sizeof(float), cudaMemcpyDeviceToHost, user_stream));
      CUCHECK(cudaStreamSynchronize(user_stream));
      cudaFree(dx); cudaFree(dy);
    }
    if (*b200soxr_num_clips(h) != 0) { fprintf(stderr, "float I/O cannot clip\n"); return 4; }
    if (ilen > 4096 && !(delay_seen > 0)) { fprintf(stderr, "delay() never reported pending output\n"); return 4; }
    CHECK(b200soxr_clear(h));
    if (b200soxr_delay(h) != 0) { fprintf(stderr, "delay() after clear\n"); return 4; }
    b200soxr_delete(h);
    if (user_stream) cudaStreamDestroy(user_stream);
    if (n_streams > 1) {                                             /* compact [stream][olen] rows to [stream][out_pos] */
      for (unsigned b = 1; b < n_streams; ++b) memmove(y + b * out_pos * ch, y + b * olen * ch, out_pos * ch * sizeof(float));
    }
    free(piece); free(cin);
  } else if (strcmp(mode, "vr") == 0) {
    /* variable rate: the create-time rates give the MAXIMUM I/O ratio; every call offers `room` frames of output */
    b200soxr_io_spec_t io = b200soxr_io_spec(B200SOXR_FLOAT32_I, B200SOXR_FLOAT32_I);
    b200soxr_quality_spec_t qv = b200soxr_quality_spec(quality, B200SOXR_VR);
    b200soxr_error_t err = NULL;
    b200soxr_t h = b200soxr_create(in_rate, out_rate, ch, &err, &io, &qv, NULL);
    CHECK(err);
    if (strcmp(b200soxr_engine(h), "b200-vr32") != 0) { fprintf(stderr, "unexpected engine %s\n", b200soxr_engine(h)); return 4; }
    const size_t chunk = 1000, room = 8192;
    size_t cap = olen * 2 + room;
    float *yv = (float *)calloc(cap * ch, sizeof(float));
    int changed = 0;
    for (size_t idx = 0;; ) {
      const int flush = idx >= ilen;
      const size_t n = flush ? 0 : (ilen - idx < chunk ? ilen - idx : chunk);
      if (!changed && idx >= ilen / 2) {
        CHECK(b200soxr_set_io_ratio(h, 0.6 * in_rate / out_rate, 500));
        changed = 1;
      }
      if (out_pos + room > cap) { cap *= 2; yv = (float *)realloc(yv, cap * ch * sizeof(float)); }
      CHECK(b200soxr_process(h, flush ? NULL : x + idx * ch, n, NULL, yv + out_pos * ch, room, &odone));
      out_pos += odone;
      if (flush && odone == 0) break;
      idx += n;
    }
    if (b200soxr_set_io_ratio(h, -1, 0) == NULL) { fprintf(stderr, "a negative ratio was accepted\n"); return 4; }
    b200soxr_delete(h);
    free(y);
    y = yv;
  } else {
    fprintf(stderr, "unknown mode %s\n", mode);
    return 1;
  }
  if (write_all(argv[7], y, out_pos * ch * n_streams)) { fprintf(stderr, "cannot write %s\n", argv[7]); return 1; }
  printf("%s: %zu frames in, %zu frames out, version %s\n", mode, ilen, out_pos, b200soxr_version());
  free(x); free(y);
  return 0;
}
