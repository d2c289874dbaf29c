#!/usr/bin/env python
"""Secondary measurements: every BASELINE.json config through the public API on one GPU, with the
unmodified reference (oracle/_ref) timed beside it on the host cores.  bench.py stays the headline
(configs[1]); this script feeds the per-config table in DESIGN.md.  One JSON line per config.

  cfg1  soxr.resample 10 s mono f32 48000->44100 HQ                 (host numpy in/out, latency bound)
  cfg2  batch 1024 x 1 s stereo f32 48000->44100 VHQ                (device resident; = bench.py)
  cfg3  ResampleStream 44100->16000, 8 ch f32, 1024-frame chunks    (host numpy chunks)
  cfg4  int16 96000->48000 HQ, 4096 mono clips (512 per GPU at 8)   (device resident batch)
  cfg5  44100->48000 VHQ f64, 64 ch                                 (host numpy + device resident)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import soxr  # noqa: E402
import ref_soxr as R  # noqa: E402  (CPU baseline leg only)

dev = torch.device("cuda", 0)


def best(f, n=7, warm=2):
    for _ in range(warm):
        f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t0)
    return min(ts), sorted(ts)[len(ts) // 2]


def dev_time(f, n=20, warm=3):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    evs = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        f()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e-3 for a, b in evs)
    return ts[0], ts[len(ts) // 2]


def rms(a, b):
    return float(np.sqrt(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2)))


def emit(**kw):
    print(json.dumps(kw), flush=True)


have_ref = R.available()
cores = os.cpu_count() or 1
only = set(sys.argv[1:])


def want(c):
    return not only or c in only


# ---- cfg1 -----------------------------------------------------------------------------------------
if want("cfg1"):
    x = np.random.default_rng(0).standard_normal(480000).astype(np.float32)
    tb, tm = best(lambda: soxr.resample(x, 48000, 44100, 'HQ'))
    y = soxr.resample(x, 48000, 44100, 'HQ')
    line = dict(cfg="cfg1", what="soxr.resample 10 s mono f32 48k->44.1k HQ, host numpy in/out",
                gpu_ms_best=tb * 1e3, gpu_ms_median=tm * 1e3, gpu_msamples_s=x.size / tb / 1e6)
    if have_ref:
        rb, rm = best(lambda: R.resample(x, 48000, 44100, 'HQ'))
        line.update(ref_ms_best=rb * 1e3, ref_msamples_s=x.size / rb / 1e6, ref_threads=1,
                    rms_vs_ref=rms(y, R.resample(x, 48000, 44100, 'HQ')))
    emit(**line)

# ---- cfg2 shape at other qualities / ratios (device resident): the k_fdr families --------------------------
if want("fam"):
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    for (ir, orr, q, ch) in [(48000, 44100, 'VHQ', 2), (48000, 44100, 'HQ', 2), (44100, 48000, 'VHQ', 2), (44100, 48000, 'HQ', 2),
                             (96000, 48000, 'HQ', 2), (48000, 96000, 'HQ', 2), (48000, 44100, 'VHQ', 1), (48000, 44100, 'MQ', 2),
                             (48000, 16000, 'HQ', 2), (16000, 48000, 'HQ', 2), (44100, 16000, 'HQ', 2), (44100, 16000, 'VHQ', 1),
                             (48000, 32000, 'HQ', 2), (32000, 48000, 'HQ', 2),
                             (48000, 8000, 'HQ', 2), (8000, 48000, 'HQ', 2), (48000, 12000, 'HQ', 2), (12000, 48000, 'HQ', 2)]:
        n_in = ir
        shape = (1024, n_in, ch) if ch > 1 else (1024, n_in)
        xd = torch.randn(shape, device=dev, generator=g)
        n_out = soxr.soxr_ext.output_frames(ir, orr, n_in)
        yd = torch.empty((1024, n_out, ch) if ch > 1 else (1024, n_out), device=dev)
        tb, tm = dev_time(lambda: soxr.resample_batch(xd, ir, orr, q, out=yd))
        alg = (xd.numel() + yd.numel()) * 4
        emit(cfg="fam", what="1024 clips x 1 s x %d ch f32 %d->%d %s, device resident" % (ch, ir, orr, q),
             gpu_ms_best=tb * 1e3, gpu_msamples_s=xd.numel() / tb / 1e6, achieved_gbs=alg / tb / 1e9,
             roofline_frac=alg / tb / 1e9 / 6550.7)
        del xd, yd

# ---- plans outside the k_fdr families (exact-grid fused kernel / generic direct-form kernels) -------------
if want("misc"):
    g = torch.Generator(device=dev)
    g.manual_seed(2)
    for (ir, orr, q, ch, n) in [(44100, 32000, 'HQ', 2, 256), (32000, 44100, 'HQ', 2, 256), (44100, 47999.5, 'VHQ', 2, 256),
                                (47999, 44100, 'HQ', 2, 256), (48000, 8000, 'HQ', 2, 256), (8000, 48000, 'HQ', 2, 256),
                                (48000, 44100, 'LQ', 2, 256), (48000, 44100, 'QQ', 2, 256), (44100, 48000, 'VHQ', 2, -256)]:
        f64 = n < 0
        n = abs(n)
        n_in = int(ir)
        xd = torch.randn((n, n_in, ch), device=dev, generator=g, dtype=torch.float64 if f64 else torch.float32)
        n_out = soxr.soxr_ext.output_frames(ir, orr, n_in)
        yd = torch.empty((n, n_out, ch), device=dev, dtype=xd.dtype)
        tb, tm = dev_time(lambda: soxr.resample_batch(xd, ir, orr, q, out=yd), n=5, warm=2)
        alg = (xd.numel() + yd.numel()) * xd.element_size()
        emit(cfg="misc", what="%d clips x 1 s x %d ch %s %s->%s %s, device resident" % (n, ch, str(xd.dtype)[6:], ir, orr, q),
             gpu_ms_best=tb * 1e3, gpu_msamples_s=xd.numel() / tb / 1e6, achieved_gbs=alg / tb / 1e9,
             roofline_frac=alg / tb / 1e9 / 6550.7)
        del xd, yd

# ---- cfg3 -----------------------------------------------------------------------------------------
if want("cfg3"):
    N = 441000
    x = np.random.default_rng(2).standard_normal((N, 8)).astype(np.float32)
    chunks = [x[i:i + 1024] for i in range(0, N, 1024)]

    def run_stream(cls):
        rs = cls(44100, 16000, 8, dtype='float32')
        outs = [rs.resample_chunk(c, last=(i == len(chunks) - 1)) for i, c in enumerate(chunks)]
        return np.concatenate(outs)

    tb, tm = best(lambda: run_stream(soxr.ResampleStream), n=3, warm=1)
    y = run_stream(soxr.ResampleStream)
    line = dict(cfg="cfg3", what="ResampleStream 44.1k->16k HQ 8 ch f32, 431 chunks of 1024 frames, host numpy",
                gpu_ms_total=tb * 1e3, gpu_us_per_chunk=tb / len(chunks) * 1e6, gpu_msamples_s=x.size / tb / 1e6,
                out_frames=int(y.shape[0]))
    if have_ref:
        rb, rm = best(lambda: run_stream(R.RefStream), n=3, warm=1)
        yr = run_stream(R.RefStream)
        line.update(ref_ms_total=rb * 1e3, ref_us_per_chunk=rb / len(chunks) * 1e6, ref_msamples_s=x.size / rb / 1e6,
                    ref_threads=1, rms_vs_ref=rms(y, yr) if y.shape == yr.shape else None)
    emit(**line)

    # low-latency mode of the same stream (B200 extension): every call returns all output whose support has arrived
    def run_ll():
        rs = soxr.ResampleStream(44100, 16000, 8, dtype='float32', low_latency=True)
        return np.concatenate([rs.resample_chunk(c, last=(i == len(chunks) - 1)) for i, c in enumerate(chunks)])
    tb, tm = best(run_ll, n=3, warm=1)
    yl = run_ll()
    rs = soxr.ResampleStream(44100, 16000, 8, dtype='float32', low_latency=True)
    first = [len(rs.resample_chunk(c)) for c in chunks[:4]]
    emit(cfg="cfg3-lowlat", what="same stream, low_latency=True", gpu_us_per_chunk=tb / len(chunks) * 1e6,
         first_call_counts=first, delay_after_4_chunks=rs.delay(), rms_vs_default_mode=rms(yl, y))

    # many streams through one handle (ResampleStreamBatch): B x 8 ch, 1024-frame chunks, one launch per chunk.
    # Chunks are laid out [chunk][stream][frame][channel] so each call's input is contiguous (as it would arrive);
    # "pinned": the chunk arrays live in page-locked memory (torch pin_memory), so the H2D copy is a true DMA.
    for B, pinned in ((16, False), (64, False), (256, False), (256, True)):
        nb = 60
        src = np.random.default_rng(5).standard_normal((nb, B, 1024, 8)).astype(np.float32)
        if pinned:
            keep = torch.from_numpy(src).pin_memory()
            xb = keep.numpy()
        else:
            xb = src

        def run_batch():
            rs = soxr.ResampleStreamBatch(44100, 16000, B, 8, dtype='float32')
            n = 0
            for i in range(nb):
                n += rs.resample_chunk(xb[i], last=(i == nb - 1)).shape[1]
            return n
        tb, tm = best(run_batch, n=3, warm=1)
        line = dict(cfg="cfg3-batch", what="%d streams x 8 ch f32 44.1k->16k HQ through one handle, %d chunks of 1024 frames each, %s host numpy"
                    % (B, nb, "pinned" if pinned else "pageable"),
                    gpu_us_per_call=tb / nb * 1e6, gpu_us_per_stream_chunk=tb / nb / B * 1e6,
                    gpu_stream_chunks_per_s=nb * B / tb, gpu_msamples_s=xb.size / tb / 1e6)
        if have_ref:
            line.update(ref_stream_chunks_per_s_one_thread=len(chunks) / rb, ref_threads=1,
                        speedup_vs_one_ref_thread=(nb * B / tb) / (len(chunks) / rb), host_cores=cores)
        emit(**line)
    # the same 256 streams with device-resident chunks (RT_DEVICE_IO through the C ABI): what the kernels sustain
    import ctypes as C
    L = soxr.soxr_ext._lib
    B, nb = 256, 60
    io = L.b200soxr_io_spec(0, 0); qs = L.b200soxr_quality_spec(4, 0); rt = L.b200soxr_runtime_spec(1)
    rt.flags |= soxr.soxr_ext.RT_DEVICE_IO
    err = C.c_char_p()
    h = L.b200soxr_create_batch(44100.0, 16000.0, B, 8, C.byref(err), C.byref(io), C.byref(qs), C.byref(rt))
    xd = torch.randn((nb, B, 1024, 8), device=dev)
    yd = torch.empty((B, 2048, 8), device=dev)
    od = C.c_size_t(0)

    def run_dev():
        L.b200soxr_clear(h)
        for i in range(nb):
            L.b200soxr_process(h, xd[i].data_ptr(), 1024, None, yd.data_ptr(), 2048, C.byref(od))
        torch.cuda.synchronize()
    tb, tm = best(run_dev, n=3, warm=1)
    L.b200soxr_delete(h)
    emit(cfg="cfg3-batch-device", what="256 streams x 8 ch, device-resident chunks (B200SOXR_RT_DEVICE_IO), %d chunks" % nb,
         gpu_us_per_call=tb / nb * 1e6, gpu_stream_chunks_per_s=nb * B / tb, gpu_msamples_s=xd.numel() / tb / 1e6,
         speedup_vs_one_ref_thread=((nb * B / tb) / (len(chunks) / rb)) if have_ref else None)

# ---- cfg4 (one GPU's share of the 8-GPU config: 512 clips; and the whole 4096 on one GPU) -------------
if want("cfg4"):
    for n_clips in (512, 4096):
        g = torch.Generator(device=dev)
        g.manual_seed(3)
        xd = (torch.randn((n_clips, 96000), device=dev, generator=g) * 5000).round().to(torch.int16)
        yd = torch.empty((n_clips, 48000), dtype=torch.int16, device=dev)
        tb, tm = dev_time(lambda: soxr.resample_batch(xd, 96000, 48000, 'HQ', out=yd))
        alg = xd.numel() * 2 + yd.numel() * 2
        line = dict(cfg="cfg4", what="int16 96k->48k HQ, %d mono clips x 1 s, device resident" % n_clips,
                    gpu_ms_best=tb * 1e3, gpu_ms_median=tm * 1e3, gpu_msamples_s=xd.numel() / tb / 1e6,
                    algorithmic_bytes=alg, achieved_gbs=alg / tb / 1e9)
        if have_ref and n_clips == 512:
            from concurrent.futures import ThreadPoolExecutor
            xs = xd[:64].cpu().numpy()
            with ThreadPoolExecutor(max_workers=cores) as ex:
                f = lambda: list(ex.map(lambda i: R.resample(xs[i % 64], 96000, 48000, 'HQ'), range(256)))
                rb, rm = best(f, n=3, warm=1)
            yr = R.resample(xs[0], 96000, 48000, 'HQ')
            yg = yd[0].cpu().numpy()
            d = np.abs(yg.astype(np.int64) - yr.astype(np.int64))
            line.update(ref_msamples_s=256 * 96000 / rb / 1e6, ref_threads=cores,
                        max_lsb_diff_vs_ref=int(d.max()), pct_equal=float((d == 0).mean() * 100))
        emit(**line)

# ---- cfg5 -----------------------------------------------------------------------------------------
if want("cfg5"):
    x = np.random.default_rng(4).standard_normal((44100, 64))
    tb, tm = best(lambda: soxr.resample(x, 44100, 48000, 'VHQ'))
    y = soxr.resample(x, 44100, 48000, 'VHQ')
    xd = torch.from_numpy(x.T.copy()).to(dev)          # 64 mono "clips" = the same 64 channels, device resident
    yd = torch.empty((64, 48000), dtype=torch.float64, device=dev)
    db, dm = dev_time(lambda: soxr.resample_batch(xd, 44100, 48000, 'VHQ', out=yd))
    alg = x.size * 8 + y.size * 8
    line = dict(cfg="cfg5", what="44.1k->48k VHQ f64, 1 s x 64 ch", gpu_host_ms_best=tb * 1e3,
                gpu_host_msamples_s=x.size / tb / 1e6, gpu_device_ms_best=db * 1e3,
                gpu_device_msamples_s=x.size / db / 1e6, achieved_gbs=alg / db / 1e9)
    if have_ref:
        rb, rm = best(lambda: R.resample(x, 44100, 48000, 'VHQ'), n=3, warm=1)
        yr = R.resample(x, 44100, 48000, 'VHQ')
        line.update(ref_ms_best=rb * 1e3, ref_msamples_s=x.size / rb / 1e6, ref_threads=1, rms_vs_ref=rms(y, yr),
                    max_abs_vs_ref=float(np.abs(y - yr).max()),
                    rms_device_vs_ref=rms(yd.cpu().numpy().T, yr))
    emit(**line)
