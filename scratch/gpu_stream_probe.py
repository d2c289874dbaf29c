"""Per-call cost of a small-chunk ResampleStream (cfg3 shape): calls that return output vs calls that do not."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200"))
import soxr
from soxr import soxr_ext
import ctypes as C
N = int(sys.argv[1]) if len(sys.argv) > 1 else 431
x = np.random.default_rng(2).standard_normal((N * 1024, 8)).astype(np.float32)
chunks = [x[i:i + 1024] for i in range(0, len(x), 1024)]
for rep in range(3):
    rs = soxr.ResampleStream(44100, 16000, 8, dtype='float32')
    t_due, t_idle, n_due, n_idle = 0.0, 0.0, 0, 0
    for i, c in enumerate(chunks):
        t0 = time.perf_counter()
        y = rs.resample_chunk(c, last=(i == len(chunks) - 1))
        dt = time.perf_counter() - t0
        if len(y): t_due += dt; n_due += 1
        else: t_idle += dt; n_idle += 1
    print("python API: due calls %d avg %.1f us, idle calls %d avg %.1f us, overall %.1f us/chunk" % (
        n_due, t_due / max(n_due, 1) * 1e6, n_idle, t_idle / max(n_idle, 1) * 1e6, (t_due + t_idle) / len(chunks) * 1e6))
# raw C ABI calls (no numpy allocation per call)
L = soxr_ext._lib
io = L.b200soxr_io_spec(0, 0); qs = L.b200soxr_quality_spec(4, 0)
err = C.c_char_p()
h = L.b200soxr_create(44100.0, 16000.0, 8, C.byref(err), C.byref(io), C.byref(qs), None)
out = np.empty((4096, 8), np.float32); od = C.c_size_t(0)
t_due = t_idle = 0.0; n_due = n_idle = 0
for i, c in enumerate(chunks):
    t0 = time.perf_counter()
    L.b200soxr_process(h, c.ctypes.data, 1024, None, out.ctypes.data, 4096, C.byref(od))
    dt = time.perf_counter() - t0
    if od.value: t_due += dt; n_due += 1
    else: t_idle += dt; n_idle += 1
print("C ABI via ctypes: due %d avg %.1f us, idle %d avg %.1f us, overall %.1f us/chunk" % (
    n_due, t_due / max(n_due, 1) * 1e6, n_idle, t_idle / max(n_idle, 1) * 1e6, (t_due + t_idle) / len(chunks) * 1e6))
L.b200soxr_delete(h)
# batched handle, device-history mode (chunk above the zero-copy size): where does a call's time go?
for B in (16, 64):
    rb = soxr.ResampleStreamBatch(44100, 16000, B, 8, dtype='float32')
    xb = np.random.default_rng(5).standard_normal((40, B, 1024, 8)).astype(np.float32)
    t0 = time.perf_counter()
    for i in range(40):
        rb.resample_chunk(xb[i], last=(i == 39))
    print("batch B=%d: %.1f us per call" % (B, (time.perf_counter() - t0) / 40 * 1e6))
    del rb
import torch
for kb in (64, 512, 2048, 8192):
    a = np.random.default_rng(1).standard_normal(kb * 256).astype(np.float32)
    d = torch.empty(kb * 256, device='cuda')
    pin = torch.empty(kb * 256).pin_memory()
    s = torch.cuda.Stream()
    for name, f in (("pageable cudaMemcpyAsync", lambda: d.copy_(torch.from_numpy(a), non_blocking=True)),
                    ("memcpy->pinned + DMA", lambda: (pin.numpy().__setitem__(slice(None), a), d.copy_(pin, non_blocking=True)))):
        with torch.cuda.stream(s):
            f(); s.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                f()
                s.synchronize()
            print("%5d kB %-26s %.1f us" % (kb, name, (time.perf_counter() - t0) / 20 * 1e6))
