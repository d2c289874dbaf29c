import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200"))
import soxr
for B in (16, 64, 256):
    xb = np.random.default_rng(5).standard_normal((30, B, 1024, 8)).astype(np.float32)
    for rep in range(2):
        rb = soxr.ResampleStreamBatch(44100, 16000, B, 8, dtype='float32')
        ts = []
        for i in range(30):
            t0 = time.perf_counter()
            y = rb.resample_chunk(xb[i], last=(i == 29))
            ts.append((time.perf_counter() - t0) * 1e6)
        print("B=%d rep %d per-call us:" % (B, rep), " ".join("%.0f" % t for t in ts))
        del rb
