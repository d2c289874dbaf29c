#!/bin/bash
# results of the whole-signal drivers in recycled page-locked buffers of the library (B200SOXR_PINNED_RESULT_MB=1024,
# default) against ordinary ndarrays (=0): whole GPU suite, then the single-clip cases of r4h
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/r4l_pinned_result.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $OUT/r4l_pinned_result.log
for order in 0 1; do for mb in 0 1024; do
  ORDER=$order B200SOXR_PINNED_RESULT_MB=$mb python - <<'P' 2>&1 | tee -a $OUT/r4l_pinned_result.log
import os, sys, time
sys.path.insert(0, "python-soxr_b200")
import numpy as np, soxr
def t(f, n=100):
    for _ in range(10): f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort(); return ts[n // 10], ts[n // 2], ts[9 * n // 10]
rng = np.random.default_rng(0)
cases = [(10, 1, np.float32, 'HQ'), (10, 2, np.int16, 'HQ'), (5, 2, np.float32, 'HQ'), (30, 2, np.int16, 'HQ'), (2.5, 2, np.float32, 'VHQ')]
if os.environ.get("ORDER") == "1": cases = cases[::-1]
out = []
for secs, ch, dt, q in cases:
    n = int(48000 * secs)
    x = rng.standard_normal((n, ch) if ch > 1 else n)
    x = (x * 5000).astype(dt) if dt == np.int16 else x.astype(dt)
    out.append("%gs x %dch %s %s %.3f/%.3f/%.3f" % ((secs, ch, np.dtype(dt).name, q) + t(lambda: soxr.resample(x, 48000, 44100, q))))
print("PINNED_RESULT_MB=%s order %s (p10/p50/p90 ms): " % (os.environ.get("B200SOXR_PINNED_RESULT_MB"), os.environ.get("ORDER")) + " | ".join(out), flush=True)
P
done; done
