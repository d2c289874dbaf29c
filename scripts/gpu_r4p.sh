#!/bin/bash
# pinned staging of single-chunk calls through rings of three slots (bounded page-locked memory): whole GPU suite,
# the single-clip cases again, and a 10-minute stereo clip
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/r4p_rings.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $OUT/r4p_rings.log
python - <<'P' 2>&1 | tee -a $OUT/r4p_rings.log
import os, sys, time
sys.path.insert(0, "python-soxr_b200")
import numpy as np, soxr
def t(f, n=40):
    for _ in range(5): f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort(); return ts[n // 10], ts[n // 2], ts[9 * n // 10]
rng = np.random.default_rng(0)
out = []
for secs, ch, dt, q, n in ((10, 1, np.float32, 'HQ', 100), (10, 2, np.int16, 'HQ', 100), (5, 2, np.float32, 'HQ', 100), (30, 2, np.int16, 'HQ', 40), (2.5, 2, np.float32, 'VHQ', 100), (60, 2, np.float32, 'HQ', 20), (600, 2, np.float32, 'HQ', 6)):
    fr = int(48000 * secs)
    x = rng.standard_normal((fr, ch) if ch > 1 else fr)
    x = (x * 5000).astype(dt) if dt == np.int16 else x.astype(dt)
    out.append("%gs x %dch %s %s %.3f/%.3f/%.3f" % ((secs, ch, np.dtype(dt).name, q) + t(lambda: soxr.resample(x, 48000, 44100, q), n)))
print("(p10/p50/p90 ms): " + " | ".join(out), flush=True)
P
