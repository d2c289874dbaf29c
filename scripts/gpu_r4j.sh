#!/bin/bash
# the same with the CALLER polling for the end of a parallel copy before it blocks (and the workers spinning)
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/r4j_pool_spin2.log
for order in 0 1; do for cfg in "0 0" "150 0" "150 2048" "0 8192"; do set -- $cfg; spin=$1; kb=$2
  ORDER=$order B200SOXR_COPY_SPIN_US=$spin B200SOXR_POOL_MIN_KB=$kb python - <<'P' 2>&1 | tee -a $OUT/r4j_pool_spin2.log
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
print("SPIN_US=%s POOL_MIN_KB=%s order %s (p10/p50/p90 ms): " % (os.environ.get("B200SOXR_COPY_SPIN_US"), os.environ.get("B200SOXR_POOL_MIN_KB"), os.environ.get("ORDER")) + " | ".join(out), flush=True)
P
done; done
