#!/bin/bash
# soxr.resample on one pageable ndarray (BASELINE configs[0] and other single-clip sizes): the single chunk staged
# through pinned slots in pieces by the host copy pool, against the driver's own pageable staging
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/r4e_single_clip.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $OUT/r4e_single_clip.log
for kb in 100000000 512 100000000 512 128; do
  B200SOXR_STAGE_ONE_KB=$kb python - <<'P' 2>&1 | tee -a $OUT/r4e_single_clip.log
import os, sys, time
sys.path.insert(0, "python-soxr_b200")
import numpy as np, soxr
def t(f, n=40):
    for _ in range(5): f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort(); return ts[0], ts[len(ts) // 2]
rng = np.random.default_rng(0)
out = []
for secs, ch, dt, q in ((10, 1, np.float32, 'HQ'), (1, 2, np.float32, 'VHQ'), (60, 2, np.float32, 'HQ'), (10, 2, np.int16, 'HQ'), (0.5, 1, np.float32, 'HQ')):
    n = int(48000 * secs)
    x = rng.standard_normal((n, ch) if ch > 1 else n)
    x = (x * 5000).astype(dt) if dt == np.int16 else x.astype(dt)
    out.append("%gs x %dch %s %s %.3f/%.3f" % ((secs, ch, np.dtype(dt).name, q) + t(lambda: soxr.resample(x, 48000, 44100, q))))
print("STAGE_ONE_KB=%s (best/median ms): " % os.environ.get("B200SOXR_STAGE_ONE_KB") + " | ".join(out), flush=True)
P
done
