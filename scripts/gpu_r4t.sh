#!/bin/bash
# result cache keeps at most one large buffer: its tests, the single-clip tests, a 10-minute clip in a loop
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_live_ref.py tests/test_gpu_c_abi.py -m gpu -x -q -k "recycled_pinned or time_split or length_match or divide_match or live or abi" 2>&1 | tail -2 | tee $OUT/r4t_result_cache.log
python - <<'P' 2>&1 | tee -a $OUT/r4t_result_cache.log
import sys, time
sys.path.insert(0, "python-soxr_b200")
import numpy as np, soxr
x = np.random.default_rng(0).standard_normal((48000 * 600, 2)).astype(np.float32)
ts = []
for i in range(8):
    t0 = time.perf_counter(); y = soxr.resample(x, 48000, 44100); ts.append((time.perf_counter() - t0) * 1e3)
print("10 min stereo, y = resample(x) in a loop (ms):", " ".join("%.1f" % t for t in ts))
P
