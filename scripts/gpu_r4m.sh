#!/bin/bash
# (1) the pinned-result test with its traceback; (2) where a 10 s mono soxr.resample call spends its GPU time
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/r4m_cfg1.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "recycled_pinned" 2>&1 | tail -30 | tee -a $OUT/r4m_cfg1.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv python - <<'P' 2>/dev/null | grep -v "^==" | awk -F'","' 'NR>1 {print substr($5,1,70), $(NF-4), $(NF-3), $NF}' | tee -a $OUT/r4m_cfg1.log
import sys
sys.path.insert(0, "python-soxr_b200")
import numpy as np, soxr
x = np.random.default_rng(0).standard_normal(480000).astype(np.float32)
for _ in range(2): soxr.resample(x, 48000, 44100, 'HQ')
xs = np.random.default_rng(0).standard_normal((240000, 2)).astype(np.float32)
for _ in range(2): soxr.resample(xs, 48000, 44100, 'HQ')
P
B200SOXR_TRACE_TIMING=1 python - <<'P' 2>&1 | tail -5 | tee -a $OUT/r4m_cfg1.log
import sys, time
sys.path.insert(0, "python-soxr_b200")
import numpy as np, soxr
from soxr import soxr_ext as E
x = np.random.default_rng(0).standard_normal(480000).astype(np.float32)
for _ in range(5): soxr.resample(x, 48000, 44100, 'HQ')
x2 = x[:, None]
for name, f in (("resample", lambda: soxr.resample(x, 48000, 44100, 'HQ')), ("divide_proc", lambda: E.csoxr_divide_proc_float32(48000, 44100, x2, 4))):
    ts = []
    for _ in range(50):
        t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort(); print(name, "p10/p50 %.3f/%.3f ms" % (ts[5], ts[25]))
P
