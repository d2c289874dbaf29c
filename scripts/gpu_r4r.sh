#!/bin/bash
# where does a k_fft_stage launch that runs alone (one stream call: 4 channel pairs x 1 block) spend its time?
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/stream_one.py <<'P'
import sys
sys.path.insert(0, "python-soxr_b200")
import numpy as np, soxr
x = np.random.default_rng(0).standard_normal((1024 * 80, 8)).astype(np.float32)
rs = soxr.ResampleStream(44100, 16000, 8, dtype='float32', quality='HQ')
for i in range(80):
    rs.resample_chunk(x[i * 1024:(i + 1) * 1024])
P
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_fft_stage -c 40 --csv python /tmp/stream_one.py 2>/dev/null | grep k_fft_stage | awk -F'","' '{print $(NF-4), $(NF-3), $NF}' | sort | uniq -c | sort -rn | head -8 | tee $OUT/r4r_stream_kernel.log
ncu --set full --clock-control none --import-source on -k regex:k_fft_stage -s 20 -c 2 -f -o $OUT/r4r_stream_prof python /tmp/stream_one.py > $OUT/r4r_ncu_full.log 2>&1
ls -la $OUT/r4r_stream_prof.ncu-rep
