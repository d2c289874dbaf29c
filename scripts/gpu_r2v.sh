#!/bin/bash
# batched streams from pageable memory: pinned-slot staging by the host copy pool vs the driver's own staging, same box
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/r2v_stream_staging.log
for mode in 0 1 0 1; do
  B200SOXR_NO_HOST_STAGING=$mode python scripts/bench_configs.py cfg3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    if d['cfg'] == 'cfg3-batch': print('no_staging=$mode', d['what'][:12], d['what'][-20:], 'us/call %.1f' % d['gpu_us_per_call'])" | tee -a $OUT/r2v_stream_staging.log
done
nproc | tee -a $OUT/r2v_stream_staging.log
