#!/bin/bash
# host-buffer pipeline of resample_batch: chunk sizes ramping up / down at the ends (B200SOXR_CHUNK_RAMP=1, default)
# against uniform chunks, pinned and pageable e2e of the headline batch, alternated
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/r4k_chunk_ramp.log
for ramp in 0 1 0 1 0 1; do
  B200SOXR_CHUNK_RAMP=$ramp python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json, sys
d = json.loads(sys.stdin.read())
e = d['e2e']
print('ramp=$ramp  e2e pinned %.3f ms  pageable %.3f ms  kernel %.4f ms' % (e['ms_per_step'], e.get('pageable', {}).get('ms_per_step', float('nan')), d['roofline']['kernel_ms_avg']))" | tee -a $OUT/r4k_chunk_ramp.log
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_live_ref.py -m gpu -x -q -k "batch or pageable or pinned or chunk" 2>&1 | tail -2 | tee -a $OUT/r4k_chunk_ramp.log
