#!/bin/bash
# e2e (host buffers through the C ABI) against the chunk size of the H2D -> compute -> D2H pipeline
OUT=gpurun_out; mkdir -p $OUT
for mb in 2 4 8 16 32 64; do
  B200SOXR_CHUNK_MB=$mb python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunk_mb $mb', 'e2e ms/step %.3f  (%.0f Msamples/s)' % (d['e2e']['ms_per_step'], d['e2e']['value']))" | tee -a $OUT/chunks.log
done
