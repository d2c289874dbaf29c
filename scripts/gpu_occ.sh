#!/bin/bash
# throughput of k_fdr against CTAs per SM (occupancy limited by padding the dynamic shared memory)
OUT=gpurun_out; mkdir -p $OUT
for pad in 0 14 34 70 180; do
  B200SOXR_FDR_PAD_KB=$pad python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pad_kb $pad', 'kernel avg %.4f min %.4f ms' % (d['roofline']['kernel_ms_avg'], d['roofline']['kernel_ms_min']))" | tee -a $OUT/occ.log
done
