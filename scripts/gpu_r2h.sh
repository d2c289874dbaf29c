#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/r2h_chunk.log
run() { python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1: kernel avg %.4f min %.4f frac %.3f' % (d['roofline']['kernel_ms_avg'], d['roofline']['kernel_ms_min'], d['roofline']['frac']))" | tee -a $OUT/r2h_chunk.log; }
run default
for cfg in "5 0 1" "8 25 2" "8 25 1" "8 15 2" "8 35 2" "8 25 3" "6 25 2" "11 30 2" "11 20 1" "8 40 3"; do
set -- $cfg
B200SOXR_FDR_CHUNK=$1 B200SOXR_FDR_TAIL_PCT=$2 B200SOXR_FDR_TAIL_CHUNK=$3 run "chunk $1 tail $2 pct by $3"
done
