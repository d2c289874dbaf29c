#!/bin/bash
# chunk-size sweep of k_fdr's ticket counter, re-run on the kernel with the tensor-memory constant store
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/r2x_chunk.log
run() {
  env "$@" python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$*: kernel avg %.4f min %.4f frac %.3f' % (d['roofline']['kernel_ms_avg'], d['roofline']['kernel_ms_min'], d['roofline']['frac']))" | tee -a $OUT/r2x_chunk.log
}
run X=default
run B200SOXR_FDR_CHUNK=4 B200SOXR_FDR_TAIL_PCT=25 B200SOXR_FDR_TAIL_CHUNK=2
run B200SOXR_FDR_CHUNK=6 B200SOXR_FDR_TAIL_PCT=25 B200SOXR_FDR_TAIL_CHUNK=2
run B200SOXR_FDR_CHUNK=8 B200SOXR_FDR_TAIL_PCT=15 B200SOXR_FDR_TAIL_CHUNK=2
run B200SOXR_FDR_CHUNK=8 B200SOXR_FDR_TAIL_PCT=15 B200SOXR_FDR_TAIL_CHUNK=1
run B200SOXR_FDR_CHUNK=8 B200SOXR_FDR_TAIL_PCT=10 B200SOXR_FDR_TAIL_CHUNK=2
run B200SOXR_FDR_CHUNK=12 B200SOXR_FDR_TAIL_PCT=20 B200SOXR_FDR_TAIL_CHUNK=2
run B200SOXR_FDR_CHUNK=16 B200SOXR_FDR_TAIL_PCT=25 B200SOXR_FDR_TAIL_CHUNK=2
run B200SOXR_FDR_STATIC_SPLIT=1
run X=default
