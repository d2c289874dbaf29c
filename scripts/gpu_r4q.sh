#!/bin/bash
# k_fft_stage: halo-correction terms of interior float32 blocks loaded four at a time (one memory latency per four
# terms instead of per term on the critical path of a block that runs alone): whole GPU suite, then cfg1 / cfg3
# against the previous build, alternated
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/r4q_halo_loads.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a $OUT/r4q_halo_loads.log
LIB=python-soxr_b200/soxr/libb200soxr.so
cp $LIB /tmp/orig.so
for v in prev orig prev orig; do
  if [ $v == orig ]; then cp /tmp/orig.so $LIB; else cp scratch/variants/$v.so $LIB; fi
  python scripts/bench_configs.py cfg1 cfg3 2>&1 | grep "^{" | python -c "
import sys, json
out = []
for l in sys.stdin:
    d = json.loads(l)
    if d['cfg'] == 'cfg1': out.append('cfg1 %.3f/%.3f ms' % (d['gpu_ms_best'], d['gpu_ms_median']))
    elif d['cfg'] == 'cfg3': out.append('cfg3 %.2f us/chunk' % d['gpu_us_per_chunk'])
    elif d['cfg'] == 'cfg3-lowlat': out.append('lowlat %.2f' % d['gpu_us_per_chunk'])
    elif d['cfg'].startswith('cfg3-batch'): out.append('%s %.1f us/call' % (d['what'][:11], d['gpu_us_per_call']))
print('$v: ' + ' | '.join(out))" | tee -a $OUT/r4q_halo_loads.log
  python scratch/gpu_stream_probe.py 2>&1 | grep "C ABI\|python API" | tail -2 | sed "s/^/$v: /" | tee -a $OUT/r4q_halo_loads.log
done
cp /tmp/orig.so $LIB
