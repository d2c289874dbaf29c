#!/bin/bash
# Run under gpurun: kernel-only bench of every scratch/variants/*.so (copied over the in-tree library in the box's scratch copy).
OUT=gpurun_out; mkdir -p $OUT
LIB=python-soxr_b200/soxr/libb200soxr.so
cp $LIB /tmp/orig.so
for v in scratch/variants/*.so; do
  n=$(basename $v .so)
  cp $v $LIB
  python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import json,sys
l=sys.stdin.read()
try:
    d=json.loads(l); print('$n', 'ms/step %.4f  kernel avg %.4f min %.4f  frac %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['roofline']['kernel_ms_min'], d['roofline']['frac']))
except Exception as e: print('$n', 'FAILED', l[-300:])
" | tee -a $OUT/ab.log
  if [ "$1" == "check" ]; then python scratch/gpu_check.py 2>&1 | sed -n 2p | tee -a $OUT/ab.log; fi
done
cp /tmp/orig.so $LIB
