#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/ab.log
LIB=python-soxr_b200/soxr/libb200soxr.so
B200SOXR_FDR_STATIC_SPLIT=1 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('static split: kernel avg %.4f min %.4f frac %.3f' % (d['roofline']['kernel_ms_avg'], d['roofline']['kernel_ms_min'], d['roofline']['frac']))" | tee -a $OUT/r2g_ab.log
python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dynamic chunks: kernel avg %.4f min %.4f frac %.3f' % (d['roofline']['kernel_ms_avg'], d['roofline']['kernel_ms_min'], d['roofline']['frac']))" | tee -a $OUT/r2g_ab.log
python scratch/gpu_check.py 2>&1 | tail -10
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fdr or batch" 2>&1 | tail -3
cp $LIB /tmp/orig.so; cp scratch/variants/t_timing.so $LIB
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra 2>/dev/null | grep FDRT | tail -740 > $OUT/fdr_timeline.txt
cp /tmp/orig.so $LIB
