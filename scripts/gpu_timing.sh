#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
LIB=python-soxr_b200/soxr/libb200soxr.so
cp $LIB /tmp/orig.so; cp scratch/variants/t_timing.so $LIB
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra 2>/dev/null | grep FDRT | tail -740 > $OUT/fdr_timeline.txt
cp /tmp/orig.so $LIB
wc -l $OUT/fdr_timeline.txt
