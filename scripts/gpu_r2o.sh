#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
ncu --set full --clock-control none -k k_fdr -s 3 -c 1 -f -o $OUT/r2o_prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > $OUT/r2o_ncu_full.log 2>&1
ls -la $OUT/r2o*
