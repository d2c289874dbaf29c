#!/bin/bash
# ncu --set full of k_fdr with the variant library scratch/variants/$1.so in place of the in-tree one
OUT=gpurun_out; mkdir -p $OUT
LIB=python-soxr_b200/soxr/libb200soxr.so
cp $LIB /tmp/orig3.so; cp scratch/variants/$1.so $LIB
ncu --set full --clock-control none --import-source on -k k_fdr -s 3 -c 1 -f -o $OUT/r2q_$1_prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > $OUT/r2q_$1_ncu_full.log 2>&1
cp /tmp/orig3.so $LIB
ls -la $OUT/r2q*
