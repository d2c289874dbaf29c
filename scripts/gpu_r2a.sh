#!/bin/bash
# round-2 call A: GPU tests on the new default, A/B of k_fdr variants, ncu of the new default
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L; nproc
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/r2a_tests.log
rm -f $OUT/ab.log
bash scripts/gpu_ab.sh check
cp $OUT/ab.log $OUT/r2a_ab.log
ncu --set full --clock-control none --import-source on -k regex:"k_fdr" -s 3 -c 2 -f -o $OUT/r2a_prof \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > $OUT/r2a_ncu_full.log 2>&1
ls -la $OUT | tail -5
