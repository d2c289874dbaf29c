#!/bin/bash
# compute-sanitizer memcheck / racecheck over the code paths of the round's last session (scratch/gpu_sanitize_r4.py)
OUT=gpurun_out; mkdir -p $OUT
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 1 python scratch/gpu_sanitize_r4.py > $OUT/r4s_memcheck.log 2>&1; echo "memcheck rc $?" | tee -a $OUT/r4s_memcheck.log
timeout 700 compute-sanitizer --tool racecheck --error-exitcode 1 python scratch/gpu_sanitize_r4.py > $OUT/r4s_racecheck.log 2>&1; echo "racecheck rc $?" | tee -a $OUT/r4s_racecheck.log
tail -n 5 $OUT/r4s_memcheck.log $OUT/r4s_racecheck.log
