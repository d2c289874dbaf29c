#!/bin/bash
# compute-sanitizer over every k_fdr family (tensor-memory constant store), the integer epilogue and the variable-rate interpreter
OUT=gpurun_out; mkdir -p $OUT
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python scratch/gpu_sanitize.py vr > $OUT/r2y_memcheck.log 2>&1; echo "memcheck rc $?" | tee -a $OUT/r2y_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python scratch/gpu_sanitize.py vr > $OUT/r2y_racecheck.log 2>&1; echo "racecheck rc $?" | tee -a $OUT/r2y_racecheck.log
tail -4 $OUT/r2y_memcheck.log $OUT/r2y_racecheck.log
