#!/bin/bash
# A/B of the tensor-memory constant store (scratch/variants/*.so), then the k_fdr parity tests on the tmem variant
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/ab.log
LIB=python-soxr_b200/soxr/libb200soxr.so
scripts/gpu_ab.sh check
cp $OUT/ab.log $OUT/r2p_ab.log
cp $LIB /tmp/orig2.so; cp scratch/variants/tmem.so $LIB
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 | tee $OUT/r2p_tests.log
cp /tmp/orig2.so $LIB
