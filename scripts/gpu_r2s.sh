#!/bin/bash
# ncu --set full of the integer epilogue (k_rint_clip) on cfg4 and of k_poly_interp on an irrational ratio
OUT=gpurun_out; mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:"k_rint_clip|k_dither_seeds" -s 4 -c 2 -f -o $OUT/r2s_k7 python scripts/bench_configs.py cfg4 > $OUT/r2s_k7.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_poly_interp" -s 2 -c 1 -f -o $OUT/r2s_poly python scratch/gpu_misc_one.py 44100 47999.5 VHQ float32 256 > $OUT/r2s_poly.log 2>&1
ls -la $OUT/r2s*
