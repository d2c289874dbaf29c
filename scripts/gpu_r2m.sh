#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
ncu --set full --clock-control none -k regex:"k_poly_interp|k_sig_copy|k_fft_stage" -s 3 -c 3 -f -o $OUT/r2m_irr python scratch/gpu_misc_one.py 44100 47999.5 VHQ float32 256 > /dev/null 2>&1
ls -la $OUT/r2m_*
