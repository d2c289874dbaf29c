#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT

B200SOXR_TRACE_TIMING=1 python scratch/gpu_stream_probe.py 2>&1 | tee $OUT/r2d_probe.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/r2d_stream_launches.csv python scratch/gpu_stream_probe.py 20 > /dev/null 2>&1
tail -12 $OUT/r2d_stream_launches.csv | cut -d, -f5,12-

