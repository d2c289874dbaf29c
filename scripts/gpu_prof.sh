#!/bin/bash
# Run under gpurun: accuracy spot-check, kernel-only bench, one ncu --set full capture of the dominant kernel.
# usage: scripts/gpu_prof.sh <tag> [noprof]
TAG=${1:-p}
OUT=gpurun_out
mkdir -p $OUT
python scratch/gpu_check.py 2>&1 | tee $OUT/${TAG}_check.log
python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | tee $OUT/${TAG}_bench.json
if [ "$2" != "noprof" ]; then
ncu --set full --clock-control none --import-source on -k regex:"k_fdr|k_fft_stage" -s 3 -c 2 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > $OUT/${TAG}_ncu_full.log 2>&1
fi
