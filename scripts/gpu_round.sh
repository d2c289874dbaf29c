#!/bin/bash
# Run under gpurun: GPU tests, bench (both arms), ncu launch list and one --set full capture of the fused kernel.
# usage: scripts/gpu_round.sh <tag> [notest]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L; nproc; lscpu | grep "Model name"
if [ "$2" != "notest" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_tests.log
fi
python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | tee $OUT/${TAG}_bench_ref.json
python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee $OUT/${TAG}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > $OUT/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_fdr|k_fft_stage" -s 3 -c 2 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -12
