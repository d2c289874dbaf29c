#!/bin/bash
# Run under gpurun: GPU tests + bench only (no ncu).  usage: scripts/gpu_quick.sh <tag> [notest]
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
if [ "$2" != "notest" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_tests.log
fi
python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee $OUT/${TAG}_bench.json
