#!/bin/bash
# bench both arms (full default run) + distributed tests on 2 GPUs when present
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -3; nproc
python bench.py --impl reference --steps 3 --warmup 3 2>&1 | tail -1 | tee $OUT/r2f_bench_ref.json
python bench.py 2>&1 | tail -1 | tee $OUT/r2f_bench.json
