#!/bin/bash
# full GPU tests + every k_fdr family and the BASELINE configs with the tensor-memory constant store
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/r2r_tests.log
python scripts/bench_configs.py fam cfg4 cfg5 2>&1 | tee $OUT/r2r_fam.jsonl | cut -c1-220
python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee $OUT/r2r_bench.json | cut -c1-400
