#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --durations=8 -k "not ref_suite" 2>&1 | tail -40 | tee $OUT/r2c_tests.log
python scripts/bench_configs.py cfg3 cfg1 2>&1 | tee $OUT/r2c_cfg3.jsonl
