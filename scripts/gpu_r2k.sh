#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q -k "not ref_suite" 2>&1 | tail -6 | tee $OUT/r2k_tests.log
python scripts/bench_configs.py misc 2>&1 | tee $OUT/r2k_misc.jsonl | cut -c1-250
