#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or irr or stream" 2>&1 | tail -3
python scripts/bench_configs.py misc 2>&1 | tee $OUT/r2k_misc.jsonl | cut -c1-250
