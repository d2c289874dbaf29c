#!/bin/bash
# after the K7 / k_poly_interp / batched-stream staging changes: full tests, cfg3 / cfg4 / misc benches
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/r2t_tests.log
python scripts/bench_configs.py cfg3 cfg4 misc 2>&1 | tee $OUT/r2t_cfg.jsonl | cut -c1-330
