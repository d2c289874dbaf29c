#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q -k "not ref_suite" 2>&1 | tail -4 | tee $OUT/r2n_tests.log
for v in 1 0; do echo "static split = $v"; B200SOXR_FDR_STATIC_SPLIT=$v python scripts/bench_configs.py misc cfg1 cfg5 2>&1 | cut -c1-230 | grep -E "LQ|float64|cfg1|cfg5|47999.5"; done | tee $OUT/r2n_fftstage_dyn.log
