#!/bin/bash
# last check of the round: smoke(), the whole GPU suite, both bench arms
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/r2_final_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/r2_final_tests.log
python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | tee $OUT/r2_final_bench_ref.json | cut -c1-300
python bench.py 2>&1 | tail -1 | tee $OUT/r2_final_bench.json | cut -c1-3000
