#!/bin/bash
# round-2 call B: full GPU test suite incl. the reference's own suite and the live full-size parity tests
OUT=gpurun_out; mkdir -p $OUT
rm -f $OUT/r2_int_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -40 | tee $OUT/r2b_tests.log
cat $OUT/r2_int_parity.jsonl
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
