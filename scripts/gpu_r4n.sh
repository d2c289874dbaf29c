#!/bin/bash
# resample_batch(numpy, out=None) returns its result in a recycled page-locked buffer: batch tests, then the bench line
# (e2e.pageable_in_result_allocated_by_call beside e2e.pageable)
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_live_ref.py tests/test_distributed.py -m gpu -x -q -k "batch or pageable or pinned or shard" 2>&1 | tail -2 | tee $OUT/r4n_tests.log
python bench.py 2>&1 | tail -1 | tee $OUT/r4n_bench.json | python -c "
import json, sys
d = json.loads(sys.stdin.read()); e = d['e2e']
print('frac %.4f  e2e pinned %.3f  pageable %.3f  pageable-in/alloc-out %.3f ms' % (d['roofline']['frac'], e['ms_per_step'], e['pageable']['ms_per_step'], e['pageable_in_result_allocated_by_call']['ms_per_step']))"
