#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/r2i_cfg4.log
for mb in 100000 80 40 20 10; do
echo "scratch MB $mb" | tee -a $OUT/r2i_cfg4.log
B200SOXR_INT_SCRATCH_MB=$mb python scripts/bench_configs.py cfg4 2>&1 | cut -c1-330 | tee -a $OUT/r2i_cfg4.log
done
timeout 600 python -m pytest tests/test_gpu_live_ref.py tests/test_gpu_parity.py -m gpu -x -q -k "cfg4 or int or k7" 2>&1 | tail -3
