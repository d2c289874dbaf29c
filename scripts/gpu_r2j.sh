#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
python scripts/bench_configs.py fam misc cfg5 2>&1 | tee $OUT/r2j_fam_misc.jsonl | cut -c1-250
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 12 --csv --log-file $OUT/r2j_cfg4_launches.csv python scripts/bench_configs.py cfg4 > /dev/null 2>&1
cut -d, -f5,7,8,12- $OUT/r2j_cfg4_launches.csv | tail -12
