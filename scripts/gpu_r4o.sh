#!/bin/bash
# every BASELINE config through the public API with the end-of-round build (feeds the per-config table of DESIGN.md)
OUT=gpurun_out; mkdir -p $OUT
python scripts/bench_configs.py cfg1 cfg3 cfg4 cfg5 fam misc 2>&1 | grep "^{" | tee $OUT/r4o_configs.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items()})"
