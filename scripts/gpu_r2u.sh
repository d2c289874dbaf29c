#!/bin/bash
# integer epilogue after the conversion-free arithmetic and the by-value block list: cfg4 step time, k_rint_clip's own duration
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/r2u_k7.log
python scripts/bench_configs.py cfg4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['what'][:40], 'best %.4f median %.4f ms' % (d['gpu_ms_best'], d.get('gpu_ms_median', 0)), {k: d[k] for k in d if 'lsb' in k or 'equal' in k})" | tee -a $OUT/r2u_k7.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_rint_clip|k_dither|k_fdr" -s 6 -c 3 --csv python scripts/bench_configs.py cfg4 2>/dev/null | grep "k_rint_clip\|k_dither\|k_fdr" | awk -F'","' '{print substr($5,1,40), $NF}' | tee -a $OUT/r2u_k7.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_live_ref.py -m gpu -x -q -k "k7 or int or dither or clip" 2>&1 | tail -3 | tee -a $OUT/r2u_k7.log
