#!/bin/bash
# integer epilogue k_rint_clip: LCG jump map built once per thread (threads walk several signals), range pre-check
# in the source precision; A/B against the previous library on cfg4 (512 and 4096 clips), then the integer tests
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/r4a_k7.log
LIB=python-soxr_b200/soxr/libb200soxr.so
cp $LIB /tmp/orig.so
for v in k7_old k7_new k7_new_nobound k7_old k7_new; do
  cp scratch/variants/$v.so $LIB
  python - <<P 2>&1 | tee -a $OUT/r4a_k7.log
import sys
sys.path.insert(0, "python-soxr_b200")
import torch, soxr
def t(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[0], ts[len(ts) // 2]
out = []
for n in (512, 4096):
    xi = (torch.randn((n, 96000), device='cuda') * 5000).round().to(torch.int16); yi = torch.empty((n, 48000), device='cuda', dtype=torch.int16)
    out.append("cfg4 %d clips best %.4f median %.4f" % ((n,) + t(lambda: soxr.resample_batch(xi, 96000, 48000, 'HQ', out=yi))))
    del xi, yi
xs = (torch.randn((1024, 48000, 2), device='cuda') * 5000).round().to(torch.int16); ys = torch.empty((1024, 44100, 2), device='cuda', dtype=torch.int16)
out.append("int16 stereo 48k->44.1k HQ 1024 clips best %.4f median %.4f" % t(lambda: soxr.resample_batch(xs, 48000, 44100, 'HQ', out=ys)))
print("$v: " + " | ".join(out), flush=True)
P
done
for v in k7_old k7_new; do
  cp scratch/variants/$v.so $LIB
  echo "ncu $v" | tee -a $OUT/r4a_k7.log
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_rint_clip|k_dither|k_fdr" -s 6 -c 3 --csv python scripts/bench_configs.py cfg4 2>/dev/null | grep "k_rint_clip\|k_dither\|k_fdr" | awk -F'","' '{print substr($5,1,40), $NF}' | tee -a $OUT/r4a_k7.log
done
cp /tmp/orig.so $LIB
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_live_ref.py tests/test_gpu_c_abi.py -m gpu -x -q -k "k7 or int or dither or clip or abi" 2>&1 | tail -3 | tee -a $OUT/r4a_k7.log
