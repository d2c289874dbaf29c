#!/bin/bash
# planar float stores and L2 prefetch for the interleaved int16 pairs of the general body; integer device batches
# return without waiting when the caller does not ask for the clip count: tests, then timing against the previous build
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/r4d_int16.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_live_ref.py tests/test_gpu_c_abi.py tests/test_distributed.py -m gpu -x -q -k "int or dither or clip or abi or batch or cfg4 or shard or gather" 2>&1 | tail -3 | tee -a $OUT/r4d_int16.log
LIB=python-soxr_b200/soxr/libb200soxr.so
cp $LIB /tmp/orig.so
for v in prev orig prev orig; do
  if [ $v == orig ]; then cp /tmp/orig.so $LIB; else cp scratch/variants/$v.so $LIB; fi
  python - <<P 2>&1 | tee -a $OUT/r4d_int16.log
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
    out.append("cfg4 %d clips %.4f/%.4f" % ((n,) + t(lambda: soxr.resample_batch(xi, 96000, 48000, 'HQ', out=yi))))
    del xi, yi
for ch, n in ((2, 1024), (4, 512)):
    xs = (torch.randn((n, 48000, ch), device='cuda') * 5000).round().to(torch.int16); ys = torch.empty((n, 44100, ch), device='cuda', dtype=torch.int16)
    out.append("int16 %dch 48k->44.1k HQ %d clips %.4f/%.4f" % ((ch, n) + t(lambda: soxr.resample_batch(xs, 48000, 44100, 'HQ', out=ys))))
    del xs, ys
x = torch.randn((1024, 48000, 2), device='cuda'); y = torch.empty((1024, 44100, 2), device='cuda')
out.append("cfg2 f32 %.4f/%.4f" % t(lambda: soxr.resample_batch(x, 48000, 44100, 'VHQ', out=y)))
print("$v (best/median ms): " + " | ".join(out), flush=True)
P
done
cp /tmp/orig.so $LIB
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_rint_clip|k_dither|k_fdr" -c 8 --csv python - <<'P' 2>/dev/null | grep "k_rint_clip\|k_dither\|k_fdr" | awk -F'","' '{print substr($5,1,60), $NF}' | tee -a $OUT/r4d_int16.log
import sys
sys.path.insert(0, "python-soxr_b200")
import torch, soxr
xs = (torch.randn((1024, 48000, 2), device='cuda') * 5000).round().to(torch.int16); ys = torch.empty((1024, 44100, 2), device='cuda', dtype=torch.int16)
for _ in range(2): soxr.resample_batch(xs, 48000, 44100, 'HQ', out=ys)
torch.cuda.synchronize()
P
