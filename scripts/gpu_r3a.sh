#!/bin/bash
# mono fast instantiations of k_fdr (MODE 2 / 3): parity tests, mono batches and cfg4 with and without them
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/r3a_mono.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fdr or cfg4" 2>&1 | tail -3 | tee -a $OUT/r3a_mono.log
for nm in 0 0; do
  B200SOXR_FDR_NO_MONO_FAST=$nm python - <<'P' 2>&1 | tee -a $OUT/r3a_mono.log
import os, sys
sys.path.insert(0, "python-soxr_b200")
import torch, soxr
def t(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
out = []
for (ir, orr, q) in [(48000, 44100, 'VHQ'), (48000, 16000, 'HQ'), (44100, 16000, 'HQ'), (16000, 48000, 'HQ')]:
    x = torch.randn((1024, ir), device='cuda'); y = torch.empty((1024, orr), device='cuda')
    out.append("%d->%d mono f32 %.4f" % (ir, orr, t(lambda: soxr.resample_batch(x, ir, orr, q, out=y))))
xi = (torch.randn((512, 96000), device='cuda') * 5000).round().to(torch.int16); yi = torch.empty((512, 48000), device='cuda', dtype=torch.int16)
out.append("cfg4 512 clips %.4f" % t(lambda: soxr.resample_batch(xi, 96000, 48000, 'HQ', out=yi)))
print("NO_MONO_FAST=%s: " % os.environ.get("B200SOXR_FDR_NO_MONO_FAST") + " | ".join(out), flush=True)
P
done
