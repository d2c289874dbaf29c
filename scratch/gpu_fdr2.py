"""k_fdr2 (B200SOXR_FDR2=1) against k_fdr on the headline shape: accuracy vs the oracle, equality of shapes, timing."""
import os, sys, subprocess, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch, soxr, soxr_oracle as O
    rng = np.random.default_rng(5)
    for shape in [(3, 20000, 2), (1, 48000, 2), (2, 2500, 2), (5, 100, 2)]:
        x = rng.standard_normal(shape).astype(np.float32)
        y = soxr.resample_batch(x, 48000, 44100, 'VHQ')
        ref = np.stack([O.resample_f64(x[i], 48000, 44100, O.VHQ) for i in range(shape[0])])
        err = y.astype(np.float64) - ref
        print(shape, 'rms %.3e max %.3e first-bad %s' % (np.sqrt(np.mean(err ** 2)), np.abs(err).max(), np.argwhere(np.abs(err) > 1e-4)[:1].tolist()), flush=True)
    xd = torch.randn((1024, 48000, 2), device='cuda'); yd = torch.empty((1024, 44100, 2), device='cuda')
    for _ in range(5): soxr.resample_batch(xd, 48000, 44100, 'VHQ', out=yd)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); soxr.resample_batch(xd, 48000, 44100, 'VHQ', out=yd); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print('1024 clips: min %.4f median %.4f ms' % (ts[0], ts[len(ts) // 2]), flush=True)
    np.save('/tmp/fdr2_%s.npy' % os.environ.get('B200SOXR_FDR2', '0'), yd[:4].cpu().numpy())
else:
    for flag in ('0', '1'):
        print('== B200SOXR_FDR2=%s' % flag, flush=True)
        subprocess.run([sys.executable, __file__, 'child'], env=dict(os.environ, B200SOXR_FDR2=flag))
    a, b = np.load('/tmp/fdr2_0.npy'), np.load('/tmp/fdr2_1.npy')
    print('k_fdr vs k_fdr2 on the same input: rms diff %.3e max %.3e' % (np.sqrt(np.mean((a - b) ** 2)), np.abs(a - b).max()))
