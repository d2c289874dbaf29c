"""k_fdr on small batches (strong-scaling shards): ticket chunk policies, per-call CUDA-event times"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200"))
    import torch, soxr
    out = []
    for n in (64, 128, 256, 512):
        xd = torch.randn((n, 48000, 2), device='cuda'); yd = torch.empty((n, 44100, 2), device='cuda')
        for _ in range(5): soxr.resample_batch(xd, 48000, 44100, 'VHQ', out=yd)
        torch.cuda.synchronize()
        ts = []
        for _ in range(30):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); soxr.resample_batch(xd, 48000, 44100, 'VHQ', out=yd); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        out.append("%d clips %.4f" % (n, ts[len(ts) // 2]))
    print(os.environ.get('TAG'), ' | '.join(out), flush=True)
else:
    for tag, env in [("default", {}), ("serial head", {"B200SOXR_SERIAL_HEAD": "1"}), ("default", {}), ("serial head", {"B200SOXR_SERIAL_HEAD": "1"})]:
        subprocess.run([sys.executable, __file__, 'child'], env=dict(os.environ, TAG=tag, **env))
