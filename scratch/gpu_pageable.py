import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200"))
import soxr, torch
x = np.random.default_rng(1).standard_normal((1024, 48000, 2)).astype(np.float32)
y = np.empty((1024, 44100, 2), np.float32)
for _ in range(2): soxr.resample_batch(x, 48000, 44100, 'VHQ', out=y)
ts = []
for _ in range(6):
    t0 = time.perf_counter(); soxr.resample_batch(x, 48000, 44100, 'VHQ', out=y); ts.append(time.perf_counter() - t0)
xp = torch.from_numpy(x).pin_memory().numpy(); yp = torch.empty(y.shape).pin_memory().numpy()
for _ in range(2): soxr.resample_batch(xp, 48000, 44100, 'VHQ', out=yp)
tp = []
for _ in range(6):
    t0 = time.perf_counter(); soxr.resample_batch(xp, 48000, 44100, 'VHQ', out=yp); tp.append(time.perf_counter() - t0)
assert np.array_equal(y, yp)
print("threads %s chunk %s MB: pageable best %.2f ms median %.2f ms; pinned best %.2f ms; ratio %.2f" % (
    os.environ.get("B200SOXR_COPY_THREADS", "default"), os.environ.get("B200SOXR_CHUNK_MB", "16"),
    min(ts) * 1e3, sorted(ts)[3] * 1e3, min(tp) * 1e3, min(ts) / min(tp)))
