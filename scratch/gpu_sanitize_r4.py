# small invocations of the code paths added in round 2's last session, for compute-sanitizer (memcheck / racecheck):
# k_rint_clip (jump map, PAIR stores, odd channel counts, clipping groups), int16-pair loads / planar stores of k_fdr's
# general body, the staged single chunk and the time-split pipeline of single-clip host calls, pinned result buffers
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200"))
import soxr
from soxr import soxr_ext
rng = np.random.default_rng(4)
for ch in (1, 2, 3, 4):
    for dt, scale in ((np.int16, 30000), (np.int32, 2.0e9)):
        planar = (rng.standard_normal((ch, 2077)) * scale).astype(np.float32)
        planar[0, 17] = 1e12
        got, clips, _ = soxr_ext.quantize(planar, np.dtype(dt).name, call_lens=[333, 0, 1000, 744], seed=123 if dt == np.int16 else None)
        print('k7', ch, np.dtype(dt).name, got.shape, clips)
for ch, n in ((2, 5), (4, 3), (1, 6), (3, 2)):
    xi = np.round(rng.standard_normal((n, 20000, ch) if ch > 1 else (n, 20000)) * 5000).astype(np.int16)
    print('int16 batch', ch, soxr.resample_batch(xi, 48000, 44100, 'HQ').shape, soxr.resample_batch(xi, 96000, 48000, 'HQ').shape)
xi32 = np.round(rng.standard_normal((3, 12000, 2)) * 5e8).astype(np.int32)
print('int32 batch', soxr.resample_batch(xi32, 48000, 44100, 'HQ').shape)
# single clips from host memory: time-split (C order, float), staged single chunk (Fortran order / integer), mono
for frames, ch, dt, q in ((400001, 2, np.float32, 'HQ'), (300017, 2, np.float64, 'VHQ'), (480000, 1, np.float32, 'HQ'), (350003, 3, np.float32, 'MQ'),
                          (300001, 2, np.int16, 'HQ'), (3000001, 2, np.float32, 'HQ')):
    x = rng.standard_normal((frames, ch))
    x = (x * 8000).astype(dt) if np.dtype(dt).kind == 'i' else x.astype(dt)
    y = soxr.resample(x if ch > 1 else x[:, 0], 48000, 44100, q)
    yf = soxr.resample(np.asfortranarray(x), 48000, 44100, q)
    print('single clip', frames, ch, np.dtype(dt).name, q, y.shape, type(y.base).__name__, bool(np.array_equal(y.reshape(yf.shape), yf)) if np.dtype(dt).kind == 'f' else '')
    del y, yf
