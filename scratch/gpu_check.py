import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import soxr
import soxr_oracle as O
Q = {'HQ': O.HQ, 'VHQ': O.VHQ}
rng = np.random.default_rng(5)
cases = [(48000, 44100, 'HQ', (12000,), 'float32'), (48000, 44100, 'VHQ', (20000, 2), 'float32'),
         (44100, 16000, 'HQ', (9000, 8), 'float32'), (96000, 48000, 'HQ', (20000, 2), 'float32'),
         (44100, 48000, 'VHQ', (9000, 4), 'float64'), (48000, 44100, 'VHQ', (9000, 3), 'float64'),
         (32000, 44100, 'HQ', (9000, 1), 'float32'), (48000, 8000, 'HQ', (30000, 2), 'float32'),
         (96000, 48000, 'HQ', (20000, 3), 'float64'), (48000, 96000, 'HQ', (5000, 2), 'float32')]
for ir, orr, q, shape, dt in cases:
    x = rng.standard_normal(shape).astype(dt)
    t = time.time()
    try:
        y = soxr.resample_batch(x[None], ir, orr, q)[0]
    except Exception as e:
        print(ir, orr, q, shape, dt, 'ERROR', e); continue
    dtm = time.time() - t
    ref = O.resample_f64(x, ir, orr, Q[q])
    err = y.astype(np.float64) - ref
    print(ir, orr, q, shape, dt, 'rms %.3e max %.3e  first-bad %s  (%.1f ms)' % (
        np.sqrt(np.mean(err ** 2)), np.abs(err).max(),
        (np.argwhere(np.abs(err) > 1e-4)[:1].tolist()), dtm * 1e3), flush=True)
