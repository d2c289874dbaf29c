import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from soxr import soxr_ext
import ref_soxr as R
def rms(a): return float(np.sqrt(np.mean(np.asarray(a, np.float64) ** 2)))
for (ir, orr, rec, dt, ch) in [(48000, 96000, 4 | 0x10, 'float32', 1), (96000, 48000, 4 | 0x10, 'float32', 1), (48000, 44100, 4 | 0x10, 'float32', 1),
                           (48000, 96000, 4 | 0x20, 'float64', 1), (44100, 48000, 6 | 0x20, 'float64', 1), (44100, 48000, 6 | 0x30, 'float64', 1),
                           (48000, 44100, 4 | 0x30, 'float32', 2), (48000, 16000, 4 | 0x10, 'float32', 1)]:
    x = np.random.default_rng(1).standard_normal((20000, ch)).astype(dt)
    code = {'float32': 0, 'float64': 1}[dt]
    a = soxr_ext.CSoxr(ir, orr, ch, code, rec)
    b = R.RefStream(ir, orr, ch, dtype=dt, quality=rec)
    ya = getattr(a, 'process_' + dt)(x, True)[:, 0].astype(np.float64)
    yb = b.resample_chunk(x, last=True)[:, 0].astype(np.float64)
    n = min(len(ya), len(yb))
    best = min(((rms(ya[2000 + s:n - 2000 + s] - yb[2000:n - 2000]), s) for s in range(-700, 701)))
    print(ir, orr, hex(rec), dt, ch, 'lens', len(ya), len(yb), 'rms %.3e' % rms(ya[:n] - yb[:n]), 'best shift', best[1], 'rms %.3e' % best[0],
          soxr_ext.plan_dump(ir, orr, rec)[0][:60])
