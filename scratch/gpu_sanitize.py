# small invocations of every k_fdr family for compute-sanitizer (memcheck / racecheck)
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200"))
import soxr
rng = np.random.default_rng(3)
for (ir, orr, q, ch) in [(48000, 44100, 'VHQ', 2), (48000, 44100, 'VHQ', 1), (44100, 48000, 'VHQ', 2), (96000, 48000, 'HQ', 1),
                         (48000, 16000, 'HQ', 2), (16000, 48000, 'HQ', 3), (44100, 16000, 'HQ', 2),
                         (48000, 32000, 'HQ', 2), (32000, 48000, 'VHQ', 1), (48000, 8000, 'HQ', 2), (48000, 8000, 'VHQ', 1),
                         (8000, 48000, 'HQ', 2), (48000, 12000, 'HQ', 3), (12000, 48000, 'HQ', 1)]:
    x = rng.standard_normal((2, 9000, ch)).astype(np.float32)
    y = soxr.resample_batch(x, ir, orr, q)
    print(ir, orr, q, ch, y.shape, float(np.abs(y).max()))
xi = np.round(rng.standard_normal((3, 20000)) * 5000).astype(np.int16)
print(soxr.resample_batch(xi, 96000, 48000, 'HQ').shape)
rs = soxr.ResampleStream(44100, 16000, 2, dtype='float32')
print(sum(rs.resample_chunk(rng.standard_normal((1024, 2)).astype(np.float32), last=(i == 5)).shape[0] for i in range(6)))
if len(sys.argv) > 1 and sys.argv[1] == 'multi':
    # more items than CTAs: consecutive blocks per CTA (exchange-buffer swap, TMA staging of the next block)
    for (ir, orr, q, ch, n) in [(48000, 44100, 'VHQ', 2, 330), (96000, 48000, 'HQ', 1, 700), (44100, 48000, 'HQ', 2, 300)]:
        x = rng.standard_normal((n, 20000, ch)).astype(np.float32)
        y = soxr.resample_batch(x, ir, orr, q)
        y1 = soxr.resample_batch(x[-1:], ir, orr, q)
        print('multi', ir, orr, q, ch, y.shape, 'last clip equals its single-clip result:', bool(np.all(y[-1] == y1[0])))
# variable-rate engine: stage switches, fades, slews, int16 output
if len(sys.argv) > 1 and sys.argv[1] in ('vr', 'multi'):
    rs = soxr.ResampleStream(6, 1, 2, 'float32', vr=True)
    n = 0
    for i, (ratio, slew) in enumerate([(None, 0), (3.3, 500), (1.7, 200), (0.8, 300), (0.35, 0), (5.9, 1000), (None, 0), (None, 0)]):
        if ratio:
            rs.set_io_ratio(ratio, 1, slew)
        n += rs.resample_chunk(rng.standard_normal((3000, 2)).astype(np.float32), last=(i == 7)).shape[0]
    print('vr float32 frames', n)
    rs = soxr.ResampleStream(3, 1, 1, 'int16', vr=True)
    print('vr int16 frames', sum(rs.resample_chunk(np.round(rng.standard_normal(2000) * 5000).astype(np.int16), last=(i == 2)).shape[0] for i in range(3)))
