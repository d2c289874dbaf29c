# accuracy of the single-pass kernel on HQ (two-stage, att 132 dB) plans, against the reference and the oracle
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import soxr, soxr_oracle as O, ref_soxr as R
rng = np.random.default_rng(11)
for (ir, orr) in [(48000, 44100), (44100, 48000)]:
    x = rng.standard_normal((4, 30000, 2)).astype(np.float32)
    y = soxr.resample_batch(x, ir, orr, 'HQ')
    ye = soxr.resample_batch(x, ir, orr, 'HQ', exact_grid=True)
    for b in range(2):
        ref = R.resample(x[b], ir, orr, 'HQ').astype(np.float64)
        ref64 = R.divide_proc(ir, orr, x[b], 'HQ', q_flags=R.DOUBLE_PRECISION).astype(np.float64)
        orc = O.resample_f64(x[b], ir, orr, O.HQ)
        r = lambda a, c: float(np.sqrt(np.mean((a.astype(np.float64) - c) ** 2)))
        print(ir, orr, 'HQ clip', b, 'fdr_all=%s' % os.environ.get('B200SOXR_FDR_ALL'),
              'default-vs-ref %.3e  default-vs-oracle %.3e  exact-vs-oracle %.3e  ref(f32 engine)-vs-oracle %.3e  default-vs-exact %.3e max %.3e' % (
                  r(y[b], ref), r(y[b], orc), r(ye[b], orc), r(ref, orc), r(y[b], ye[b]), float(np.abs(y[b].astype(np.float64) - orc).max())))
