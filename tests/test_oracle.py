"""CPU tests: pin the oracle restatement (oracle/soxr_oracle.py + .c) against the
reference's golden vectors (tests/golden, generated from the unmodified
reference) and -- when oracle/_ref is present -- against the reference live."""
import os
import sys

import numpy as np
import pytest

from _golden import Q, case_input, load_meta, load_plans, load_vectors, rms

import soxr_oracle as O
import ref_soxr as R

META = load_meta()
VEC = load_vectors()
PLANS = load_plans()

# tolerance vs the reference *as shipped*: f64 engine (VHQ) results agree to
# rounding; f32 engine (HQ/LQ/MQ/QQ) results carry the reference's own float
# noise (1.2-2.7e-7 RMS on unit-variance noise).
TOL_F64_ENGINE = 1e-12
TOL_F32_ENGINE = 5e-7


@pytest.mark.parametrize("key", sorted(PLANS))
def test_plan_matches_reference_trace(key):
    i, o, q = key.split(":")
    P = O.make_plan(float(i), float(o), Q[q])
    assert P.trace == PLANS[key]


@pytest.mark.parametrize("name", sorted(META))
def test_oracle_matches_golden_vectors(name):
    m = META[name]
    x = case_input(m)
    y_ref = VEC[name]
    y = O.resample_f64(x, m['in_rate'], m['out_rate'], Q[m['quality']])
    assert y.shape == y_ref.shape
    f64_engine = m['quality'] == 'VHQ'
    if np.dtype(m['dtype']).kind == 'f':
        tol = TOL_F64_ENGINE if (f64_engine and m['dtype'] == 'float64') else TOL_F32_ENGINE
        assert rms(y - y_ref) <= tol * max(1.0, rms(y_ref))
    else:
        # compare against the reference engine's float stream (exact power-of-two io scale)
        eng = VEC[name + '__engine'].astype(np.float64)
        tol = TOL_F64_ENGINE if f64_engine else TOL_F32_ENGINE
        assert rms(y - eng) <= tol * rms(eng)
        # rounding-only output: within 1 LSB of round(oracle) for int16
        if m['dtype'] == 'int16':
            nd = VEC[name + '__nodither'].astype(np.float64)
            d = np.abs(np.rint(y) - nd)
            assert d.max() <= 1
            assert (d == 0).mean() > 0.99


@pytest.mark.parametrize("name", [n for n in sorted(META) if META[n]['dtype'].startswith('int')])
def test_k7_restatement_bit_exact(name):
    """rint/clip/TPDF-dither restatement (oracle/soxr_oracle.c) reproduces the
    reference's integer output bit-for-bit when fed the reference's own float
    stream with the reference's per-call block partition."""
    m = META[name]
    eng = VEC[name + '__engine']
    eng2 = eng[:, None] if eng.ndim == 1 else eng
    ch = eng2.shape[1]
    io_ratio = m['in_rate'] / m['out_rate']
    # (b) no dither
    got, clips, _ = O.rint_clip_interleave(np.ascontiguousarray(eng2.T), m['dtype'], seed=None)
    want = VEC[name + '__nodither']
    want2 = want[:, None] if want.ndim == 1 else want
    assert np.array_equal(got, want2)
    # (a) dither (int16 only), replaying the call partition
    if m['dtype'] == 'int16':
        seed = O.dither_seed0(ch, io_ratio)
        pos = 0
        outs = []
        for n in m['calls']:
            blk = np.ascontiguousarray(eng2[pos:pos + n].T)
            o, c, seed = O.rint_clip_interleave(blk, 'int16', seed=seed)
            outs.append(o)
            pos += n
        got = np.concatenate(outs)
        want = VEC[name]
        want2 = want[:, None] if want.ndim == 1 else want
        assert np.array_equal(got, want2)


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libsoxr_ref.so not built")
@pytest.mark.parametrize("case", [
    (48000, 44100, 'VHQ', 3000), (44100, 48000, 'VHQ', 3000), (96000, 48000, 'VHQ', 5000),
    (44100, 47999.5, 'VHQ', 2500), (48000, 8000, 'VHQ', 6000), (8000, 48000, 'VHQ', 1500),
    (32000, 44100, 'VHQ', 999), (44100, 32000, 'VHQ', 1), (44100, 32000, 'VHQ', 0),
])
def test_oracle_matches_reference_live(case):
    ir, orate, q, n = case
    x = np.random.default_rng(n + 17).standard_normal((n, 2))
    y_ref = R.divide_proc(ir, orate, x, q)
    y = O.resample_f64(x, ir, orate, Q[q])
    assert y.shape == y_ref.shape
    assert rms(y - y_ref) <= TOL_F64_ENGINE


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libsoxr_ref.so not built")
@pytest.mark.parametrize("seed", range(16))
def test_oracle_matches_reference_live_on_random_rates(seed):
    # the oracle's closed form against the reference's f64 engine (VHQ) at rates the fixtures do not hold,
    # rational and irrational, up and down
    rng = np.random.default_rng(3000 + seed)
    rates = [8000, 11025, 12000, 16000, 22050, 24000, 32000, 44100, 48000, 88200, 96000]
    ir = float(rng.choice(rates))
    orate = float(rng.choice(rates)) + (0.5 if seed % 4 == 0 else 0.0)
    if ir == orate:
        orate = ir * 1.5
    n = int(rng.integers(1, 4000))
    x = rng.standard_normal((n, 1))
    y_ref = R.divide_proc(ir, orate, x, 'VHQ')
    y = O.resample_f64(x, ir, orate, O.VHQ)
    assert y.shape == y_ref.shape, (ir, orate, n)
    assert rms(y - y_ref) <= TOL_F64_ENGINE, (ir, orate, n)
