"""CPU tests of the product's host side: C-ABI symbols, planner vs the reference's
SOXR_TRACE plans, call-schedule emulation vs the reference's stream patterns,
front-end argument validation.  No compute call is made (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from _golden import Q, load_plans, load_streams

import soxr
from soxr import soxr_ext

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLANS = load_plans()
STREAMS = load_streams()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200soxr.h")).read()
    names = set(re.findall(r"\b(b200soxr_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 18
    lib = C.CDLL(os.path.join(ROOT, "python-soxr_b200", "soxr", "libb200soxr.so"))
    for n in sorted(names):
        assert hasattr(lib, n), n


def test_spec_constructors_match_reference_presets():
    # soxr.c:106-145: HQ -> 20 bit, VHQ -> 28 bit, Fp0 values from SURVEY.md section 8a-R1
    lib = soxr_ext._lib
    hq = lib.b200soxr_quality_spec(soxr.HQ, 0)
    vhq = lib.b200soxr_quality_spec(soxr.VHQ, 0)
    assert (hq.precision, vhq.precision) == (20, 28)
    assert hq.phase_response == 50 and hq.stopband_begin == 1
    assert abs(hq.passband_end - 0.913744) < 1e-6 and abs(vhq.passband_end - 0.911328) < 1e-6
    rt = lib.b200soxr_runtime_spec(1)
    assert (rt.log2_min_dft_size, rt.log2_large_dft_size, rt.coef_size_kbytes) == (10, 17, 400)
    io = lib.b200soxr_io_spec(3, 3)
    assert io.scale == 1 and not io.e
    assert lib.b200soxr_io_spec(9, 0).e


@pytest.mark.parametrize("key", sorted(PLANS))
def test_planner_matches_reference_trace(key):
    i, o, q = key.split(":")
    assert soxr_ext.plan_dump(float(i), float(o), Q[q]) == PLANS[key]


@pytest.mark.parametrize("name", sorted(STREAMS))
def test_call_emulation_matches_reference_stream(name):
    s = STREAMS[name]
    n, chunk = s['frames'], s['chunk']
    lens = [min(chunk, n - i) for i in range(0, n, chunk)]
    # every chunk but the last is a plain process() call
    od, dl = soxr_ext.emulate_calls(s['in_rate'], s['out_rate'], Q[s['quality']], lens[:-1])
    assert od == s['counts'][:-1]
    assert np.allclose(dl, s['delays'][:-1], rtol=0, atol=1e-9)
    # the last chunk is followed by flush calls until empty: total must be floor(n*out/in+.5)
    od2, _ = soxr_ext.emulate_calls(s['in_rate'], s['out_rate'], Q[s['quality']], lens + [0, 0])
    assert od2[-1] == 0
    assert sum(od2) == soxr_ext.output_frames(s['in_rate'], s['out_rate'], n)
    assert od2[-3] + od2[-2] == s['counts'][-1]


@pytest.mark.parametrize("n", [0, 1, 2, 99, 100, 101, 31999, 32000, 32001, 266151])
@pytest.mark.parametrize("rates", [(44100, 32000), (32000, 44100), (48000, 44100), (44100, 47999.5)])
def test_output_frames(n, rates):
    assert soxr_ext.output_frames(rates[0], rates[1], n) == int(n / (rates[0] / rates[1]) + .5)


@pytest.mark.parametrize('in_rate, out_rate', [(100, 0), (50, -1), (0, 100.5), (-1.5, 100)])
def test_bad_sr(in_rate, out_rate):
    with pytest.raises(ValueError):
        soxr.resample(np.zeros(100), in_rate, out_rate)
    with pytest.raises(ValueError):
        soxr.ResampleStream(in_rate, out_rate, 1)


@pytest.mark.parametrize('dtype', [np.complex64, np.complex128, np.int8, np.int64])
def test_bad_dtype(dtype):
    with pytest.raises((TypeError, ValueError)):
        soxr.resample(np.zeros(100, dtype=dtype), 100, 200)
    with pytest.raises(TypeError):
        soxr.ResampleStream(100, 200, 1, dtype=dtype)


def test_bad_quality_channels_ndim():
    with pytest.raises(ValueError):
        soxr.resample(np.zeros(100, np.float32), 100, 200, quality='ultra')
    with pytest.raises(ValueError):
        soxr.resample(np.zeros((4, 4, 4), np.float32), 100, 200)
    with pytest.raises(ValueError):
        soxr.ResampleStream(100, 200, 0)
    with pytest.raises(ValueError):
        soxr.ResampleStream(100, 200, 65537)
    assert soxr._quality_to_enum('SOXR_MQ') == soxr.MQ and soxr._quality_to_enum(soxr.VHQ) == soxr.VHQ


def test_no_cpu_fallback_without_gpu():
    if soxr_ext.device_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        soxr.resample(np.zeros(100, np.float32), 48000, 44100)
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        soxr.ResampleStream(48000, 44100, 1)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "python-soxr_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "soxr_oracle" not in txt and "ref_soxr" not in txt and "libsoxr_ref" not in txt, f


# ---- randomized, against the live reference (oracle/_ref, built by __graft_entry__.build() where the reference
# sources exist): the planner and the integer call emulator over rates / chunkings the fixtures do not hold ---------
def _ref_stream_counts(in_rate, out_rate, quality, lens):
    import ref_soxr as R
    rs = R.RefStream(in_rate, out_rate, 1, dtype='float32', quality=quality)
    counts, delays = [], []
    for n in lens:
        y = rs.resample_chunk(np.zeros(n, np.float32), last=False)
        counts.append(len(y))
        delays.append(rs.delay())
    return counts, delays


def _ref_available():
    try:
        import ref_soxr as R
        return R.available()
    except Exception:
        return False


@pytest.mark.skipif(not _ref_available(), reason="oracle/_ref/libsoxr_ref.so not built")
@pytest.mark.parametrize("seed", range(40))
def test_call_emulation_matches_live_reference_on_random_partitions(seed):
    rng = np.random.default_rng(1000 + seed)
    rates = [8000, 11025, 16000, 22050, 24000, 32000, 44100, 48000, 88200, 96000, 47999, 44100.5]
    in_rate, out_rate = (float(r) for r in rng.choice(rates, 2, replace=False))
    quality = ['QQ', 'LQ', 'MQ', 'HQ', 'VHQ'][int(rng.integers(0, 5))]
    lens = [int(v) for v in rng.integers(1, 3000, size=int(rng.integers(3, 25)))]
    want_counts, want_delays = _ref_stream_counts(in_rate, out_rate, quality, lens)
    got_counts, got_delays = soxr_ext.emulate_calls(in_rate, out_rate, Q[quality], lens)
    assert got_counts == want_counts, (in_rate, out_rate, quality, lens)
    assert np.allclose(got_delays, want_delays, rtol=0, atol=1e-9)


@pytest.mark.skipif(not _ref_available(), reason="oracle/_ref/libsoxr_ref.so not built")
@pytest.mark.parametrize("seed", range(24))
def test_planner_matches_live_reference_trace_on_random_rates(seed):
    import ref_soxr as R
    rng = np.random.default_rng(2000 + seed)
    in_rate = float(rng.integers(4000, 200000))
    out_rate = float(rng.integers(4000, 200000)) + (0.5 if seed % 3 == 0 else 0.0)
    quality = ['LQ', 'MQ', 'HQ', 'VHQ'][seed % 4]
    want = [l for l in R.trace_plan(in_rate, out_rate, quality) if l.startswith('fir_len') or '>>' in l]
    assert soxr_ext.plan_dump(in_rate, out_rate, Q[quality]) == want, (in_rate, out_rate, quality)


# ---- non-default quality specs and libsoxr's run-time environment overrides (soxr.c:426-433) ----
STEEP = 0x40                      # SOXR_STEEP_FILTER (soxr.h:299)
ROLLOFF_MEDIUM, ROLLOFF_NONE = 1, 2


@pytest.mark.skipif(not _ref_available(), reason="oracle/_ref/libsoxr_ref.so not built")
@pytest.mark.parametrize("in_rate,out_rate,recipe,q_flags", [
    (48000, 44100, 4 | STEEP, 0), (44100, 48000, 6 | STEEP, 0), (96000, 48000, 4 | STEEP, 0), (44100, 47999.5, 6 | STEEP, 0),
    (48000, 44100, 4, ROLLOFF_MEDIUM), (48000, 44100, 6, ROLLOFF_NONE), (44100, 16000, 4, ROLLOFF_NONE),
    (48000, 8000, 5, 0), (8000, 48000, 7, 0), (44100, 48000, 3, 0),
    # non-linear phase recipes: the tap counts after lsx_fir_to_phase are part of the trace
    (48000, 44100, 4 | 0x10, 0), (48000, 44100, 4 | 0x30, 0), (44100, 48000, 6 | 0x20, 0), (96000, 48000, 4 | 0x30, 0),
    (44100, 47999.5, 6 | 0x10, 0), (48000, 8000, 4 | 0x30, 0), (48000, 44100, 1 | 0x30, 0)])
def test_planner_follows_non_default_quality_specs(in_rate, out_rate, recipe, q_flags):
    import ref_soxr as R
    want = [l for l in R.trace_plan(in_rate, out_rate, recipe, q_flags=q_flags) if l.startswith('fir_len') or '>>' in l]
    assert soxr_ext.plan_dump(in_rate, out_rate, recipe, q_flags) == want


@pytest.mark.skipif(not _ref_available(), reason="oracle/_ref/libsoxr_ref.so not built")
@pytest.mark.parametrize("env", [{"SOXR_MIN_DFT_SIZE": "12"}, {"SOXR_LARGE_DFT_SIZE": "9", "SOXR_MIN_DFT_SIZE": "8"},
                                 {"SOXR_COEFS_SIZE": "100"}, {"SOXR_COEF_INTERP": "2"}, {"SOXR_NOSMALLINTOPT": "1"},
                                 {"SOXR_MIN_DFT_SIZE": "99"}])
@pytest.mark.parametrize("in_rate,out_rate,quality", [(48000, 44100, 'HQ'), (44100, 47999.5, 'VHQ'), (8000, 48000, 'HQ')])
def test_planner_reads_libsoxr_environment_overrides(env, in_rate, out_rate, quality, monkeypatch):
    import ref_soxr as R
    want = [l for l in R.trace_plan(in_rate, out_rate, quality, env=env) if l.startswith('fir_len') or '>>' in l]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    assert soxr_ext.plan_dump(in_rate, out_rate, Q[quality]) == want, (env, in_rate, out_rate, quality)
