"""The C ABI driven from plain C (tests/c_abi/test_abi.c, compiled by __graft_entry__.build()): the calls a
re-pointed python-soxr binding would make -- b200soxr_oneshot, b200soxr_process with split (*_S) pointer arrays,
interleaved streaming, B200SOXR_RT_DEVICE_IO, a caller-supplied cudaStream_t, b200soxr_create_batch -- each checked
against the oracle (<= 1e-6 RMS, the north_star float32 tolerance), against the committed reference golden where one
exists for the shape, and bit for bit against the python front-end on the same input (same library, same block grid).
"""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import soxr
import soxr_oracle as O
from _golden import Q, case_input, load_meta, load_vectors, rms

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "c_abi", "test_abi")


def _run(mode, x, in_rate, out_rate, quality, tmp_path, n_streams=1):
    assert os.path.exists(EXE), "tests/c_abi/test_abi is not built: run __graft_entry__.build()"
    ch = x.shape[-1]
    fin, fout = str(tmp_path / "in.f32"), str(tmp_path / "out.f32")
    np.ascontiguousarray(x, dtype=np.float32).tofile(fin)
    r = subprocess.run([EXE, mode, repr(float(in_rate)), repr(float(out_rate)), str(ch), str(Q[quality]), fin, fout],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    y = np.fromfile(fout, dtype=np.float32)
    return y.reshape(n_streams, -1, ch) if n_streams > 1 else y.reshape(-1, ch)


@pytest.mark.parametrize("mode", ["oneshot", "split", "stream", "device", "cstream"])
@pytest.mark.parametrize("in_rate,out_rate,quality,frames,ch", [(48000, 44100, 'VHQ', 6000, 2), (44100, 16000, 'HQ', 30000, 3),
                                                                (44100, 47999.5, 'HQ', 9000, 1), (48000, 48000, 'HQ', 1000, 2),
                                                                (96000, 48000, 'HQ', 0, 2)])
def test_c_harness_modes(mode, in_rate, out_rate, quality, frames, ch, tmp_path):
    x = np.random.default_rng(frames + ch).standard_normal((frames, ch)).astype(np.float32)
    y = _run(mode, x, in_rate, out_rate, quality, tmp_path)
    want = O.resample_f64(x, in_rate, out_rate, Q[quality]) if frames else np.zeros((0, ch))
    assert y.shape == want.shape
    if frames:
        assert rms(y.astype(np.float64) - want) <= 1e-6
    # the python front-end makes the same calls through ctypes: bit-identical
    assert np.array_equal(y, soxr._resample_oneshot(x, in_rate, out_rate, quality))


def test_c_harness_against_reference_golden(tmp_path):
    meta, vec = load_meta(), load_vectors()
    names = [n for n in sorted(meta) if meta[n]['dtype'] == 'float32' and len(meta[n]['shape']) == 2]
    assert names
    for name in names[:6]:
        m = meta[name]
        x = case_input(m)
        for mode in ("oneshot", "split", "device"):
            y = _run(mode, x, m['in_rate'], m['out_rate'], m['quality'], tmp_path)
            assert y.shape == vec[name].shape
            assert rms(y.astype(np.float64) - vec[name]) <= 1e-6, (name, mode)


def test_c_harness_batched_streams(tmp_path):
    x = np.random.default_rng(8).standard_normal((3, 7000, 2)).astype(np.float32)
    y = _run("batch", x, 44100, 16000, 'HQ', tmp_path, n_streams=3)
    for b in range(3):
        assert np.array_equal(y[b], soxr._resample_oneshot(x[b], 44100, 16000, 'HQ'))


def test_c_harness_variable_rate(tmp_path):
    """`vr` mode of the harness (B200SOXR_VR + b200soxr_set_io_ratio from plain C) against the same calls made on the
    unmodified reference through ctypes: bit for bit."""
    import ctypes as C
    import ref_soxr as R
    if not R.available():
        pytest.skip("oracle/_ref not built")
    in_rate, out_rate, ch = 5.0, 1.0, 2
    x = np.random.default_rng(21).standard_normal((12000, ch)).astype(np.float32)
    y = _run("vr", x, in_rate, out_rate, 'HQ', tmp_path)
    rl = R.lib()
    h = R._create(in_rate, out_rate, ch, 0, 0, 'HQ', 0, 32)
    room, chunk, ilen = 8192, 1000, x.shape[0]
    buf = np.zeros((room, ch), np.float32)
    od = C.c_size_t(0)
    pieces, idx, changed = [], 0, False
    while True:
        flush = idx >= ilen
        n = 0 if flush else min(chunk, ilen - idx)
        if not changed and idx >= ilen // 2:
            assert not rl.soxr_set_io_ratio(h, 0.6 * in_rate / out_rate, 500)
            changed = True
        err = rl.soxr_process(h, None if flush else x[idx:].ctypes.data, n, None, buf.ctypes.data, room, C.byref(od))
        assert not err, err
        pieces.append(buf[:od.value].copy())
        if flush and od.value == 0:
            break
        idx += n
    rl.soxr_delete(h)
    want = np.concatenate(pieces)
    assert y.shape == want.shape
    assert np.array_equal(y, want)
