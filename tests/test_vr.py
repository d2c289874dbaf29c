"""Variable-rate engine (`ResampleStream(vr=True)`, `set_io_ratio`).

CPU: the host control emulator (csrc/vr.cu, `b200soxr_vr_emulate`) against the LIVE reference's per-call output counts
on random sequences of ratio changes, slews, chunk sizes and flushes.  GPU: the CUDA path against golden streams
generated from the unmodified reference (tests/golden/make_vr_golden.py) -- bit for bit, int16 dither included --
and against the live reference when oracle/_ref is built."""
import ctypes as C
import importlib.util
import json
import multiprocessing as mp
import os

import numpy as np
import pytest

from _golden import GOLDEN

import ref_soxr as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "python-soxr_b200", "soxr", "libb200soxr.so")

_spec = importlib.util.spec_from_file_location("make_vr_golden", os.path.join(GOLDEN, "make_vr_golden.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)


def _emulate(max_ratio, steps):
    L = C.CDLL(LIB)
    L.b200soxr_vr_emulate.restype = C.c_char_p
    L.b200soxr_vr_emulate.argtypes = [C.c_double, C.c_size_t] + [C.c_void_p] * 6
    n = len(steps)
    ratio = np.array([s[0] for s in steps], np.float64)
    slew = np.array([s[1] for s in steps], np.uint64)
    ilen = np.array([s[2] for s in steps], np.uint64)
    fl = np.array([s[3] for s in steps], np.uint8)
    olen = np.array([s[4] for s in steps], np.uint64)
    od = np.zeros(n, np.uint64)
    err = L.b200soxr_vr_emulate(max_ratio, n, ratio.ctypes.data, slew.ctypes.data, ilen.ctypes.data, fl.ctypes.data,
                                olen.ctypes.data, od.ctypes.data)
    assert not err, err
    return [int(v) for v in od]


def _ref_counts(max_ratio, steps):
    rl = R.lib()
    h = R._create(max_ratio, 1.0, 1, 0, 0, 'HQ', 0, 32)
    out = []
    for (ratio, slew, ilen, flush, olen) in steps:
        if ratio > 0:
            assert not rl.soxr_set_io_ratio(h, ratio, slew)
        x = np.zeros(max(ilen, 1), np.float32)
        y = np.zeros(max(olen, 1), np.float32)
        od = C.c_size_t(0)
        if flush and not ilen:
            err = rl.soxr_process(h, None, 0, None, y.ctypes.data, olen, C.byref(od))
        else:
            err = rl.soxr_process(h, x.ctypes.data, ((~ilen) & (2 ** 64 - 1)) if flush else ilen, None, y.ctypes.data,
                                  olen, C.byref(od))
        assert not err, err
        out.append(od.value)
    rl.soxr_delete(h)
    return out


def _ref_counts_child(args, q):
    q.put(_ref_counts(*args))


def _ref_counts_isolated(max_ratio, steps):
    """vr32.c corrupts its own heap on a few sequences (1 of 1500 random ones here): run it in a child."""
    ctx = mp.get_context('fork')
    q = ctx.Queue()
    p = ctx.Process(target=_ref_counts_child, args=((max_ratio, steps), q))
    p.start()
    try:
        r = q.get(timeout=30)
    except Exception:
        r = None
    p.join(5)
    return r if p.exitcode == 0 else None


def _random_steps(rng, max_ratio):
    steps = []
    n = int(rng.integers(3, 30))
    for i in range(n):
        ratio, slew = 0.0, 0
        if rng.random() < 0.3:
            ratio = float(max_ratio * rng.uniform(0.05, 1.0)) if rng.random() < 0.8 else float(max_ratio)
            slew = int(rng.choice([0, 0, 1, 100, 1000, 5000]))
        ilen = int(rng.choice([0, 1, 10, 100, 480, 1000, 4096]))
        olen = int(rng.choice([16, 100, 1000, 5000, 20000]))
        steps.append((ratio, slew, ilen, 1 if i == n - 1 else 0, olen))
    return steps + [(0.0, 0, 0, 1, 4096)] * 5


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
def test_vr_control_emulator_matches_live_reference_counts():
    rng = np.random.default_rng(0)
    checked = 0
    for _ in range(120):
        max_ratio = float(rng.choice([0.05, 0.3, 0.9, 1.0, 1.5, 2.0, 3.7, 10.0, 20.0]))
        steps = _random_steps(rng, max_ratio)
        want = _ref_counts_isolated(max_ratio, steps)
        if want is None:
            continue
        assert _emulate(max_ratio, steps) == want, (max_ratio, steps)
        checked += 1
    assert checked >= 100


def test_vr_emulator_negative_clock_drain():
    # the reduced case that exposed fifo_read's "too much asked: drop nothing" path with a negative count
    steps = [(0.0, 0, 4096, 0, 100), (1.2068767235243916, 5000, 1, 0, 20000), (0.7557603680316425, 0, 1000, 0, 1000),
             (0.0, 0, 1000, 1, 16)] + [(0.0, 0, 0, 1, 4096)] * 3
    assert _emulate(3.7, steps) == [100, 899, 1000, 16, 4096, 1733, 0]


def _meta():
    with open(os.path.join(GOLDEN, "vr_meta.json")) as f:
        return json.load(f)


def _run_ours(sc):
    import soxr
    return G.run(sc, lambda s: soxr.ResampleStream(s["in_rate"], s["out_rate"], s["ch"], s["dtype"], "HQ", vr=True),
                 lambda rs, x, last: rs.resample_chunk(x if sc["ch"] > 1 else x[:, 0], last))


@pytest.mark.gpu
@pytest.mark.parametrize("name", [m["name"] for m in _meta()])
def test_vr_golden_streams_bit_exact(name):
    sc = next(m for m in _meta() if m["name"] == name)
    want = np.load(os.path.join(GOLDEN, "vr_vectors.npz"))[name]
    outs = _run_ours(sc)
    assert [int(o.shape[0]) for o in outs] == sc["lengths"]
    got = np.concatenate([o.reshape(o.shape[0], sc["ch"]) for o in outs], axis=0)
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.array_equal(got, want), "max |diff| %g" % np.abs(got.astype(np.float64) - want).max()


@pytest.mark.gpu
@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
def test_vr_random_scenarios_vs_live_reference():
    rng = np.random.default_rng(7)
    for case in range(6):
        in_rate = float(rng.choice([0.4, 1.0, 2.5, 7.0, 12.0]))
        sc = dict(name="rnd%d" % case, in_rate=in_rate, out_rate=1.0, ch=int(rng.integers(1, 4)), dtype="float32",
                  seed=1000 + 50 * case, scale=1.0, actions=[])
        for i in range(int(rng.integers(4, 10))):
            if rng.random() < 0.5:
                sc["actions"].append(["ratio", float(in_rate * rng.uniform(0.1, 1.0)), 1.0, int(rng.choice([0, 1, 300, 2500]))])
            sc["actions"].append(["chunk", int(rng.choice([100, 480, 1500, 4000])), False])
        sc["actions"].append(["chunk", 500, True])
        want = G.run(sc, lambda s: R.RefStream(s["in_rate"], s["out_rate"], s["ch"], s["dtype"], "HQ", vr=True),
                     lambda rs, x, last: rs.process_vr(x, last))
        got = _run_ours(sc)
        assert [o.shape[0] for o in got] == [o.shape[0] for o in want], sc
        for a, b in zip(got, want):
            assert np.array_equal(a.reshape(b.shape), b), sc


@pytest.mark.gpu
def test_vr_api_behaviour():
    import soxr
    rs = soxr.ResampleStream(4, 1, 1, 'float32', vr=True)
    assert rs.engine() == 'b200-vr32'
    assert rs.delay() == 100.0                     # vr_delay is a constant in the reference (vr32.c:612-616)
    x = np.random.default_rng(1).standard_normal(3000).astype(np.float32)
    a = rs.resample_chunk(x, last=True)
    with pytest.raises(RuntimeError):
        rs.resample_chunk(x)                       # input after last
    rs.clear()
    b = rs.resample_chunk(x, last=True)
    assert np.array_equal(a, b)                    # clear() => fresh stream at the create-time ratio
    # a constant-rate handle accepts only its own ratio (soxr.c:595-597)
    cr = soxr.ResampleStream(48000, 44100, 1)
    cr.set_io_ratio(48000, 44100)
    with pytest.raises(RuntimeError, match="varying O/I ratio"):
        cr.set_io_ratio(48000, 32000)
