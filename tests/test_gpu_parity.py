"""GPU parity tests (-m gpu): the CUDA path, called through the python front-end
and the C ABI, against (1) the committed golden vectors produced by the
unmodified reference and (2) the oracle restatement on the same seeded inputs.

Tolerances (BASELINE.json north_star): float32 I/O within 1e-6 RMS of libsoxr on
unit-variance noise; float64 I/O (f64 engine) within 1e-11 RMS; integers: see
SURVEY.md section 7-H2 -- K7 itself is bit-exact given the reference's float
stream, end-to-end ints are compared after rounding (|delta| <= 1 and >= 99.7 % equal
on these short vectors without dither -- the full-size figure, >= 99.9 %, is asserted in
test_gpu_live_ref.py -- and |delta| <= 2 with the reference's data-salted dither).
"""
import numpy as np
import pytest

from _golden import Q, case_input, load_meta, load_streams, load_vectors, make_input, rms

pytestmark = pytest.mark.gpu

import soxr
from soxr import soxr_ext
import soxr_oracle as O

META = load_meta()
VEC = load_vectors()
STREAMS = load_streams()

TOL_F32 = 1e-6
TOL_F64 = 1e-11


def _tol(dtype, quality):
    return TOL_F64 if (dtype == 'float64' and quality == 'VHQ') else TOL_F32


@pytest.mark.parametrize("name", [n for n in sorted(META) if META[n]['dtype'].startswith('float')])
def test_float_matches_reference_golden_and_oracle(name):
    m = META[name]
    x = case_input(m)
    y = soxr.resample(x, m['in_rate'], m['out_rate'], m['quality'])
    ref = VEC[name]
    assert y.dtype == x.dtype and y.shape == ref.shape
    tol = _tol(m['dtype'], m['quality'])
    assert rms(y.astype(np.float64) - ref) <= tol
    orc = O.resample_f64(x, m['in_rate'], m['out_rate'], Q[m['quality']])
    assert rms(y.astype(np.float64) - orc) <= tol


@pytest.mark.parametrize("name", [n for n in sorted(META) if META[n]['dtype'] == 'float32'])
def test_batch_kernel_matches_reference_golden(name):
    # soxr.resample_batch: the single-pass frequency-domain kernel where the plan has a ratio family (the fam_*
    # goldens, one per family), the exact-grid kernels otherwise -- against the unmodified reference's output
    m = META[name]
    x = case_input(m)
    xb = np.stack([x, x[::-1].copy()])            # two clips: the golden one and its time reverse
    y = soxr.resample_batch(xb, m['in_rate'], m['out_rate'], m['quality'])
    ref = VEC[name]
    assert y.dtype == np.float32 and y[0].shape == ref.shape
    assert rms(y[0].astype(np.float64) - ref) <= TOL_F32
    assert np.abs(y[0].astype(np.float64) - ref).max() <= 8e-6


@pytest.mark.parametrize("name", [n for n in sorted(META) if META[n]['dtype'].startswith('int')])
def test_int_matches_reference_after_rounding(name):
    m = META[name]
    x = case_input(m)
    y = soxr.resample(x, m['in_rate'], m['out_rate'], m['quality'])
    ref = VEC[name]
    assert y.dtype == x.dtype and y.shape == ref.shape
    d = np.abs(y.astype(np.int64) - ref.astype(np.int64))
    if m['dtype'] == 'int16':
        assert d.max() <= 2                       # reference's own tolerance (tests/test_resample.py:130)
        nd = VEC[name + '__nodither'].astype(np.int64)
        eng = VEC[name + '__engine'].astype(np.float64)
        # rounding-only comparison through the C ABI (NO_DITHER)
        x2 = x[:, None] if x.ndim == 1 else x
        y2 = np.empty((ref.shape[0], x2.shape[1]), np.int16)
        soxr_ext.resample_batch_raw(np.ascontiguousarray(x2).ctypes.data, y2.ctypes.data,
                                    in_rate=m['in_rate'], out_rate=m['out_rate'], quality=Q[m['quality']],
                                    dtype_code=3, n_clips=1, channels=x2.shape[1], ilen=x2.shape[0],
                                    in_strides=(0, x2.shape[1], 1), out_strides=(0, x2.shape[1], 1),
                                    olen=ref.shape[0], flags=soxr_ext.BATCH_NO_DITHER)
        nd2 = nd[:, None] if nd.ndim == 1 else nd
        dd = np.abs(y2.astype(np.int64) - nd2)
        assert dd.max() <= 1 and (dd == 0).mean() >= 0.997
    else:
        # int32 at 2^31 scale: VHQ runs the f64 engine -> exact after rounding up to 1 LSB;
        full = float(np.abs(ref).max())
        assert d.max() <= max(1, full * 2e-7)


@pytest.mark.parametrize("name", [n for n in sorted(META) if META[n]['dtype'].startswith('int')])
def test_device_k7_bit_exact_on_reference_stream(name):
    m = META[name]
    eng = VEC[name + '__engine']
    eng2 = eng[:, None] if eng.ndim == 1 else eng
    planar = np.ascontiguousarray(eng2.T)
    want = VEC[name + '__nodither']
    want = want[:, None] if want.ndim == 1 else want
    got, clips, _ = soxr_ext.quantize(planar, m['dtype'], seed=None)
    assert np.array_equal(got, want)
    if m['dtype'] == 'int16':
        seed = O.dither_seed0(planar.shape[0], m['in_rate'] / m['out_rate'])
        got, clips, _ = soxr_ext.quantize(planar, 'int16', call_lens=m['calls'], seed=seed)
        want = VEC[name]
        want = want[:, None] if want.ndim == 1 else want
        assert np.array_equal(got, want)
        assert clips == m['clips']


def test_device_k7_saturation_and_clip_count():
    rng = np.random.default_rng(5)
    planar = (rng.standard_normal((2, 1000)) * 30000).astype(np.float32)
    planar[0, 17] = 1e9
    planar[1, 999] = -1e9
    planar[1, 3] = np.nan
    for seed in (None, 12345):
        got, clips, s2 = soxr_ext.quantize(planar, 'int16', call_lens=[300, 0, 700], seed=seed)
        pos = 0
        outs, want_clips, sd = [], 0, seed
        for n in [300, 0, 700]:
            o, c, sd = O.rint_clip_interleave(np.ascontiguousarray(planar[:, pos:pos + n]), 'int16', seed=sd)
            outs.append(o)
            want_clips += c
            pos += n
        assert np.array_equal(got, np.concatenate(outs))
        assert clips == want_clips and s2 == sd


@pytest.mark.parametrize('channels', [1, 2, 3, 4, 6])
@pytest.mark.parametrize('out_dtype', ['int16', 'int32'])
@pytest.mark.parametrize('src_dtype', [np.float32, np.float64])
def test_device_k7_channel_layouts(channels, out_dtype, src_dtype):
    # interleaved outputs with an even channel count leave k_rint_clip as channel pairs (one 4- / 8-byte element per
    # frame, 16-byte vectors for stereo), the others sample by sample: same values, clip counts and seeds either way,
    # with saturating samples and NaNs in whole groups, tails and at call boundaries
    rng = np.random.default_rng(100 * channels + (out_dtype == 'int32'))
    scale = 30000 if out_dtype == 'int16' else 2.0e9
    n = 2000 + 16 * channels + 5
    planar = (rng.standard_normal((channels, n)) * scale).astype(src_dtype)
    planar[0, 17] = 1e12
    planar[channels - 1, n - 1] = -1e12
    planar[channels // 2, 3] = np.nan
    planar[:, 640:700] *= 1e-3          # quiet stretch: groups that never come near the range ends
    calls = [333, 0, 1000, n - 1333]
    for seed in ((None, 987654321) if out_dtype == 'int16' else (None,)):
        got, clips, s2 = soxr_ext.quantize(planar, out_dtype, call_lens=calls, seed=seed)
        pos, outs, want_clips, sd = 0, [], 0, seed
        for m in calls:
            o, c, sd = O.rint_clip_interleave(np.ascontiguousarray(planar[:, pos:pos + m]), out_dtype, seed=sd)
            outs.append(o)
            want_clips += c
            pos += m
        assert np.array_equal(got, np.concatenate(outs))
        assert clips == want_clips and s2 == sd


# (44100, 32000) and (32000, 44100) are the reference's own pairs; the others are the ratio families of the
# single-pass frequency-domain kernel, where the same bit-for-bit contract holds (the whole-signal drivers and
# streams all evaluate the reference's block grid; only soxr.resample_batch uses the faster kernel by default)
RATE_PAIRS = [(44100, 32000), (32000, 44100), (48000, 44100), (44100, 48000), (96000, 48000), (48000, 16000)]


@pytest.mark.parametrize('in_rate, out_rate', RATE_PAIRS)
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_divide_match(in_rate, out_rate, dtype):
    # reference tests/test_resample.py:46-57: three drivers agree bit-exactly
    x = np.random.default_rng(1).standard_normal((25999, 2)).astype(dtype)
    y_oneshot = soxr._resample_oneshot(x, in_rate, out_rate)
    y_divide = soxr.resample(x, in_rate, out_rate)
    y_split = soxr.resample(np.asfortranarray(x), in_rate, out_rate)
    assert np.all(y_oneshot == y_divide)
    assert np.all(y_oneshot == y_split)
    assert y_split.strides[0] == y_split.itemsize


@pytest.mark.parametrize('length', [0, 1, 2, 99, 100, 101, 31999, 32000, 32001, 34828, 34829, 34830, 44099, 44100, 44101,
                                    47999, 48000, 48001, 66149, 66150, 266151])
def test_length_match(length):
    # reference tests/test_resample.py:60-71 (all of its lengths)
    x = np.random.default_rng(2).standard_normal((266151, 2)).astype(np.float32)
    for in_rate, out_rate in [(44100, 32000), (32000, 44100), (48000, 44100)]:
        y_oneshot = soxr._resample_oneshot(x[:length], in_rate, out_rate)
        y_divide = soxr.resample(x[:length], in_rate, out_rate)
        y_split = soxr.resample(np.asfortranarray(x)[:length], in_rate, out_rate)
        assert y_divide.shape == (int(length * out_rate / in_rate + .5), 2)
        assert np.all(y_oneshot == y_divide)
        assert np.all(y_oneshot == y_split)


@pytest.mark.parametrize('in_rate, out_rate, quality, dtype, channels, frames', [
    (48000, 44100, 'HQ', np.float32, 2, 400001),
    (48000, 44100, 'HQ', np.float32, 1, 480000),      # BASELINE configs[0]
    (44100, 48000, 'VHQ', np.float64, 2, 300017),
    (47999, 44100, 'HQ', np.float32, 1, 600000),      # irrational: interpolated-coefficient stage behind the DFT stage
    (96000, 16000, 'MQ', np.float32, 3, 350003),
    (44100, 48000, 'QQ', np.float32, 2, 500000),
    (8000, 48000, 'HQ', np.float32, 2, 200000),
    (48000, 44100, 'HQ', np.float32, 2, 5000003),     # 40 MB: more pieces than pinned ring slots, both pipelines
    (44100, 48000, 'HQ', np.int16, 2, 3000001),       # integer output: the staged single chunk (12 MB in, rings reused)
])
def test_time_split_single_clip(in_rate, out_rate, quality, dtype, channels, frames):
    # one large C-ordered ndarray through soxr.resample runs cut in time (piece k + 1 copied in while piece k is
    # computed and copied back, api.cu); the Fortran-ordered twin of the same data takes the unsplit path and a
    # ResampleStream fed odd chunks is the third evaluation of the same block grid: all bit for bit equal
    x = np.random.default_rng(frames).standard_normal((frames, channels))
    x = (x * 8000).astype(dtype) if np.dtype(dtype).kind == 'i' else x.astype(dtype)
    y = soxr.resample(x if channels > 1 else x[:, 0], in_rate, out_rate, quality)
    y = y.reshape(len(y), channels)
    assert y.shape[0] == int(frames * out_rate / in_rate + .5)
    if np.dtype(dtype).kind == 'i':
        # integer output: the dither partition differs between the drivers by design (as in the reference); the same
        # clip as a device-resident batch of one has no host staging at all and replays resample's partition
        import torch
        yd = soxr.resample_batch(torch.from_numpy(x[None]).cuda(), in_rate, out_rate, quality, exact_grid=True)
        assert np.array_equal(y, yd[0].cpu().numpy())
        return
    if channels > 1:
        y_unsplit = soxr.resample(np.asfortranarray(x), in_rate, out_rate, quality)
        assert np.array_equal(y, y_unsplit)
    y_stream, _, _ = _stream(x, in_rate, out_rate, 70001, dtype, quality)
    assert np.array_equal(y, y_stream)
    y_one = soxr._resample_oneshot(x, in_rate, out_rate, quality)
    assert np.array_equal(y, y_one.reshape(y.shape))


def test_results_live_in_recycled_pinned_buffers():
    # results of the whole-signal drivers from 128 kB are ndarrays over page-locked buffers of the library (the
    # device-to-host copy lands in them directly); views keep the buffer alive, the last one gives it back, and the
    # next result of that size gets the same buffer again
    import gc

    def owner_of(a):
        o = a.base
        while isinstance(o, np.ndarray):
            o = o.base
        return o

    x = np.random.default_rng(11).standard_normal((150000, 2)).astype(np.float32)
    y = soxr.resample(x, 48000, 44100)
    assert type(owner_of(y)).__name__ == '_PinnedResult' and y.flags.writeable and not y.flags.owndata
    want = y.copy()
    ptr = owner_of(y).ptr
    view = y[1000:2000, 1]
    del y
    gc.collect()
    assert np.array_equal(view, want[1000:2000, 1])          # still backed by the buffer
    y2 = soxr.resample(x, 48000, 44100)                       # the first buffer is still held by `view`
    assert owner_of(y2).ptr != ptr and np.array_equal(y2, want)
    del view
    gc.collect()
    y3 = soxr._resample_oneshot(x, 48000, 44100)              # (a recycled buffer: this one or an older one of its size)
    assert type(owner_of(y3)).__name__ == '_PinnedResult' and np.array_equal(y3, want)
    y3 += 1                                                   # ordinary writable memory
    assert np.array_equal(y3, want + 1)
    small = soxr.resample(x[:1000], 48000, 44100)             # small results stay in ordinary memory
    assert type(owner_of(small)).__name__ != '_PinnedResult'
    # the cache itself, in a size class nothing else in the suite uses: freed buffers are handed out again
    p1 = soxr_ext._lib.b200soxr_result_alloc(13 << 20)
    assert p1
    soxr_ext._lib.b200soxr_result_free(p1)
    p2 = soxr_ext._lib.b200soxr_result_alloc((13 << 20) - 4096)   # same size class
    assert p2 == p1
    soxr_ext._lib.b200soxr_result_free(p2)
    # large buffers (64 MB and up): at most one stays cached
    L = soxr_ext._lib
    a, b = L.b200soxr_result_alloc(100 << 20), L.b200soxr_result_alloc(100 << 20)
    assert a and b and a != b
    L.b200soxr_result_free(a)
    L.b200soxr_result_free(b)          # drops `a` from the cache
    c = L.b200soxr_result_alloc(100 << 20)
    assert c == b
    L.b200soxr_result_free(c)
    # 1-D input takes the split-channel driver (as in the reference); its result is pinned too
    m = soxr.resample(x[:, 0].copy(), 48000, 44100)
    assert m.ndim == 1 and type(owner_of(m)).__name__ in ('_PinnedResult', 'DummyArray')
    assert np.array_equal(m, soxr.resample(x[:, :1].copy(), 48000, 44100)[:, 0])


@pytest.mark.parametrize('channels', [1, 2, 3, 5, 7, 24, 49])
def test_channel_match(channels):
    # reference tests/test_resample.py:74-84
    x = np.random.default_rng(3).standard_normal((15013, 49)).astype(np.float32)
    y_oneshot = soxr._resample_oneshot(x[:, :channels], 44100, 32000)
    y_divide = soxr.resample(x[:, :channels], 44100, 32000)
    y_split = soxr.resample(np.asfortranarray(x)[:, :channels], 44100, 32000)
    assert np.all(y_oneshot == y_divide)
    assert np.all(y_oneshot == y_split)
    orc = O.resample_f64(x[:, :channels], 44100, 32000, O.HQ)
    assert rms(y_divide - orc) <= TOL_F32


def _stream(x, in_rate, out_rate, chunk, dtype, quality='HQ'):
    rs = soxr.ResampleStream(in_rate, out_rate, x.shape[1], dtype=dtype, quality=quality)
    ys, counts, delays = [np.ndarray([0, x.shape[1]], dtype=dtype)], [], []
    for idx in range(0, len(x), chunk):
        end = idx + chunk
        eof = len(x) <= end
        y = rs.resample_chunk(x[idx:min(end, len(x))], last=eof)
        ys.append(y)
        counts.append(len(y))
        delays.append(rs.delay())
    return np.concatenate(ys), counts, delays


@pytest.mark.parametrize('in_rate, out_rate', [(44100, 32000), (32000, 44100), (48000, 44100), (96000, 48000)])
@pytest.mark.parametrize('chunk_size', [17, 509, 44100])
@pytest.mark.parametrize('length', [0, 100, 31999, 44100, 166151])
@pytest.mark.parametrize('dtype', ['float32', np.float64])
def test_stream_length(in_rate, out_rate, chunk_size, length, dtype):
    # reference tests/test_resample.py:105-116 (all of its lengths and chunk sizes): stream == one-shot, bit-exact
    x = np.random.default_rng(4).standard_normal((length, 1)).astype(dtype)
    y_oneshot = soxr._resample_oneshot(x, in_rate, out_rate)
    y_stream, _, _ = _stream(x, in_rate, out_rate, chunk_size, dtype)
    assert np.all(y_oneshot == y_stream)


@pytest.mark.parametrize("name", sorted(STREAMS))
def test_stream_counts_delays_match_reference(name):
    s = STREAMS[name]
    is_int = np.dtype(s['dtype']).kind == 'i'
    x = make_input(s['seed'], (s['frames'], s['channels']), s['dtype'], 5000.0 if is_int else 1.0)
    y, counts, delays = _stream(x, s['in_rate'], s['out_rate'], s['chunk'], s['dtype'], s['quality'])
    assert counts == s['counts']
    assert np.allclose(delays, s['delays'], rtol=0, atol=1e-9)
    orc = O.resample_f64(x, s['in_rate'], s['out_rate'], Q[s['quality']])
    assert y.shape == orc.shape
    if is_int:
        # rounding (.5) + TPDF dither (< 1) on top of the engine's float result; the reference's
        # own stream-vs-oneshot tolerance for ints is atol=2 (tests/test_resample.py:130)
        assert np.abs(y - orc).max() <= 1.5 + 1e-3 * np.abs(orc).max()
        assert np.allclose(y, soxr.resample(x, s['in_rate'], s['out_rate'], s['quality']), atol=2)
    else:
        assert rms(y - orc) <= TOL_F32
        # bit-identical to the one-shot drivers (tests/test_resample.py:46-57,109-116), at every ratio
        assert np.all(y == soxr._resample_oneshot(x, s['in_rate'], s['out_rate'], s['quality']))
        assert np.all(y == soxr.resample(x, s['in_rate'], s['out_rate'], s['quality']))


@pytest.mark.parametrize('in_rate, out_rate', [(48000, 22050), (8000, 48000)])
@pytest.mark.parametrize('chunk_size', [50, 101])
@pytest.mark.parametrize('length', [1, 101, 12000])
@pytest.mark.parametrize('dtype', ['int32', np.int16])
def test_stream_int(in_rate, out_rate, chunk_size, length, dtype):
    # reference tests/test_resample.py:119-130
    x = (np.random.default_rng(5).standard_normal((length, 2)) * 5000).astype(dtype)
    y_oneshot = soxr._resample_oneshot(x, in_rate, out_rate)
    y_stream, _, _ = _stream(x, in_rate, out_rate, chunk_size, dtype)
    assert np.allclose(y_oneshot, y_stream, atol=2)


@pytest.mark.parametrize('dtype,in_rate,out_rate,quality', [('float32', 44100, 16000, 'HQ'), ('float32', 48000, 44100, 'VHQ'),
                                                            ('int16', 48000, 22050, 'HQ'), ('float64', 44100, 48000, 'VHQ'),
                                                            ('float32', 44100, 47999.5, 'HQ')])
def test_stream_batch_equals_separate_streams(dtype, in_rate, out_rate, quality):
    # ResampleStreamBatch: B streams through ONE handle == B ResampleStreams, per call and bit for bit (ints incl. dither)
    B, ch, chunk, n = 5, 2, 1024, 9000
    scale = 5000.0 if np.dtype(dtype).kind == 'i' else 1.0
    x = make_input(21, (B, n, ch), dtype, scale)
    rb = soxr.ResampleStreamBatch(in_rate, out_rate, B, ch, dtype=dtype, quality=quality)
    singles = [soxr.ResampleStream(in_rate, out_rate, ch, dtype=dtype, quality=quality) for _ in range(B)]
    for i in range(0, n, chunk):
        last = i + chunk >= n
        yb = rb.resample_chunk(np.ascontiguousarray(x[:, i:i + chunk]), last=last)
        assert yb.shape[0] == B and yb.shape[2] == ch
        for b in range(B):
            ys = singles[b].resample_chunk(x[b, i:i + chunk], last=last)
            assert ys.shape == yb[b].shape
            assert np.array_equal(ys, yb[b])
        assert rb.delay() == singles[0].delay()
    if np.dtype(dtype).kind == 'i':
        assert rb.num_clips() == sum(s.num_clips() for s in singles)
    rb.clear()
    y2 = rb.resample_chunk(np.ascontiguousarray(x[:, :3000]), last=True)
    s2 = soxr.ResampleStream(in_rate, out_rate, ch, dtype=dtype, quality=quality)
    if np.dtype(dtype).kind != 'i':   # (after clear() the reference restarts the dither seed from 0, not from its create-time value)
        assert np.array_equal(y2[3], s2.resample_chunk(x[3, :3000], last=True))
    with pytest.raises(ValueError):
        soxr.ResampleStreamBatch(in_rate, out_rate, B, ch, dtype=dtype).resample_chunk(x[:2, :100])


def test_stream_history_moves_from_host_to_device_on_a_large_chunk():
    # small chunks keep their history in mapped pinned host memory; the first chunk above the zero-copy size moves it
    # to the device for good.  The result must not depend on where the history lives.
    x = np.random.default_rng(31).standard_normal((300000, 2)).astype(np.float32)
    want = soxr._resample_oneshot(x, 44100, 32000)
    rs = soxr.ResampleStream(44100, 32000, 2, dtype='float32')
    cuts = [0, 700, 1724, 2000, 2000 + 40000, 2000 + 40000 + 300, 180000, 180017, 300000]   # 40000 x 2 x 4 B = 320 kB > 256 kB
    ys = [rs.resample_chunk(x[a:b], last=(b == 300000)) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate(ys), want)


@pytest.mark.parametrize('in_rate,out_rate,quality,dtype', [(44100, 16000, 'HQ', 'float32'), (48000, 44100, 'VHQ', 'float32'),
                                                            (44100, 47999.5, 'VHQ', 'float64'), (48000, 8000, 'HQ', 'int16'),
                                                            (16000, 44100, 'MQ', 'float32')])
def test_low_latency_stream(in_rate, out_rate, quality, dtype):
    # low_latency=True: every call returns all output whose filter support has arrived (direct-form kernels)
    n, ch, chunk = 30000, 2, 1024
    scale = 5000.0 if np.dtype(dtype).kind == 'i' else 1.0
    x = make_input(41, (n, ch), dtype, scale)
    want = O.resample_f64(x, in_rate, out_rate, Q[quality])
    rs = soxr.ResampleStream(in_rate, out_rate, ch, dtype=dtype, quality=quality, low_latency=True)
    ref_mode = soxr.ResampleStream(in_rate, out_rate, ch, dtype=dtype, quality=quality)
    ys, got_total, fed = [], 0, 0
    for i in range(0, n, chunk):
        last = i + chunk >= n
        y = rs.resample_chunk(x[i:i + chunk], last=last)
        ref_mode.resample_chunk(x[i:i + chunk], last=last)
        fed = min(n, i + chunk)
        got_total += len(y)
        ys.append(y)
        if not last:
            # pending output = what the zero-phase filters still wait for: well under the reference's FFT-block latency
            assert abs(rs.delay() - (fed * out_rate / in_rate - got_total)) < 1e-6
            assert rs.delay() < 700 * max(1.0, out_rate / in_rate)
            assert rs.delay() <= ref_mode.delay() + 1e-9 or ref_mode.delay() < 1
    y = np.concatenate(ys)
    assert y.shape == want.shape and rs.delay() == 0
    if np.dtype(dtype).kind == 'i':
        assert np.abs(y - want).max() <= 1.5 + 1e-3 * np.abs(want).max()
    else:
        assert rms(y.astype(np.float64) - want) <= (TOL_F64 if dtype == 'float64' and quality == 'VHQ' else TOL_F32)
    # independent of the chunking (position-keyed direct form), so also equal to one big chunk
    rs2 = soxr.ResampleStream(in_rate, out_rate, ch, dtype=dtype, quality=quality, low_latency=True)
    if np.dtype(dtype).kind != 'i':
        assert np.array_equal(rs2.resample_chunk(x, last=True), y)


def test_stream_clear_resets_clip_counter_int16():
    # soxr_clear zeroes the clip counter (soxr.c:609-628); the device-side accumulator must follow
    x = np.where(np.sin(2 * np.pi * np.arange(4000) / 100.0) >= 0, 32767, -32768).astype(np.int16)[:, None]
    # full-scale 480 Hz square wave: the band-limited result overshoots (Gibbs) and saturates
    rs = soxr.ResampleStream(48000, 44100, 1, dtype='int16')
    rs.resample_chunk(x, last=True)
    assert rs.num_clips() > 0
    rs.clear()
    assert rs.num_clips() == 0
    rs.resample_chunk(np.zeros((4000, 1), np.int16), last=True)
    assert rs.num_clips() == 0
    n1 = soxr.ResampleStream(48000, 44100, 1, dtype='int16')
    n1.resample_chunk(x, last=True)
    rs.clear()
    rs.resample_chunk(x, last=True)
    assert rs.num_clips() == n1.num_clips()


def test_stream_clear_delay_and_input_after_last():
    rs = soxr.ResampleStream(9600, 4410, 1, dtype='float32')   # libsoxr/tests/1-delay-clear.c:19-64
    x = np.random.default_rng(6).standard_normal(5000).astype(np.float32)
    a = np.concatenate([rs.resample_chunk(x[:2500]), rs.resample_chunk(x[2500:], last=True)])
    with pytest.raises(RuntimeError):
        rs.resample_chunk(x[:10])
    rs.clear()
    assert rs.delay() == 0
    b = rs.resample_chunk(x, last=True)
    assert np.all(a == b)
    assert np.all(a == soxr.resample(x, 9600, 4410))
    with pytest.raises(TypeError):
        soxr.ResampleStream(9600, 4410, 1, dtype='float32').resample_chunk(x.astype(np.float64))
    assert 'b200' in rs.engine()


def make_tone(freq, sr, duration):
    length = int(sr * duration)
    sig = np.sin(2 * np.pi * freq / sr * np.arange(length)) * np.hanning(length)
    return np.stack([sig, np.zeros_like(sig)], axis=-1)


@pytest.mark.parametrize('in_rate,out_rate', [(44100, 22050), (22050, 32000)])
@pytest.mark.parametrize('quality', [soxr.VHQ, 'HQ', 'SOXR_MQ', 'lq', 'soxr_qq'])
def test_quality_sine(in_rate, out_rate, quality):
    # reference tests/test_resample.py:133-156
    x = make_tone(32.0, in_rate, 2.0)
    y = make_tone(32.0, out_rate, 2.0)
    y_pred = soxr.resample(x, in_rate, out_rate, quality=quality)
    y_split = soxr.resample(np.asfortranarray(x), in_rate, out_rate, quality=quality)
    assert np.allclose(y, y_pred, atol=1e-4)
    assert np.allclose(y, y_split, atol=1e-4)


@pytest.mark.parametrize('in_rate,out_rate', [(48000, 24000), (32000, 44100)])
@pytest.mark.parametrize('dtype', [np.int32, np.int16])
def test_int_sine(in_rate, out_rate, dtype):
    # reference tests/test_resample.py:159-176
    x = (make_tone(32.0, in_rate, 2.0) * 16384).astype(dtype)
    y = (make_tone(32.0, out_rate, 2.0) * 16384).astype(dtype)
    y_pred = soxr.resample(x, in_rate, out_rate)
    y_split = soxr.resample(np.asfortranarray(x), in_rate, out_rate)
    y_oneshot = soxr._resample_oneshot(x, in_rate, out_rate)
    assert np.allclose(y, y_pred, atol=2)
    assert np.allclose(y, y_split, atol=2)
    assert np.allclose(y_oneshot, y_pred, atol=2)
    assert np.allclose(y_oneshot, y_split, atol=2)


@pytest.mark.parametrize('num_task', [2, 5, 12])
def test_multithread(num_task):
    # reference tests/test_resample.py:179-191
    from concurrent.futures import ThreadPoolExecutor
    from functools import partial
    x = np.random.default_rng(7).standard_normal((75999, 2)).astype(np.float32)
    with ThreadPoolExecutor() as p:
        results = list(p.map(partial(soxr.resample, in_rate=44100, out_rate=32000), [x] * num_task))
    assert all(np.all(results[0] == r) for r in results[1:])
    xi = (np.random.default_rng(8).standard_normal((70001, 2)) * 5000).astype(np.int16)
    with ThreadPoolExecutor() as p:
        results = list(p.map(partial(soxr.resample, in_rate=32000, out_rate=48000), [xi] * num_task))
    assert all(np.all(results[0] == r) for r in results[1:])   # deterministic seed (soxr.c:439)


def test_batch_matches_per_clip_and_torch_device_path():
    import torch
    x = np.random.default_rng(9).standard_normal((5, 7000, 2)).astype(np.float32)
    yb = soxr.resample_batch(x, 48000, 44100, 'VHQ')                     # single-pass frequency-domain kernel
    ye = soxr.resample_batch(x, 48000, 44100, 'VHQ', exact_grid=True)    # the reference's block grid
    for b in range(5):
        r = soxr.resample(x[b], 48000, 44100, 'VHQ')
        assert np.all(ye[b] == r)
        assert rms(yb[b].astype(np.float64) - r) <= 5e-7
    yt = soxr.resample_batch(torch.from_numpy(x).cuda(), 48000, 44100, 'VHQ')
    torch.cuda.synchronize()
    assert np.all(yt.cpu().numpy() == yb)
    xm = x[:, :, 0].copy()
    ym = soxr.resample_batch(xm, 48000, 44100, 'HQ', exact_grid=True)
    assert np.all(ym[3] == soxr.resample(xm[3], 48000, 44100, 'HQ'))
    # int16: the batch replays each clip's call partition, so exact-grid batches equal soxr.resample bit for bit
    xi = np.round(x[:, :, 0] * 5000).astype(np.int16)
    yi = soxr.resample_batch(xi, 96000, 48000, 'HQ', exact_grid=True)
    for b in range(5):
        assert np.all(yi[b] == soxr.resample(xi[b], 96000, 48000, 'HQ'))
    # a caller-supplied `out` must be exactly what the call would allocate
    good = np.empty_like(yb)
    assert soxr.resample_batch(x, 48000, 44100, 'VHQ', out=good) is good and np.all(good == yb)
    for bad in (np.empty((5, 6430, 2), np.float32), np.empty(yb.shape, np.float64), np.empty((5, 2, 6431), np.float32).transpose(0, 2, 1)):
        with pytest.raises(ValueError):
            soxr.resample_batch(x, 48000, 44100, 'VHQ', out=bad)
    with pytest.raises(ValueError):
        soxr.resample_batch(torch.from_numpy(x).cuda(), 48000, 44100, 'VHQ', out=torch.empty(yb.shape, dtype=torch.float32))
    with pytest.raises(TypeError):
        soxr.resample_batch(x, 48000, 44100, 'VHQ', stream="0")


def test_linearity_and_shift_properties_full_size():
    # size-independent properties at a BASELINE-sized clip (1 s stereo, 48k -> 44.1k VHQ)
    rng = np.random.default_rng(10)
    a = rng.standard_normal((48000, 2)).astype(np.float32)
    b = rng.standard_normal((48000, 2)).astype(np.float32)
    ya, yb = soxr.resample(a, 48000, 44100, 'VHQ'), soxr.resample(b, 48000, 44100, 'VHQ')
    yab = soxr.resample(a + b, 48000, 44100, 'VHQ')
    assert ya.shape == (44100, 2)
    assert rms(yab.astype(np.float64) - ya - yb) <= 2e-6
    # delaying the input by 160 frames delays the output by exactly 147 frames
    a2 = np.concatenate([np.zeros((160, 2), np.float32), a])
    ya2 = soxr.resample(a2, 48000, 44100, 'VHQ')
    assert rms(ya2[147 + 400:44100 - 400].astype(np.float64) - ya[400:44100 - 147 - 400]) <= 1e-6


# ---- single-pass frequency-domain kernel (fdr_stage.cuh): whole-clip float32 VHQ at 160:147 --------
@pytest.mark.parametrize("length", [1, 2, 7, 100, 2239, 2240, 2241, 2400, 4480, 4481, 12345, 48000])
@pytest.mark.parametrize("channels", [1, 2, 3])
def test_fdr_matches_exact_grid_and_oracle(length, channels):
    x = np.random.default_rng(100 + length + channels).standard_normal((3, length, channels)).astype(np.float32)
    y = soxr.resample_batch(x, 48000, 44100, 'VHQ')
    ye = soxr.resample_batch(x, 48000, 44100, 'VHQ', exact_grid=True)
    assert y.shape == ye.shape == (3, soxr_ext.output_frames(48000, 44100, length), channels)
    for b in range(3):
        want = O.resample_f64(x[b], 48000, 44100, O.VHQ)
        assert rms(y[b].astype(np.float64) - want) <= TOL_F32
        assert rms(ye[b].astype(np.float64) - want) <= TOL_F32
        assert np.max(np.abs(y[b].astype(np.float64) - want)) <= 4e-6
        # the exact-grid path is the stream's: bit-identical to a ResampleStream fed the same clip
        rs = soxr.ResampleStream(48000, 44100, channels, dtype='float32', quality='VHQ')
        assert np.all(rs.resample_chunk(x[b], last=True) == ye[b])


def test_fdr_layouts_and_one_shot_drivers():
    x = np.random.default_rng(77).standard_normal((30011, 2)).astype(np.float32)
    want = O.resample_f64(x, 48000, 44100, O.VHQ)
    y_div = soxr.resample(x, 48000, 44100, 'VHQ')                        # C order: divide_proc driver
    y_split = soxr.resample(np.asfortranarray(x), 48000, 44100, 'VHQ')   # Fortran order: split_ch driver
    y_one = soxr._resample_oneshot(x, 48000, 44100, 'VHQ')
    assert np.all(y_div == y_split) and np.all(y_div == y_one)            # reference block grid in all three
    y_fdr = soxr.resample_batch(x[None], 48000, 44100, 'VHQ')[0]          # single-pass kernel, interleaved stereo
    y_fdr_split = np.empty((y_one.shape[0], 2), np.float32, order='F')    # ... and planar (split-channel) buffers
    xf = np.asfortranarray(x)
    soxr_ext.resample_batch_raw(xf.ctypes.data, y_fdr_split.ctypes.data, in_rate=48000, out_rate=44100, quality=soxr.VHQ,
                                dtype_code=0, n_clips=1, channels=2, ilen=len(x), in_strides=(0, 1, len(x)),
                                out_strides=(0, 1, y_one.shape[0]), olen=y_one.shape[0])
    assert np.all(y_fdr == y_fdr_split)
    for y in (y_div, y_fdr):
        assert rms(y.astype(np.float64) - want) <= TOL_F32
    assert rms(y_fdr.astype(np.float64) - y_one) <= 5e-7
    y_m = soxr.resample_batch(x[None, :, 0].copy(), 48000, 44100, 'VHQ')[0]   # mono: block-paired transforms
    assert rms(y_m.astype(np.float64) - want[:, 0]) <= TOL_F32


# ---- the other ratio families of the single-pass kernel, and block-paired mono signals ----------------------
FDR_FAMILIES = [(44100, 48000, 'VHQ'), (96000, 48000, 'HQ'), (96000, 48000, 'VHQ'), (48000, 96000, 'HQ'),
                (22050, 44100, 'VHQ'), (44100, 22050, 'MQ'), (48000, 44100, 'HQ'), (48000, 16000, 'HQ'),
                (48000, 16000, 'VHQ'), (16000, 48000, 'HQ'), (8000, 24000, 'VHQ'), (44100, 16000, 'HQ'),
                (44100, 16000, 'VHQ'), (48000, 32000, 'HQ'), (32000, 48000, 'VHQ'),
                (48000, 8000, 'HQ'), (48000, 8000, 'VHQ'), (8000, 48000, 'HQ'), (48000, 12000, 'HQ'), (12000, 48000, 'HQ'),
                (44100, 32000, 'HQ'), (44100, 32000, 'VHQ'), (32000, 44100, 'HQ'), (88200, 48000, 'HQ'), (44100, 24000, 'VHQ'),
                (48000, 88200, 'HQ'), (96000, 44100, 'HQ'), (48000, 22050, 'VHQ'), (44100, 96000, 'HQ'), (16000, 44100, 'HQ')]


@pytest.mark.parametrize("in_rate,out_rate,quality", FDR_FAMILIES)
@pytest.mark.parametrize("length", [1, 5, 300, 2143, 2144, 2145, 2646, 3215, 3217, 4289, 9999, 30000])
@pytest.mark.parametrize("channels", [1, 2, 3])
def test_fdr_families_match_exact_grid_and_oracle(in_rate, out_rate, quality, length, channels):
    x = np.random.default_rng(length + 7 * channels + in_rate).standard_normal((2, length, channels)).astype(np.float32)
    y = soxr.resample_batch(x, in_rate, out_rate, quality)
    ye = soxr.resample_batch(x, in_rate, out_rate, quality, exact_grid=True)
    assert y.shape == ye.shape == (2, soxr_ext.output_frames(in_rate, out_rate, length), channels)
    q = {'HQ': O.HQ, 'VHQ': O.VHQ, 'MQ': O.MQ}[quality]
    for b in range(2):
        want = O.resample_f64(x[b], in_rate, out_rate, q)
        assert rms(y[b].astype(np.float64) - want) <= TOL_F32
        assert rms(ye[b].astype(np.float64) - want) <= TOL_F32
        assert rms(y[b].astype(np.float64) - ye[b]) <= TOL_F32
    # a mono clip's result (two of its own blocks per transform) does not depend on its batch neighbours
    if channels == 1:
        y0 = soxr.resample_batch(x[1:2], in_rate, out_rate, quality)
        assert np.all(y0[0] == y[1])


def test_fdr_mono_int16_batch_cfg4_shape():
    # BASELINE config 4 in small: int16 96k -> 48k HQ mono clips; against the reference golden drivers' tolerance
    x = np.round(np.random.default_rng(3).standard_normal((6, 20000)) * 5000).astype(np.int16)
    y = soxr.resample_batch(x, 96000, 48000, 'HQ')
    ye = soxr.resample_batch(x, 96000, 48000, 'HQ', exact_grid=True)
    assert y.dtype == np.int16 and y.shape == (6, 10000)
    for b in range(6):
        want = O.resample_f64(x[b].astype(np.float64), 96000, 48000, O.HQ)
        assert np.max(np.abs(y[b] - want)) <= 2.0          # rounding + TPDF dither (<= 1 LSB each side)
        assert np.max(np.abs(y[b].astype(np.int64) - ye[b])) <= 2
    import torch
    yt = soxr.resample_batch(torch.from_numpy(x).cuda(), 96000, 48000, 'HQ')
    assert np.all(yt.cpu().numpy() == y)


@pytest.mark.parametrize("in_rate,out_rate,quality,channels,n_clips", [(48000, 44100, 'VHQ', 2, 330), (96000, 48000, 'HQ', 1, 700),
                                                                       (44100, 48000, 'HQ', 2, 300)])
def test_fdr_many_blocks_per_cta_is_batch_invariant(in_rate, out_rate, quality, channels, n_clips):
    # more (pair, block) items than resident CTAs: every CTA walks several consecutive blocks (exchange buffers
    # swap roles, the next block is staged by TMA).  A clip's result must not depend on where in the batch it sits.
    x = np.random.default_rng(5).standard_normal((n_clips, 20000, channels)).astype(np.float32)
    y = soxr.resample_batch(x, in_rate, out_rate, quality)
    for b in (0, n_clips // 2, n_clips - 1):
        assert np.all(y[b] == soxr.resample_batch(x[b:b + 1], in_rate, out_rate, quality)[0])
    q = {'HQ': O.HQ, 'VHQ': O.VHQ}[quality]
    want = O.resample_f64(x[n_clips - 1], in_rate, out_rate, q)
    assert rms(y[n_clips - 1].astype(np.float64) - want) <= TOL_F32


def test_describe_path_names_the_kernels():
    # no silent fall-through: which kernels serve a plan is one call away
    d = soxr_ext.describe_path(48000, 44100, soxr.VHQ, 'float32')
    assert 'k_fft_stage' in d and 'vpoly0 fused' in d and 'k_fdr[2560 -> 2352' in d
    assert 'b200-cr64' in soxr_ext.describe_path(48000, 44100, soxr.VHQ, 'float64')
    assert 'k_fdr[3528 -> 2560' in soxr_ext.describe_path(44100, 32000, soxr.HQ, 'float32')
    assert 'k_poly_interp' in soxr_ext.describe_path(44100, 47999.5, soxr.VHQ, 'float32')
    assert 'same as exact grid' in soxr_ext.describe_path(44100, 47999.5, soxr.VHQ, 'float32')


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,in_rate,out_rate,quality", [("float32", 48000, 44100, "VHQ"), ("int16", 96000, 48000, "HQ")])
def test_pageable_host_batches_equal_pinned_and_device(dtype, in_rate, out_rate, quality):
    """A host batch large enough for the chunked H2D -> compute -> D2H pipeline (several 16 MB chunks): ordinary
    pageable arrays go through the pinned slots of the host copy pool (non-temporal copies), page-locked arrays are
    DMA'd in place, torch CUDA tensors stay on the device -- the three must agree bit for bit."""
    import torch
    rng = np.random.default_rng(17)
    n, frames = 260, (24000 if dtype == "float32" else 48000)      # 50 MB / 25 MB in: several chunks
    if dtype == "float32":
        x = rng.standard_normal((n, frames, 2)).astype(np.float32)
    else:
        x = np.round(rng.standard_normal((n, frames)) * 5000).astype(np.int16)
    y_page = soxr.resample_batch(x, in_rate, out_rate, quality)
    xp = torch.from_numpy(x).pin_memory().numpy()
    yp = torch.empty(y_page.shape, dtype=torch.from_numpy(y_page).dtype).pin_memory().numpy()
    soxr.resample_batch(xp, in_rate, out_rate, quality, out=yp)
    y_dev = soxr.resample_batch(torch.from_numpy(x).cuda(), in_rate, out_rate, quality).cpu().numpy()
    assert np.array_equal(y_page, yp)
    assert np.array_equal(y_page, y_dev)
