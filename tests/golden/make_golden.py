"""tests/golden/make_golden.py -- regenerate the committed golden fixtures.

Runs the UNMODIFIED reference (oracle/_ref/libsoxr_ref.so, compiled from
/root/reference/libsoxr/src by oracle/Makefile) through the ctypes re-statement
of python-soxr's drivers (oracle/ref_soxr.py) on seeded inputs and stores

  plans.json    SOXR_TRACE plan lines per (in_rate, out_rate, quality)
  vectors.npz   reference outputs (inputs are regenerated from the seeds below)
  streams.json  per-chunk output counts and delay() values of ResampleStream

Usage (only where /root/reference or a prebuilt oracle/_ref exists):
    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_soxr as R  # noqa: E402

PLAN_CASES = [
    (48000, 44100, 'HQ'), (48000, 44100, 'VHQ'), (44100, 16000, 'HQ'), (96000, 48000, 'HQ'),
    (44100, 48000, 'VHQ'), (44100, 47999.5, 'VHQ'), (47999, 44100, 'HQ'), (48000, 8000, 'HQ'),
    (8000, 48000, 'HQ'), (44100, 32000, 'HQ'), (32000, 44100, 'HQ'), (48000, 44100, 'LQ'),
    (48000, 44100, 'MQ'), (44100, 22050, 'VHQ'), (22050, 32000, 'HQ'), (48000, 22050, 'HQ'),
    (48000, 24000, 'HQ'), (100, 200, 'HQ'), (44100, 22050, 'MQ'), (44100, 22050, 'LQ'),
    (22050, 32000, 'LQ'), (22050, 32000, 'MQ'), (44100, 22050, 'QQ'), (22050, 32000, 'QQ'),
    (44100, 192000, 'VHQ'), (192000, 44100, 'HQ'), (44100, 65537, 'HQ'), (65537, 44100, 'VHQ'),
    (48000, 47999, 'VHQ'), (8000, 44100, 'HQ'), (44100, 8000, 'VHQ'), (16000, 48000, 'HQ'),
    (48000, 16000, 'VHQ'), (44100, 88200, 'HQ'), (96000, 44100, 'VHQ'), (11025, 48000, 'HQ'),
    (192000, 48000, 'VHQ'), (48000, 32000, 'VHQ'), (32000, 48000, 'VHQ'), (96000, 24000, 'HQ'),
    (44100, 44100, 'HQ'), (48000, 12000, 'HQ'), (8000, 44100, 'LQ'), (8000, 48000, 'LQ'),
]

# name -> (seed, shape, dtype, in_rate, out_rate, quality, scale)
VECTOR_CASES = {
    # BASELINE.json configs, shortened so the fixtures stay small
    'cfg1_f32_48k_44k1_HQ_mono':    (0, (12000,), 'float32', 48000, 44100, 'HQ', 1.0),
    'cfg2_f32_48k_44k1_VHQ_stereo': (1, (6000, 2), 'float32', 48000, 44100, 'VHQ', 1.0),
    'cfg3_f32_44k1_16k_HQ_8ch':     (2, (5000, 8), 'float32', 44100, 16000, 'HQ', 1.0),
    'cfg4_i16_96k_48k_HQ_mono':     (3, (9600,), 'int16', 96000, 48000, 'HQ', 5000.0),
    'cfg5_f64_44k1_48k_VHQ_4ch':    (4, (4410, 4), 'float64', 44100, 48000, 'VHQ', 1.0),
    # widening rows
    'i32_48k_44k1_VHQ_stereo':      (5, (4000, 2), 'int32', 48000, 44100, 'VHQ', 3.0e8),
    'i16_48k_44k1_HQ_stereo':       (6, (4000, 2), 'int16', 48000, 44100, 'HQ', 5000.0),
    'irr_f64_44k1_47999p5_VHQ':     (7, (4000, 2), 'float64', 44100, 47999.5, 'VHQ', 1.0),
    'irr_f32_47999_44k1_HQ':        (8, (4000,), 'float32', 47999, 44100, 'HQ', 1.0),
    'half_f32_48k_8k_HQ':           (9, (8000, 2), 'float32', 48000, 8000, 'HQ', 1.0),
    'up6_f32_8k_48k_HQ':            (10, (2000,), 'float32', 8000, 48000, 'HQ', 1.0),
    'f64_32k_44k1_VHQ':             (11, (3000, 2), 'float64', 32000, 44100, 'VHQ', 1.0),
    'f32_44k1_32k_HQ':              (12, (5000, 2), 'float32', 44100, 32000, 'HQ', 1.0),
    'lq_f32_48k_44k1':              (13, (4000,), 'float32', 48000, 44100, 'LQ', 1.0),
    'mq_f32_44k1_48k':              (14, (4000,), 'float32', 44100, 48000, 'MQ', 1.0),
    'qq_f32_44k1_22k05':            (15, (4000,), 'float32', 44100, 22050, 'QQ', 1.0),
    'tiny_f32_48k_44k1_HQ':         (16, (7,), 'float32', 48000, 44100, 'HQ', 1.0),
    # ratio families of the single-pass frequency-domain kernel, long enough for several blocks
    'fam_f32_44k1_48k_VHQ_stereo':  (40, (9000, 2), 'float32', 44100, 48000, 'VHQ', 1.0),
    'fam_f32_48k_44k1_HQ_stereo':   (41, (9000, 2), 'float32', 48000, 44100, 'HQ', 1.0),
    'fam_f32_96k_48k_VHQ_mono':     (42, (12000,), 'float32', 96000, 48000, 'VHQ', 1.0),
    'fam_f32_48k_96k_HQ_stereo':    (43, (6000, 2), 'float32', 48000, 96000, 'HQ', 1.0),
    'fam_f32_48k_16k_HQ_stereo':    (44, (12000, 2), 'float32', 48000, 16000, 'HQ', 1.0),
    'fam_f32_16k_48k_VHQ_mono':     (45, (5000,), 'float32', 16000, 48000, 'VHQ', 1.0),
    'fam_f32_44k1_16k_VHQ_3ch':     (46, (12000, 3), 'float32', 44100, 16000, 'VHQ', 1.0),
    'fam_i16_48k_16k_HQ_stereo':    (47, (9000, 2), 'int16', 48000, 16000, 'HQ', 5000.0),
    # round-2 families: 441:320, 320:441, 147:80, 80:147, 320:147, 147:320, 160:441
    'fam_f32_44k1_32k_VHQ_stereo':  (50, (14000, 2), 'float32', 44100, 32000, 'VHQ', 1.0),
    'fam_f32_32k_44k1_HQ_stereo':   (51, (9000, 2), 'float32', 32000, 44100, 'HQ', 1.0),
    'fam_f32_88k2_48k_HQ_stereo':   (52, (12000, 2), 'float32', 88200, 48000, 'HQ', 1.0),
    'fam_f32_48k_88k2_HQ_mono':     (53, (6000,), 'float32', 48000, 88200, 'HQ', 1.0),
    'fam_f32_96k_44k1_HQ_stereo':   (54, (12000, 2), 'float32', 96000, 44100, 'HQ', 1.0),
    'fam_f32_22k05_48k_VHQ_3ch':    (55, (5000, 3), 'float32', 22050, 48000, 'VHQ', 1.0),
    'fam_f32_16k_44k1_HQ_mono':     (56, (5000,), 'float32', 16000, 44100, 'HQ', 1.0),
}

STREAM_CASES = {
    # name -> (seed, frames, channels, dtype, in_rate, out_rate, quality, chunk)
    'cfg3_stream_1024': (2, 20000, 8, 'float32', 44100, 16000, 'HQ', 1024),
    'vhq_480':          (21, 12000, 2, 'float32', 48000, 44100, 'VHQ', 480),
    'up_17':            (22, 3000, 1, 'float64', 32000, 44100, 'HQ', 17),
    # one per remaining plan shape: F-domain /2, half-band chain, irrational (vpoly2/3), x6 post stage, LQ, QQ
    'fdec_i16_1000':    (23, 20000, 1, 'int16', 96000, 48000, 'HQ', 1000),
    'half_777':         (24, 30000, 2, 'float32', 48000, 8000, 'HQ', 777),
    'irr_vhq_509':      (25, 9000, 1, 'float64', 44100, 47999.5, 'VHQ', 509),
    'irr_hq_1024':      (26, 9000, 1, 'float32', 47999, 44100, 'HQ', 1024),
    'up6_101':          (27, 5000, 1, 'float32', 8000, 48000, 'HQ', 101),
    'lq_480':           (28, 9000, 2, 'float32', 48000, 44100, 'LQ', 480),
    'qq_300':           (29, 6000, 1, 'float32', 44100, 22050, 'QQ', 300),
    'vhq_dn3_2000':     (30, 30000, 1, 'float32', 48000, 16000, 'VHQ', 2000),
    'hq_up_4096':       (31, 30000, 2, 'float32', 44100, 48000, 'HQ', 4096),
}


def make_input(seed, shape, dtype, scale):
    x = np.random.default_rng(seed).standard_normal(shape)
    if np.dtype(dtype).kind == 'i':
        return np.round(x * scale).astype(dtype)
    return (x * scale).astype(dtype)


def main():
    plans = {}
    for (i, o, q) in PLAN_CASES:
        lines = [l for l in R.trace_plan(i, o, q) if l.startswith('fir_len') or '>>' in l]
        plans[f"{i}:{o}:{q}"] = lines
    with open(os.path.join(HERE, "plans.json"), "w") as f:
        json.dump(plans, f, indent=1)

    out = {}
    meta = {}
    for name, (seed, shape, dtype, ir, orate, q, scale) in VECTOR_CASES.items():
        x = make_input(seed, shape, dtype, scale)
        x2 = x[:, None] if x.ndim == 1 else x
        y, calls, clips = R.divide_proc(ir, orate, x2, q, return_calls=True)
        out[name] = y[:, 0] if x.ndim == 1 else y
        meta[name] = dict(seed=seed, shape=list(shape), dtype=dtype, in_rate=ir, out_rate=orate,
                          quality=q, scale=scale, calls=calls, clips=int(clips))
        if np.dtype(dtype).kind == 'i':
            # (b) rounding-only variant and (c) the engine's float stream (exact: the
            # int->float io scale is a power of two, soxr.c:435-436)
            ynd = R.divide_proc(ir, orate, x2, q, io_flags=R.NO_DITHER)
            out[name + '__nodither'] = ynd[:, 0] if x.ndim == 1 else ynd
            eng_dt = np.float32 if q in ('HQ', 'LQ', 'MQ', 'QQ') else np.float64
            yf = R.divide_proc(ir, orate, x2, q, out_dtype=eng_dt)
            full = 32768.0 if dtype == 'int16' else 65536.0 * 32768.0
            yf = (yf.astype(np.float64) * full).astype(eng_dt)
            out[name + '__engine'] = yf[:, 0] if x.ndim == 1 else yf
    np.savez(os.path.join(HERE, "vectors.npz"), **out)

    streams = {}
    for name, (seed, n, ch, dtype, ir, orate, q, chunk) in STREAM_CASES.items():
        x = make_input(seed, (n, ch), dtype, 5000.0 if np.dtype(dtype).kind == 'i' else 1.0)
        rs = R.RefStream(ir, orate, ch, dtype, q)
        counts, delays = [], []
        for idx in range(0, n, chunk):
            last = idx + chunk >= n
            y = rs.resample_chunk(x[idx:idx + chunk], last)
            counts.append(int(y.shape[0]))
            delays.append(float(rs.delay()))
        streams[name] = dict(seed=seed, frames=n, channels=ch, dtype=dtype, in_rate=ir, out_rate=orate,
                             quality=q, chunk=chunk, counts=counts, delays=delays)
    with open(os.path.join(HERE, "streams.json"), "w") as f:
        json.dump(streams, f)
    with open(os.path.join(HERE, "vectors_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", len(plans), "plans,", len(out), "vectors,", len(streams), "streams")


if __name__ == "__main__":
    main()
