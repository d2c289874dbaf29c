"""Generates tests/golden/vr_vectors.npz + vr_meta.json: variable-rate streams (`ResampleStream(vr=True)`,
`set_io_ratio`) run through the UNMODIFIED reference (oracle/_ref/libsoxr_ref.so, vr32.c) by oracle/ref_soxr.py's
exact mirror of the binding's CSoxr::process.  Run here (needs oracle/_ref): python tests/golden/make_vr_golden.py

A scenario is a list of actions: ["ratio", in_rate, out_rate, slew_len] | ["chunk", frames, last]; chunk k of a
scenario is seeded noise (seed = scenario seed + k).  Stored: every call's output, concatenated, and its length.
The first two scenarios follow the reference's own tests/vr.py (rates, ratio changes, chunk sizes)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_soxr as R  # noqa: E402

SCENARIOS = [
    # tests/vr.py:62-110 (10:1 maximum, two channels, ratio 5 -> 10 -> 5, 2500-frame chunks)
    dict(name="vrpy_down", in_rate=10, out_rate=1, ch=2, dtype="float32", seed=100, scale=1.0, actions=[
        ["ratio", 5, 1, 1], ["chunk", 0, False], ["chunk", 10, False], ["chunk", 0, False],
        ["chunk", 2500, False], ["chunk", 2500, False], ["ratio", 10, 1, 1], ["chunk", 2500, False],
        ["chunk", 2500, False], ["ratio", 5, 1, 1], ["chunk", 2500, False], ["chunk", 2500, False],
        ["chunk", 2500, False], ["chunk", 2500, True]]),
    # tests/vr.py:14-52 (1:20 maximum, mono, ratio 10 -> 1 -> 10, 1000-frame chunks)
    dict(name="vrpy_up", in_rate=1, out_rate=20, ch=1, dtype="float32", seed=200, scale=1.0, actions=[
        ["ratio", 10, 1, 1], ["chunk", 1000, False], ["chunk", 1000, False], ["ratio", 1, 1, 1],
        ["chunk", 1000, False], ["chunk", 1000, False], ["ratio", 10, 1, 1], ["chunk", 1000, False],
        ["chunk", 1000, True]]),
    # slews across octaves in both directions: stage switches, fades and cross-fades between the decimators
    dict(name="slew_octaves", in_rate=6, out_rate=1, ch=2, dtype="float32", seed=300, scale=1.0, actions=[
        ["chunk", 3000, False], ["ratio", 3.3, 1, 2000], ["chunk", 4000, False], ["chunk", 4000, False],
        ["ratio", 1.7, 1, 500], ["chunk", 3000, False], ["ratio", 0.8, 1, 3000], ["chunk", 3000, False],
        ["chunk", 3000, False], ["ratio", 0.35, 1, 0], ["chunk", 2000, False], ["ratio", 5.9, 1, 4000],
        ["chunk", 6000, False], ["chunk", 6000, False], ["chunk", 6000, True]]),
    # small real-time chunks, immediate ratio steps
    dict(name="rt_chunks", in_rate=48000, out_rate=44100, ch=1, dtype="float32", seed=400, scale=1.0, actions=(
        [["chunk", 480, False]] * 12 + [["ratio", 48000, 47000, 0]] + [["chunk", 480, False]] * 12 +
        [["ratio", 48000, 46000, 960]] + [["chunk", 480, False]] * 12 + [["chunk", 480, True]])),
    # integer I/O: the engine runs in float32 whatever the datatype; int16 is dithered with the reference's
    # data-salted generator, so equality here needs every float bit upstream to be equal
    dict(name="int16_down", in_rate=3, out_rate=1, ch=2, dtype="int16", seed=500, scale=5000.0, actions=[
        ["chunk", 3000, False], ["ratio", 2.2, 1, 1000], ["chunk", 3000, False], ["chunk", 3000, True]]),
    dict(name="int32_up", in_rate=1, out_rate=2, ch=1, dtype="int32", seed=600, scale=1.0e8, actions=[
        ["chunk", 2000, False], ["ratio", 0.75, 1, 0], ["chunk", 2000, True]]),
]


def chunk_input(sc, k, frames):
    x = np.random.default_rng(sc["seed"] + k).standard_normal((frames, sc["ch"])) * sc["scale"]
    dt = np.dtype(sc["dtype"])
    return np.round(x).astype(dt) if dt.kind == "i" else x.astype(dt)


def run(sc, make_stream, chunk_fn):
    rs = make_stream(sc)
    outs, k = [], 0
    for a in sc["actions"]:
        if a[0] == "ratio":
            rs.set_io_ratio(a[1], a[2], int(a[3]))
        else:
            outs.append(chunk_fn(rs, chunk_input(sc, k, a[1]), a[2]))
            k += 1
    return outs


def main():
    vec, meta = {}, []
    for sc in SCENARIOS:
        outs = run(sc, lambda s: R.RefStream(s["in_rate"], s["out_rate"], s["ch"], s["dtype"], "HQ", vr=True),
                   lambda rs, x, last: rs.process_vr(x, last))
        vec[sc["name"]] = np.concatenate(outs, axis=0)
        m = dict(sc)
        m["lengths"] = [int(o.shape[0]) for o in outs]
        meta.append(m)
        print(sc["name"], m["lengths"])
    np.savez_compressed(os.path.join(HERE, "vr_vectors.npz"), **vec)
    with open(os.path.join(HERE, "vr_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
