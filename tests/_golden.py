"""Shared helpers for reading tests/golden fixtures (inputs are re-generated from seeds)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
Q = {'QQ': 0, 'LQ': 1, 'MQ': 2, 'HQ': 4, 'VHQ': 6}


def make_input(seed, shape, dtype, scale):
    x = np.random.default_rng(seed).standard_normal(tuple(shape))
    if np.dtype(dtype).kind == 'i':
        return np.round(x * scale).astype(dtype)
    return (x * scale).astype(dtype)


def load_meta():
    with open(os.path.join(GOLDEN, "vectors_meta.json")) as f:
        return json.load(f)


def load_vectors():
    return np.load(os.path.join(GOLDEN, "vectors.npz"))


def load_plans():
    with open(os.path.join(GOLDEN, "plans.json")) as f:
        return json.load(f)


def load_streams():
    with open(os.path.join(GOLDEN, "streams.json")) as f:
        return json.load(f)


def case_input(meta):
    return make_input(meta['seed'], meta['shape'], meta['dtype'], meta['scale'])


def rms(a):
    a = np.asarray(a, dtype=np.float64)
    return float(np.sqrt(np.mean(a * a))) if a.size else 0.0
