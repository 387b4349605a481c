"""Shared helpers for the parity tests (oracle = checker, libmaprast = thing under test)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def golden():
    return json.load(open(os.path.join(HERE, "golden", "runtests_vectors.json")))


def golden_palette(tol=None):
    g = golden()
    cats = g["categories"]
    arr = lambda k: np.array([[c[ch][k] for ch in "rgb"] for c in cats])  # noqa: E731
    return arr("mean"), arr("min"), arr("max"), arr("sd"), (g["tol"] if tol is None else tol), g["vectors"]


def random_palette(rng, K, C=3, spread=0.08):
    """Random but plausible stats: means in [0,1], boxes of random half-width, random tolerances."""
    mean = rng.random((K, C))
    half = rng.random((K, C)) * spread + 0.01
    mn, mx = mean - half, mean + half
    sd = rng.random((K, C)) * spread / 2 + 0.005
    tol = rng.random(K) * 3.0
    return mean, mn, mx, sd, tol


def mismatches(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return int((a != b).sum())
