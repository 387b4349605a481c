"""Synthetic workloads of SURVEY.md 8(d): palettes, category statistics, scan shapes.

Host-side only (tiny): the palettes and statistics are *inputs* handed to both the CUDA path
and the CPU oracle, so nothing here needs a CPU/GPU twin.  The scans themselves are produced
on the device by `Context.synth_scan_dev` (csrc/mr_synth.cu); the oracle owns the identical
CPU generator used by the tests.
"""
import numpy as np

from .api import Palette, component_stats, rgb8_to_lab

SEED_BASE = 0x4D520000   # + config number

# BASELINE.json configs: (n1, n2, K, R, colourspace, n_sheets)
CONFIGS = {
    1: dict(n1=2000, n2=2000, K=8, R=1, colourspace="rgb", n_sheets=1),
    2: dict(n1=20000, n2=20000, K=16, R=2, colourspace="rgb", n_sheets=1),
    3: dict(n1=60000, n2=40000, K=16, R=2, colourspace="rgb", n_sheets=1),
    4: dict(n1=20000, n2=20000, K=64, R=3, colourspace="lab", n_sheets=1),
    5: dict(n1=4096, n2=4096, K=16, R=2, colourspace="rgb", n_sheets=256),
}


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _hash_u64(seed, n, stream=0):
    with np.errstate(over="ignore"):
        base = np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(stream) * np.uint64(0x9E3779B1)
        return _splitmix64(base + np.arange(n, dtype=np.uint64))


def make_palette(seed, K, kind="farthest"):
    """K N0f8 colours.  'farthest': seeded farthest-point sampling among 4096 hashed candidates
    (integer squared distances, ties -> lowest candidate index).  'lattice': colours on a
    spacing-32 lattice so that many colours are exactly equidistant from two means."""
    if kind == "lattice":
        g = np.array([32, 96, 160, 224], dtype=np.uint8)
        pts = np.array([(r, gg, b) for b in g for gg in g for r in g], dtype=np.uint8)
        order = np.argsort(_hash_u64(seed, len(pts), 7), kind="stable")
        return np.ascontiguousarray(pts[order[:K]])
    h = _hash_u64(seed, 4096, 1)
    cand = np.stack([(h >> np.uint64(s)) & np.uint64(255) for s in (0, 8, 16)], axis=1).astype(np.int64)
    chosen = [0]
    d = ((cand - cand[0]) ** 2).sum(axis=1)
    for _ in range(1, K):
        nxt = int(np.argmax(d))
        chosen.append(nxt)
        d = np.minimum(d, ((cand - cand[nxt]) ** 2).sum(axis=1))
    return np.ascontiguousarray(cand[chosen].astype(np.uint8))


def _noise(seed, n, stream):
    """Irwin-Hall noise with sigma ~ 6 N0f8 units (same shape as the scan generator's)."""
    h = _hash_u64(seed, n, stream)
    s = sum(((h >> np.uint64(8 * k)) & np.uint64(255)).astype(np.int64) for k in range(4)) - 510
    return ((s * 333 + 4096 + (1 << 20)) >> 13) - 128


def make_stats(seed, palette_rgb, n_samples=32, tol=2.0, colourspace="rgb", exact_means=False):
    """Category statistics as `_component_stats` would produce from `n_samples` clicked pixels
    per category (noisy samples around each palette colour) -> Palette."""
    pal = np.asarray(palette_rgb, dtype=np.int64)
    K = pal.shape[0]
    if exact_means:   # adversarial: means exactly on N0f8 values (exact distance ties)
        m = pal / 255.0
        return Palette(m, (pal - 12) / 255.0, (pal + 12) / 255.0, np.full((K, 3), 6.0 / 255.0),
                       np.full(K, tol), "rgb")
    mean, mn, mx, sd = (np.empty((K, 3)) for _ in range(4))
    for c in range(K):
        nz = np.stack([_noise(seed, n_samples, 100 + 3 * c + ch) for ch in range(3)], axis=1)
        samples = np.clip(pal[c][None, :] + nz, 0, 255).astype(np.uint8)
        vals = rgb8_to_lab(samples) if colourspace == "lab" else samples.astype(np.float64) / 255.0
        st = component_stats(vals)
        mean[c], mn[c], mx[c], sd[c] = st["mean"], st["min"], st["max"], st["sd"]
    return Palette(mean, mn, mx, sd, np.full(K, tol), colourspace)


def config_workload(cfg):
    """(seed, palette_rgb, Palette, spec) of BASELINE.json config number `cfg`."""
    spec = CONFIGS[cfg]
    seed = SEED_BASE + cfg
    pal = make_palette(seed, spec["K"])
    stats = make_stats(seed, pal, colourspace=spec["colourspace"])
    return seed, pal, stats, spec
