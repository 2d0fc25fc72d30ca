"""Seeded synthetic inputs shared by the tests and bench.py.

uniform_mt reproduces, in numpy, libstdc++'s
    std::mt19937 g(seed); std::uniform_real_distribution<float> u(lo, hi);
    for each particle: x = u(g), y = u(g), z = u(g)
so the tree-statistics known answers of SURVEY 8c (N = 16 384 -> 4 793 nodes /
4 194 leaves / depth 6) can be checked without the reference.
"""
import numpy as np


def mt19937_u32(seed, count):
    bg = np.random.MT19937()
    bg._legacy_seeding(seed)          # init_genrand(seed) == std::mt19937(seed)
    return bg.random_raw(count).astype(np.uint32)


def uniform_mt(n, seed=42, lo=-50.0, hi=50.0):
    raw = mt19937_u32(seed, 3 * n)
    # generate_canonical<float, 24>: one 32-bit draw, float(draw) / 2^32, clamped below 1
    c = raw.astype(np.float32) / np.float32(4294967296.0)
    c = np.where(c >= np.float32(1.0), np.nextafter(np.float32(1.0), np.float32(0.0)), c).astype(np.float32)
    v = c * np.float32(hi - lo) + np.float32(lo)
    return v.astype(np.float32).reshape(n, 3)


def uniform_np(n, seed=0, lo=-50.0, hi=50.0):
    rng = np.random.default_rng(seed)
    return rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)


def masses_np(n, seed=1, lo=0.5, hi=1.5):
    rng = np.random.default_rng(seed)
    return rng.uniform(lo, hi, size=n).astype(np.float32)


def clustered_np(n, seed=2, box=100.0, nblobs=32, sigma=2.0):
    """Gaussian blobs: stresses deep trees and max-depth overflow leaves."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-0.4 * box, 0.4 * box, size=(nblobs, 3))
    which = rng.integers(0, nblobs, size=n)
    p = c[which] + rng.normal(0.0, sigma, size=(n, 3))
    return np.clip(p, -0.5 * box + 1e-3, 0.5 * box - 1e-3).astype(np.float32)


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))
