"""CPU check of the walk's accept-test screening (tree.cu: accept_cell_d).  The reference decides
`size / sqrt(dx*dx + dy*dy + dz*dz) < theta` with one rounding per operation (tree_force_computer.cpp:302-310).
The kernel first looks at `size*size - theta^2 * d2` with a CONTRACTED d2 and only falls back to the reference's
sequence when that difference is within 3e-5 of theta^2 d2.  This restates both in numpy float32 and checks that
whenever the screen decides, it decides as the reference does -- on random cells and on cells placed within a few
ulps of the threshold.  (The kernel itself is held to the oracle's interaction counters in tests/test_gpu_tree.py.)"""
import numpy as np

F = np.float32


def fma(a, b, c):                       # float32 fused multiply-add: the product of two floats is exact in double
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def reference_accept(size, dx, dy, dz, theta):
    d2 = ((dx * dx).astype(F) + (dy * dy).astype(F)).astype(F)
    d2 = (d2 + (dz * dz).astype(F)).astype(F)
    with np.errstate(divide="ignore"):
        return (size / np.sqrt(d2).astype(F)).astype(F) < theta


def screen(size, dx, dy, dz, theta):
    """(decided, accept): the kernel's fast path; ptxas may or may not contract size*size - t, so both are tried."""
    d2 = fma(dz, dz, fma(dy, dy, (dx * dx).astype(F)))
    t = (F(theta) * F(theta) * d2).astype(F)
    out = []
    for diff in (fma(size, size, -t), ((size * size).astype(F) - t).astype(F)):
        decided = np.abs(diff) > (F(3.0e-5) * t).astype(F)
        out.append((decided, diff < 0))
    return out


def _check(size, dx, dy, dz, theta):
    want = reference_accept(size, dx, dy, dz, F(theta))
    n_decided = 0
    for decided, acc in screen(size, dx, dy, dz, theta):
        assert np.array_equal(acc[decided], want[decided])
        n_decided += int(decided.sum())
    return n_decided


def test_screen_agrees_on_random_cells():
    rng = np.random.default_rng(0)
    n = 2_000_000
    size = (F(100.0) / (2.0 ** rng.integers(0, 12, n))).astype(F)          # cell edges of a 100 box
    d = rng.normal(0.0, 1.0, (3, n)) * (size / 0.5 * rng.uniform(0.2, 5.0, n))
    dx, dy, dz = (d[k].astype(F) for k in range(3))
    for theta in (0.3, 0.5, 0.7, 1.0):
        assert _check(size, dx, dy, dz, theta) > 1.9 * n                   # nearly everything is decided fast


def test_screen_never_decides_wrongly_near_the_threshold():
    rng = np.random.default_rng(1)
    n = 1_000_000
    theta = 0.5
    d = rng.normal(0.0, 10.0, (3, n))
    dx, dy, dz = (d[k].astype(F) for k in range(3))
    r = np.sqrt(dx.astype(np.float64) ** 2 + dy.astype(np.float64) ** 2 + dz.astype(np.float64) ** 2)
    for scale in (1e-7, 1e-6, 1e-5, 2e-5, 5e-5, 1e-4):                     # relative distance from size = theta |d|
        size = (theta * r * (1.0 + scale * rng.uniform(-1.0, 1.0, n))).astype(F)
        _check(size, dx, dy, dz, theta)
    # at the threshold to the last ulp the screen must abstain, not guess
    size = (theta * r).astype(F)
    for decided, _ in screen(size, dx, dy, dz, theta):
        assert decided.sum() == 0


def test_zero_distance_opens():
    z = np.zeros(4, F)
    size = np.array([100.0, 1.0, 1e-3, 50.0], F)
    assert not reference_accept(size, z, z, z, F(0.5)).any()               # size / 0 = inf: never accepted
    for decided, acc in screen(size, z, z, z, 0.5):
        assert decided.all() and not acc.any()
