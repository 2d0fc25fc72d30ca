"""CPU suite: the numpy restatement of the initial-conditions step (oracle/ics_np.py) against
the reference's own IC scalars (tests/golden/ic_scalars.npz, made from oracle/_ref) and against
the defining properties of a Zel'dovich field."""
import numpy as np
import pytest

from conftest import golden
from oracle import ics_np


def test_power_spectrum_and_growth_match_reference():
    g = golden("ic_scalars.npz")
    k = g["k"]
    # the reference interpolates a 1000-point log-log table (initial_conditions.cpp:173-199): 1e-5 apart
    assert np.max(np.abs(ics_np.power(k) / g["pk_z49"] - 1.0)) < 5e-5
    assert np.array_equal(g["pk_z49"], g["pk_z0"])          # normalised at z = 0 whatever z_initial is
    for z in (49, 9, 0):
        a = 1.0 / (1.0 + z)
        d, f, h = g[f"dfh_z{z}"]
        assert abs(ics_np.growth(a) / d - 1.0) < 1e-13
        assert abs(ics_np.rate(a) / f - 1.0) < 1e-13
        assert abs(ics_np.hubble(a) / h - 1.0) < 1e-13


def test_sigma8_of_normalised_spectrum():
    l0, l1, n = np.log(0.001), np.log(100.0), 1000
    dl = (l1 - l0) / n
    k = np.exp(l0 + (np.arange(n) + 0.5) * dl)
    kr = 8.0 * k
    w = 3.0 * (np.sin(kr) - kr * np.cos(kr)) / kr ** 3
    s8 = np.sqrt((ics_np.power(k) * w * w * k ** 3 * dl).sum() / (2 * np.pi ** 2))
    assert abs(s8 - 0.81) < 1e-12


def test_white_noise_statistics():
    w = ics_np.white_noise(32, 12345).ravel()
    n = w.size
    assert abs(w.mean()) < 4.0 / np.sqrt(n) and abs(w.var() - 1.0) < 4.0 * np.sqrt(2.0 / n)
    assert abs(np.mean(w[1:] * w[:-1])) < 4.0 / np.sqrt(n)           # neighbouring cells uncorrelated
    assert not np.array_equal(w, ics_np.white_noise(32, 12346).ravel())
    assert np.all(np.isfinite(w)) and np.abs(w).max() < 6.0


def test_displacement_is_minus_inverse_divergence_of_delta():
    """div psi = -delta, checked with a spectral derivative taken independently of the construction:
    FFT the real-space psi back and contract with i k."""
    G, box = 16, 100.0
    psi, delta_k = ics_np.displacement_field(G, box, 7)
    V = box ** 3
    n = np.fft.fftfreq(G, 1.0 / G)
    nz = np.arange(G // 2 + 1)
    dk = 2 * np.pi / box
    ks = np.meshgrid(n * dk, n * dk, nz * dk, indexing="ij")
    div_k = sum(1j * ks[a] * np.fft.rfftn(psi[a]) * V / G ** 3 for a in range(3))
    # Nyquist planes were dropped from psi: compare away from them
    keep = np.ones(div_k.shape, bool)
    keep[G // 2, :, :] = keep[:, G // 2, :] = keep[:, :, G // 2] = False
    keep[0, 0, 0] = False
    assert np.max(np.abs(div_k[keep] + delta_k[keep])) < 1e-9 * np.abs(delta_k[keep]).max()


def test_measured_power_matches_input():
    """|delta_k|^2 / V averaged in k-shells follows P(k) (sample variance ~ 1/sqrt(modes))."""
    G, box = 32, 100.0
    acc = {}
    for seed in (1, 2, 3, 4):
        _, dk_ = ics_np.displacement_field(G, box, seed)
        n = np.fft.fftfreq(G, 1.0 / G)
        nz = np.arange(G // 2 + 1)
        kx, ky, kz = np.meshgrid(n, n, nz, indexing="ij")
        kk = np.sqrt(kx ** 2 + ky ** 2 + kz ** 2)
        for lo in (2, 4, 6, 8, 10, 12, 14):
            m = (kk >= lo) & (kk < lo + 2)
            acc.setdefault(lo, []).append((np.abs(dk_[m]) ** 2 / box ** 3, kk[m] * 2 * np.pi / box))
    for lo, parts in acc.items():
        p = np.concatenate([a for a, _ in parts])
        k = np.concatenate([b for _, b in parts])
        ratio = p.mean() / ics_np.power(k).mean()
        assert abs(ratio - 1.0) < 5.0 / np.sqrt(p.size), (lo, ratio, p.size)


def test_zeldovich_particles_layout():
    pos, vel, (rms, D, vfac) = ics_np.zeldovich(16, n_particles=512, shift=50.0)
    assert pos.shape == (512, 3) and pos.dtype == np.float32
    assert pos.min() >= -50.0 and pos.max() < 50.0
    assert abs(D - 0.0199996388) < 1e-9                                 # golden: growth factor at z = 49
    # stride-8 subsample of a 16^3 grid: particle 1 is grid point 8 = (0, 0, 8)
    full, _, _ = ics_np.zeldovich(16, shift=50.0)
    assert np.array_equal(pos, full[::8])
    assert np.allclose(vel, (pos - (np.stack(np.unravel_index(np.arange(512) * 8, (16, 16, 16)), 1) + 0.5)
                             .astype(np.float32) * np.float32(6.25) + 50.0) * np.float32(vfac), atol=2e-2)
    assert 0.05 < rms < 0.5


def _delta_k_of(field, box):
    """delta_k in the generator's convention, delta(x) = (1/V) sum_k delta_k e^{ikx}."""
    G = field.shape[0]
    return np.fft.rfftn(field) * box ** 3 / G ** 3


def test_second_order_source_analytic():
    """lap phi = delta.  One plane wave has no second-order term; two crossed waves
    delta = A cos(k1 x) + B cos(k2 y) give phi,xx = A cos, phi,yy = B cos, S = A B cos(k1 x) cos(k2 y)."""
    G, box = 32, 100.0
    x = (np.arange(G) * box / G)
    k1, k2 = 2 * np.pi / box * 2, 2 * np.pi / box * 3
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    one = 0.3 * np.cos(k1 * X + 0.4)
    assert np.abs(ics_np.second_order_source(_delta_k_of(one, box), G, box)).max() < 1e-14
    two = 0.3 * np.cos(k1 * X) + 0.2 * np.cos(k2 * Y)
    S = ics_np.second_order_source(_delta_k_of(two, box), G, box)
    assert np.abs(S - 0.06 * np.cos(k1 * X) * np.cos(k2 * Y)).max() < 1e-14
    # oblique waves exercise the off-diagonal terms: delta = A cos(k.x) + B cos(q.x),
    # S = A B cos cos (1 - (k.q)^2 / (k^2 q^2))
    kv = 2 * np.pi / box * np.array([1, 2, 0]); qv = 2 * np.pi / box * np.array([2, -1, 3])
    pk, pq = kv[0] * X + kv[1] * Y + kv[2] * Z, qv[0] * X + qv[1] * Y + qv[2] * Z
    obl = 0.3 * np.cos(pk) + 0.2 * np.cos(pq)
    mu2 = (kv @ qv) ** 2 / ((kv @ kv) * (qv @ qv))
    kq = 2 * np.pi / box * np.array([3, 1, 0])
    obl2 = 0.3 * np.cos(pk) + 0.2 * np.cos(kq[0] * X + kq[1] * Y)
    for f, q in ((obl, qv), (obl2, kq)):
        mu2 = (kv @ q) ** 2 / ((kv @ kv) * (q @ q))
        S = ics_np.second_order_source(_delta_k_of(f, box), G, box)
        expect = 0.06 * np.cos(pk) * np.cos(q[0] * X + q[1] * Y + q[2] * Z) * (1.0 - mu2)
        assert np.abs(S - expect).max() < 1e-13, mu2


def test_second_order_displacement_is_gradient_of_inverse_laplacian():
    G, box = 32, 100.0
    x = (np.arange(G) * box / G)
    k1, k2 = 2 * np.pi / box * 2, 2 * np.pi / box * 3
    X, Y, _ = np.meshgrid(x, x, x, indexing="ij")
    S = np.cos(k1 * X) * np.cos(k2 * Y)                       # phi2 = -S / (k1^2 + k2^2)
    psi2 = ics_np.second_order_displacement(S, box)
    kk = k1 * k1 + k2 * k2
    assert np.abs(psi2[0] - k1 * np.sin(k1 * X) * np.cos(k2 * Y) / kk).max() < 1e-12
    assert np.abs(psi2[1] - k2 * np.cos(k1 * X) * np.sin(k2 * Y) / kk).max() < 1e-12
    assert np.abs(psi2[2]).max() < 1e-12


def test_2lpt_correction_is_second_order():
    """D2 psi2 scales as D1^2: small against the first order at z = 49, four times larger (relative) at twice the
    growth; the velocity factor uses f2 = 2 Omega_m^(6/11)."""
    z1 = ics_np.zeldovich(16, z_init=49.0, seed=3)
    l1 = ics_np.zeldovich(16, z_init=49.0, seed=3, use_2lpt=True)
    z2 = ics_np.zeldovich(16, z_init=24.0, seed=3)
    l2 = ics_np.zeldovich(16, z_init=24.0, seed=3, use_2lpt=True)

    def wrapped(a, b):
        d = np.abs(a.astype(np.float64) - b)
        return np.minimum(d, 100.0 - d)
    c1 = np.sqrt((wrapped(l1[0], z1[0]) ** 2).sum(1).mean())
    c2 = np.sqrt((wrapped(l2[0], z2[0]) ** 2).sum(1).mean())
    assert 0 < c1 < 0.05 * z1[2][0]
    ratio = (c2 / c1) / (ics_np.growth(1 / 25.0) / ics_np.growth(1 / 50.0)) ** 2
    assert abs(ratio - 1.0) < 0.02, ratio
    D2, vf2 = ics_np.growth2(1 / 50.0)
    assert abs(D2 / (-3.0 / 7.0 * ics_np.growth(1 / 50.0) ** 2) - 1.0) < 1e-3
    assert abs(vf2 / (2.0 * z1[2][2]) - 1.0) < 0.01           # f2 ~ 2 f1 in the matter era
