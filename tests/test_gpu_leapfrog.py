"""-m gpu parity tests, Lambda-CDM leapfrog (rows L1-L3).  Gate: after 100 KDK
steps positions agree within 1e-4 of the box size (north_star)."""
import numpy as np
import pytest

from conftest import golden
from inputs import masses_np, rel_l2, uniform_mt

pytestmark = pytest.mark.gpu


def _dev(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()


@pytest.mark.parametrize("n", [1, 3, 4, 5, 1001, 40000])
def test_kick_and_drift_bit_exact(engine, oracle, n):
    import torch
    rng = np.random.default_rng(n)
    p = rng.uniform(0, 100, size=(n, 3)).astype(np.float32)
    v = rng.normal(0, 100, size=(n, 3)).astype(np.float32)
    acc = rng.normal(0, 50, size=(n, 3)).astype(np.float32)
    m = masses_np(n)
    posm = _dev(np.concatenate([p, m[:, None]], 1))
    vel, a3 = _dev(v), _dev(acc)
    a, dt = 1.37, np.float32(2e-3)
    # kick, kick, drift in one pass
    engine.leapfrog_dev(posm, vel, a3, n, 2, dt * np.float32(0.5), a, dt, 100.0)
    torch.cuda.synchronize()
    vo, po = v.copy(), p.copy()
    oracle.kick(vo, acc, m, dt * np.float32(0.5), a)
    oracle.kick(vo, acc, m, dt * np.float32(0.5), a)
    oracle.drift(po, vo, dt, 100.0)
    assert np.array_equal(vel.cpu().numpy(), vo)
    assert np.array_equal(posm.cpu().numpy()[:, :3], po)
    assert np.array_equal(posm.cpu().numpy()[:, 3], m)
    # drift without wrap, no kick
    engine.leapfrog_dev(posm, vel, a3, n, 0, 0.0, a, dt, 0.0)
    torch.cuda.synchronize()
    oracle.drift(po, vo, dt, 0.0)
    assert np.array_equal(posm.cpu().numpy()[:, :3], po)


def _max_dev_min_image(a, b, box):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    if box > 0:
        d = np.minimum(d, box - d)
    return float(d.max())


def test_kdk_100_steps_direct(engine, oracle):
    """C1-style run (scaled to 4096 particles so the CPU oracle finishes in seconds):
    100 KDK steps, dt = 1e-4, a: 1 -> 1.84."""
    import b200grav
    g = golden("random_2048.npz")           # reference generate_random_particles: [0,100) + N(0,100) velocities
    p, v, m = g["pos"], g["vel"], g["mass"]
    sim = b200grav.LambdaCDMSimulation(engine, p, v, m, box=100.0, force="direct", eps=0.01)
    for _ in range(100):
        sim.step(1e-4)
    po, vo, ao = oracle.kdk_run(p, v, m, lambda x: oracle.direct_f32(x, m, eps=0.01), 100, 1e-4, box=100.0)
    assert abs(sim.get_scale_factor() - ao) < 1e-12 and 1.8 < ao < 2.0
    assert _max_dev_min_image(sim.positions(), po, 100.0) < 1e-4 * 100.0
    assert rel_l2(sim.velocities(), vo) < 1e-4


def test_kdk_100_steps_tree(engine, oracle):
    import b200grav
    n = 4096
    p = uniform_mt(n, seed=9)
    rng = np.random.default_rng(10)
    v = rng.normal(0, 100, size=(n, 3)).astype(np.float32)
    m = np.ones(n, np.float32)
    sim = b200grav.LambdaCDMSimulation(engine, p, v, m, box=100.0, force="tree", theta=0.5, wrap=False)
    for _ in range(100):
        sim.step(1e-4)

    def f(x):
        t = oracle.tree_build(x, m)
        return oracle.tree_forces(t, x, 0.5)

    po, vo, ao = oracle.kdk_run(p, v, m, f, 100, 1e-4, box=0.0)
    assert abs(sim.get_scale_factor() - ao) < 1e-12
    assert _max_dev_min_image(sim.positions(), po, 0.0) < 1e-4 * 100.0


def test_c1_config_16k_10_steps(engine, oracle):
    """BASELINE config 1: DirectForceComputer, 16 384 uniform particles, 10 leapfrog steps."""
    import b200grav
    n = 16384
    p = uniform_mt(n, seed=42)
    rng = np.random.default_rng(12345)
    v = rng.normal(0, 100, size=(n, 3)).astype(np.float32)
    m = np.ones(n, np.float32)
    sim = b200grav.LambdaCDMSimulation(engine, p, v, m, box=100.0, force="direct", eps=0.01, wrap=False)
    for _ in range(10):
        sim.step(1e-3)
    po, vo, ao = oracle.kdk_run(p, v, m, lambda x: oracle.direct_f32(x, m, eps=0.01), 10, 1e-3, box=0.0)
    assert abs(sim.get_scale_factor() - ao) < 1e-12
    assert _max_dev_min_image(sim.positions(), po, 0.0) < 1e-4 * 100.0


def test_kdk_100_steps_tree_256k(engine, oracle):
    """North-star trajectory gate at scale ("parity run" for config 4): 100 KDK steps of the Barnes-Hut
    path on 2^18 particles, dt = 1e-4, a: 1 -> 1.84, against the restated CPU tree (OpenMP over targets, ~1 s per
    step); positions within 1e-4 of the box."""
    import b200grav
    n = 1 << 18
    p = uniform_mt(n, seed=19)
    rng = np.random.default_rng(20)
    v = rng.normal(0, 100, size=(n, 3)).astype(np.float32)
    m = np.ones(n, np.float32)
    sim = b200grav.LambdaCDMSimulation(engine, p, v, m, box=100.0, force="tree", theta=0.5, wrap=False)
    for _ in range(100):
        sim.step(1e-4)

    def f(x):
        return oracle.tree_forces(oracle.tree_build(x, m), x, 0.5)

    po, vo, ao = oracle.kdk_run(p, v, m, f, 100, 1e-4, box=0.0)
    assert abs(sim.get_scale_factor() - ao) < 1e-12 and 1.8 < ao < 2.0
    assert _max_dev_min_image(sim.positions(), po, 0.0) < 1e-4 * 100.0
    assert rel_l2(sim.velocities(), vo) < 1e-4


def test_kdk_100_steps_direct_16k_vs_reference(engine, oracle):
    """BASELINE config 1's particle count, 100 KDK steps of the direct path.  CPU side: the restated pair loop on all
    cores every step -- bitwise the reference's own (oracle/_ref = TreeForceComputer with one root leaf), which is
    asserted here on the first and on the last positions of the run (the reference itself is single-threaded:
    1.3 s per evaluation)."""
    import b200grav
    from oracle.pyoracle import Ref
    n = 16384
    p = uniform_mt(n, seed=42)
    rng = np.random.default_rng(12345)
    v = rng.normal(0, 100, size=(n, 3)).astype(np.float32)
    m = np.ones(n, np.float32)
    sim = b200grav.LambdaCDMSimulation(engine, p, v, m, box=100.0, force="direct", eps=0.01, wrap=False)
    for _ in range(100):
        sim.step(1e-4)
    po, vo, ao = oracle.kdk_run(p, v, m, lambda x: oracle.direct_f32(x, None, eps=0.01), 100, 1e-4, box=0.0)
    if Ref.available():
        r = Ref()
        for x in (p, po):
            assert np.array_equal(r.direct(x), oracle.direct_f32(x, None, eps=0.01))
    assert abs(sim.get_scale_factor() - ao) < 1e-12
    assert _max_dev_min_image(sim.positions(), po, 0.0) < 1e-4 * 100.0
    assert rel_l2(sim.velocities(), vo) < 1e-4
