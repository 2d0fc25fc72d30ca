"""-m gpu: the "fixed physics" Barnes-Hut mode: no orphans, data-fitted root cube,
real masses in leaf pairs, eps parameter.  Topology bit-exact against the oracle's fixed builder,
forces against the oracle's fixed walk, and -- the point of the mode -- against the FP64 direct sum."""
import numpy as np
import pytest
import torch

from inputs import clustered_np, masses_np, rel_l2, uniform_mt, uniform_np

pytestmark = pytest.mark.gpu

KEYS = ("level", "first_child", "part_off", "part_idx", "center", "size", "com", "mass")


def _posm(pos, mass):
    return torch.from_numpy(np.concatenate([pos, mass[:, None]], 1).astype(np.float32)).cuda()


def _cases():
    yield "uniform", uniform_mt(20000, seed=3), masses_np(20000, seed=4), 8, 20
    yield "box_convention", uniform_np(6000, seed=5, lo=0.0, hi=100.0), masses_np(6000, seed=6), 8, 20
    yield "clustered_small_leaves", clustered_np(15000, seed=7), masses_np(15000, seed=8), 3, 20
    yield "shallow", clustered_np(5000, seed=9), np.ones(5000, np.float32), 4, 5
    p = uniform_mt(3000, seed=10)
    p[100:140] = p[50]                                  # 41 coincident particles: a max-depth chain
    yield "duplicates", p, masses_np(3000, seed=11), 8, 20
    yield "tiny", uniform_mt(5, seed=12), masses_np(5, seed=13), 8, 20
    yield "one", uniform_mt(1, seed=14), masses_np(1, seed=15), 8, 20


@pytest.mark.parametrize("name,pos,mass,cap,depth", list(_cases()), ids=[c[0] for c in _cases()])
def test_fixed_tree_topology_and_forces(engine, oracle, name, pos, mass, cap, depth):
    n = pos.shape[0]
    posm = _posm(pos, mass)
    engine.tree_build_fixed_dev(posm, n, cap, depth, eps=0.02)
    acc = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    engine.tree_set_counting(True)
    engine.tree_walk_dev(acc, 0, n, theta=0.5)
    torch.cuda.synchronize()
    cnt = engine.tree_counters()
    engine.tree_set_counting(False)
    t = oracle.tree_build_fixed(pos, mass, cap, depth)
    exp = engine.tree_export()
    for k in KEYS:
        assert np.array_equal(exp[k], getattr(t, k)), k               # bit-exact, centres of mass included
    assert int(t.part_off[-1]) == n                                    # every particle is a leaf member
    want, c0 = oracle.tree_forces_fixed(t, pos, mass, 0.5, 0.02, counters=True)
    assert [int(x) for x in cnt] == [int(x) for x in c0]               # same visits / cells / pairs
    if n > 1:
        assert rel_l2(acc.cpu().numpy(), want) < 1e-5


def test_fixed_tree_converges_to_direct_sum(engine, oracle):
    n = 30000
    pos, mass = clustered_np(n, seed=21), masses_np(n, seed=22)
    ref = oracle.direct_f64(pos, mass, eps=0.01)
    posm = _posm(pos, mass)
    engine.tree_build_fixed_dev(posm, n, 8, 20, eps=0.01)
    acc = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    errs = []
    for theta in (0.7, 0.5, 0.3, 0.2):
        engine.tree_walk_dev(acc, 0, n, theta=theta)
        torch.cuda.synchronize()
        errs.append(rel_l2(acc.cpu().numpy(), ref))
    assert errs[1] < 4e-3 and errs[3] < 2e-4, errs                      # monopole Barnes-Hut error at theta 0.5 / 0.2
    assert errs[0] > errs[1] > errs[2] > errs[3], errs
    # the reference's own tree semantics on the same particles, for the record: ~0.3-0.6 away
    engine.tree_build_dev(posm, n, 100.0, 8, 20)
    engine.tree_walk_dev(acc, 0, n, theta=0.5)
    torch.cuda.synchronize()
    assert rel_l2(acc.cpu().numpy(), ref) > 0.1


def test_fixed_tree_host_entry_and_shards(engine, oracle):
    n = 9001
    pos, mass = uniform_mt(n, seed=31), masses_np(n, seed=32)
    full = engine.tree_forces_fixed_host(pos, mass, theta=0.4, leaf_cap=8, max_depth=20, eps=0.05)
    t = oracle.tree_build_fixed(pos, mass, 8, 20)
    assert rel_l2(full, oracle.tree_forces_fixed(t, pos, mass, 0.4, 0.05)) < 1e-5
    posm = _posm(pos, mass)
    engine.tree_build_fixed_dev(posm, n, 8, 20, eps=0.05)
    parts = []
    for lo, hi in ((0, 3000), (3000, 3001), (3001, n)):                 # target shards: bitwise the same forces
        a = torch.empty((hi - lo, 3), dtype=torch.float32, device="cuda")
        engine.tree_walk_dev(a, lo, hi - lo, theta=0.4)
        parts.append(a.cpu().numpy())
    assert np.array_equal(np.concatenate(parts), full)
    unit = engine.tree_forces_fixed_host(pos, None, theta=0.4, eps=0.05)                 # mass == NULL -> 1
    assert rel_l2(unit, oracle.tree_forces_fixed(oracle.tree_build_fixed(pos, np.ones(n, np.float32)), pos,
                                                 np.ones(n, np.float32), 0.4, 0.05)) < 1e-5


def test_fixed_tree_rejects_bad_eps(engine):
    import b200grav
    posm = _posm(uniform_mt(100, seed=1), np.ones(100, np.float32))
    with pytest.raises(b200grav.B200Error):
        engine.tree_build_fixed_dev(posm, 100, 8, 20, eps=0.0)


def test_fixed_tree_periodic_walk(engine, oracle):
    """Minimum-image walk in a periodic box: same decisions and forces as the oracle's periodic walk, and
    convergence to the periodic (minimum-image) direct sum as theta shrinks."""
    n, box = 20000, 100.0
    pos, mass = uniform_np(n, seed=41, lo=0.0, hi=box), masses_np(n, seed=42)
    posm = _posm(pos, mass)
    engine.tree_build_fixed_dev(posm, n, 8, 20, eps=0.02)
    acc = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    t = oracle.tree_build_fixed(pos, mass, 8, 20)
    try:
        engine.tree_set_periodic(box)
        engine.tree_set_counting(True)
        engine.tree_walk_dev(acc, 0, n, theta=0.5)
        torch.cuda.synchronize()
        cnt = engine.tree_counters()
        engine.tree_set_counting(False)
        want, c0 = oracle.tree_forces_fixed(t, pos, mass, 0.5, 0.02, counters=True, box=box)
        assert [int(x) for x in cnt] == [int(x) for x in c0]
        # in a periodic box the net force is the small residual of many large terms: the two FP32
        # summation orders differ by ~2e-5 of it (1e-6 in the open-boundary cases above)
        assert rel_l2(acc.cpu().numpy(), want) < 1e-4
        ref = oracle.direct_periodic_f32(pos, mass, 0.02, box)
        errs = []
        for theta in (0.5, 0.25):
            engine.tree_walk_dev(acc, 0, n, theta=theta)
            torch.cuda.synchronize()
            errs.append(rel_l2(acc.cpu().numpy(), ref))
        assert errs[0] < 6e-3 and errs[1] < errs[0] * 0.5, errs
        # and it differs from the open-boundary walk (particles near a face feel the far side)
        engine.tree_set_periodic(0.0)
        engine.tree_walk_dev(acc, 0, n, theta=0.5)
        torch.cuda.synchronize()
        assert rel_l2(acc.cpu().numpy(), want) > 1e-2
    finally:
        engine.tree_set_periodic(0.0)
        engine.tree_set_counting(False)


@pytest.mark.parametrize("box", [0.0, 100.0])
def test_fixed_tree_potential_and_energy(engine, oracle, box):
    """Tree potential (the walk accumulating m / r): against the oracle's walk, against the direct-sum
    potential, and as the energy diagnostic of a tree run."""
    n = 30000
    pos = uniform_np(n, seed=51, lo=0.0, hi=100.0) if box > 0 else clustered_np(n, seed=51)
    mass = masses_np(n, seed=52)
    rng = np.random.default_rng(53)
    vel = rng.normal(0, 50, (n, 3)).astype(np.float32)
    posm = _posm(pos, mass)
    engine.tree_build_fixed_dev(posm, n, 8, 20, eps=0.05)
    phi = torch.empty(n, dtype=torch.float32, device="cuda")
    t = oracle.tree_build_fixed(pos, mass, 8, 20)
    try:
        engine.tree_set_periodic(box)
        engine.tree_potential_dev(phi, 0, n, theta=0.5)
        torch.cuda.synchronize()
        got = phi.cpu().numpy()
        want = oracle.tree_potential_fixed(t, pos, mass, 0.5, 0.05, box)
        assert np.max(np.abs(got - want) / want) < 2e-5
        phi_d = torch.empty(n, dtype=torch.float32, device="cuda")
        engine.direct_potential_dev(posm, phi_d, eps=0.05, box=box)
        torch.cuda.synchronize()
        d = phi_d.cpu().numpy()
        assert np.sqrt(((got - d) ** 2).sum() / (d ** 2).sum()) < 2e-3          # monopole error at theta 0.5
        ke, pe = engine.tree_energy_dev(torch.from_numpy(vel).cuda(), 0, n, theta=0.5)
        ke_d, pe_d = engine.energy_dev(posm, torch.from_numpy(vel).cuda(), eps=0.05, box=box)
        assert abs(ke - ke_d) <= 1e-12 * ke_d and abs(pe / pe_d - 1.0) < 1e-3
        # target shards add up
        k2 = p2 = 0.0
        for lo, hi in ((0, 10000), (10000, n)):
            a, b = engine.tree_energy_dev(torch.from_numpy(vel[lo:hi].copy()).cuda(), lo, hi - lo, theta=0.5)
            k2 += a
            p2 += b
        assert abs(k2 - ke) <= 1e-12 * ke and abs(p2 - pe) <= 1e-9 * abs(pe)
        import b200grav
        with pytest.raises(b200grav.B200Error):
            engine.tree_potential_dev(phi, 0, n, theta=0.7)                     # a target could accept its own cell
    finally:
        engine.tree_set_periodic(0.0)
    engine.tree_build_dev(posm, n, 100.0, 8, 20)
    import b200grav
    with pytest.raises(b200grav.B200Error):
        engine.tree_potential_dev(phi, 0, n, theta=0.5)                         # reference-faithful tree: unsupported


def test_fixed_tree_overflow_fails_loudly(engine):
    """The fixed tree's node count has no useful a-priori bound (tight groups of leaf_cap + 1 particles make
    20-level chains); the table holds N/2 internal nodes.  A build that runs out must not yield a plausible
    force: the walk writes NaN and the host entry point reports B200_ERR_NOMEM."""
    import b200grav
    rng = np.random.default_rng(61)
    groups = rng.uniform(-40, 40, (12, 3))
    pos = np.repeat(groups, 9, axis=0) + rng.uniform(-2e-5, 2e-5, (108, 3))        # 12 knots of 9 particles
    pos = pos.astype(np.float32)
    mass = np.ones(len(pos), np.float32)
    with pytest.raises(b200grav.B200Error, match="allocation"):
        engine.tree_forces_fixed_host(pos, mass, theta=0.5, leaf_cap=8, max_depth=20, eps=0.01)
    posm = _posm(pos, mass)
    engine.tree_build_fixed_dev(posm, len(pos), 8, 20, eps=0.01)
    acc = torch.zeros((len(pos), 3), dtype=torch.float32, device="cuda")
    engine.tree_walk_dev(acc, 0, len(pos), theta=0.5)
    torch.cuda.synchronize()
    assert torch.isnan(acc).all()
    with pytest.raises(b200grav.B200Error):
        engine.tree_stats()
    # the same particles with room to spare (leaf capacity 16: no chains) are fine
    out = engine.tree_forces_fixed_host(pos, mass, theta=0.5, leaf_cap=16, max_depth=20, eps=0.01)
    assert np.isfinite(out).all()
