"""-m gpu: the energy diagnostic -- b200_direct_potential_dev / b200_energy_dev
against the CPU restatement of the reference's compute_energy kernel
(src/physics/lambda_cdm_kernels.cu:338-408)."""
import numpy as np
import pytest
import torch

from inputs import masses_np, uniform_mt

pytestmark = pytest.mark.gpu

REL = 2e-6          # FP32 pair terms (rsqrt.approx ~2^-22), FP64 sums on both sides


def _tol(pe0, mass, eps):
    # the i == i term m_i/eps rides in a target's FP32 tile sum before it is subtracted again:
    # an absolute error of ~1e-7 m_i/eps per target, visible only when phi_i << 1/eps (tiny N)
    return REL * abs(pe0) + 1e-7 * float((mass.astype(np.float64) ** 2).sum()) / eps


def _setup(n, unit, seed=7):
    pos = uniform_mt(n, seed=seed)
    mass = np.ones(n, np.float32) * np.float32(2.5) if unit else masses_np(n, seed=seed + 1)
    rng = np.random.default_rng(seed + 2)
    vel = rng.normal(0.0, 100.0, (n, 3)).astype(np.float32)
    posm = torch.from_numpy(np.concatenate([pos, mass[:, None]], 1).astype(np.float32)).cuda()
    return pos, mass, vel, posm


@pytest.mark.parametrize("n,unit,box", [(3000, False, 0.0), (3000, True, 0.0), (5000, False, 100.0),
                                        (20000, True, 100.0), (20000, False, 0.0), (1, False, 0.0), (2, True, 0.0)])
def test_energy_vs_oracle(engine, oracle, n, unit, box):
    pos, mass, vel, posm = _setup(n, unit)
    if box > 0:                      # K6 semantics: positions inside [0, box)
        posm[:, :3] += 50.0
        pos = pos + np.float32(50.0)
    ke, pe = engine.energy_dev(posm, torch.from_numpy(vel).cuda(), eps=0.01, box=box)
    ke0, pe0 = oracle.energy(pos, vel, mass, 0.01, box)
    assert abs(ke - ke0) <= REL * abs(ke0)
    assert abs(pe - pe0) <= _tol(pe0, mass, 0.01)


def test_periodic_energy_power_of_two_box_equal_masses(engine, oracle):
    """box = 64 and n not a multiple of the 512-source tile: the padding slots (parked at 1e18) wrap to distance
    exactly 0 under the FP32 minimum image, so only their zero mass may silence them."""
    n, box = 3001, 64.0
    rng = np.random.default_rng(21)
    pos = rng.uniform(0.0, box, (n, 3)).astype(np.float32)
    mass = np.full(n, 2.5, np.float32)
    vel = rng.normal(0.0, 10.0, (n, 3)).astype(np.float32)
    posm = torch.from_numpy(np.concatenate([pos, mass[:, None]], 1)).cuda()
    ke, pe = engine.energy_dev(posm, torch.from_numpy(vel).cuda(), eps=0.01, box=box)
    ke0, pe0 = oracle.energy(pos, vel, mass, 0.01, box)
    assert abs(ke - ke0) <= REL * abs(ke0)
    assert abs(pe - pe0) <= _tol(pe0, mass, 0.01)


def test_potential_per_particle_and_shards(engine, oracle):
    n = 7001
    pos, mass, vel, posm = _setup(n, unit=False, seed=11)
    phi = torch.empty(n, dtype=torch.float32, device="cuda")
    engine.direct_potential_dev(posm, phi, eps=0.05)
    torch.cuda.synchronize()
    d = pos[None, :64, :].astype(np.float64) - pos[:, None, :].astype(np.float64)       # [j, i, 3]
    r = np.sqrt((d ** 2).sum(-1) + 0.05 ** 2)
    want = (mass[:, None] / r).sum(0) - mass[:64] / 0.05                                  # drop j == i
    got = phi[:64].cpu().numpy()
    assert np.max(np.abs(got - want) / want) < 2e-6
    # three target shards add up to the whole
    ke_all, pe_all = engine.energy_dev(posm, torch.from_numpy(vel).cuda(), eps=0.05)
    ke_s = pe_s = 0.0
    for lo, hi in ((0, 2000), (2000, 2001), (2001, n)):
        k, p = engine.energy_dev(posm, torch.from_numpy(vel[lo:hi].copy()).cuda(), i0=lo, n_targets=hi - lo, eps=0.05)
        ke_s += k
        pe_s += p
    assert abs(ke_s - ke_all) <= 1e-12 * abs(ke_all) and abs(pe_s - pe_all) <= 1e-9 * abs(pe_all)
    ke0, pe0 = oracle.energy(pos, vel, mass, 0.05, 0.0)
    assert abs(pe_all - pe0) <= _tol(pe0, mass, 0.05)
