"""CPU suite: the oracle's periodic direct sum and leapfrog against tests/golden/ref_gpu_kernels.npz -- outputs of
the reference's OWN CUDA kernels K2 (compute_forces_tiled) and K4 (leapfrog_update), recompiled for sm_100a and run
on a B200 by tests/golden/make_golden_gpu.py.  They were built as the reference builds them (-O3 --use_fast_math):
approximate rsqrt / reciprocals and contracted multiply-adds, so agreement is to rounding, not bit-for-bit."""
import importlib.util
import os

import numpy as np

from conftest import GOLDEN, golden
from inputs import rel_l2


def _gen():
    spec = importlib.util.spec_from_file_location("make_golden_gpu", os.path.join(GOLDEN, "make_golden_gpu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_oracle_periodic_direct_matches_reference_gpu_kernel(oracle):
    mg, g = _gen(), golden("ref_gpu_kernels.npz")
    posm, _, _ = mg.inputs()
    a = oracle.direct_periodic_f32(posm[:, :3].copy(), posm[:, 3].copy(), mg.EPS, mg.BOX)
    # both are sequential FP32 sums over 10 240 sources (measured 1.4e-5); K2 stores a_i * m_i (:217-219)
    assert rel_l2(a * posm[:, 3:4], g["forces"]) < 3e-5


def test_oracle_leapfrog_matches_reference_gpu_kernel(oracle):
    mg, g = _gen(), golden("ref_gpu_kernels.npz")
    posm, vel, force = mg.inputs()
    n = mg.N_LEAP
    m = posm[:n, 3].copy()
    vo = vel.copy()
    oracle.kick(vo, (force / m[:, None]).astype(np.float32), m, np.float32(mg.DT * 0.5), mg.A)
    assert np.abs(vo - g["kick_vel"]).max() <= 4e-7 * np.abs(g["kick_vel"]).max()
    po = posm[:n, :3].copy()
    oracle.drift(po, g["kick_vel"].copy(), np.float32(mg.DT), mg.BOX)
    d = np.abs(po - g["drift_pos"])
    assert np.minimum(d, mg.BOX - d).max() <= 2e-5              # 2 ulp at 100
    assert g["drift_pos"].min() >= 0.0 and g["drift_pos"].max() < mg.BOX


def test_oracle_energy_matches_reference_gpu_kernel(oracle):
    """K6 compute_energy behind launch_energy_computation (lambda_cdm_kernels.cu:338-408, 492-516), run on a B200:
    float pair terms, float per-thread sums over up to 10 239 partners, float atomics across blocks -- the restated
    energy (FP64 sums) agrees to the reference's own round-off."""
    mg, g = _gen(), golden("ref_gpu_energy.npz")
    posm, _, _ = mg.inputs()
    vel = mg.energy_velocities()
    pos, mass = posm[:, :3].copy(), posm[:, 3].copy()
    for tag, box in (("periodic", mg.BOX), ("open", 0.0)):
        ke, pe = oracle.energy(pos, vel, mass, mg.EPS, box)
        assert abs(ke - float(g["ke_" + tag])) <= 2e-5 * abs(ke)
        assert abs(pe - float(g["pe_" + tag])) <= 1e-4 * abs(pe)
