"""Generate tests/golden/ref_gpu_kernels.npz from the reference's OWN CUDA kernels, recompiled for sm_100a
(oracle/_ref/liblcdm_ref_gpu.so, `make -C oracle refgpu`), run on a B200:

    gpurun -- python tests/golden/make_golden_gpu.py          # writes gpurun_out/ref_gpu_kernels.npz
    cp gpurun_out/ref_gpu_kernels.npz tests/golden/

Pins, as the reference's GPU path computes them (-O3 --use_fast_math, CMakeLists.txt:93):
  K2 compute_forces_tiled (lambda_cdm_kernels.cu:144-221): periodic forces (= a_i * m_i) of 10 240 particles,
      box 100, eps 0.01, masses in [0.5, 1.5);
  K4 leapfrog_update (:290-335): one kick (dt/2) and one drift (dt) of 4 096 particles, a = 1.37, box 100.
  K6 compute_energy (:338-408, launch_energy_computation :492-516): kinetic and potential energy of the same
      10 240 particles with N(0,100) velocities, periodic (box 100) and with the minimum image switched off
      (box 1e9), eps 0.01  ->  gpurun_out/ref_gpu_energy.npz.
Inputs are regenerated from the seeds below (inputs()), so only the outputs are stored.  Needs numpy and ctypes only."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

N_DIRECT, N_LEAP = 10240, 4096
BOX, EPS, A, DT = 100.0, 0.01, 1.37, 2e-3


def inputs():
    rng = np.random.default_rng(20240)
    posm = np.empty((N_DIRECT, 4), np.float32)
    posm[:, :3] = rng.uniform(0.0, BOX, (N_DIRECT, 3))
    posm[:, 3] = rng.uniform(0.5, 1.5, N_DIRECT)
    vel = rng.normal(0.0, 100.0, (N_LEAP, 3)).astype(np.float32)
    force = rng.normal(0.0, 50.0, (N_LEAP, 3)).astype(np.float32)
    return posm, vel, force


def energy_velocities():
    return np.random.default_rng(20241).normal(0.0, 100.0, (N_DIRECT, 3)).astype(np.float32)


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "liblcdm_ref_gpu.so"))
    fp = C.POINTER(C.c_float)
    lib.refgpu_direct.argtypes = [fp, fp, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, fp]
    lib.refgpu_leapfrog.argtypes = [fp, fp, fp, C.c_int, C.c_float, C.c_float, C.c_double, C.c_int]
    p = lambda a: a.ctypes.data_as(fp)                                             # noqa: E731
    posm, vel, force = inputs()
    forces = np.empty((N_DIRECT, 3), np.float32)
    ms = np.zeros(1, np.float32)
    assert lib.refgpu_direct(p(posm), p(forces), N_DIRECT, BOX, EPS, 0, 1, p(ms)) == 0
    lp = np.ascontiguousarray(posm[:N_LEAP]).copy()
    lv = vel.copy()
    assert lib.refgpu_leapfrog(p(lp), p(lv), p(force), N_LEAP, DT * 0.5, BOX, A, 1) == 0      # kick dt/2
    kick_vel = lv.copy()
    assert lib.refgpu_leapfrog(p(lp), p(lv), p(force), N_LEAP, DT, BOX, A, 0) == 0            # drift dt
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    np.savez_compressed(os.path.join(out, "ref_gpu_kernels.npz"), forces=forces, kick_vel=kick_vel, drift_pos=lp[:, :3])
    print("wrote", os.path.join(out, "ref_gpu_kernels.npz"), float(np.abs(forces).max()))
    lib.refgpu_energy.argtypes = [fp, fp, C.c_int, C.c_float, C.c_float, fp, fp]
    ve = energy_velocities()
    res = {}
    for tag, box in (("periodic", BOX), ("open", 1.0e9)):
        ke, pe = np.zeros(1, np.float32), np.zeros(1, np.float32)
        assert lib.refgpu_energy(p(posm), p(ve), N_DIRECT, box, EPS, p(ke), p(pe)) == 0
        res["ke_" + tag], res["pe_" + tag] = ke[0], pe[0]
    np.savez_compressed(os.path.join(out, "ref_gpu_energy.npz"), **res)
    print("wrote", os.path.join(out, "ref_gpu_energy.npz"), res)


if __name__ == "__main__":
    sys.exit(main())
