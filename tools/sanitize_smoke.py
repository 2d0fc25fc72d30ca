"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Sizes are tiny on purpose (the tools slow kernels down 10-100x)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lambda-cdm-raytracing_b200", "python"))


def main():
    import torch
    import b200grav
    eng = b200grav.Engine(0)
    rng = np.random.default_rng(1)
    for n in (1, 700, 5000):
        pos = rng.uniform(-50, 50, (n, 3)).astype(np.float32)
        mass = rng.uniform(0.5, 1.5, n).astype(np.float32)
        posm = torch.from_numpy(np.concatenate([pos, mass[:, None]], 1)).cuda()
        acc = torch.empty((n, 3), dtype=torch.float32, device="cuda")
        vel = torch.zeros((n, 3), dtype=torch.float32, device="cuda")
        eng.direct_forces_dev(posm, acc, 0, n, eps=0.01)
        eng.direct_forces_dev(posm, acc, 0, n, eps=0.01, box=100.0)
        eng.energy_dev(posm, vel, eps=0.01)
        for fixed in (False, True):
            if fixed:
                eng.tree_build_fixed_dev(posm, n, 8, 20, eps=0.01)
            else:
                eng.tree_build_dev(posm, n, 100.0, 8, 20)
            eng.tree_walk_dev(acc, 0, n, theta=0.5)
            eng.tree_set_counting(True)
            eng.tree_walk_dev(acc, 0, n, theta=0.5)
            eng.tree_set_counting(False)
        eng.leapfrog_dev(posm, vel, acc, n, 2, np.float32(5e-4), 1.0, np.float32(1e-3), 0.0)
        torch.cuda.synchronize()
    posm = torch.empty((4096, 4), dtype=torch.float32, device="cuda")
    vel = torch.empty((4096, 3), dtype=torch.float32, device="cuda")
    eng.zeldovich_ics_dev(posm, vel, grid=16)
    torch.cuda.synchronize()
    eng.close()
    print("sanitize smoke done")


if __name__ == "__main__":
    main()
