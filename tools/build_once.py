import os, sys
import numpy as np, torch
sys.path.insert(0, "lambda-cdm-raytracing_b200/python")
import b200grav
n = int(sys.argv[1])
rng = np.random.default_rng(42)
pos = rng.uniform(-50, 50, (n, 3))
posm = torch.from_numpy(np.concatenate([pos, np.ones((n, 1))], 1).astype(np.float32)).cuda()
eng = b200grav.Engine(0)
for _ in range(2):
    eng.tree_build_dev(posm, n, 100.0, 8, 20)
torch.cuda.synchronize()
print(eng.tree_stats())
