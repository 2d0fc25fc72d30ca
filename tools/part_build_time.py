"""Times one rank's share of the octant-sharded build (part 0 of P) on one GPU: 2^24 uniform particles stored in
Hilbert order with the original index order as arrival order -- the C4 configuration -- without the NCCL exchange.
Usage: python tools/part_build_time.py [--parts 8] [--n 16777216]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lambda-cdm-raytracing_b200", "python"))


def main():
    import torch
    import b200grav
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 24)
    ap.add_argument("--parts", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    n, P = args.n, args.parts
    g = torch.Generator(device="cuda")
    g.manual_seed(4242)
    posm = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    posm[:, :3] = torch.rand((n, 3), generator=g, device="cuda") * 100.0 - 50.0
    posm[:, 3] = 1.0
    eng = b200grav.Engine(0)
    perm = torch.empty(n, dtype=torch.int32, device="cuda")
    eng.spatial_order_dev(posm, n, 100.0, perm)
    stored = posm[perm.long()].contiguous()
    arrival = torch.empty_like(perm)
    arrival[perm.long()] = torch.arange(n, dtype=torch.int32, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for part in range(min(P, 2)):
        for _ in range(3):
            eng.tree_build_part_dev(stored, n, part, P, 100.0, 8, 20, arrival=arrival)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(args.reps):
            eng.tree_build_part_dev(stored, n, part, P, 100.0, 8, 20, arrival=arrival)
        ev1.record()
        torch.cuda.synchronize()
        print(f"n={n} part {part} of {P}: build {ev0.elapsed_time(ev1) / args.reps:.3f} ms  stats {eng.tree_stats()}")
    eng.close()


if __name__ == "__main__":
    main()
