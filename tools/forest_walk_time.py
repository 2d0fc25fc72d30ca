"""One GPU plays all P ranks of the C4 configuration: 2^24 uniform particles stored in Hilbert order, P part builds
published into one forest, then the walk of every rank's slot range timed on its own -- per-rank cost spread and the
forest walk against the single-table walk of the same targets.  Usage: python tools/forest_walk_time.py [--parts 8]"""
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
    ap.add_argument("--drift", type=float, default=0.0, help="move every particle by N(0, drift) per axis after the "
                    "storage order has been fixed (the C4 run after a few steps)")
    ap.add_argument("--clamp", action="store_true", help="after the drift, wrap the particles back into the root cube")
    ap.add_argument("--single-only", action="store_true")
    ap.add_argument("--cold", action="store_true", help="flush the L2 (512 MB write) before every timed walk; time "
                    "the whole call (target order included) with CUDA events, once")
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
    if args.drift > 0:
        stored[:, :3] += args.drift * torch.randn((n, 3), generator=g, device="cuda")
    if args.clamp:
        stored[:, :3] = torch.remainder(stored[:, :3] + 50.0, 100.0) - 50.0
    nl = n // P
    acc = torch.empty((nl, 3), dtype=torch.float32, device="cuda")
    eng.set_timing(True)

    flush = torch.empty(128 << 20, dtype=torch.float32, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def walks(tag):
        ts = []
        for r in range(P):
            best = 1e30
            if args.cold:
                eng.tree_walk_dev(acc, r * nl, nl, 0.5)          # target order of this range computed here, untimed
                flush.zero_()
                torch.cuda.synchronize()
                e0.record()
                eng.tree_walk_dev(acc, r * nl, nl, 0.5)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
                continue
            for _ in range(3):
                eng.tree_walk_dev(acc, r * nl, nl, 0.5)
                torch.cuda.synchronize()
                best = min(best, eng.last_kernel_ms())
            ts.append(best)
        print(f"{tag}: per-rank walk ms " + " ".join(f"{t:.3f}" for t in ts) + f"  max {max(ts):.3f} mean {np.mean(ts):.3f}")

    eng.tree_build_part_dev(stored, n, 0, 1, 100.0, 8, 20, arrival=arrival)       # the whole tree, one table
    print("tree", eng.tree_stats(), " particles outside the root cube:",
          int((stored[:, :3].abs() >= 50.0).any(dim=1).sum()))
    walks("single table")
    eng.tree_set_counting(True)
    for r in range(P):
        eng.tree_walk_dev(acc, r * nl, nl, 0.5)
        torch.cuda.synchronize()
        st = [int(x) for x in eng.tree_walk_stats()]
        print(f"  range {r}: per target {st[0] / nl:.0f} visits {st[1] / nl:.0f} cells {st[2] / nl:.0f} pairs; pair-row slots "
              f"{st[3] / nl:.0f} (useful {st[2] / max(st[3], 1):.3f}), node-visit slots {st[4] / nl:.0f} (awake {st[5] / max(st[4], 1):.3f})")
    eng.tree_set_counting(False)
    for r in range(P):          # the same targets as an explicit list in storage order (no re-sort by current keys)
        lst = torch.arange(r * nl, (r + 1) * nl, dtype=torch.int32, device="cuda")
        best = 1e30
        for _ in range(2):
            flush.zero_()
            torch.cuda.synchronize()
            eng.tree_walk_list_dev(acc, lst, nl, 0.5)
            torch.cuda.synchronize()
            best = min(best, eng.last_kernel_ms())
        print(f"  range {r} as a list in storage order: {best:.3f} ms")
        # the same targets as a list in CURRENT Hilbert order (what the range walk sorts them into), results by slot
        sub = stored[r * nl:(r + 1) * nl].contiguous()
        prm = torch.empty(nl, dtype=torch.int32, device="cuda")
        eng.spatial_order_dev(sub, nl, 100.0, prm)
        lst2 = (prm + r * nl).contiguous()
        best = 1e30
        for _ in range(2):
            flush.zero_()
            torch.cuda.synchronize()
            eng.tree_walk_list_dev(acc, lst2, nl, 0.5)
            torch.cuda.synchronize()
            best = min(best, eng.last_kernel_ms())
        moved = (prm.long() - torch.arange(nl, device="cuda")).abs().float()
        print(f"  range {r} as a list in current Hilbert order: {best:.3f} ms   |perm - identity| mean {moved.mean().item():.1f} "
              f"max {moved.max().item():.0f}")
    if args.single_only:
        eng.close()
        return
    for q in range(P):
        eng.tree_build_part_dev(stored, n, q, P, 100.0, 8, 20, arrival=arrival)
        eng.tree_forest_publish()
    walks(f"forest of {P}")
    eng.close()


if __name__ == "__main__":
    main()
