"""-m gpu: the peer-memory direct sum.  Two processes share GPU 0; each packs half of
the sources into a tile buffer it exports over CUDA IPC, maps the other's, and runs
b200_direct_forces_parts_dev over [own part, peer part].  (Host-side gloo barrier;
no kernel waits on another process.)"""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, q):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "lambda-cdm-raytracing_b200", "python")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import b200grav
    from inputs import masses_np, uniform_mt
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    eng = b200grav.Engine(0)
    pos = uniform_mt(n, seed=5)
    mass = masses_np(n, seed=6)
    full = np.concatenate([pos, mass[:, None]], 1).astype(np.float32)
    lo, hi = b200grav.shard_range(n, rank, world)
    shard = torch.from_numpy(full[lo:hi].copy()).cuda()

    def exchange(mine):
        out = [None] * world
        dist.all_gather_object(out, mine)
        return out

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()

    peers = b200grav.PeerSources(eng, n, rank, world, barrier, exchange)
    acc = torch.empty((hi - lo, 3), dtype=torch.float32, device="cuda")
    res = None
    for step in range(3):                       # exercises both buffers
        parts = peers.publish(shard)
        eng.direct_forces_parts_dev(parts, peers.lens, shard, hi - lo, acc, eps=0.01)
        torch.cuda.synchronize()
        res = acc.cpu().numpy()
    out = [None] * world
    dist.all_gather_object(out, (lo, hi, res))
    dist.barrier()
    peers.close()
    eng.close()
    if rank == 0:
        r = np.empty((n, 3), np.float32)
        for a, b, part in out:
            r[a:b] = part
        q.put(r)
    dist.destroy_process_group()


def test_peer_mapped_sources_two_processes(oracle):
    import torch.multiprocessing as mp
    from inputs import masses_np, rel_l2, uniform_mt
    n = 6000
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = oracle.direct_f32(uniform_mt(n, seed=5), masses_np(n, seed=6), eps=0.01)
    assert rel_l2(res, want) < 1e-5


def _forest_worker(rank, world, port, n, q):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "lambda-cdm-raytracing_b200", "python")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import b200grav
    from inputs import masses_np, uniform_mt
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    eng = b200grav.Engine(rank)
    box = [eng.shard_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    eng.shard_init(box[0], rank, world)
    pos = uniform_mt(n, seed=31)
    mass = masses_np(n, seed=32)
    posm0 = torch.from_numpy(np.concatenate([pos, mass[:, None]], 1).astype(np.float32)).to(dev)
    # particles STORED along a Hilbert curve (rank r owns the r-th run of slots), inserted in original index order
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    eng.spatial_order_dev(posm0, n, 100.0, perm)
    posm = posm0[perm.long()].contiguous()
    arrival = torch.empty_like(perm)
    arrival[perm.long()] = torch.arange(n, dtype=torch.int32, device=dev)
    nl = n // world
    acc = torch.empty((nl, 3), dtype=torch.float32, device=dev)
    eng.tree_set_counting(True)
    for _ in range(2):                              # twice: slots are refilled, the build graph is replayed
        eng.tree_build_part_dev(posm, n, rank, world, arrival=arrival)
        eng.tree_forest_publish()
        eng.tree_walk_dev(acc, rank * nl, nl, theta=0.5)
        torch.cuda.synchronize()
    cnt = eng.tree_counters()
    own = perm[rank * nl:(rank + 1) * nl]
    exp = eng.tree_export()
    exp["part_idx"] = perm.cpu().numpy()[exp["part_idx"]].astype(np.int32)      # slots -> original indices
    out = [None] * world
    dist.all_gather_object(out, (own.cpu().numpy(), acc.cpu().numpy(), cnt, exp, eng.tree_forest_root()))
    dist.barrier()
    eng.shard_finalize()
    eng.close()
    if rank == 0:
        q.put(out)
    dist.destroy_process_group()


def test_forest_two_ranks(oracle):
    """Octant-sharded build with the NCCL table exchange, two ranks on two GPUs (skips on a 1-GPU box): merged
    topology bit-exact, summed counters equal to the oracle's, forces of both target lists."""
    import torch
    import torch.multiprocessing as mp
    from forest_util import merge_parts
    from inputs import masses_np, rel_l2, uniform_mt
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (NCCL refuses two ranks on one device)")
    n = 100000
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_forest_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    pos, mass = uniform_mt(n, seed=31), masses_np(n, seed=32)
    o = oracle.tree_build(pos, mass)
    merged = merge_parts([out[0][3], out[1][3]], out[0][4])
    for k in ("level", "center", "size", "first_child", "arrivals", "part_off", "part_idx", "mass", "com"):
        assert np.array_equal(merged[k], getattr(o, k)), k
    want, ocnt = oracle.tree_forces(o, pos, 0.5, counters=True)
    assert np.array_equal(out[0][2] + out[1][2], ocnt)
    assert np.array_equal(np.sort(np.concatenate([out[0][0], out[1][0]])), np.arange(n))
    for own, acc, _, _, _ in out:
        assert rel_l2(acc, want[own]) < 1e-5
