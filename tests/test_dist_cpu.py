"""CPU suite, part 3: the N > 1 host logic (target sharding + source all-gather)
on world_size 2 with the gloo backend.  The per-rank force evaluation is played
by the CPU oracle here; on GPUs the same SourceGather feeds b200_direct_forces_dev."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "lambda-cdm-raytracing_b200", "python")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import b200grav
    from inputs import masses_np, uniform_mt
    from oracle.pyoracle import Oracle
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    o = Oracle()
    pos = uniform_mt(n, seed=77)
    mass = masses_np(n, seed=78)
    full = np.concatenate([pos, mass[:, None]], 1).astype(np.float32)
    lo, hi = b200grav.shard_range(n, rank, world)
    # each rank only knows its own shard; the rest starts as garbage
    posm = torch.full((n, 4), float("nan"))
    posm[lo:hi] = torch.from_numpy(full[lo:hi])
    gather = b200grav.SourceGather(n, rank, world)
    for step in range(2):                       # two "steps": drift the local shard, gather again
        gather(posm)
        got = posm.numpy()
        assert np.array_equal(got, full), "all-gather did not reproduce the replicated source array"
        acc = o.direct_f32(got[:, :3].copy(), got[:, 3].copy(), eps=0.01, i0=lo, n_targets=hi - lo)
        full[:, :3] += np.float32(0.25)         # every rank applies the same deterministic drift to its shard
        posm[lo:hi] = torch.from_numpy(full[lo:hi])
    out = [None] * world
    dist.all_gather_object(out, (lo, hi, acc))
    if rank == 0:
        res = np.empty((n, 3), np.float32)
        for a, b, part in out:
            res[a:b] = part
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1024, 1001])
def test_sharded_targets_gloo_world2(oracle, n):
    import torch.multiprocessing as mp
    from inputs import masses_np, uniform_mt
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pos = uniform_mt(n, seed=77) + np.float32(0.25)       # positions at the second step
    mass = masses_np(n, seed=78)
    assert np.array_equal(res, oracle.direct_f32(pos, mass, eps=0.01))


def test_shard_ranges_cover():
    import b200grav
    for n in (1, 7, 1000, 1 << 20):
        for w in (1, 2, 3, 8):
            r = [b200grav.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))


def test_bench_input_helpers():
    """Host-side helpers of bench.py: Zel'dovich grid choice and Morton storage order."""
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.zeldovich_grid(1 << 20) == 128 and bench.zeldovich_grid(1 << 24) == 256
    assert bench.zeldovich_grid(64) == 4 and bench.zeldovich_grid(65) == 8
    import numpy as np
    p, m = bench.make_particles(20000, order="morton")
    q, _ = bench.make_particles(20000)
    assert np.array_equal(np.sort(p.view([("x", "f4"), ("y", "f4"), ("z", "f4")]).ravel(), order=("x", "y", "z")),
                          np.sort(q.view([("x", "f4"), ("y", "f4"), ("z", "f4")]).ravel(), order=("x", "y", "z")))
    # neighbours in storage order are neighbours in space
    assert np.abs(np.diff(p, axis=0)).mean() < 0.2 * np.abs(np.diff(q, axis=0)).mean()
    assert np.all(m == 1.0)
