"""-m gpu: the octant-sharded octree build and the forest walk (multi-GPU Barnes-Hut), played on ONE GPU:
the parts are built one after another in the same context and published into its own slots -- the kernels, tables
and walk are those of the N-rank run, only the NCCL broadcast is a device copy (the collective itself is covered by
tests/test_gpu_peer.py::test_forest_two_ranks on a 2-GPU box and by bench.py's c4_summary parity check).
Gates: merged topology, stored lists and centres of mass BIT-EXACT against the oracle; walk counters equal; forces
equal to the unsharded walk."""
import numpy as np
import pytest

from forest_util import merge_parts
from inputs import clustered_np, masses_np, rel_l2, uniform_mt, uniform_np

pytestmark = pytest.mark.gpu
TREE_KEYS = ("level", "center", "size", "first_child", "arrivals", "part_off", "part_idx", "mass", "com")


def _posm(p, m):
    import torch
    return torch.from_numpy(np.ascontiguousarray(np.concatenate([p, m[:, None]], 1), np.float32)).cuda()


def _build_forest(engine, posm, n, n_parts, arrival=None, **kw):
    import torch
    exports = []
    for q in range(n_parts):
        engine.tree_build_part_dev(posm, n, q, n_parts, arrival=arrival, **kw)
        if n_parts > 1:
            engine.tree_forest_publish()
        torch.cuda.synchronize()
        exports.append(engine.tree_export())
    return exports


@pytest.mark.parametrize("gen,n_parts,stored", [("uniform", 2, "index"), ("uniform", 8, "hilbert"), ("clustered", 4, "hilbert"),
                                                ("box", 8, "index"), ("uniform", 3, "hilbert"), ("uniform", 1, "hilbert")])
def test_forest_topology_forces_counters(engine, oracle, gen, n_parts, stored):
    """stored = "hilbert": the particles are STORED along a Hilbert curve and the build gets the arrival order
    (inverse permutation) -- the tree, its stored lists (mapped back to original indices) and the forces must still
    be the reference's, which inserts in original index order."""
    import torch
    n = 60000
    if gen == "uniform":
        p = uniform_mt(n, seed=21)
    elif gen == "box":
        p = uniform_np(n, seed=22, lo=0.0, hi=100.0)        # 7/8 of the particles outside the root cube: one octant chain
    else:
        p = clustered_np(n, seed=23)
    p[100:130] = p[0:30]                                    # duplicates
    p[200:230, 0] = 0.0                                     # on the root's x boundary (strict > sends them low)
    m = masses_np(n, seed=24)
    posm0 = _posm(p, m)
    # the unsharded build + walk on the particles in their original order
    engine.tree_build_dev(posm0, n, 100.0, 8, 20)
    engine.tree_set_counting(True)
    acc0 = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    engine.tree_walk_dev(acc0, 0, n, theta=0.5)
    torch.cuda.synchronize()
    cnt0 = engine.tree_counters()
    # the same targets as an explicit list on the unsharded tree
    shuffle = torch.from_numpy(np.random.default_rng(5).permutation(n).astype(np.int32)).cuda()
    acc_l = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    engine.tree_walk_list_dev(acc_l, shuffle, theta=0.5)
    torch.cuda.synchronize()
    assert np.array_equal(acc_l.cpu().numpy(), acc0.cpu().numpy()[shuffle.cpu().numpy()])
    assert np.array_equal(engine.tree_counters(), cnt0)
    # storage order and arrival order
    if stored == "hilbert":
        perm = torch.empty(n, dtype=torch.int32, device="cuda")
        engine.spatial_order_dev(posm0, n, 100.0, perm)           # perm[slot] = original index
        posm = posm0[perm.long()].contiguous()
        arrival = torch.empty_like(perm)
        arrival[perm.long()] = torch.arange(n, dtype=torch.int32, device="cuda")      # arrival[original] = slot
        slot_to_orig = perm.cpu().numpy()
    else:
        posm, arrival, slot_to_orig = posm0, None, np.arange(n)
    exports = _build_forest(engine, posm, n, n_parts, arrival=arrival, box=100.0, leaf_cap=8, max_depth=20)
    o = oracle.tree_build(p, m)
    if n_parts > 1:
        merged = merge_parts(exports, engine.tree_forest_root())
    else:
        merged = exports[0]
    merged["part_idx"] = slot_to_orig[merged["part_idx"]].astype(np.int32)
    for k in TREE_KEYS:
        assert np.array_equal(merged[k], getattr(o, k)), k
    acc1 = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    engine.tree_walk_dev(acc1, 0, n, theta=0.5)             # on a part-built context this walks the forest
    torch.cuda.synchronize()
    cnt1 = engine.tree_counters()
    engine.tree_set_counting(False)
    want, ocnt = oracle.tree_forces(o, p, 0.5, counters=True)
    assert np.array_equal(cnt1, ocnt) and np.array_equal(cnt0, ocnt), (cnt0, cnt1, ocnt)
    a0, a1 = acc0.cpu().numpy()[slot_to_orig], acc1.cpu().numpy()
    assert rel_l2(a1, a0) < 1e-6, rel_l2(a1, a0)
    assert rel_l2(a1, want[slot_to_orig]) < 1e-5
    # the shipped (non-counting) instance, explicit target list, gives the same numbers
    acc2 = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    engine.tree_walk_list_dev(acc2, shuffle, theta=0.5)
    torch.cuda.synchronize()
    assert np.array_equal(acc2.cpu().numpy(), a1[shuffle.cpu().numpy()])


def test_forest_state_errors(engine):
    import torch
    import b200grav
    n = 5000
    p = uniform_mt(n, seed=3)
    posm = _posm(p, np.ones(n, np.float32))
    acc = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    lst = torch.arange(n, dtype=torch.int32, device="cuda")
    engine.tree_build_part_dev(posm, n, 0, 2)
    engine.tree_forest_publish()
    with pytest.raises(b200grav.B200Error):                 # part 1 of THIS forest was never published (slots of
        engine.tree_walk_list_dev(acc, lst)                 # earlier forests do not count)
    with pytest.raises(b200grav.B200Error):                 # a part alone is not the tree
        engine.tree_walk_dev(acc, 0, n)
    engine.tree_build_part_dev(posm, n, 1, 2)
    engine.tree_forest_publish()
    engine.tree_walk_list_dev(acc, lst)
    torch.cuda.synchronize()
    assert np.isfinite(acc.cpu().numpy()).all()
    engine.tree_build_part_dev(posm, n, 0, 2)               # part 0 rebuilt: its slot is stale until published again
    with pytest.raises(b200grav.B200Error):
        engine.tree_walk_list_dev(acc, lst)
    with pytest.raises(b200grav.B200Error):                 # the root does not split: nothing to shard
        engine.tree_build_part_dev(posm, 8, 0, 2)
    engine.tree_build_dev(posm, n, 100.0, 8, 20)            # back to a whole tree
    engine.tree_walk_dev(acc, 0, n)
    torch.cuda.synchronize()


def test_forest_full_size_16m_topology(engine, oracle):
    """BASELINE config 4 size through the sharded build: 2^24 particles in 8 parts, merged node table bit-exact."""
    import torch
    n = 1 << 24
    rng = np.random.default_rng(4242)
    p = rng.uniform(-50.0, 50.0, size=(n, 3)).astype(np.float32)
    m = np.ones(n, np.float32)
    posm = _posm(p, m)
    exports = _build_forest(engine, posm, n, 8, box=100.0, leaf_cap=8, max_depth=20)
    merged = merge_parts(exports, engine.tree_forest_root())
    del exports
    o = oracle.tree_build(p, m)
    for k in TREE_KEYS:
        assert np.array_equal(merged[k], getattr(o, k)), k
    lst = torch.arange(3000000, 3000000 + 65536, dtype=torch.int32, device="cuda")
    acc = torch.empty((65536, 3), dtype=torch.float32, device="cuda")
    engine.tree_walk_list_dev(acc, lst, theta=0.5)
    torch.cuda.synchronize()
    want = oracle.tree_forces(o, p, 0.5, i0=3000000, n_targets=65536)
    assert rel_l2(acc.cpu().numpy(), want) < 1e-3
    engine.tree_build_dev(posm, 1024, 100.0, 8, 20)        # leave a small whole tree behind for the next tests
