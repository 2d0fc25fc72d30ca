"""CPU restatement of the device walk's data structure and control flow (tree.cu: com_kernel's subtree record
counts, pack_level_kernel's depth-first ids, walk_warp*_kernel's visit loop), checked against the oracle's
recursive walk (tree_force_computer.cpp:257-347): one record per internal node that carries mass, stored in
depth-first order; "descend" = id + 1, `skip` = id + records in the subtree; a target that accepts a cell adds the
monopole and continues at `skip`, one that opens it sums over the particles of ALL leaf children (orphans left out)
and continues at id + 1.  The per-target counters (nodes visited, cell, pair interactions) must EQUAL the oracle's,
the forces agree to float rounding.  The CUDA kernels are held to the same oracle in tests/test_gpu_tree.py; this
test pins the table design itself."""
import numpy as np
import pytest

from inputs import clustered_np, masses_np, rel_l2, uniform_mt, uniform_np

F = np.float32


def build_tables(t):
    """Depth-first walk records from the oracle's canonical (breadth-first) node table."""
    fc, mass = np.asarray(t.first_child), np.asarray(t.mass)
    nn = len(fc)
    level = np.asarray(t.level)
    sub = np.zeros(nn, np.int64)                      # records in the subtree (com_kernel, bottom-up by level)
    for k in np.argsort(-level, kind="stable"):
        if fc[k] >= 0:
            sub[k] = 0 if mass[k] == 0 else 1 + sub[fc[k]:fc[k] + 8].sum()
    pre = np.full(nn, -1, np.int64)                   # ids handed down level by level (pack_level_kernel)
    if fc[0] >= 0 and sub[0] > 0:
        pre[0] = 0
    for k in np.argsort(level, kind="stable"):
        if fc[k] < 0 or pre[k] < 0:
            continue
        nxt = pre[k] + 1
        for c in range(fc[k], fc[k] + 8):
            if sub[c] > 0:
                pre[c] = nxt
                nxt += sub[c]
    n_rec = int(sub[0])
    node_of = np.zeros(n_rec, np.int64)
    for k in range(nn):
        if pre[k] >= 0:
            node_of[pre[k]] = k
    skip = np.array([pre[k] + sub[k] for k in node_of], np.int64)
    off, idx = np.asarray(t.part_off), np.asarray(t.part_idx)
    leaves = []                                       # leaf-child particles of every record, children in order
    for k in node_of:
        ids = [idx[off[c]:off[c + 1]] for c in range(fc[k], fc[k] + 8) if fc[c] < 0]
        leaves.append(np.concatenate(ids) if ids else np.zeros(0, np.int32))
    return node_of, skip, leaves


def table_walk(t, tables, pos, i, theta, eps=F(0.01)):
    node_of, skip, leaves = tables
    com, mass, size, fc = np.asarray(t.com, F), np.asarray(t.mass, F), np.asarray(t.size, F), np.asarray(t.first_child)
    p = pos[i]
    acc = np.zeros(3, np.float64)
    vis, cells, pairs = 1, 0, 0                       # the root is visited by everyone

    def leaf_sum(ids):
        nonlocal pairs
        for j in ids:
            if j == i:
                continue
            d = pos[j] - p
            r2 = F(F(F(d[0] * d[0]) + F(d[1] * d[1])) + F(d[2] * d[2])) + F(eps * eps)
            acc[:] += d / (np.float64(r2) ** 1.5)     # unit mass (:340)
            pairs += 1

    if mass[0] == 0:
        return acc, (vis, cells, pairs)
    if fc[0] < 0:
        off, idx = np.asarray(t.part_off), np.asarray(t.part_idx)
        leaf_sum(idx[off[0]:off[1]])
        return acc, (vis, cells, pairs)
    k = 0
    while k < len(node_of):
        n = node_of[k]
        d = com[n] - p
        d2 = F(F(F(d[0] * d[0]) + F(d[1] * d[1])) + F(d[2] * d[2]))
        with np.errstate(divide="ignore"):
            accept = F(size[n] / F(np.sqrt(d2))) < F(theta)          # :302-310
        if accept:
            r2 = F(d2 + F(eps * eps))
            acc[:] += np.float64(mass[n]) * d / (np.float64(r2) ** 1.5)
            cells += 1
            k = skip[k]                               # asleep until the walk leaves this subtree
        else:
            vis += 8                                  # its 8 children, leaves included
            leaf_sum(leaves[k])
            k += 1                                    # the next record in depth-first order
    return acc, (vis, cells, pairs)


@pytest.mark.parametrize("gen,cap", [("uniform", 8), ("clustered", 8), ("box", 8), ("uniform", 1), ("tiny", 8)])
def test_depth_first_tables_reproduce_the_reference_walk(oracle, gen, cap):
    n = 3000
    if gen == "uniform":
        pos = uniform_mt(n, seed=5)
    elif gen == "clustered":
        pos = clustered_np(n, seed=6)
    elif gen == "box":
        pos = uniform_np(n, seed=7, lo=0.0, hi=100.0)               # the [0, box) convention: deep corner chains
    else:
        n = 6
        pos = uniform_np(n, seed=8)                                  # the whole tree is one leaf
    m = masses_np(n, seed=9)
    t = oracle.tree_build(pos, m, leaf_cap=cap)
    tables = build_tables(t)
    node_of, skip, _ = tables
    assert np.all(skip > np.arange(len(skip))) and (len(skip) == 0 or skip[0] == len(skip))
    rng = np.random.default_rng(1)
    for i in rng.choice(n, size=min(n, 40), replace=False):
        ref, cnt = oracle.tree_forces(t, pos, 0.5, i0=int(i), n_targets=1, counters=True)
        got, mine = table_walk(t, tables, pos, int(i), 0.5)
        assert tuple(int(c) for c in cnt[:3]) == mine, (gen, int(i))
        if np.abs(ref).max() > 0:
            assert rel_l2(got[None, :].astype(np.float32), ref) < 2e-5
