"""CPU suite: tests/forest_util.merge_parts (the checker of the octant-sharded build's topology) against the oracle
tree itself -- part exports are emulated by pruning the oracle's canonical table to the octants a part owns."""
import numpy as np
import pytest

from forest_util import FIELDS, merge_parts, part_owner
from inputs import masses_np, uniform_mt


def _prune(T, q, P):
    owner = part_owner(P)
    fc = T["first_child"]
    order, newid = list(range(9)), {i: i for i in range(9)}
    cur = [1 + d for d in range(8) if owner[d] == q]
    while cur:
        nxt = []
        for k in cur:
            if fc[k] >= 0:
                for j in range(8):
                    c = fc[k] + j
                    newid[c] = len(order)
                    order.append(c)
                    nxt.append(c)
        cur = nxt
    rn = np.array(order)
    E = {k: T[k][rn].copy() for k in ("level", "center", "size", "mass", "com", "arrivals")}
    mine = lambda k: k == 0 or k > 8 or owner[k - 1] == q                      # noqa: E731
    E["first_child"] = np.array([newid[fc[k]] if fc[k] >= 0 and mine(k) else -1 for k in order], np.int32)
    lists = [T["part_idx"][T["part_off"][k]:T["part_off"][k + 1]] if mine(k) else np.zeros(0, np.int32) for k in order]
    for d in range(8):
        if owner[d] != q:
            E["mass"][1 + d] = 0
            E["com"][1 + d] = 0
            E["arrivals"][1 + d] = 0
    E["part_idx"] = np.concatenate(lists).astype(np.int32)
    E["part_off"] = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int64)
    return E


@pytest.mark.parametrize("n_parts", [1, 2, 3, 8])
def test_merge_parts_reproduces_the_canonical_table(oracle, n_parts):
    n = 20000
    t = oracle.tree_build(uniform_mt(n, seed=21), masses_np(n, seed=24))
    T = {k: np.array(getattr(t, k)) for k in FIELDS}
    root = np.concatenate([T["com"][0], [T["mass"][0]], [0, 0, T["size"][0], 0]]).astype(np.float32)
    M = merge_parts([_prune(T, q, n_parts) for q in range(n_parts)], root)
    for k in FIELDS:
        assert np.array_equal(M[k], T[k]), k
