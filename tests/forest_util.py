"""Merges the per-part exports of an octant-sharded octree build (b200_tree_build_part_dev + b200_tree_export, one
export per part) into the canonical breadth-first node table the oracle emits (oracle.h: root 0, the 8 children of
a node contiguous, children of earlier parents first), so that the sharded build can be held to the same bit-exact
array compare as the unsharded one.  Pure numpy; levels are contiguous node ranges in every part, so the merge is
a concatenation per (level, part) plus an id shift for the child links."""
import numpy as np

FIELDS = ("level", "center", "size", "first_child", "arrivals", "part_off", "part_idx", "mass", "com")


def part_owner(n_parts):
    return [q for q in range(n_parts) for _ in range(q * 8 // n_parts, (q + 1) * 8 // n_parts)]


def merge_parts(parts, root_record):
    """parts: list of Engine.tree_export() dicts, part q at index q.  root_record: Engine.tree_forest_root()."""
    P = len(parts)
    owner = part_owner(P)
    depth = max(int(t["level"].max()) for t in parts) + 1
    # local node ranges per (part, level)
    beg = np.zeros((P, depth + 1), np.int64)
    for q, t in enumerate(parts):
        beg[q, :] = np.searchsorted(t["level"], np.arange(depth + 1))
    cnt = beg[:, 1:] - beg[:, :-1]                       # [P, depth]
    cnt[:, 0] = 0
    cnt[:, 1] = 0                                        # levels 0 and 1 are handled apart
    g_beg = np.zeros(depth + 1, np.int64)                # global first id of each level
    g_beg[1] = 1
    if depth > 1:
        g_beg[2] = 9
    for L in range(2, depth):
        g_beg[L + 1] = g_beg[L] + cnt[:, L].sum()
    off = np.zeros((P, depth), np.int64)                 # global id of part q's first level-L node
    for L in range(2, depth):
        off[:, L] = g_beg[L] + np.concatenate([[0], np.cumsum(cnt[:, L])[:-1]])
    nn = int(g_beg[depth]) if depth > 1 else 1

    def gid_of_child(q, L_child, local):                 # local ids of level-L_child nodes of part q -> global ids
        return off[q, L_child] + (local - beg[q, L_child])

    out = {k: [] for k in FIELDS if k not in ("part_off",)}
    lens = []

    def take(t, lo, hi, q, L):
        for k in ("level", "center", "size", "arrivals", "mass", "com"):
            out[k].append(t[k][lo:hi])
        fc = t["first_child"][lo:hi].astype(np.int64)
        if L + 1 < depth + 1:
            fc = np.where(fc >= 0, gid_of_child(q, min(L + 1, depth - 1), fc) if L >= 1 else fc, -1)
        out["first_child"].append(fc.astype(np.int32))
        po = t["part_off"]
        out["part_idx"].append(t["part_idx"][po[lo]:po[hi]])
        lens.append(np.diff(po[lo:hi + 1]))

    # root: topology from any part (its orphans and arrivals are the same everywhere), mass / com merged
    t0 = parts[0]
    take(t0, 0, 1, 0, 0)
    out["first_child"][-1] = np.array([1 if t0["first_child"][0] >= 0 else -1], np.int32)
    out["mass"][-1] = np.array([root_record[3]], np.float32)
    out["com"][-1] = np.asarray(root_record[:3], np.float32).reshape(1, 3)
    if depth > 1:
        for d in range(8):
            take(parts[owner[d]], 1 + d, 2 + d, owner[d], 1)
        for L in range(2, depth):
            for q in range(P):
                if cnt[q, L]:
                    take(parts[q], int(beg[q, L]), int(beg[q, L + 1]), q, L)
    res = {k: np.concatenate(v) for k, v in out.items()}
    res["part_off"] = np.concatenate([[0], np.cumsum(np.concatenate(lens))]).astype(np.int64)
    assert res["level"].shape[0] == nn
    return res
