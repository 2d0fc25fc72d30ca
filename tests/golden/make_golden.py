"""Generate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref, i.e. the
reference's own CPU sources compiled where they lie under /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The fixtures pin: direct forces, tree topology / centres of mass / forces,
Morton keys, H(a) and the Zel'dovich generator output, exactly as the reference
computes them.  They are what tests/test_oracle.py holds the restated oracle
to, and what the -m gpu tests hold the CUDA path to.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from inputs import masses_np, uniform_mt, uniform_np, clustered_np  # noqa: E402
from oracle.pyoracle import Ref, build  # noqa: E402


def ic_scalars(r):
    """Normalised P(k) and (D, f, H)(a_init) of the reference's IC generator / CosmologyModel."""
    k = np.exp(np.linspace(np.log(0.005), np.log(20.0), 48))
    out = {"k": k}
    for z in (49.0, 9.0, 0.0):
        pk, d, f, h = r.ic_scalars(k, z)
        out[f"pk_z{int(z)}"] = pk
        out[f"dfh_z{int(z)}"] = np.array([d, f, h])
    np.savez_compressed(os.path.join(HERE, "ic_scalars.npz"), **out)


def main():
    build(ref=True)
    r = Ref()
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "ic":
        ic_scalars(r)
        print("wrote ic_scalars.npz")
        return
    ic_scalars(r)
    out = {}

    # D1: the reference's only CPU direct sum (one root leaf), unit masses
    p = uniform_mt(768, seed=42)
    np.savez_compressed(os.path.join(HERE, "direct_768.npz"), pos=p, acc=r.direct(p))

    # T3-T6: centred uniform, real masses, theta 0.5 leaf 8
    p = uniform_np(4096, seed=7)
    m = masses_np(4096, seed=8)
    acc, st = r.tree_forces(p, m, stats=True)
    d = r.tree_dump(p, m)
    np.savez_compressed(os.path.join(HERE, "tree_centred_4096.npz"), pos=p, mass=m, acc=acc,
                        stats=st, theta=0.5, leaf_cap=8, max_depth=20, box=100.0, **{"t_" + k: v for k, v in d.items()})

    # box convention [0,100): particles outside the root cube, max-depth overflow leaves
    p = uniform_np(3000, seed=9, lo=0.0, hi=100.0)
    m = np.ones(3000, np.float32)
    acc, st = r.tree_forces(p, m, stats=True)
    d = r.tree_dump(p, m)
    np.savez_compressed(os.path.join(HERE, "tree_box_3000.npz"), pos=p, mass=m, acc=acc,
                        stats=st, theta=0.5, leaf_cap=8, max_depth=20, box=100.0, **{"t_" + k: v for k, v in d.items()})

    # clustered, small leaves, shallow max depth, different theta
    p = clustered_np(2500, seed=10)
    m = masses_np(2500, seed=11)
    acc, st = r.tree_forces(p, m, theta=0.7, leaf_cap=3, max_depth=6, stats=True)
    d = r.tree_dump(p, m, leaf_cap=3, max_depth=6)
    np.savez_compressed(os.path.join(HERE, "tree_clustered_2500.npz"), pos=p, mass=m, acc=acc,
                        stats=st, theta=0.7, leaf_cap=3, max_depth=6, box=100.0, **{"t_" + k: v for k, v in d.items()})

    # Zel'dovich generator (examples/zeldovich_test.cpp parameters), shifted to the centred convention
    zp, zv, zm = r.zeldovich(4096, grid=64, box=100.0, z_init=49.0, seed=12345)
    pc = (zp - np.float32(50.0)).astype(np.float32)
    acc, st = r.tree_forces(pc, zm, stats=True)
    np.savez_compressed(os.path.join(HERE, "zeldovich_4096.npz"), pos_box=zp, vel=zv, mass=zm, pos=pc,
                        acc=acc, stats=st, keys=r.morton_keys(zp, 100.0))

    # reference random particles (initial_conditions.cpp:800-821): box-convention positions + velocities
    rp, rv, rm = r.random_particles(2048, 100.0, 12345)
    np.savez_compressed(os.path.join(HERE, "random_2048.npz"), pos=rp, vel=rv, mass=rm,
                        keys=r.morton_keys(rp, 100.0))

    # scalars
    a = np.array([0.02, 0.1, 0.25, 0.5, 0.75, 1.0, 1.5, 2.0])
    kat_xyz = np.array([(0.5, 0.5, 0.5), (0.999, 0, 0), (0, 0.999, 0), (0, 0, 0.999),
                        (0.25, 0.75, 0.125), (1, 1, 1)], np.float32)
    st16 = r.tree_forces(uniform_mt(16384), np.ones(16384, np.float32), stats=True)[1]
    two = r.tree_forces(np.array([[0, 0, 0], [1, 0, 0]], np.float32), np.array([3, 5], np.float32))
    f1, f2 = r.newtonian_pair([0, 0, 0], [1, 0, 0])
    z10 = r.zeldovich(10000, grid=64, box=100.0, z_init=49.0, seed=12345)[0]   # examples/zeldovich_test.cpp:57
    np.savez_compressed(os.path.join(HERE, "scalars.npz"), a=a,
                        hubble=np.array([r.hubble_a(x) for x in a]),
                        morton_xyz=kat_xyz,
                        morton=np.array([r.morton3d(*map(float, x)) for x in kat_xyz], np.uint32),
                        expand_3ff=np.uint32(r.expand_bits(0x3ff)),
                        stats_uniform_mt_16384=st16, two_body=two, newtonian_pair=np.stack([f1, f2]),
                        zeldovich_10000_head=z10[:4], zeldovich_10000_com=z10.astype(np.float64).mean(0))
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(" ", f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
