"""ctypes bindings for the CPU checkers under oracle/.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product
(lambda-cdm-raytracing_b200/) must never import this module.

  Oracle  -> oracle/liboracle.so        (oracle.c, the restatement)
  Ref     -> oracle/_ref/liblcdm_ref.so (the reference's own CPU sources,
             compiled by oracle/Makefile where /root/reference exists)
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")


def build(ref=True):
    """Compile the checkers (make decides whether the reference tree exists)."""
    subprocess.run(["make", "-s", "-C", _HERE, "liboracle.so"], check=True)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


def _mass_arg(mass):
    if mass is None:
        return None
    return np.ascontiguousarray(mass, np.float32).ctypes.data_as(C.c_void_p)


class _Tree(C.Structure):
    _fields_ = [("n_nodes", C.c_size_t), ("n_particles", C.c_size_t),
                ("level", C.POINTER(C.c_int32)), ("center", C.POINTER(C.c_float)),
                ("size", C.POINTER(C.c_float)), ("first_child", C.POINTER(C.c_int32)),
                ("arrivals", C.POINTER(C.c_int64)), ("part_off", C.POINTER(C.c_int64)),
                ("part_idx", C.POINTER(C.c_int32)), ("mass", C.POINTER(C.c_float)),
                ("com", C.POINTER(C.c_float))]


class Tree:
    """Owning view of an orc_tree (canonical breadth-first node table)."""

    def __init__(self, lib, ptr):
        self._lib, self._ptr = lib, ptr
        t = ptr.contents
        nn = t.n_nodes
        self.n_nodes = nn
        as_np = np.ctypeslib.as_array
        self.level = as_np(t.level, (nn,))
        self.center = as_np(t.center, (nn, 3))
        self.size = as_np(t.size, (nn,))
        self.first_child = as_np(t.first_child, (nn,))
        self.arrivals = as_np(t.arrivals, (nn,))
        self.part_off = as_np(t.part_off, (nn + 1,))
        self.part_idx = as_np(t.part_idx, (max(int(self.part_off[nn]), 1),))[: int(self.part_off[nn])]
        self.mass = as_np(t.mass, (nn,))
        self.com = as_np(t.com, (nn, 3))

    @property
    def n_leaves(self):
        return int((self.first_child < 0).sum())

    @property
    def depth(self):
        return int(self.level.max()) + 1

    def __del__(self):
        if getattr(self, "_ptr", None):
            self._lib.orc_tree_free(self._ptr)
            self._ptr = None


class Oracle:
    def __init__(self, path=None):
        path = path or os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        sz = C.c_size_t
        L.orc_direct_f32.argtypes = [_f32p, C.c_void_p, sz, sz, sz, C.c_float, _f32p]
        L.orc_direct_f64.argtypes = [_f32p, C.c_void_p, sz, sz, sz, C.c_double, _f64p]
        L.orc_direct_periodic_f32.argtypes = [_f32p, C.c_void_p, sz, sz, sz, C.c_float, C.c_float, _f32p]
        L.orc_energy.argtypes = [_f32p, _f32p, C.c_void_p, sz, C.c_float, C.c_float, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_expand_bits.argtypes = [C.c_uint32]; L.orc_expand_bits.restype = C.c_uint32
        L.orc_morton3d.argtypes = [C.c_float] * 3; L.orc_morton3d.restype = C.c_uint32
        L.orc_morton_keys.argtypes = [_f32p, sz, C.c_float, _u32p]
        L.orc_sort_pairs.argtypes = [_u32p, sz, _u32p, _i32p]
        for f in (L.orc_tree_build_insert, L.orc_tree_build_levels):
            f.argtypes = [_f32p, _f32p, sz, C.c_float, sz, C.c_int]
            f.restype = C.POINTER(_Tree)
        L.orc_tree_build_fixed.argtypes = [_f32p, _f32p, sz, sz, C.c_int]
        L.orc_tree_build_fixed.restype = C.POINTER(_Tree)
        L.orc_tree_forces_fixed.argtypes = [C.POINTER(_Tree), _f32p, _f32p, C.c_float, C.c_float, sz, sz, _f32p,
                                            C.c_void_p]
        L.orc_tree_forces_fixed_periodic.argtypes = [C.POINTER(_Tree), _f32p, _f32p, C.c_float, C.c_float, C.c_float,
                                                     sz, sz, _f32p, C.c_void_p]
        L.orc_tree_potential_fixed.argtypes = [C.POINTER(_Tree), _f32p, _f32p, C.c_float, C.c_float, C.c_float, sz, sz,
                                               _f32p]
        L.orc_tree_free.argtypes = [C.POINTER(_Tree)]
        L.orc_tree_equal.argtypes = [C.POINTER(_Tree)] * 2; L.orc_tree_equal.restype = C.c_int
        L.orc_tree_forces.argtypes = [C.POINTER(_Tree), _f32p, C.c_float, sz, sz, _f32p, C.c_void_p]
        L.orc_hubble_a.argtypes = [C.c_double] * 5; L.orc_hubble_a.restype = C.c_double
        L.orc_scale_factor_step.argtypes = [C.c_double] * 6; L.orc_scale_factor_step.restype = C.c_double
        L.orc_kick.argtypes = [_f32p, _f32p, C.c_void_p, sz, C.c_float, C.c_double]
        L.orc_drift.argtypes = [_f32p, _f32p, sz, C.c_float, C.c_float]
        L.orc_num_threads.restype = C.c_int

    # -- direct --------------------------------------------------------------
    def direct_f32(self, pos, mass=None, eps=0.01, i0=0, n_targets=None):
        pos = np.ascontiguousarray(pos, np.float32)
        n = pos.shape[0]
        nt = n - i0 if n_targets is None else n_targets
        out = np.empty((nt, 3), np.float32)
        self.lib.orc_direct_f32(pos, _mass_arg(mass), n, i0, nt, eps, out)
        return out

    def direct_f64(self, pos, mass=None, eps=0.01, i0=0, n_targets=None):
        pos = np.ascontiguousarray(pos, np.float32)
        n = pos.shape[0]
        nt = n - i0 if n_targets is None else n_targets
        out = np.empty((nt, 3), np.float64)
        self.lib.orc_direct_f64(pos, _mass_arg(mass), n, i0, nt, eps, out)
        return out

    def direct_periodic_f32(self, pos, mass, eps, box, i0=0, n_targets=None):
        pos = np.ascontiguousarray(pos, np.float32)
        n = pos.shape[0]
        nt = n - i0 if n_targets is None else n_targets
        out = np.empty((nt, 3), np.float32)
        self.lib.orc_direct_periodic_f32(pos, _mass_arg(mass), n, i0, nt, eps, box, out)
        return out

    def energy(self, pos, vel, mass, eps, box=0.0):
        """(kinetic, potential) as the reference's compute_energy defines them; eps is the softening LENGTH."""
        pos = np.ascontiguousarray(pos, np.float32)
        vel = np.ascontiguousarray(vel, np.float32)
        ke, pe = C.c_double(), C.c_double()
        self.lib.orc_energy(pos, vel, _mass_arg(mass), pos.shape[0], np.float32(eps) * np.float32(eps), box,
                            C.byref(ke), C.byref(pe))
        return ke.value, pe.value

    # -- keys / sort ---------------------------------------------------------
    def expand_bits(self, v):
        return int(self.lib.orc_expand_bits(v))

    def morton3d(self, x, y, z):
        return int(self.lib.orc_morton3d(x, y, z))

    def morton_keys(self, pos, box):
        pos = np.ascontiguousarray(pos, np.float32)
        keys = np.empty(pos.shape[0], np.uint32)
        self.lib.orc_morton_keys(pos, pos.shape[0], box, keys)
        return keys

    def sort_pairs(self, keys):
        keys = np.ascontiguousarray(keys, np.uint32)
        sk = np.empty_like(keys)
        perm = np.empty(keys.shape[0], np.int32)
        self.lib.orc_sort_pairs(keys, keys.shape[0], sk, perm)
        return sk, perm

    # -- tree ----------------------------------------------------------------
    def tree_build(self, pos, mass, box=100.0, leaf_cap=8, max_depth=20, method="levels"):
        pos = np.ascontiguousarray(pos, np.float32)
        mass = np.ascontiguousarray(mass, np.float32)
        f = self.lib.orc_tree_build_levels if method == "levels" else self.lib.orc_tree_build_insert
        return Tree(self.lib, f(pos, mass, pos.shape[0], box, leaf_cap, max_depth))

    def tree_build_fixed(self, pos, mass, leaf_cap=8, max_depth=20):
        """The fixed-physics tree (no orphans, data-fitted root cube)."""
        pos = np.ascontiguousarray(pos, np.float32)
        mass = np.ascontiguousarray(mass, np.float32)
        return Tree(self.lib, self.lib.orc_tree_build_fixed(pos, mass, pos.shape[0], leaf_cap, max_depth))

    def tree_forces_fixed(self, tree, pos, mass, theta=0.5, eps=0.01, i0=0, n_targets=None, counters=False, box=0.0):
        """box > 0: minimum-image separations (periodic box)."""
        pos = np.ascontiguousarray(pos, np.float32)
        mass = np.ascontiguousarray(mass, np.float32)
        n = pos.shape[0]
        nt = n - i0 if n_targets is None else n_targets
        out = np.empty((nt, 3), np.float32)
        cnt = np.zeros(3, np.uint64)
        self.lib.orc_tree_forces_fixed_periodic(tree._ptr, pos, mass, theta, eps, box, i0, nt, out,
                                                cnt.ctypes.data_as(C.c_void_p))
        return (out, cnt) if counters else out

    def tree_potential_fixed(self, tree, pos, mass, theta=0.5, eps=0.01, box=0.0, i0=0, n_targets=None):
        pos = np.ascontiguousarray(pos, np.float32)
        mass = np.ascontiguousarray(mass, np.float32)
        n = pos.shape[0]
        nt = n - i0 if n_targets is None else n_targets
        phi = np.empty(nt, np.float32)
        self.lib.orc_tree_potential_fixed(tree._ptr, pos, mass, theta, eps, box, i0, nt, phi)
        return phi

    def tree_equal(self, a, b):
        return bool(self.lib.orc_tree_equal(a._ptr, b._ptr))

    def tree_forces(self, tree, pos, theta=0.5, i0=0, n_targets=None, counters=False):
        pos = np.ascontiguousarray(pos, np.float32)
        n = pos.shape[0]
        nt = n - i0 if n_targets is None else n_targets
        out = np.empty((nt, 3), np.float32)
        cnt = np.zeros(3, np.uint64)
        self.lib.orc_tree_forces(tree._ptr, pos, theta, i0, nt, out, cnt.ctypes.data_as(C.c_void_p))
        return (out, cnt) if counters else out

    # -- leapfrog ------------------------------------------------------------
    def hubble_a(self, a, om=0.31, ok=0.0, ol=0.69, h=0.67):
        return float(self.lib.orc_hubble_a(a, om, ok, ol, h))

    def scale_factor_step(self, a, dt, om=0.31, ok=0.0, ol=0.69, h=0.67):
        return float(self.lib.orc_scale_factor_step(a, dt, om, ok, ol, h))

    def kick(self, vel, acc, mass, dt, a):
        self.lib.orc_kick(vel, np.ascontiguousarray(acc, np.float32), _mass_arg(mass), vel.shape[0], dt, a)

    def drift(self, pos, vel, dt, box):
        self.lib.orc_drift(pos, vel, pos.shape[0], dt, box)

    def kdk_run(self, pos, vel, mass, force_fn, steps, dt, a0=1.0, box=100.0,
                om=0.31, ok=0.0, ol=0.69, h=0.67, acc0=None):
        """KDK order of LambdaCDMSimulationImpl::step (lambda_cdm_impl.cu:167-213):
        kick(dt/2, a) -> drift(dt) -> a += a*H(a)*dt -> forces(x_new) -> kick(dt/2, a_new).
        The reference's first half-kick reads uninitialised forces; the oracle
        defines F(step 0) = forces at the initial positions."""
        pos = np.array(pos, np.float32, copy=True)
        vel = np.array(vel, np.float32, copy=True)
        a = float(a0)
        acc = force_fn(pos) if acc0 is None else acc0
        half = np.float32(dt * 0.5)
        for _ in range(steps):
            self.kick(vel, acc, mass, half, a)
            self.drift(pos, vel, np.float32(dt), box)
            a = self.scale_factor_step(a, dt, om, ok, ol, h)
            acc = force_fn(pos)
            self.kick(vel, acc, mass, half, a)
        return pos, vel, a

    def num_threads(self):
        return int(self.lib.orc_num_threads())


class Ref:
    """The reference's own compiled CPU code (oracle/_ref/liblcdm_ref.so)."""

    @staticmethod
    def available():
        return os.path.exists(os.path.join(_HERE, "_ref", "liblcdm_ref.so"))

    def __init__(self):
        L = self.lib = C.CDLL(os.path.join(_HERE, "_ref", "liblcdm_ref.so"))
        sz = C.c_size_t
        L.ref_tree_forces.argtypes = [_f32p, _f32p, sz, C.c_float, sz, C.c_int, C.c_float, _f32p, C.c_void_p]
        L.ref_factory_tree_forces.argtypes = [_f32p, _f32p, sz, _f32p]
        L.ref_tree_dump.argtypes = [_f32p, _f32p, sz, sz, C.c_int, C.c_float,
                                    C.POINTER(sz), C.POINTER(sz)] + [C.c_void_p] * 8
        L.ref_hubble_a.argtypes = [C.c_double] * 5; L.ref_hubble_a.restype = C.c_double
        L.ref_zeldovich.argtypes = [sz, C.c_float, C.c_double, C.c_uint32, sz, _f32p, _f32p, _f32p]
        if hasattr(L, "ref_ic_scalars"):
            L.ref_ic_scalars.argtypes = [C.c_double, sz, _f64p, _f64p] + [C.POINTER(C.c_double)] * 3
        L.ref_random_particles.argtypes = [sz, C.c_float, C.c_uint32, _f32p, _f32p, _f32p]
        L.ref_expand_bits.argtypes = [C.c_uint32]; L.ref_expand_bits.restype = C.c_uint32
        L.ref_morton3d.argtypes = [C.c_float] * 3; L.ref_morton3d.restype = C.c_uint32
        L.ref_morton_keys.argtypes = [_f32p, sz, C.c_float, _u32p]
        L.ref_newtonian_pair.argtypes = [_f32p, _f32p, C.c_float, C.c_float, _f32p, _f32p]

    def tree_forces(self, pos, mass, theta=0.5, leaf_cap=8, max_depth=20, box=100.0, stats=False):
        pos = np.ascontiguousarray(pos, np.float32)
        mass = np.ascontiguousarray(mass, np.float32)
        out = np.empty_like(pos)
        st = np.zeros(3, np.uint64)
        rc = self.lib.ref_tree_forces(pos, mass, pos.shape[0], theta, leaf_cap, max_depth, box, out,
                                      st.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
        return (out, st) if stats else out

    def direct(self, pos, mass=None):
        """The reference's only CPU direct sum: one root leaf holding everything."""
        pos = np.ascontiguousarray(pos, np.float32)
        m = np.ones(pos.shape[0], np.float32) if mass is None else mass
        return self.tree_forces(pos, m, 0.5, pos.shape[0] + 1, 20, 100.0)

    def factory_tree_forces(self, pos, mass):
        pos = np.ascontiguousarray(pos, np.float32)
        mass = np.ascontiguousarray(mass, np.float32)
        out = np.empty_like(pos)
        rc = self.lib.ref_factory_tree_forces(pos, mass, pos.shape[0], out)
        assert rc == 0, rc
        return out

    def tree_dump(self, pos, mass, leaf_cap=8, max_depth=20, box=100.0):
        pos = np.ascontiguousarray(pos, np.float32)
        mass = np.ascontiguousarray(mass, np.float32)
        nn, ns = C.c_size_t(0), C.c_size_t(0)
        self.lib.ref_tree_dump(pos, mass, pos.shape[0], leaf_cap, max_depth, box,
                               C.byref(nn), C.byref(ns), *([None] * 8))
        nn, ns = nn.value, ns.value
        d = dict(level=np.empty(nn, np.int32), center=np.empty((nn, 3), np.float32),
                 size=np.empty(nn, np.float32), first_child=np.empty(nn, np.int32),
                 part_off=np.empty(nn + 1, np.int64), part_idx=np.empty(max(ns, 1), np.int32),
                 mass=np.empty(nn, np.float32), com=np.empty((nn, 3), np.float32))
        a, b = C.c_size_t(0), C.c_size_t(0)
        self.lib.ref_tree_dump(pos, mass, pos.shape[0], leaf_cap, max_depth, box, C.byref(a), C.byref(b),
                               *[d[k].ctypes.data_as(C.c_void_p) for k in
                                 ("level", "center", "size", "first_child", "part_off", "part_idx", "mass", "com")])
        d["part_idx"] = d["part_idx"][:ns]
        return d

    def hubble_a(self, a, om=0.31, ol=0.69, ok=0.0, h=0.67):
        return float(self.lib.ref_hubble_a(a, om, ol, ok, h))

    def zeldovich(self, n, grid=64, box=100.0, z_init=49.0, seed=12345):
        pos = np.empty((n, 3), np.float32); vel = np.empty((n, 3), np.float32); m = np.empty(n, np.float32)
        rc = self.lib.ref_zeldovich(grid, box, z_init, seed, n, pos, vel, m)
        assert rc == 0, rc
        return pos, vel, m

    def ic_scalars(self, k, z_init=49.0):
        """Normalised P(k) of the IC generator and (D, f, H) at a_init, from the reference's own code."""
        k = np.ascontiguousarray(k, np.float64)
        pk = np.empty_like(k)
        d, f, h = C.c_double(), C.c_double(), C.c_double()
        rc = self.lib.ref_ic_scalars(z_init, k.size, k, pk, C.byref(d), C.byref(f), C.byref(h))
        assert rc == 0, rc
        return pk, d.value, f.value, h.value

    def random_particles(self, n, box=100.0, seed=12345):
        pos = np.empty((n, 3), np.float32); vel = np.empty((n, 3), np.float32); m = np.empty(n, np.float32)
        self.lib.ref_random_particles(n, box, seed, pos, vel, m)
        return pos, vel, m

    def morton3d(self, x, y, z):
        return int(self.lib.ref_morton3d(x, y, z))

    def expand_bits(self, v):
        return int(self.lib.ref_expand_bits(v))

    def morton_keys(self, pos, box):
        pos = np.ascontiguousarray(pos, np.float32)
        keys = np.empty(pos.shape[0], np.uint32)
        self.lib.ref_morton_keys(pos, pos.shape[0], box, keys)
        return keys

    def newtonian_pair(self, p1, p2, m1=1.0, m2=1.0):
        f1 = np.empty(3, np.float32); f2 = np.empty(3, np.float32)
        self.lib.ref_newtonian_pair(np.asarray(p1, np.float32), np.asarray(p2, np.float32), m1, m2, f1, f2)
        return f1, f2
