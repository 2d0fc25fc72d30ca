/*
 * oracle.h -- CPU restatement of the reference's N-body force hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and there only as the checker or the
 * CPU baseline.  The product path (lambda-cdm-raytracing_b200/) never links,
 * imports or falls back to anything here.
 *
 * Every function cites the reference file:line it restates (paths relative
 * to the reference tree).  Parity of this restatement against the reference's
 * own compiled sources (oracle/_ref) is pinned by tests/test_oracle_vs_ref.py
 * (run where /root/reference exists) and by the committed fixtures in
 * tests/golden/ (generated from oracle/_ref by tests/golden/make_golden.py).
 */
#ifndef B200GRAV_ORACLE_H
#define B200GRAV_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- D1: direct sum (src/forces/tree_force_computer.cpp:312-347) ---------
 * acc_i = sum_{j != i} m_j * d / (|d|^2 + eps^2)^{3/2},  d = x_j - x_i,
 * FP32, sequential j order, sqrtf + one divide per pair, G = 1.
 * mass == NULL reproduces the reference's unit-mass leaf quirk (:340).
 * Targets are [i0, i0 + n_targets); sources are all n particles.
 * out is float[3 * n_targets]. */
void orc_direct_f32(const float* pos3, const float* mass, size_t n,
                    size_t i0, size_t n_targets, float eps, float* out3);

/* Same sum in FP64 (inputs still float): the truth used above 64 K particles,
 * where the FP32 oracle's own round-off exceeds the 1e-5 gate. */
void orc_direct_f64(const float* pos3, const float* mass, size_t n,
                    size_t i0, size_t n_targets, double eps, double* out3);

/* Periodic K1 variant (src/physics/lambda_cdm_kernels.cu:14-56):
 * d -= box * roundf(d / box) per axis, r2 = |d|^2 + eps2, f = m_j / (r2*sqrtf(r2)). */
void orc_direct_periodic_f32(const float* pos3, const float* mass, size_t n,
                             size_t i0, size_t n_targets, float eps, float box,
                             float* out3);

/* ---- T1: Morton keys (include/forces/barnes_hut_tree.hpp:11-27,
 *          src/forces/barnes_hut_tree.cu:33-55) ----------------------------- */
/* K6: compute_energy (src/physics/lambda_cdm_kernels.cu:338-408); box <= 0 = open boundary */
void orc_energy(const float* pos3, const float* vel3, const float* mass, size_t n,
                float eps2, float box, double* kinetic, double* potential);

uint32_t orc_expand_bits(uint32_t v);
uint32_t orc_morton3d(float x, float y, float z);          /* inputs in [0,1] */
void orc_morton_keys(const float* pos3, size_t n, float box, uint32_t* keys);

/* ---- T2: stable ascending sort of (key, original index)
 *          (src/forces/barnes_hut_tree.cu:383-401; thrust LSD radix = stable) */
void orc_sort_pairs(const uint32_t* keys, size_t n, uint32_t* sorted_keys,
                    int32_t* perm);

/* ---- T3/T4: octree (src/forces/tree_force_computer.cpp:130-243) ----------
 * Canonical breadth-first node table: node 0 is the root; the 8 children of
 * an internal node are contiguous, in octant order; children of earlier
 * parents (in BFS order) come first.  Both builders below emit this form. */
typedef struct orc_tree {
    size_t    n_nodes;
    size_t    n_particles;      /* number of inserted particles              */
    int32_t*  level;            /* [n_nodes]                                 */
    float*    center;           /* [n_nodes*3]                               */
    float*    size;             /* [n_nodes]  full edge length               */
    int32_t*  first_child;      /* [n_nodes]  -1 for a leaf                  */
    int64_t*  arrivals;         /* [n_nodes]  particles that reached the node*/
    int64_t*  part_off;         /* [n_nodes+1] into part_idx                 */
    int32_t*  part_idx;         /* stored particles (leaf members, or the
                                   orphans of an internal node), arrival order*/
    float*    mass;             /* [n_nodes]  total_mass after T4            */
    float*    com;              /* [n_nodes*3]                               */
} orc_tree;

/* Line-by-line restatement: sequential insert_particle / subdivide_node /
 * get_octant (:144-194) into a pointer-free node pool, then BFS-renumbered. */
orc_tree* orc_tree_build_insert(const float* pos3, const float* mass, size_t n,
                                float box, size_t leaf_cap, int max_depth);
/* Closed form: level-by-level stable 8-way split with the
 * "first leaf_cap arrivals stay" orphan rule.  Must equal the above. */
orc_tree* orc_tree_build_levels(const float* pos3, const float* mass, size_t n,
                                float box, size_t leaf_cap, int max_depth);
/* Fixed-physics tree and walk: no orphans, data-fitted root cube, real masses in
 * leaf pairs, eps parameter.  Not the reference's tree; validated against orc_direct_f64. */
void orc_fixed_root(const float* pos3, size_t n, float center[3], float* size);
orc_tree* orc_tree_build_fixed(const float* pos3, const float* mass, size_t n,
                               size_t leaf_cap, int max_depth);
void orc_tree_forces_fixed(const orc_tree* t, const float* pos3, const float* mass, float theta,
                           float eps, size_t i0, size_t n_targets, float* out3,
                           uint64_t* counters);
void orc_tree_forces_fixed_periodic(const orc_tree* t, const float* pos3, const float* mass, float theta,
                                    float eps, float box, size_t i0, size_t n_targets, float* out3,
                                    uint64_t* counters);
void orc_tree_potential_fixed(const orc_tree* t, const float* pos3, const float* mass, float theta,
                              float eps, float box, size_t i0, size_t n_targets, float* phi);
void orc_tree_free(orc_tree* t);
/* 1 if the two canonical tables are identical (topology, stored particles,
 * centres, sizes, mass, com bit patterns), else 0. */
int orc_tree_equal(const orc_tree* a, const orc_tree* b);
size_t orc_tree_leaf_count(const orc_tree* t);
int orc_tree_depth(const orc_tree* t);   /* reference convention: root-only = 1 */

/* ---- T5/T6: walk (src/forces/tree_force_computer.cpp:245-347) ------------
 * Depth-first, children 0..7, accept iff size / |com - x| < theta (unsoftened
 * r, IEEE sqrt and divide); monopole and leaf pairs with eps = 0.01f literal;
 * leaf pairs use unit mass (masses == nullptr at :253).
 * counters (nullable): [0] nodes visited, [1] monopole, [2] pair interactions,
 * summed over the targets. */
void orc_tree_forces(const orc_tree* t, const float* pos3, float theta,
                     size_t i0, size_t n_targets, float* out3,
                     uint64_t* counters);

/* ---- L1-L3: leapfrog + scale factor --------------------------------------
 * H(a) (include/physics/cosmology_model.hpp:49-61), km/s/Mpc, FP64. */
double orc_hubble_a(double a, double omega_m, double omega_k,
                    double omega_lambda, double h);
/* a <- a + a*H(a)*dt  (src/physics/lambda_cdm_impl.cu:261-269) */
double orc_scale_factor_step(double a, double dt, double omega_m,
                             double omega_k, double omega_lambda, double h);
/* kick (src/physics/lambda_cdm_kernels.cu:310-318):
 * v += F * (1/m) * dt * (1/a^2), F = acc * m (K2 output convention, :217-219) */
void orc_kick(float* vel3, const float* acc3, const float* mass, size_t n,
              float dt, double a);
/* drift (:321-333): x += v*dt; x = fmodf(x + box, box).  box <= 0: no wrap. */
void orc_drift(float* pos3, const float* vel3, size_t n, float dt, float box);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
