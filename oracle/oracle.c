/*
 * oracle.c -- CPU restatement of the reference's N-body force hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Build with
 *     gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared
 * -ffp-contract=off matters: the reference's FP32 expressions are evaluated
 * with one rounding per operation here, independent of the host ISA.
 *
 * Parity status: pinned against the reference's own compiled CPU sources
 * (oracle/_ref, built by oracle/Makefile from /root/reference) -- bitwise for
 * tree topology, centres of mass and forces -- and against the known-answer
 * values the reference's docs/examples hold.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ D1 --- */
/* src/forces/tree_force_computer.cpp:312-347 (compute_node_particle_interaction)
 * applied to one leaf holding every particle. */
void orc_direct_f32(const float* pos3, const float* mass, size_t n,
                    size_t i0, size_t n_targets, float eps, float* out3) {
#pragma omp parallel for schedule(static)
    for (long long t = 0; t < (long long)n_targets; ++t) {
        size_t i = i0 + (size_t)t;
        float px = pos3[3 * i + 0], py = pos3[3 * i + 1], pz = pos3[3 * i + 2];
        float fx = 0.0f, fy = 0.0f, fz = 0.0f;
        for (size_t j = 0; j < n; ++j) {
            if (j == i) continue;                               /* :321 */
            float dx = pos3[3 * j + 0] - px;                    /* :328-331 */
            float dy = pos3[3 * j + 1] - py;
            float dz = pos3[3 * j + 2] - pz;
            float r2 = dx * dx + dy * dy + dz * dz;             /* :333 */
            r2 += eps * eps;                                    /* :334-335 */
            float r = sqrtf(r2);                                /* :337 */
            float r3 = r2 * r;                                  /* :338 */
            float m = mass ? mass[j] : 1.0f;                    /* :340 */
            float f = m / r3;                                   /* :341 */
            fx += f * dx;                                       /* :343-345 */
            fy += f * dy;
            fz += f * dz;
        }
        out3[3 * t + 0] = fx;
        out3[3 * t + 1] = fy;
        out3[3 * t + 2] = fz;
    }
}

void orc_direct_f64(const float* pos3, const float* mass, size_t n,
                    size_t i0, size_t n_targets, double eps, double* out3) {
    const double eps2 = eps * eps;
#pragma omp parallel for schedule(static)
    for (long long t = 0; t < (long long)n_targets; ++t) {
        size_t i = i0 + (size_t)t;
        double px = pos3[3 * i + 0], py = pos3[3 * i + 1], pz = pos3[3 * i + 2];
        double fx = 0.0, fy = 0.0, fz = 0.0;
        for (size_t j = 0; j < n; ++j) {
            if (j == i) continue;
            double dx = (double)pos3[3 * j + 0] - px;
            double dy = (double)pos3[3 * j + 1] - py;
            double dz = (double)pos3[3 * j + 2] - pz;
            double r2 = dx * dx + dy * dy + dz * dz + eps2;
            double m = mass ? (double)mass[j] : 1.0;
            double f = m / (r2 * sqrt(r2));
            fx += f * dx;
            fy += f * dy;
            fz += f * dz;
        }
        out3[3 * t + 0] = fx;
        out3[3 * t + 1] = fy;
        out3[3 * t + 2] = fz;
    }
}

/* src/physics/lambda_cdm_kernels.cu:14-56 (compute_forces_direct, K1) */
void orc_direct_periodic_f32(const float* pos3, const float* mass, size_t n,
                             size_t i0, size_t n_targets, float eps, float box,
                             float* out3) {
    const float eps2 = eps * eps;
#pragma omp parallel for schedule(static)
    for (long long t = 0; t < (long long)n_targets; ++t) {
        size_t i = i0 + (size_t)t;
        float px = pos3[3 * i + 0], py = pos3[3 * i + 1], pz = pos3[3 * i + 2];
        float fx = 0.0f, fy = 0.0f, fz = 0.0f;
        for (size_t j = 0; j < n; ++j) {
            if (j == i) continue;                               /* :28-29 */
            float dx = pos3[3 * j + 0] - px;                    /* :34-36 */
            float dy = pos3[3 * j + 1] - py;
            float dz = pos3[3 * j + 2] - pz;
            dx = dx - box * roundf(dx / box);                   /* :39-41 */
            dy = dy - box * roundf(dy / box);
            dz = dz - box * roundf(dz / box);
            float r2 = dx * dx + dy * dy + dz * dz + eps2;      /* :43 */
            float r = sqrtf(r2);
            float r3 = r2 * r;
            float m = mass ? mass[j] : 1.0f;
            float f = m / r3;                                   /* :48 */
            fx += f * dx;
            fy += f * dy;
            fz += f * dz;
        }
        out3[3 * t + 0] = fx;
        out3[3 * t + 1] = fy;
        out3[3 * t + 2] = fz;
    }
}

/* ------------------------------------------------------------------ K6 --- */
/* compute_energy (src/physics/lambda_cdm_kernels.cu:338-408): kinetic
 * 1/2 m v^2 (:364) and potential -G m_i m_j / sqrt(|d|^2 + eps^2) over pairs
 * j > i (:367-386), d by minimum_image (:122-141, half-box compare) when
 * box > 0.  The reference adds float partial sums with atomics (:402-407),
 * which has no defined order; this restatement keeps the FP32 pair arithmetic
 * (:375-385) and accumulates in double, so it is the exact value the
 * reference's sum scatters around.  eps2 is the reference's `softening2`. */
void orc_energy(const float* pos3, const float* vel3, const float* mass, size_t n,
                float eps2, float box, double* kinetic, double* potential) {
    double ke = 0.0, pe = 0.0;
    const float half_box = box * 0.5f;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : ke, pe)
    for (long long ii = 0; ii < (long long)n; ++ii) {
        size_t i = (size_t)ii;
        float mi = mass ? mass[i] : 1.0f;
        float vx = vel3[3 * i], vy = vel3[3 * i + 1], vz = vel3[3 * i + 2];
        ke += (double)(0.5f * mi * (vx * vx + vy * vy + vz * vz));      /* :364 */
        double p = 0.0;
        for (size_t j = i + 1; j < n; ++j) {                            /* :367 */
            float dx = pos3[3 * j + 0] - pos3[3 * i + 0];               /* :375-378 */
            float dy = pos3[3 * j + 1] - pos3[3 * i + 1];
            float dz = pos3[3 * j + 2] - pos3[3 * i + 2];
            if (box > 0.0f) {                                           /* :122-141 */
                if (dx > half_box) dx -= box; else if (dx < -half_box) dx += box;
                if (dy > half_box) dy -= box; else if (dy < -half_box) dy += box;
                if (dz > half_box) dz -= box; else if (dz < -half_box) dz += box;
            }
            float r2 = dx * dx + dy * dy + dz * dz + eps2;              /* :381 */
            float r = sqrtf(r2);                                        /* :382 */
            float mj = mass ? mass[j] : 1.0f;
            p += (double)(-1.0f * mi * mj / r);                         /* :385, G_CONSTANT = 1 (:114) */
        }
        pe += p;
    }
    *kinetic = ke;
    *potential = pe;
}

/* ------------------------------------------------------------------ T1 --- */
/* include/forces/barnes_hut_tree.hpp:11-17 */
uint32_t orc_expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

/* include/forces/barnes_hut_tree.hpp:19-27 */
uint32_t orc_morton3d(float x, float y, float z) {
    x = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
    y = fminf(fmaxf(y * 1024.0f, 0.0f), 1023.0f);
    z = fminf(fmaxf(z * 1024.0f, 0.0f), 1023.0f);
    uint32_t xx = orc_expand_bits((uint32_t)x);
    uint32_t yy = orc_expand_bits((uint32_t)y);
    uint32_t zz = orc_expand_bits((uint32_t)z);
    return xx * 4 + yy * 2 + zz;
}

/* src/forces/barnes_hut_tree.cu:33-55 (compute_morton_codes_kernel), IEEE divide */
void orc_morton_keys(const float* pos3, size_t n, float box, uint32_t* keys) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) {
        float x = pos3[3 * i + 0] / box;                        /* :44-46 */
        float y = pos3[3 * i + 1] / box;
        float z = pos3[3 * i + 2] / box;
        x = x - floorf(x);                                      /* :49-51 */
        y = y - floorf(y);
        z = z - floorf(z);
        keys[i] = orc_morton3d(x, y, z);                        /* :54 */
    }
}

/* ------------------------------------------------------------------ T2 --- */
/* Stable ascending sort of (key, index): what thrust::sort_by_key's LSD radix
 * yields at src/forces/barnes_hut_tree.cu:383-401 after thrust::sequence (:358). */
void orc_sort_pairs(const uint32_t* keys, size_t n, uint32_t* sorted_keys,
                    int32_t* perm) {
    uint32_t* ka = (uint32_t*)malloc(n * sizeof(uint32_t));
    uint32_t* kb = (uint32_t*)malloc(n * sizeof(uint32_t));
    int32_t* va = (int32_t*)malloc(n * sizeof(int32_t));
    int32_t* vb = (int32_t*)malloc(n * sizeof(int32_t));
    for (size_t i = 0; i < n; ++i) { ka[i] = keys[i]; va[i] = (int32_t)i; }
    for (int pass = 0; pass < 4; ++pass) {
        size_t hist[257];
        memset(hist, 0, sizeof hist);
        int sh = pass * 8;
        for (size_t i = 0; i < n; ++i) hist[((ka[i] >> sh) & 0xFFu) + 1]++;
        for (int b = 0; b < 256; ++b) hist[b + 1] += hist[b];
        for (size_t i = 0; i < n; ++i) {
            size_t d = hist[(ka[i] >> sh) & 0xFFu]++;
            kb[d] = ka[i];
            vb[d] = va[i];
        }
        uint32_t* tk = ka; ka = kb; kb = tk;
        int32_t* tv = va; va = vb; vb = tv;
    }
    memcpy(sorted_keys, ka, n * sizeof(uint32_t));
    memcpy(perm, va, n * sizeof(int32_t));
    free(ka); free(kb); free(va); free(vb);
}

/* --------------------------------------------------------------- T3/T4 --- */
static orc_tree* tree_alloc(size_t n_nodes, size_t n_parts_stored, size_t n) {
    orc_tree* t = (orc_tree*)calloc(1, sizeof(orc_tree));
    t->n_nodes = n_nodes;
    t->n_particles = n;
    t->level = (int32_t*)calloc(n_nodes, sizeof(int32_t));
    t->center = (float*)calloc(n_nodes * 3, sizeof(float));
    t->size = (float*)calloc(n_nodes, sizeof(float));
    t->first_child = (int32_t*)calloc(n_nodes, sizeof(int32_t));
    t->arrivals = (int64_t*)calloc(n_nodes, sizeof(int64_t));
    t->part_off = (int64_t*)calloc(n_nodes + 1, sizeof(int64_t));
    t->part_idx = (int32_t*)calloc(n_parts_stored ? n_parts_stored : 1, sizeof(int32_t));
    t->mass = (float*)calloc(n_nodes, sizeof(float));
    t->com = (float*)calloc(n_nodes * 3, sizeof(float));
    return t;
}

void orc_tree_free(orc_tree* t) {
    if (!t) return;
    free(t->level); free(t->center); free(t->size); free(t->first_child);
    free(t->arrivals); free(t->part_off); free(t->part_idx); free(t->mass);
    free(t->com); free(t);
}

/* src/forces/tree_force_computer.cpp:188-194 (get_octant): strict > */
static inline int octant_of(const float* p, const float* c) {
    int o = 0;
    if (p[0] > c[0]) o |= 1;
    if (p[1] > c[1]) o |= 2;
    if (p[2] > c[2]) o |= 4;
    return o;
}

/* src/forces/tree_force_computer.cpp:173-186 (subdivide_node): child centre */
static inline void child_center(const float* c, float size, int i, float* out,
                                float* child_size) {
    float half_size = size * 0.5f;                              /* :175 */
    out[0] = c[0] + ((i & 1) ? half_size : -half_size) * 0.5f;  /* :180-182 */
    out[1] = c[1] + ((i & 2) ? half_size : -half_size) * 0.5f;
    out[2] = c[2] + ((i & 4) ? half_size : -half_size) * 0.5f;
    *child_size = half_size;                                    /* :184 */
}

/* src/forces/tree_force_computer.cpp:196-243 (compute_center_of_mass) on the
 * canonical table: children have larger ids than their parent, so a reverse
 * sweep is a valid post-order. */
static void tree_com(orc_tree* t, const float* pos3, const float* mass) {
    for (long long k = (long long)t->n_nodes - 1; k >= 0; --k) {
        float total_mass = 0.0f;
        float wx = 0.0f, wy = 0.0f, wz = 0.0f;
        if (t->first_child[k] < 0) {                            /* leaf :197-216 */
            for (int64_t q = t->part_off[k]; q < t->part_off[k + 1]; ++q) {
                size_t p = (size_t)t->part_idx[q];
                float m = mass[p];
                total_mass += m;
                wx += pos3[3 * p + 0] * m;
                wy += pos3[3 * p + 1] * m;
                wz += pos3[3 * p + 2] * m;
            }
        } else {                                                /* internal :217-241 */
            for (int i = 0; i < 8; ++i) {
                size_t c = (size_t)t->first_child[k] + (size_t)i;
                float cm = t->mass[c];
                if (cm > 0.0f) {
                    total_mass += cm;
                    wx += t->com[3 * c + 0] * cm;
                    wy += t->com[3 * c + 1] * cm;
                    wz += t->com[3 * c + 2] * cm;
                }
            }
        }
        if (total_mass > 0.0f) {
            t->com[3 * k + 0] = wx / total_mass;
            t->com[3 * k + 1] = wy / total_mass;
            t->com[3 * k + 2] = wz / total_mass;
        }
        t->mass[k] = total_mass;
    }
}

/* -- builder 1: sequential insertion, restated line by line ---------------- */
typedef struct {
    float center[3];
    float size;
    int level;
    int is_leaf;
    long long child0;          /* index of child 0 in the pool, -1 if none */
    int32_t* parts;
    size_t nparts, cap;
    long long arrivals;
} pool_node;

typedef struct { pool_node* a; size_t n, cap; } pool;

static long long pool_new(pool* P, const float* c, float size, int level) {
    if (P->n == P->cap) {
        P->cap = P->cap ? P->cap * 2 : 1024;
        P->a = (pool_node*)realloc(P->a, P->cap * sizeof(pool_node));
    }
    pool_node* nd = &P->a[P->n];
    memset(nd, 0, sizeof *nd);
    nd->center[0] = c[0]; nd->center[1] = c[1]; nd->center[2] = c[2];
    nd->size = size; nd->level = level; nd->is_leaf = 1; nd->child0 = -1;
    return (long long)P->n++;
}

static void node_push(pool_node* nd, int32_t p) {
    if (nd->nparts == nd->cap) {
        nd->cap = nd->cap ? nd->cap * 2 : 8;
        nd->parts = (int32_t*)realloc(nd->parts, nd->cap * sizeof(int32_t));
    }
    nd->parts[nd->nparts++] = p;
}

orc_tree* orc_tree_build_insert(const float* pos3, const float* mass, size_t n,
                                float box, size_t leaf_cap, int max_depth) {
    pool P = {0, 0, 0};
    float c0[3] = {0.0f, 0.0f, 0.0f};                           /* :132 */
    pool_new(&P, c0, box, 0);                                   /* :133, ctor level=0 */
    for (size_t i = 0; i < n; ++i) {                            /* :136-138 */
        long long k = 0;
        for (;;) {                                              /* insert_particle :144-171 */
            pool_node* nd = &P.a[k];
            nd->arrivals++;
            if (nd->is_leaf) {
                if (nd->nparts < leaf_cap && nd->level < max_depth) {   /* :148 */
                    node_push(nd, (int32_t)i);
                    break;
                } else if (nd->level < max_depth) {             /* :151-153 */
                    float cc[3] = {nd->center[0], nd->center[1], nd->center[2]};
                    float sz = nd->size; int lv = nd->level;
                    nd->is_leaf = 0;                            /* subdivide_node :173-186 */
                    long long first = -1;
                    for (int ch = 0; ch < 8; ++ch) {
                        float cen[3], csz;
                        child_center(cc, sz, ch, cen, &csz);
                        long long id = pool_new(&P, cen, csz, lv + 1);
                        if (ch == 0) first = id;
                    }
                    P.a[k].child0 = first;                      /* pool may have moved */
                } else {                                        /* :154-158 */
                    node_push(nd, (int32_t)i);
                    break;
                }
            }
            pool_node* cur = &P.a[k];
            int o = octant_of(&pos3[3 * i], cur->center);       /* :167 */
            k = cur->child0 + o;                                /* :170 */
        }
    }
    /* breadth-first renumbering into the canonical table */
    size_t nn = P.n;
    long long* order = (long long*)malloc(nn * sizeof(long long));
    long long* newid = (long long*)malloc(nn * sizeof(long long));
    size_t head = 0, tail = 0, stored = 0;
    order[tail++] = 0;
    while (head < tail) {
        long long k = order[head];
        newid[k] = (long long)head;
        ++head;
        stored += P.a[k].nparts;
        if (!P.a[k].is_leaf)
            for (int ch = 0; ch < 8; ++ch) order[tail++] = P.a[k].child0 + ch;
    }
    orc_tree* t = tree_alloc(nn, stored, n);
    int64_t off = 0;
    for (size_t q = 0; q < nn; ++q) {
        pool_node* nd = &P.a[order[q]];
        t->level[q] = nd->level;
        t->center[3 * q + 0] = nd->center[0];
        t->center[3 * q + 1] = nd->center[1];
        t->center[3 * q + 2] = nd->center[2];
        t->size[q] = nd->size;
        t->first_child[q] = nd->is_leaf ? -1 : (int32_t)newid[nd->child0];
        t->arrivals[q] = nd->arrivals;
        t->part_off[q] = off;
        for (size_t s = 0; s < nd->nparts; ++s) t->part_idx[off + (int64_t)s] = nd->parts[s];
        off += (int64_t)nd->nparts;
    }
    t->part_off[nn] = off;
    for (size_t q = 0; q < nn; ++q) free(P.a[q].parts);
    free(P.a); free(order); free(newid);
    tree_com(t, pos3, mass);                                    /* :141 */
    return t;
}

/* -- builder 2: closed form, level by level -------------------------------- */
static orc_tree* build_levels_impl(const float* pos3, const float* mass, size_t n,
                                   const float root_center[3], float root_size, size_t leaf_cap,
                                   int max_depth, size_t keep) {
    /* keep = arrivals a splitting node retains: leaf_cap in the reference's tree (they are
     * never redistributed, :144-171), 0 in the fixed-physics tree (everything moves down) */
    /* growing node arrays */
    size_t ncap = 1024, nn = 0;
    int32_t* level = (int32_t*)malloc(ncap * sizeof(int32_t));
    float* center = (float*)malloc(ncap * 3 * sizeof(float));
    float* size = (float*)malloc(ncap * sizeof(float));
    int32_t* first_child = (int32_t*)malloc(ncap * sizeof(int32_t));
    int64_t* arrivals = (int64_t*)malloc(ncap * sizeof(int64_t));
    int64_t* seg_start = (int64_t*)malloc(ncap * sizeof(int64_t));  /* into cur[] */
    int64_t* st_off = (int64_t*)malloc(ncap * sizeof(int64_t));
    int64_t* st_cnt = (int64_t*)malloc(ncap * sizeof(int64_t));
#define GROW(need)                                                              \
    while ((need) > ncap) {                                                     \
        ncap *= 2;                                                              \
        level = (int32_t*)realloc(level, ncap * sizeof(int32_t));               \
        center = (float*)realloc(center, ncap * 3 * sizeof(float));             \
        size = (float*)realloc(size, ncap * sizeof(float));                     \
        first_child = (int32_t*)realloc(first_child, ncap * sizeof(int32_t));   \
        arrivals = (int64_t*)realloc(arrivals, ncap * sizeof(int64_t));         \
        seg_start = (int64_t*)realloc(seg_start, ncap * sizeof(int64_t));       \
        st_off = (int64_t*)realloc(st_off, ncap * sizeof(int64_t));             \
        st_cnt = (int64_t*)realloc(st_cnt, ncap * sizeof(int64_t));             \
    }
    int32_t* cur = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
    int32_t* nxt = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
    int32_t* stored = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
    int64_t nstored = 0;
    for (size_t i = 0; i < n; ++i) cur[i] = (int32_t)i;
    level[0] = 0; center[0] = root_center[0]; center[1] = root_center[1]; center[2] = root_center[2];
    size[0] = root_size;
    arrivals[0] = (int64_t)n; seg_start[0] = 0; nn = 1;
    size_t lv_begin = 0, lv_end = 1;
    for (int lv = 0; lv_begin < lv_end; ++lv) {
        int64_t nxt_fill = 0;
        size_t next_begin = nn;
        for (size_t k = lv_begin; k < lv_end; ++k) {
            int64_t cnt = arrivals[k];
            const int32_t* seg = cur + seg_start[k];
            /* the (leaf_cap+1)-th arrival splits a node below max_depth (:148-153) */
            int split = (cnt > (int64_t)leaf_cap) && (lv < max_depth);
            st_off[k] = nstored;
            if (!split) {
                first_child[k] = -1;
                for (int64_t s = 0; s < cnt; ++s) stored[nstored++] = seg[s];
                st_cnt[k] = cnt;
                continue;
            }
            for (int64_t s = 0; s < (int64_t)keep; ++s) stored[nstored++] = seg[s];
            st_cnt[k] = (int64_t)keep;
            GROW(nn + 8);
            first_child[k] = (int32_t)nn;
            int64_t ccount[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int64_t s = (int64_t)keep; s < cnt; ++s)
                ccount[octant_of(&pos3[3 * (size_t)seg[s]], &center[3 * k])]++;
            int64_t cbase[8];
            for (int ch = 0; ch < 8; ++ch) {
                size_t c = nn + (size_t)ch;
                float csz;
                child_center(&center[3 * k], size[k], ch, &center[3 * c], &csz);
                size[c] = csz;
                level[c] = lv + 1;
                arrivals[c] = ccount[ch];
                seg_start[c] = nxt_fill;
                cbase[ch] = nxt_fill;
                nxt_fill += ccount[ch];
            }
            for (int64_t s = (int64_t)keep; s < cnt; ++s) {   /* stable split */
                int o = octant_of(&pos3[3 * (size_t)seg[s]], &center[3 * k]);
                nxt[cbase[o]++] = seg[s];
            }
            nn += 8;
        }
        int32_t* tmp = cur; cur = nxt; nxt = tmp;
        lv_begin = next_begin; lv_end = nn;
    }
#undef GROW
    orc_tree* t = tree_alloc(nn, (size_t)nstored, n);
    memcpy(t->level, level, nn * sizeof(int32_t));
    memcpy(t->center, center, nn * 3 * sizeof(float));
    memcpy(t->size, size, nn * sizeof(float));
    memcpy(t->first_child, first_child, nn * sizeof(int32_t));
    memcpy(t->arrivals, arrivals, nn * sizeof(int64_t));
    /* stored[] was filled in BFS node order, so offsets are already canonical */
    for (size_t k = 0; k < nn; ++k) t->part_off[k] = st_off[k];
    t->part_off[nn] = nstored;
    memcpy(t->part_idx, stored, (size_t)nstored * sizeof(int32_t));
    (void)st_cnt;
    free(level); free(center); free(size); free(first_child); free(arrivals);
    free(seg_start); free(st_off); free(st_cnt); free(cur); free(nxt); free(stored);
    tree_com(t, pos3, mass);
    return t;
}

orc_tree* orc_tree_build_levels(const float* pos3, const float* mass, size_t n,
                                float box, size_t leaf_cap, int max_depth) {
    const float origin[3] = {0.0f, 0.0f, 0.0f};                /* :132-133 */
    return build_levels_impl(pos3, mass, n, origin, box, leaf_cap, max_depth, leaf_cap);
}

/* -- "fixed physics" tree: NOT the reference's tree.  Same octant rule
 * (:188-194), same child geometry (:173-186), same centre-of-mass pass (:196-243), but
 *   - a node that splits hands ALL its particles to its children (no orphans),
 *   - the root cube is fitted to the data: centre = midpoint of the bounding box, edge =
 *     largest extent * 1.00001f (1.0f when every particle coincides).
 * Its purpose is a physically meaningful Barnes-Hut force, validated against the FP64
 * direct sum rather than against the reference's tree. */
void orc_fixed_root(const float* pos3, size_t n, float center[3], float* size) {
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            float v = pos3[3 * i + k];
            if (i == 0 || v < lo[k]) lo[k] = v;
            if (i == 0 || v > hi[k]) hi[k] = v;
        }
    float ext = 0.0f;
    for (int k = 0; k < 3; ++k) {
        center[k] = (lo[k] + hi[k]) * 0.5f;
        float e = hi[k] - lo[k];
        if (e > ext) ext = e;
    }
    *size = ext > 0.0f ? ext * 1.00001f : 1.0f;
}

orc_tree* orc_tree_build_fixed(const float* pos3, const float* mass, size_t n,
                               size_t leaf_cap, int max_depth) {
    float c[3], sz;
    orc_fixed_root(pos3, n, c, &sz);
    return build_levels_impl(pos3, mass, n, c, sz, leaf_cap, max_depth, 0);
}

int orc_tree_equal(const orc_tree* a, const orc_tree* b) {
    if (a->n_nodes != b->n_nodes || a->n_particles != b->n_particles) return 0;
    size_t nn = a->n_nodes;
    if (memcmp(a->level, b->level, nn * sizeof(int32_t))) return 0;
    if (memcmp(a->center, b->center, nn * 3 * sizeof(float))) return 0;
    if (memcmp(a->size, b->size, nn * sizeof(float))) return 0;
    if (memcmp(a->first_child, b->first_child, nn * sizeof(int32_t))) return 0;
    if (memcmp(a->arrivals, b->arrivals, nn * sizeof(int64_t))) return 0;
    if (memcmp(a->part_off, b->part_off, (nn + 1) * sizeof(int64_t))) return 0;
    if (memcmp(a->part_idx, b->part_idx, (size_t)a->part_off[nn] * sizeof(int32_t))) return 0;
    if (memcmp(a->mass, b->mass, nn * sizeof(float))) return 0;
    if (memcmp(a->com, b->com, nn * 3 * sizeof(float))) return 0;
    return 1;
}

size_t orc_tree_leaf_count(const orc_tree* t) {
    size_t c = 0;
    for (size_t k = 0; k < t->n_nodes; ++k) c += (t->first_child[k] < 0);
    return c;
}

/* src/forces/tree_force_computer.cpp:452-464 (compute_tree_depth) */
int orc_tree_depth(const orc_tree* t) {
    int mx = 0;
    for (size_t k = 0; k < t->n_nodes; ++k)
        if (t->level[k] > mx) mx = t->level[k];
    return mx + 1;
}

/* --------------------------------------------------------------- T5/T6 --- */
void orc_tree_forces(const orc_tree* t, const float* pos3, float theta,
                     size_t i0, size_t n_targets, float* out3,
                     uint64_t* counters) {
    uint64_t c_vis = 0, c_pc = 0, c_pp = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : c_vis, c_pc, c_pp)
    for (long long tt = 0; tt < (long long)n_targets; ++tt) {
        size_t i = i0 + (size_t)tt;
        float px = pos3[3 * i + 0], py = pos3[3 * i + 1], pz = pos3[3 * i + 2];
        float fx = 0.0f, fy = 0.0f, fz = 0.0f;                  /* :246-249 */
        int32_t stack[8 * 64];
        int sp = 0;
        stack[sp++] = 0;
        while (sp > 0) {
            size_t k = (size_t)stack[--sp];
            ++c_vis;
            if (t->mass[k] == 0.0f) continue;                   /* :260 */
            if (t->first_child[k] < 0) {                        /* leaf :268-270 -> :312-347 */
                for (int64_t q = t->part_off[k]; q < t->part_off[k + 1]; ++q) {
                    size_t j = (size_t)t->part_idx[q];
                    if (j == i) continue;
                    float dx = pos3[3 * j + 0] - px;
                    float dy = pos3[3 * j + 1] - py;
                    float dz = pos3[3 * j + 2] - pz;
                    float r2 = dx * dx + dy * dy + dz * dz;
                    float softening = 0.01f;
                    r2 += softening * softening;
                    float r = sqrtf(r2);
                    float r3 = r2 * r;
                    float f = 1.0f / r3;                        /* masses == nullptr :253,:340 */
                    fx += f * dx; fy += f * dy; fz += f * dz;
                    ++c_pp;
                }
                continue;
            }
            float dx = t->com[3 * k + 0] - px;                  /* :302-310 */
            float dy = t->com[3 * k + 1] - py;
            float dz = t->com[3 * k + 2] - pz;
            float r = sqrtf(dx * dx + dy * dy + dz * dz);
            if ((t->size[k] / r) < theta) {                     /* :309 */
                float r2 = dx * dx + dy * dy + dz * dz;         /* :280-290 */
                float softening = 0.01f;
                r2 += softening * softening;
                float rr = sqrtf(r2);
                float r3 = r2 * rr;
                float f = t->mass[k] / r3;
                fx += f * dx; fy += f * dy; fz += f * dz;
                ++c_pc;
            } else {                                            /* :293-297, children 0..7 */
                for (int ch = 7; ch >= 0; --ch) stack[sp++] = t->first_child[k] + ch;
            }
        }
        out3[3 * tt + 0] = fx;
        out3[3 * tt + 1] = fy;
        out3[3 * tt + 2] = fz;
    }
    if (counters) { counters[0] = c_vis; counters[1] = c_pc; counters[2] = c_pp; }
}

/* Fixed-physics walk: the reference's walk (:257-310) with the two leaf quirks removed --
 * leaf pairs use the sources' real masses and the softening is a parameter -- on a tree
 * without orphans.  No j == i test: with eps > 0 the self pair adds exactly 0 (it IS
 * counted in counters[2]). */
void orc_tree_forces_fixed(const orc_tree* t, const float* pos3, const float* mass, float theta,
                           float eps, size_t i0, size_t n_targets, float* out3,
                           uint64_t* counters) {
    orc_tree_forces_fixed_periodic(t, pos3, mass, theta, eps, 0.0f, i0, n_targets, out3, counters);
}

/* ... with box > 0: every separation (cell and particle) taken to its nearest periodic image,
 * d -= box * roundf(d / box), the minimum image of compute_forces_direct
 * (src/physics/lambda_cdm_kernels.cu:39-41) that the reference's GPU tree kernel also applies to its
 * cells (src/forces/barnes_hut_tree.cu:247-254).  No Ewald sum. */
void orc_tree_forces_fixed_periodic(const orc_tree* t, const float* pos3, const float* mass, float theta,
                                    float eps, float box, size_t i0, size_t n_targets, float* out3,
                                    uint64_t* counters) {
    uint64_t c_vis = 0, c_pc = 0, c_pp = 0;
    const float eps2 = eps * eps;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : c_vis, c_pc, c_pp)
    for (long long tt = 0; tt < (long long)n_targets; ++tt) {
        size_t i = i0 + (size_t)tt;
        float px = pos3[3 * i + 0], py = pos3[3 * i + 1], pz = pos3[3 * i + 2];
        float fx = 0.0f, fy = 0.0f, fz = 0.0f;
        int32_t stack[8 * 64];
        int sp = 0;
        stack[sp++] = 0;
        while (sp > 0) {
            size_t k = (size_t)stack[--sp];
            ++c_vis;
            if (t->mass[k] == 0.0f) continue;
            if (t->first_child[k] < 0) {
                for (int64_t q = t->part_off[k]; q < t->part_off[k + 1]; ++q) {
                    size_t j = (size_t)t->part_idx[q];
                    float dx = pos3[3 * j + 0] - px;
                    float dy = pos3[3 * j + 1] - py;
                    float dz = pos3[3 * j + 2] - pz;
                    if (box > 0.0f) {
                        dx = dx - box * roundf(dx / box);
                        dy = dy - box * roundf(dy / box);
                        dz = dz - box * roundf(dz / box);
                    }
                    float r2 = dx * dx + dy * dy + dz * dz + eps2;
                    float r = sqrtf(r2);
                    float f = (mass ? mass[j] : 1.0f) / (r2 * r);
                    fx += f * dx; fy += f * dy; fz += f * dz;
                    ++c_pp;
                }
                continue;
            }
            float dx = t->com[3 * k + 0] - px;
            float dy = t->com[3 * k + 1] - py;
            float dz = t->com[3 * k + 2] - pz;
            if (box > 0.0f) {
                dx = dx - box * roundf(dx / box);
                dy = dy - box * roundf(dy / box);
                dz = dz - box * roundf(dz / box);
            }
            float d2 = dx * dx + dy * dy + dz * dz;
            /* periodic: a cell whose particles may wrap differently from its centre of mass (it reaches
             * across the half-box distance from the target) cannot be a monopole: it is opened */
            int wraps = 0;
            if (box > 0.0f) {
                const float hb = box * 0.5f, sz = t->size[k];
                wraps = (fabsf(dx) + sz > hb) || (fabsf(dy) + sz > hb) || (fabsf(dz) + sz > hb);
            }
            if (!wraps && (t->size[k] / sqrtf(d2)) < theta) {
                float r2 = d2 + eps2;
                float f = t->mass[k] / (r2 * sqrtf(r2));
                fx += f * dx; fy += f * dy; fz += f * dz;
                ++c_pc;
            } else {
                for (int ch = 7; ch >= 0; --ch) stack[sp++] = t->first_child[k] + ch;
            }
        }
        out3[3 * tt + 0] = fx;
        out3[3 * tt + 1] = fy;
        out3[3 * tt + 2] = fz;
    }
    if (counters) { counters[0] = c_vis; counters[1] = c_pc; counters[2] = c_pp; }
}

/* Potential from the fixed-physics tree: the walk of orc_tree_forces_fixed_periodic accumulating
 * phi_i = sum m / sqrt(|d|^2 + eps^2) (M / r for an accepted cell), j == i skipped. */
void orc_tree_potential_fixed(const orc_tree* t, const float* pos3, const float* mass, float theta,
                              float eps, float box, size_t i0, size_t n_targets, float* phi) {
    const float eps2 = eps * eps;
#pragma omp parallel for schedule(dynamic, 256)
    for (long long tt = 0; tt < (long long)n_targets; ++tt) {
        size_t i = i0 + (size_t)tt;
        float px = pos3[3 * i + 0], py = pos3[3 * i + 1], pz = pos3[3 * i + 2];
        double acc = 0.0;
        int32_t stack[8 * 64];
        int sp = 0;
        stack[sp++] = 0;
        while (sp > 0) {
            size_t k = (size_t)stack[--sp];
            if (t->mass[k] == 0.0f) continue;
            if (t->first_child[k] < 0) {
                for (int64_t q = t->part_off[k]; q < t->part_off[k + 1]; ++q) {
                    size_t j = (size_t)t->part_idx[q];
                    if (j == i) continue;
                    float dx = pos3[3 * j + 0] - px, dy = pos3[3 * j + 1] - py, dz = pos3[3 * j + 2] - pz;
                    if (box > 0.0f) {
                        dx = dx - box * roundf(dx / box);
                        dy = dy - box * roundf(dy / box);
                        dz = dz - box * roundf(dz / box);
                    }
                    float r2 = dx * dx + dy * dy + dz * dz + eps2;
                    acc += (double)((mass ? mass[j] : 1.0f) / sqrtf(r2));
                }
                continue;
            }
            float dx = t->com[3 * k + 0] - px, dy = t->com[3 * k + 1] - py, dz = t->com[3 * k + 2] - pz;
            if (box > 0.0f) {
                dx = dx - box * roundf(dx / box);
                dy = dy - box * roundf(dy / box);
                dz = dz - box * roundf(dz / box);
            }
            float d2 = dx * dx + dy * dy + dz * dz;
            int wraps = 0;
            if (box > 0.0f) {
                const float hb = box * 0.5f, sz = t->size[k];
                wraps = (fabsf(dx) + sz > hb) || (fabsf(dy) + sz > hb) || (fabsf(dz) + sz > hb);
            }
            if (!wraps && (t->size[k] / sqrtf(d2)) < theta) {
                acc += (double)(t->mass[k] / sqrtf(d2 + eps2));
            } else {
                for (int ch = 7; ch >= 0; --ch) stack[sp++] = t->first_child[k] + ch;
            }
        }
        phi[tt] = (float)acc;
    }
}

/* --------------------------------------------------------------- L1-L3 --- */
/* include/physics/cosmology_model.hpp:49-61 */
double orc_hubble_a(double a, double omega_m, double omega_k,
                    double omega_lambda, double h) {
    double z = 1.0 / a - 1.0;                                   /* :59 */
    double aa = 1.0 / (1.0 + z);                                /* :50 */
    double e2 = omega_m * pow(aa, -3) + omega_k * pow(aa, -2) + omega_lambda;
    return 100.0 * h * sqrt(e2);                                /* :54 */
}

/* src/physics/lambda_cdm_impl.cu:261-269 */
double orc_scale_factor_step(double a, double dt, double omega_m,
                             double omega_k, double omega_lambda, double h) {
    double H = orc_hubble_a(a, omega_m, omega_k, omega_lambda, h);
    return a + a * H * dt;
}

/* src/physics/lambda_cdm_kernels.cu:307-318 with F = acc * m_i (:217-219) */
void orc_kick(float* vel3, const float* acc3, const float* mass, size_t n,
              float dt, double a) {
    const float a2_inv = (float)(1.0f / (a * a));               /* :308 */
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) {
        float m = mass ? mass[i] : 1.0f;
        const float mass_inv = 1.0f / m;                        /* :307 */
        for (int k = 0; k < 3; ++k) {
            float force = acc3[3 * i + k] * m;
            vel3[3 * i + k] += force * mass_inv * dt * a2_inv;  /* :312-314 */
        }
    }
}

/* src/physics/lambda_cdm_kernels.cu:321-333 */
void orc_drift(float* pos3, const float* vel3, size_t n, float dt, float box) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) {
        for (int k = 0; k < 3; ++k) {
            float p = pos3[3 * i + k];
            p += vel3[3 * i + k] * dt;                          /* :322-324 */
            if (box > 0.0f) p = fmodf(p + box, box);            /* :327-329 */
            pos3[3 * i + k] = p;
        }
    }
}
