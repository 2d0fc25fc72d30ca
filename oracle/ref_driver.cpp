// ref_driver.cpp -- C-ABI shim over the reference's OWN compiled CPU sources.
//
// TEST INFRASTRUCTURE ONLY.  This translation unit contains no reference code:
// it #includes the reference's public headers from where they lie
// (-I$(REF)/include) and is linked against the reference's own .cpp files,
// compiled in place by oracle/Makefile into oracle/_ref/ (git-ignored).
// It exists so the restated oracle (oracle.c) and the CUDA path can be checked
// against what the reference really computes.
//
// Reference entry points exercised:
//   forces::TreeForceComputer            include/forces/tree_force_computer.hpp:32-148
//   forces::ForceComputerFactory         include/forces/force_computer_factory.hpp:94-156
//   physics::CosmologyModel              include/physics/cosmology_model.hpp:35-61
//   physics::InitialConditionsGenerator  include/physics/initial_conditions.hpp:58-100
#include <any>
#include <array>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <typeindex>
#include <unordered_map>
#include <vector>
#include <complex>
#include <random>

// TreeForceComputer::root_ is private (tree_force_computer.hpp:35); the
// topology dump needs it.  Standard headers are included above so the macro
// only touches the reference headers.
#define private public
#include "forces/tree_force_computer.hpp"
#undef private
#include "core/simulation_context.hpp"
#include "physics/cosmology_model.hpp"
#include "physics/initial_conditions.hpp"

namespace {

struct Quiet {   // the reference prints on initialize/finalize; keep test logs clean
    std::streambuf *o, *e;
    std::ostringstream so, se;
    Quiet() : o(std::cout.rdbuf(so.rdbuf())), e(std::cerr.rdbuf(se.rdbuf())) {}
    ~Quiet() { std::cout.rdbuf(o); std::cerr.rdbuf(e); }
};

std::unique_ptr<forces::TreeForceComputer> make_tree(float theta, size_t leaf_cap,
                                                     int max_depth, float box) {
    auto t = std::make_unique<forces::TreeForceComputer>("oracle", theta, leaf_cap, max_depth);
    t->set_box_size(box);
    core::SimulationContext ctx;
    t->initialize(ctx);          // no HAVE_CUDA: downgrades to the CPU path
    return t;
}

}  // namespace

extern "C" {

// TreeForceComputer::compute_forces on host arrays; leaf_cap > n gives the
// reference's only CPU direct sum (unit masses).
int ref_tree_forces(const float* pos3, const float* mass, size_t n, float theta,
                    size_t leaf_cap, int max_depth, float box, float* out3,
                    size_t* stats /* nodes, leaves, depth; nullable */) {
    try {
        Quiet q;
        auto t = make_tree(theta, leaf_cap, max_depth, box);
        t->compute_forces(pos3, mass, out3, n);
        if (stats) {
            stats[0] = t->get_node_count();
            stats[1] = t->get_leaf_count();
            stats[2] = t->get_tree_depth();
        }
        t->finalize();
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

// Through the factory, the way examples/basic_simulation.cpp:12-13 does.
int ref_factory_tree_forces(const float* pos3, const float* mass, size_t n, float* out3) {
    try {
        Quiet q;
        forces::ForceComputerFactory::register_all_builtin_computers();
        auto c = forces::ForceComputerFactory::create_force_computer("TreeForceComputer", "oracle");
        if (!c) return 2;
        core::SimulationContext ctx;
        if (!c->initialize(ctx)) return 3;
        c->compute_forces(pos3, mass, out3, n);
        c->finalize();
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

// Canonical breadth-first dump of the reference's private pointer tree (same
// table layout as orc_tree in oracle.h).  Call once with null arrays to size.
int ref_tree_dump(const float* pos3, const float* mass, size_t n, size_t leaf_cap,
                  int max_depth, float box, size_t* n_nodes, size_t* n_stored,
                  int32_t* level, float* center, float* size, int32_t* first_child,
                  int64_t* part_off, int32_t* part_idx, float* node_mass, float* com) {
    Quiet q;
    auto t = make_tree(0.5f, leaf_cap, max_depth, box);
    t->build_tree_cpu(pos3, mass, n);
    std::vector<const forces::OctreeNode*> order;
    order.push_back(t->root_.get());
    for (size_t h = 0; h < order.size(); ++h)
        if (!order[h]->is_leaf)
            for (int c = 0; c < 8; ++c) order.push_back(order[h]->children[c].get());
    size_t stored = 0;
    for (auto* nd : order) stored += nd->particle_indices.size();
    *n_nodes = order.size();
    *n_stored = stored;
    if (!level) return 0;
    // id of the first child = number of nodes queued before it
    size_t next_child = 1;
    int64_t off = 0;
    for (size_t k = 0; k < order.size(); ++k) {
        const auto* nd = order[k];
        level[k] = nd->level;
        center[3 * k + 0] = nd->center.x;
        center[3 * k + 1] = nd->center.y;
        center[3 * k + 2] = nd->center.z;
        size[k] = nd->size;
        if (nd->is_leaf) {
            first_child[k] = -1;
        } else {
            first_child[k] = (int32_t)next_child;
            next_child += 8;
        }
        part_off[k] = off;
        for (size_t p : nd->particle_indices) part_idx[off++] = (int32_t)p;
        node_mass[k] = nd->total_mass;
        com[3 * k + 0] = nd->center_of_mass.x;
        com[3 * k + 1] = nd->center_of_mass.y;
        com[3 * k + 2] = nd->center_of_mass.z;
    }
    part_off[order.size()] = off;
    return 0;
}

double ref_hubble_a(double a, double omega_m, double omega_lambda, double omega_k, double h) {
    physics::CosmologyParams p;
    p.omega_m = omega_m; p.omega_lambda = omega_lambda; p.omega_k = omega_k; p.h = h;
    physics::CosmologyModel m(p);
    return m.hubble_parameter_a(a);
}

// InitialConditionsGenerator::generate_particles with the parameters of
// examples/zeldovich_test.cpp:12-29 (grid/seed/box/z configurable).
int ref_zeldovich(size_t grid, float box, double z_init, uint32_t seed, size_t n,
                  float* pos3, float* vel3, float* mass) {
    try {
        Quiet q;
        physics::CosmologyParams cp;
        cp.omega_m = 0.31; cp.omega_lambda = 0.69; cp.h = 0.67; cp.sigma_8 = 0.81; cp.n_s = 0.965;
        physics::CosmologyModel cosmo(cp);
        physics::InitialConditionsParams ip;
        ip.grid_size = grid; ip.box_size = box; ip.z_initial = z_init;
        ip.ps_type = physics::PowerSpectrumType::EISENSTEIN_HU;
        ip.random_seed = seed; ip.normalize_at_z0 = true;
        physics::InitialConditionsGenerator gen(ip, cosmo);
        std::vector<float3> p, v; std::vector<float> m;
        gen.generate_particles(n, p, v, m);
        if (p.size() < n) return 2;
        for (size_t i = 0; i < n; ++i) {
            pos3[3 * i] = p[i].x; pos3[3 * i + 1] = p[i].y; pos3[3 * i + 2] = p[i].z;
            vel3[3 * i] = v[i].x; vel3[3 * i + 1] = v[i].y; vel3[3 * i + 2] = v[i].z;
            mass[i] = m[i];
        }
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

// Scalars of the IC generator: normalised P(k) (InitialConditionsGenerator::get_power_spectrum,
// initial_conditions.cpp:173-199, after normalize_power_spectrum :131-144) and the growth
// factor / growth rate / H at a_init (cosmology_model.hpp:49-97).
int ref_ic_scalars(double z_init, size_t nk, const double* k, double* pk, double* growth,
                   double* rate, double* hubble) {
    try {
        Quiet q;
        physics::CosmologyParams cp;
        cp.omega_m = 0.31; cp.omega_lambda = 0.69; cp.h = 0.67; cp.sigma_8 = 0.81; cp.n_s = 0.965;
        physics::CosmologyModel cosmo(cp);
        physics::InitialConditionsParams ip;
        ip.grid_size = 8; ip.box_size = 100.0f; ip.z_initial = z_init;
        ip.ps_type = physics::PowerSpectrumType::EISENSTEIN_HU;
        ip.normalize_at_z0 = true;
        physics::InitialConditionsGenerator gen(ip, cosmo);
        for (size_t i = 0; i < nk; ++i) pk[i] = gen.get_power_spectrum(k[i]);
        const double a = cosmo.z_to_a(z_init);
        *growth = cosmo.growth_factor(a);
        *rate = cosmo.growth_rate(a);
        *hubble = cosmo.hubble_parameter_a(a);
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

// initial_conditions_utils::generate_random_particles (initial_conditions.cpp:800-821)
int ref_random_particles(size_t n, float box, uint32_t seed, float* pos3, float* vel3, float* mass) {
    std::vector<float3> p, v; std::vector<float> m;
    physics::initial_conditions_utils::generate_random_particles(n, box, p, v, m, seed);
    for (size_t i = 0; i < n; ++i) {
        pos3[3 * i] = p[i].x; pos3[3 * i + 1] = p[i].y; pos3[3 * i + 2] = p[i].z;
        vel3[3 * i] = v[i].x; vel3[3 * i + 1] = v[i].y; vel3[3 * i + 2] = v[i].z;
        mass[i] = m[i];
    }
    return 0;
}

// NewtonianGravityKernel sign quirk (force_computer_factory.cpp:150-178), documented only.
void ref_newtonian_pair(const float* p1, const float* p2, float m1, float m2, float* f1, float* f2) {
    forces::NewtonianGravityKernel k;
    float3 a = make_float3(p1[0], p1[1], p1[2]), b = make_float3(p2[0], p2[1], p2[2]);
    float3 fa, fb;
    k.compute_pairwise_force(a, b, m1, m2, fa, fb);
    f1[0] = fa.x; f1[1] = fa.y; f1[2] = fa.z;
    f2[0] = fb.x; f2[1] = fb.y; f2[2] = fb.z;
}

}  // extern "C"
