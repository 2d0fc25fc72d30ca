// b200_force_computers.cpp -- see the header.  Host C++ only; every numerical
// call goes through the C ABI (include/b200grav.h).
#include "b200_force_computers.hpp"

#include <iostream>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "b200grav.h"
#include "core/simulation_context.hpp"

namespace {

[[noreturn]] void fail(const char* where, int status) {
    throw std::runtime_error(std::string(where) + ": " + b200_error_string(status));
}

// A caller may hand a ForceComputeParameters through compute_forces' std::any
// (the factory accepts one and ignores it: force_computer_factory.cpp:14-28).
const forces::ForceComputeParameters* params_of(const std::any& a) {
    return std::any_cast<forces::ForceComputeParameters>(&a);
}

}  // namespace

namespace forces {

B200ComputerBase::~B200ComputerBase() {
    unpin_all();
    if (ctx_) b200_ctx_destroy(ctx_);
    ctx_ = nullptr;
}

void B200ComputerBase::pin(const void* ptr, size_t bytes) {
    if (!pin_host_arrays_ || !ctx_ || !ptr || bytes < (1u << 16)) return;        // small arrays: not worth a registration
    for (const auto& r : pinned_)
        if (r.first == ptr && r.second >= bytes) return;
    for (auto it = pinned_.begin(); it != pinned_.end();)                       // same base, grown: register afresh
        if (it->first == ptr) { b200_host_unregister(ctx_, const_cast<void*>(ptr)); it = pinned_.erase(it); } else ++it;
    if (pinned_.size() >= 8) { b200_host_unregister(ctx_, const_cast<void*>(pinned_.front().first)); pinned_.erase(pinned_.begin()); }
    if (b200_host_register(ctx_, const_cast<void*>(ptr), bytes) == B200_OK) pinned_.emplace_back(ptr, bytes);
}

void B200ComputerBase::unpin_all() {
    if (ctx_) for (const auto& r : pinned_) b200_host_unregister(ctx_, const_cast<void*>(r.first));
    pinned_.clear();
}

bool B200ComputerBase::initialize(const core::SimulationContext& context) {
    if (ctx_) return true;
    const int ctx_dev = context.get_cuda_device_id();
    if (ctx_dev > 0) device_ = ctx_dev;
    const int rc = b200_ctx_create(device_, 0, &ctx_);
    if (rc != B200_OK) {
        // the reference prints and downgrades to its CPU path here
        // (tree_force_computer.cpp:54-62); this plugin has none.
        std::cerr << get_type() << " initialization failed: " << b200_error_string(rc) << std::endl;
        ctx_ = nullptr;
        return false;
    }
    std::cout << get_type() << " initialized: " << name_ << " (B200 device " << device_ << ")" << std::endl;
    return true;
}

void B200ComputerBase::finalize() {
    unpin_all();
    if (ctx_) {
        b200_ctx_destroy(ctx_);
        ctx_ = nullptr;
        std::cout << get_type() << " finalized: " << name_ << std::endl;
    }
}

// ForceComputeParameters::cuda_device_id (force_computer_factory.hpp:39): a caller that hands parameters
// through compute_forces' std::any names the device there; the context moves to it (its scratch is
// reallocated on first use).  Fails loudly if that device is not a usable sm_100 part.
void B200ComputerBase::use_device(int device) {
    if (device == device_ && ctx_) return;
    if (device < 0) return;
    b200_ctx* fresh = nullptr;
    const int rc = b200_ctx_create(device, 0, &fresh);
    if (rc != B200_OK) fail((get_type() + ": cuda_device_id " + std::to_string(device)).c_str(), rc);
    unpin_all();
    if (ctx_) b200_ctx_destroy(ctx_);
    ctx_ = fresh;
    device_ = device;
}

void B200ComputerBase::require_ctx() const {
    if (!ctx_) throw std::runtime_error(get_type() + ": not initialized (no B200 context; there is no CPU fallback)");
}

void DirectForceComputer::compute_forces(const float* positions, const float* masses, float* forces,
                                         size_t num_particles, const std::any& params) {
    if (num_particles == 0) return;                    // tree_force_computer.cpp:83
    require_ctx();
    float eps = softening_;
    if (const ForceComputeParameters* p = params_of(params)) { eps = p->softening_length; use_device(p->cuda_device_id); }
    pin(positions, num_particles * 12); pin(masses, num_particles * 4); pin(forces, num_particles * 12);
    const int rc = b200_direct_forces_host(ctx_, positions, masses, forces, num_particles, eps, box_size_);
    if (rc != B200_OK) {
        std::cerr << "DirectForceComputer::compute_forces failed: " << b200_error_string(rc) << std::endl;
        fail("DirectForceComputer::compute_forces", rc);
    }
    force_evaluations_ += num_particles;
}

void B200TreeForceComputer::compute_forces(const float* positions, const float* masses, float* forces,
                                           size_t num_particles, const std::any& params) {
    if (num_particles == 0) return;
    require_ctx();
    float theta = theta_;
    size_t cap = leaf_capacity_;
    int depth = max_depth_;
    float eps = softening_;
    if (const ForceComputeParameters* p = params_of(params)) {
        theta = p->theta; cap = p->leaf_capacity; depth = p->tree_max_depth; eps = p->softening_length;
        use_device(p->cuda_device_id);
    }
    pin(positions, num_particles * 12); pin(masses, num_particles * 4); pin(forces, num_particles * 12);
    const int rc = fixed_physics_
                       ? b200_tree_forces_fixed_host(ctx_, positions, masses, forces, num_particles, theta, (int)cap,
                                                     depth, eps)
                       : b200_tree_forces_host(ctx_, positions, masses, forces, num_particles, theta, (int)cap,
                                               depth, box_size_);
    if (rc != B200_OK) {
        std::cerr << "TreeForceComputer::compute_forces failed: " << b200_error_string(rc) << std::endl;
        fail("TreeForceComputer::compute_forces", rc);
    }
    force_evaluations_ += num_particles;
    tree_traversals_ += num_particles;
}

void B200TreeForceComputer::build_tree(const float* positions, const float* masses, size_t num_particles) {
    if (num_particles == 0) return;
    require_ctx();
    const int rc = b200_tree_build_host(ctx_, positions, masses, num_particles, box_size_, (int)leaf_capacity_, max_depth_);
    if (rc != B200_OK) fail("TreeForceComputer::build_tree", rc);
}

void B200TreeForceComputer::compute_tree_forces(const float*, float* forces, size_t num_particles) const {
    if (num_particles == 0) return;
    require_ctx();
    const int rc = b200_tree_walk_host(ctx_, forces, num_particles, theta_);
    if (rc != B200_OK) fail("TreeForceComputer::compute_tree_forces", rc);
}

void B200TreeForceComputer::count_interactions(size_t num_particles, unsigned long long counts[3]) const {
    require_ctx();
    std::vector<float> scratch(3 * num_particles);
    uint64_t c[3] = {0, 0, 0};
    int rc = b200_tree_set_counting(ctx_, 1);
    if (rc == B200_OK) rc = b200_tree_walk_host(ctx_, scratch.data(), num_particles, theta_);
    if (rc == B200_OK) rc = b200_tree_counters(ctx_, c);
    b200_tree_set_counting(ctx_, 0);
    if (rc != B200_OK) fail("TreeForceComputer::count_interactions", rc);
    for (int k = 0; k < 3; ++k) counts[k] = c[k];
}

size_t B200TreeForceComputer::get_node_count() const {
    size_t n = 0;
    return (ctx_ && b200_tree_stats(ctx_, &n, nullptr, nullptr, nullptr) == B200_OK) ? n : 0;
}
size_t B200TreeForceComputer::get_leaf_count() const {
    size_t n = 0;
    return (ctx_ && b200_tree_stats(ctx_, nullptr, &n, nullptr, nullptr) == B200_OK) ? n : 0;
}
size_t B200TreeForceComputer::get_tree_depth() const {
    size_t n = 0;
    return (ctx_ && b200_tree_stats(ctx_, nullptr, nullptr, &n, nullptr) == B200_OK) ? n : 0;
}

void register_b200_force_computers(bool replace_cpu_tree) {
    ForceComputerFactory::register_force_computer<DirectForceComputer>("DirectForceComputer");
    ForceComputerFactory::register_force_computer<DirectForceComputer>("B200DirectForceComputer");
    ForceComputerFactory::register_force_computer<B200TreeForceComputer>("B200TreeForceComputer");
    if (replace_cpu_tree)
        ForceComputerFactory::register_force_computer<B200TreeForceComputer>("TreeForceComputer");
}

}  // namespace forces

namespace physics {

B200LeapfrogIntegrator::~B200LeapfrogIntegrator() {
    if (ctx_) b200_ctx_destroy(ctx_);
    ctx_ = nullptr;
}

bool B200LeapfrogIntegrator::initialize(const core::SimulationContext& context) {
    if (ctx_) return true;
    if (context.get_cuda_device_id() > 0) device_ = context.get_cuda_device_id();
    const int rc = b200_ctx_create(device_, 0, &ctx_);
    if (rc != B200_OK) {
        std::cerr << "LeapfrogIntegrator initialization failed: " << b200_error_string(rc) << std::endl;
        ctx_ = nullptr;
        return false;
    }
    return true;
}

void B200LeapfrogIntegrator::finalize() {
    if (ctx_) b200_ctx_destroy(ctx_);
    ctx_ = nullptr;
}

void B200LeapfrogIntegrator::step(float* positions, float* velocities, const float* forces,
                                  size_t num_particles, double dt, const std::any& params) {
    if (num_particles == 0) return;
    if (!ctx_) throw std::runtime_error("LeapfrogIntegrator: not initialized (no B200 context)");
    LeapfrogStepParams p;
    if (const LeapfrogStepParams* q = std::any_cast<LeapfrogStepParams>(&params)) p = *q;
    // the same parameters as a plain map, for hosts that do not want a plugin type in their translation units
    // (integration/engine_wiring.patch): "scale_factor", "n_kicks", "drift" (0/1), "box_size"; unit masses
    else if (const auto* m = std::any_cast<std::unordered_map<std::string, double>>(&params)) {
        auto get = [m](const char* k, double dflt) { auto it = m->find(k); return it == m->end() ? dflt : it->second; };
        p.scale_factor = get("scale_factor", 1.0);
        p.n_kicks = (int)get("n_kicks", 1.0);
        p.drift = get("drift", 1.0) != 0.0;
        p.box_size = (float)get("box_size", 0.0);
    }
    // lambda_cdm_impl.cu:170-189: kick by dt*0.5 (double, narrowed to the kernel's float dt), drift by dt
    const int rc = b200_leapfrog_host(ctx_, positions, velocities, forces, p.masses, num_particles, p.n_kicks,
                                      (float)(dt * 0.5), p.scale_factor, p.drift ? (float)dt : 0.0f, p.box_size);
    if (rc != B200_OK) fail("LeapfrogIntegrator::step", rc);
}

}  // namespace physics
