// b200_force_computers.hpp -- the reference-facing plugin layer, in the
// reference's own language: IForceComputer / IIntegrator / ICosmologyModel
// implementations that forward to the C ABI of libb200grav.so.
//
// Compiled against the HOST APPLICATION's headers (-I<reference>/include), the
// way any plugin is: nothing from the reference is copied here.
//
//   forces::DirectForceComputer    fills the forward declaration at
//                                  include/forces/force_computer_factory.hpp:14
//                                  (the reference never defines it; its
//                                  registration is commented out at
//                                  src/forces/force_computer_factory.cpp:135-138)
//   forces::B200TreeForceComputer  same interface, setters and getters as
//                                  forces::TreeForceComputer
//                                  (include/forces/tree_force_computer.hpp:32-106),
//                                  same tree node for node, on the GPU
//   physics::B200LeapfrogIntegrator  core::IIntegrator (interfaces.hpp:42-49)
//   physics::LambdaCDMModel          core::ICosmologyModel (interfaces.hpp:51-59)
//                                    over physics::CosmologyModel
//
// Error behaviour follows the reference (src/forces/tree_force_computer.cpp:
// 46-96): initialize() returns false on failure -- but never downgrades to a
// CPU path -- and compute_forces() throws std::runtime_error.
#pragma once

#include <any>
#include <cstddef>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "core/interfaces.hpp"
#include "forces/force_computer_factory.hpp"
#include "physics/cosmology_model.hpp"

struct b200_ctx;

namespace forces {

class B200ComputerBase : public core::IForceComputer {
protected:
    std::string name_;
    b200_ctx* ctx_ = nullptr;
    int device_ = 0;
    size_t max_particles_ = 1000000;           // tree_force_computer.cpp:37
    mutable size_t force_evaluations_ = 0;

    explicit B200ComputerBase(const std::string& name) : name_(name) {}
    void require_ctx() const;
    void use_device(int device);               // ForceComputeParameters::cuda_device_id
    // Opt-in (set_pin_host_arrays): page-lock the caller's arrays the first time they are seen, so that the host
    // entry points copy at DMA speed instead of through the driver's pageable staging.  Only for callers whose
    // arrays outlive this object's finalize() -- the engine's particle arrays do (simulation_engine.hpp:60-63).
    bool pin_host_arrays_ = false;
    std::vector<std::pair<const void*, size_t>> pinned_;
    void pin(const void* ptr, size_t bytes);
    void unpin_all();

public:
    ~B200ComputerBase() override;
    bool initialize(const core::SimulationContext& context) override;
    void finalize() override;                   // idempotent (registry + dtor both call it)
    std::string get_name() const override { return name_; }
    std::string get_version() const override { return "1.0.0-b200"; }
    bool supports_gpu() const override { return true; }
    bool supports_mpi() const override { return false; }   // sharding is NCCL, not MPI
    size_t get_max_particles() const override { return max_particles_; }
    void set_max_particles(size_t n) { max_particles_ = n; }
    void set_cuda_device(int device) { device_ = device; }
    void set_pin_host_arrays(bool on) { pin_host_arrays_ = on; if (!on) unpin_all(); }
    size_t get_force_evaluations() const { return force_evaluations_; }
    b200_ctx* native_handle() const { return ctx_; }
};

// "DirectForceComputer": softened direct sum, acceleration output, G = 1.
class DirectForceComputer : public B200ComputerBase {
    float softening_ = 0.01f;                   // ForceComputeParameters::softening_length default
    float box_size_ = 0.0f;                     // 0 = open boundaries (the CPU leaf loop); >0 = minimum image
public:
    explicit DirectForceComputer(const std::string& name) : B200ComputerBase(name) {}
    std::string get_type() const override { return "DirectForceComputer"; }
    void compute_forces(const float* positions, const float* masses, float* forces,
                        size_t num_particles, const std::any& params = {}) override;
    void set_softening(float eps) { softening_ = eps; }
    float get_softening() const { return softening_; }
    void set_periodic_box(float box) { box_size_ = box; }
};

// "TreeForceComputer" on the GPU: identical octree, theta walk, quirks included.
class B200TreeForceComputer : public B200ComputerBase {
    float theta_ = 0.5f;
    size_t leaf_capacity_ = 8;
    int max_depth_ = 20;
    float box_size_ = 100.0f;
    mutable size_t tree_traversals_ = 0;
    bool fixed_physics_ = false;                // see set_fixed_physics
    float softening_ = 0.01f;                   // fixed-physics mode only (the reference hard-codes 0.01)
public:
    explicit B200TreeForceComputer(const std::string& name) : B200ComputerBase(name) {}
    B200TreeForceComputer(const std::string& name, float theta, size_t leaf_capacity = 8, int max_depth = 20)
        : B200ComputerBase(name), theta_(theta), leaf_capacity_(leaf_capacity), max_depth_(max_depth) {}
    std::string get_type() const override { return "TreeForceComputer"; }
    void compute_forces(const float* positions, const float* masses, float* forces,
                        size_t num_particles, const std::any& params = {}) override;
    // tree_force_computer.hpp:78-80: the two phases of compute_forces (the tree stays on the device in between;
    // `positions` of compute_tree_forces must be the array build_tree saw, as in the reference's own call :125)
    void build_tree(const float* positions, const float* masses, size_t num_particles);
    void compute_tree_forces(const float* positions, float* forces, size_t num_particles) const;
    // tree_force_computer.hpp:83-93
    void set_opening_angle(float theta) { theta_ = theta; }
    void set_leaf_capacity(size_t capacity) { leaf_capacity_ = capacity; }
    void set_max_depth(int depth) { max_depth_ = depth; }
    void set_box_size(float size) { box_size_ = size; }
    float get_opening_angle() const { return theta_; }
    size_t get_leaf_capacity() const { return leaf_capacity_; }
    int get_max_depth() const { return max_depth_; }
    float get_box_size() const { return box_size_; }
    // Not in the reference: Barnes-Hut without the CPU tree's quirks (no orphaned particles, real
    // masses in leaf pairs, root cube fitted to the data, softening from set_softening() or
    // ForceComputeParameters::softening_length) -- b200_tree_build_fixed_dev.  Off by default:
    // the default is the reference's tree, node for node.
    void set_fixed_physics(bool on) { fixed_physics_ = on; }
    bool get_fixed_physics() const { return fixed_physics_; }
    void set_softening(float eps) { softening_ = eps; }
    // tree_force_computer.hpp:96-98 (the reference adds N to both per compute_forces call, :126-127)
    size_t get_tree_traversals() const { return tree_traversals_; }
    void reset_statistics() const { force_evaluations_ = 0; tree_traversals_ = 0; }
    float get_tree_efficiency() const {            // tree_force_computer.cpp:420-423
        const size_t n2 = get_max_particles() * get_max_particles();
        return n2 > 0 ? static_cast<float>(force_evaluations_) / n2 : 0.0f;
    }
    // What the reference's counters do not say: nodes visited, cell and pair interactions of one more walk
    // over the current tree (valid after a compute_forces / build_tree call).
    void count_interactions(size_t num_particles, unsigned long long counts[3]) const;
    // tree_force_computer.hpp:100-103 (valid after a compute_forces call)
    size_t get_tree_depth() const;
    size_t get_node_count() const;
    size_t get_leaf_count() const;
};

// Registers the two computers under the reference's convenience names
// ("DirectForceComputer", "TreeForceComputer" -- the latter replaces the CPU
// one) and under "B200DirectForceComputer" / "B200TreeForceComputer".
void register_b200_force_computers(bool replace_cpu_tree = true);

}  // namespace forces

namespace physics {

// Parameters a caller may pass through IIntegrator::step's std::any.
struct LeapfrogStepParams {
    const float* masses = nullptr;     // nullptr = unit masses
    double scale_factor = 1.0;
    float box_size = 0.0f;             // <= 0: no wrap
    int n_kicks = 1;                   // half-kicks of dt/2 applied before the drift
    bool drift = true;
};

// KDK pieces of LambdaCDMSimulationImpl::step (src/physics/lambda_cdm_impl.cu:
// 167-213) on host arrays: step() = n_kicks x kick(dt/2) then drift(dt).
class B200LeapfrogIntegrator : public core::IIntegrator {
    std::string name_;
    b200_ctx* ctx_ = nullptr;
    int device_ = 0;
public:
    explicit B200LeapfrogIntegrator(const std::string& name) : name_(name) {}
    ~B200LeapfrogIntegrator() override;
    bool initialize(const core::SimulationContext& context) override;
    void finalize() override;
    std::string get_type() const override { return "LeapfrogIntegrator"; }
    std::string get_name() const override { return name_; }
    std::string get_version() const override { return "1.0.0-b200"; }
    void step(float* positions, float* velocities, const float* forces, size_t num_particles,
              double dt, const std::any& params = {}) override;
    double get_recommended_timestep() const override { return 1e-3; }   // cuda_nbody_test.cpp:53
    bool is_symplectic() const override { return true; }
};

// core::ICosmologyModel over the reference's physics::CosmologyModel.
class LambdaCDMModel : public core::ICosmologyModel {
    std::string name_;
    CosmologyModel model_;
public:
    explicit LambdaCDMModel(const std::string& name, const CosmologyParams& p = CosmologyParams())
        : name_(name), model_(p) {}
    bool initialize(const core::SimulationContext&) override { return true; }
    void finalize() override {}
    std::string get_type() const override { return "LambdaCDMModel"; }
    std::string get_name() const override { return name_; }
    std::string get_version() const override { return "1.0.0-b200"; }
    double hubble_function(double a) const override { return model_.hubble_parameter_a(a); }
    double growth_factor(double a) const override { return model_.growth_factor(a); }
    double omega_matter(double a) const override { return model_.omega_matter_a(a); }
    double omega_lambda(double a) const override { return model_.omega_lambda_a(a); }
    // lambda_cdm_impl.cu:261-269
    void update_scale_factor(double& a, double dt) const override { a += a * model_.hubble_parameter_a(a) * dt; }
};

}  // namespace physics
