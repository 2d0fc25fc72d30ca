// b200_simulation.hpp -- device-resident Lambda-CDM leapfrog driver with the public
// surface of the reference's physics::LambdaCDMSimulation
// (include/physics/lambda_cdm.hpp:22-75; behaviour src/physics/lambda_cdm_impl.cu).
// State (float4 positions+mass, velocities, accelerations) stays in HBM between
// steps; every numerical call goes through the C ABI (include/b200grav.h).
//
// Differences from the reference class, on purpose:
//  * step() orders kick -> drift -> forces -> kick on ONE stream (the reference
//    races its kick and drift on two streams, lambda_cdm_impl.cu:170-189) and the
//    first half-kick uses forces computed at the initial positions (the reference
//    reads uninitialised memory there); the closing half-kick is folded into the next
//    step's single kick-kick-drift pass (one leapfrog launch per step; velocity
//    read-backs apply it first), and step() does not synchronise the host;
//  * initialize_particles() is seeded (the reference seeds curand from the clock);
//  * the force method is selectable: Direct (periodic minimum image, like the
//    reference's K1/K2 kernels) or Tree (the CPU TreeForceComputer's semantics);
//  * compute_energy() runs the O(N^2) pair sum in the direct-sum kernel's potential
//    instance and returns doubles (the reference keeps float atomics, lambda_cdm_kernels.cu:402-407);
//    on a sharded run the two sums are all-reduced, so every rank reports the totals;
//  * enable_sharding(): target-sharded data parallelism over several GPUs.
//    One object per GPU (one process or one thread each); rank r integrates particles
//    [r*N/G, (r+1)*N/G) and the float4 positions are all-gathered with NCCL every step
//    (b200_allgather_sources_dev).  Every rank holds all positions; velocities and
//    accelerations exist only for the local range.
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

#include "physics/cosmology_model.hpp"
#include "physics/lambda_cdm.hpp"        // physics::Particle

struct b200_ctx;

namespace physics {

// Direct: minimum image + wrap (the reference's K1/K2); DirectOpen: open boundary; Tree: the CPU
// TreeForceComputer's tree, open boundary; TreeFixed: b200_tree_build_fixed_dev, open boundary;
// TreeFixedPeriodic: the same with minimum-image walks (b200_tree_set_periodic) and wrapped drifts.
enum class B200ForceMethod { Direct, DirectOpen, Tree, TreeFixed, TreeFixedPeriodic };

class B200LambdaCDMSimulation {
    b200_ctx* ctx_ = nullptr;
    void* stream_ = nullptr;                // the context's own cudaStream_t
    CosmologyParams params_;
    CosmologyModel cosmology_;
    size_t num_particles_;
    float box_size_;
    double scale_factor_ = 1.0;
    float softening_ = 0.01f;               // lambda_cdm_impl.cu:95
    size_t current_step_ = 0;
    B200ForceMethod method_ = B200ForceMethod::Direct;
    float theta_ = 0.5f;
    int leaf_capacity_ = 8, max_depth_ = 20;
    bool have_forces_ = false;
    double kinetic_energy_ = 0.0, potential_energy_ = 0.0;
    std::vector<int> order_;                // set_particles_spatially_ordered
    int rank_ = 0, world_ = 1;              // enable_sharding()
    size_t i0_ = 0, n_local_ = 0;           // this rank's particle range
    void* d_posm_ = nullptr;                // float4[N]
    void* d_vel_ = nullptr;                 // float[3N]
    void* d_acc_ = nullptr;                 // float[3N]
    void* d_tmp3_ = nullptr;                // float[3N] staging
    void* d_arrival_ = nullptr;             // int32[N]: slot of the caller's particle i (spatially ordered storage)

    // closing half-kick of the last step(), folded into the next step's kick-kick-drift pass
    mutable bool pending_kick_ = false;
    mutable float pending_dt_kick_ = 0.0f;
    mutable double pending_a_ = 1.0;

    enum class Frame { Unknown, Box, Centred };
    Frame generated_frame_ = Frame::Unknown;    // frame of the particles this object generated itself
    static bool needs_centred_frame(B200ForceMethod m);
    void check_tree_overflow() const;

    void check(int status, const char* where) const;
    void flush_pending_kick() const;
    float wrap_box() const;

public:
    B200LambdaCDMSimulation(size_t num_particles, float box_size, const CosmologyParams& params = CosmologyParams(),
                            int cuda_device = 0);
    ~B200LambdaCDMSimulation();
    B200LambdaCDMSimulation(const B200LambdaCDMSimulation&) = delete;
    B200LambdaCDMSimulation& operator=(const B200LambdaCDMSimulation&) = delete;

    // Initialization (lambda_cdm.hpp:41-44)
    // uniform in the box, v ~ N(0, 100*sqrt(omega_m)), m = 1; positions in [0, box), or in [-box/2, box/2) for the
    // tree methods, whose root cube is centred on the origin.  Set the force method FIRST.
    void initialize_particles(uint32_t seed = 12345);
    // host arrays of all N particles (every rank passes the same data), mass may be null
    void set_particles(const float* pos3, const float* vel3, const float* mass);
    // Same, but the particles are first put in space-filling-curve order (b200_spatial_order_dev; the cube is
    // centred on the origin for the tree methods and on box/2 otherwise): contiguous index ranges -- the shards
    // of a multi-GPU run -- become compact regions.  get_particle_order()[k] is the caller's index of stored
    // particle k.  The Tree method still builds the reference's tree: particles are inserted in the caller's index
    // order (arrival order handed to b200_tree_build_part_dev), wherever they are stored.
    void set_particles_spatially_ordered(const float* pos3, const float* vel3, const float* mass);
    const std::vector<int>& get_particle_order() const { return order_; }
    // Multi-GPU: call once, before set_particles/initialize_particles, on every rank with the
    // 128-byte id rank 0 obtained from b200_shard_unique_id().  Collective (blocks until all
    // `world` ranks have called it).
    void enable_sharding(const unsigned char* nccl_unique_id, int rank, int world);
    // lambda_cdm.hpp:42 -- declared by the reference, defined nowhere in it.  Here: Zel'dovich particles
    // generated on the device (b200_zeldovich_ics_dev: the reference generator's spectrum, growth and
    // velocity conventions, with the inverse FFT its displacement step lacks), grid = the smallest power
    // of two with grid^3 >= N, positions in [0, box) -- or origin-centred for the tree methods, whose root
    // cube is centred on the origin.  Every rank of a sharded run generates the same particles.
    void set_initial_conditions_from_power_spectrum(uint32_t seed = 12345, double z_initial = 49.0,
                                                    bool use_2lpt = false);   // InitialConditionsParams::use_2lpt
    void set_softening(float softening) { softening_ = softening; have_forces_ = false; }
    void set_force_method(B200ForceMethod m, float theta = 0.5f, int leaf_capacity = 8, int max_depth = 20);

    // Simulation methods (lambda_cdm.hpp:46-50)
    void step(double dt);
    void compute_forces();
    void update_scale_factor(double dt);
    void compute_energy();                                     // lambda_cdm_impl.cu:222-241

    // Cosmology functions (lambda_cdm.hpp:52-55)
    double hubble_function(double a) const { return cosmology_.hubble_parameter_a(a); }
    double growth_factor(double a) const { return cosmology_.growth_factor(a); }

    // Data access (lambda_cdm.hpp:57-60)
    void copy_positions_to_host(float* positions) const;       // float[3N]
    void copy_velocities_to_host(float* velocities) const;     // float[3*local count]
    void copy_forces_to_host(float* accelerations) const;      // float[3*local count]
    void copy_particles_to_host(std::vector<Particle>& particles) const;   // lambda_cdm.hpp:58; the local range
    // analysis::PowerSpectrumAnalyzer::compute_power_spectrum (src/analysis/power_spectrum.cu:53-84) on the
    // device-resident particles: grid/2 bins of width 2 pi / box
    void power_spectrum(int grid, std::vector<float>& k, std::vector<float>& pk, std::vector<int>& modes,
                        bool shot_noise_correction = true) const;
    size_t get_local_offset() const { return i0_; }            // first particle this rank integrates
    size_t get_local_count() const { return n_local_; }        // == N unless sharded
    int get_rank() const { return rank_; }
    int get_world_size() const { return world_; }

    // Accessors (lambda_cdm.hpp:62-71)
    double get_scale_factor() const { return scale_factor_; }
    double get_redshift() const { return 1.0 / scale_factor_ - 1.0; }
    size_t get_num_particles() const { return num_particles_; }
    float get_box_size() const { return box_size_; }
    size_t get_current_step() const { return current_step_; }
    double get_kinetic_energy() const { return kinetic_energy_; }        // after compute_energy()
    double get_potential_energy() const { return potential_energy_; }
    double get_total_energy() const { return kinetic_energy_ + potential_energy_; }
};

}  // namespace physics
