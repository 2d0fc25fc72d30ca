// b200_simulation.cpp -- see the header.
#include "b200_simulation.hpp"

#include <cmath>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200grav.h"

namespace physics {

void B200LambdaCDMSimulation::check(int status, const char* where) const {
    if (status != B200_OK) throw std::runtime_error(std::string(where) + ": " + b200_error_string(status));
}

B200LambdaCDMSimulation::B200LambdaCDMSimulation(size_t num_particles, float box_size, const CosmologyParams& params,
                                                 int cuda_device)
    : params_(params), cosmology_(params), num_particles_(num_particles), box_size_(box_size),
      n_local_(num_particles) {
    // lambda_cdm_impl.cu:96-99: no suitable device -> throw (there is no CPU path)
    check(b200_ctx_create(cuda_device, num_particles, &ctx_), "B200LambdaCDMSimulation");
    stream_ = b200_ctx_stream(ctx_);
    const size_t n = num_particles_ ? num_particles_ : 1;
    check(b200_device_alloc(ctx_, n * 16, &d_posm_), "alloc positions");
    check(b200_device_alloc(ctx_, n * 12 + 16, &d_vel_), "alloc velocities");
    check(b200_device_alloc(ctx_, n * 12 + 16, &d_acc_), "alloc accelerations");
    check(b200_device_alloc(ctx_, n * 12 + 16, &d_tmp3_), "alloc staging");
}

B200LambdaCDMSimulation::~B200LambdaCDMSimulation() {
    if (!ctx_) return;
    b200_device_free(ctx_, d_posm_);
    b200_device_free(ctx_, d_vel_);
    b200_device_free(ctx_, d_acc_);
    b200_device_free(ctx_, d_tmp3_);
    if (d_arrival_) b200_device_free(ctx_, d_arrival_);
    b200_ctx_destroy(ctx_);
}

void B200LambdaCDMSimulation::enable_sharding(const unsigned char* nccl_unique_id, int rank, int world) {
    check(b200_shard_init(ctx_, nccl_unique_id, rank, world), "shard init");
    rank_ = rank;
    world_ = world;
    check(b200_shard_range(num_particles_, rank, world, &i0_, &n_local_), "shard range");
    have_forces_ = false;
}

void B200LambdaCDMSimulation::set_particles(const float* pos3, const float* vel3, const float* mass) {
    const size_t n = num_particles_;
    if (n == 0) return;
    if (d_arrival_) { b200_device_free(ctx_, d_arrival_); d_arrival_ = nullptr; }      // stored in insertion order
    void* d_mass = nullptr;
    check(b200_memcpy_h2d(ctx_, d_tmp3_, pos3, n * 12, stream_), "upload positions");
    if (mass) {   // velocities' buffer doubles as the mass staging area before the velocities land
        check(b200_memcpy_h2d(ctx_, d_vel_, mass, n * 4, stream_), "upload masses");
        d_mass = d_vel_;
    }
    check(b200_pack_posm_dev(ctx_, d_tmp3_, d_mass, n, d_posm_, stream_), "pack");
    if (n_local_)
        check(b200_memcpy_h2d(ctx_, d_vel_, vel3 + 3 * i0_, n_local_ * 12, stream_), "upload velocities");
    check(b200_ctx_sync(ctx_, stream_), "sync");
    generated_frame_ = Frame::Unknown;
    have_forces_ = false;
    pending_kick_ = false;
    current_step_ = 0;
}

void B200LambdaCDMSimulation::set_particles_spatially_ordered(const float* pos3, const float* vel3, const float* mass) {
    const size_t n = num_particles_;
    order_.assign(n, 0);
    if (n == 0) return;
    // keys need positions relative to a cube centred on the origin
    const bool centred = needs_centred_frame(method_);
    const float shift = centred ? 0.0f : 0.5f * box_size_;
    std::vector<float> p(3 * n);
    for (size_t i = 0; i < 3 * n; ++i) p[i] = pos3[i] - shift;
    void* d_perm = nullptr;
    check(b200_device_alloc(ctx_, n * 4, &d_perm), "alloc order");
    int rc = b200_memcpy_h2d(ctx_, d_tmp3_, p.data(), n * 12, stream_);
    if (rc == B200_OK) rc = b200_pack_posm_dev(ctx_, d_tmp3_, nullptr, n, d_posm_, stream_);
    if (rc == B200_OK) rc = b200_spatial_order_dev(ctx_, d_posm_, n, box_size_, d_perm, stream_);
    if (rc == B200_OK) rc = b200_memcpy_d2h(ctx_, order_.data(), d_perm, n * 4, stream_);
    b200_device_free(ctx_, d_perm);
    check(rc, "spatial order");
    std::vector<float> ps(3 * n), vs(3 * n), ms(mass ? n : 0);
    std::vector<int> arrival(n);
    for (size_t k = 0; k < n; ++k) {
        const size_t i = (size_t)order_[k];
        for (int c = 0; c < 3; ++c) { ps[3 * k + c] = pos3[3 * i + c]; vs[3 * k + c] = vel3[3 * i + c]; }
        if (mass) ms[k] = mass[i];
        arrival[i] = (int)k;                    // the caller's particle i sits in slot k
    }
    set_particles(ps.data(), vs.data(), mass ? ms.data() : nullptr);
    // the reference-faithful tree depends on insertion order = the caller's index order: hand it to the build
    if (!d_arrival_) check(b200_device_alloc(ctx_, n * sizeof(int), &d_arrival_), "alloc arrival order");
    check(b200_memcpy_h2d(ctx_, d_arrival_, arrival.data(), n * sizeof(int), stream_), "upload arrival order");
    check(b200_ctx_sync(ctx_, stream_), "sync");
}

void B200LambdaCDMSimulation::initialize_particles(uint32_t seed) {
    // generate_initial_conditions (lambda_cdm_impl.cu:26-49): uniform positions in the box, Gaussian
    // velocities with dispersion 100*sqrt(omega_m) (:153), unit masses -- seeded here.
    // The tree methods' root cube is centred on the origin (tree_force_computer.cpp:132-133), the periodic
    // methods wrap into [0, box): positions are drawn in the frame the force method works in.
    const float shift = needs_centred_frame(method_) ? 0.5f * box_size_ : 0.0f;
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> uni(0.0f, box_size_);
    std::normal_distribution<float> nrm(0.0f, 100.0f * (float)std::sqrt(params_.omega_m));
    std::vector<float> pos(3 * num_particles_), vel(3 * num_particles_);
    for (size_t i = 0; i < num_particles_; ++i) {
        for (int k = 0; k < 3; ++k) pos[3 * i + k] = uni(rng) - shift;
        for (int k = 0; k < 3; ++k) vel[3 * i + k] = nrm(rng);
    }
    set_particles(pos.data(), vel.data(), nullptr);
    generated_frame_ = shift > 0.0f ? Frame::Centred : Frame::Box;
}

void B200LambdaCDMSimulation::set_initial_conditions_from_power_spectrum(uint32_t seed, double z_initial,
                                                                         bool use_2lpt) {
    const size_t n = num_particles_;
    if (n == 0) return;
    b200_ic_params p;
    b200_ic_params_default(&p);
    int grid = 4;
    while ((size_t)grid * grid * grid < n) grid *= 2;
    p.grid = grid;
    p.box = box_size_;
    p.z_initial = z_initial;
    p.seed = seed;
    p.use_2lpt = use_2lpt ? 1 : 0;
    p.omega_m = params_.omega_m; p.omega_lambda = params_.omega_lambda; p.omega_k = params_.omega_k;
    p.h = params_.h; p.sigma_8 = params_.sigma_8; p.n_s = params_.n_s;
    const bool centred = needs_centred_frame(method_);
    p.origin_shift = centred ? 0.5f * box_size_ : 0.0f;
    // all N velocities land in the staging buffer; this rank keeps its own range
    void* d_vel_all = nullptr;
    check(b200_device_alloc(ctx_, n * 12 + 16, &d_vel_all), "alloc IC velocities");
    const int rc = b200_zeldovich_ics_dev(ctx_, &p, n, d_posm_, d_vel_all, nullptr, stream_);
    if (rc == B200_OK && n_local_) {
        std::vector<float> v(3 * n_local_);
        const int r2 = b200_memcpy_d2h(ctx_, v.data(), (char*)d_vel_all + i0_ * 12, n_local_ * 12, stream_);
        if (r2 == B200_OK) b200_memcpy_h2d(ctx_, d_vel_, v.data(), n_local_ * 12, stream_);
        b200_ctx_sync(ctx_, stream_);
    }
    b200_device_free(ctx_, d_vel_all);
    check(rc, "Zel'dovich initial conditions");
    scale_factor_ = 1.0 / (1.0 + z_initial);
    generated_frame_ = centred ? Frame::Centred : Frame::Box;
    have_forces_ = false;
    pending_kick_ = false;
    current_step_ = 0;
}

void B200LambdaCDMSimulation::copy_particles_to_host(std::vector<Particle>& particles) const {
    particles.resize(n_local_);
    if (!n_local_) return;
    flush_pending_kick();
    std::vector<float> posm(4 * n_local_), vel(3 * n_local_);
    check(b200_memcpy_d2h(ctx_, posm.data(), (const char*)d_posm_ + i0_ * 16, n_local_ * 16, stream_), "download particles");
    check(b200_memcpy_d2h(ctx_, vel.data(), d_vel_, n_local_ * 12, stream_), "download velocities");
    for (size_t i = 0; i < n_local_; ++i) {
        Particle& q = particles[i];
        q.position.x = posm[4 * i]; q.position.y = posm[4 * i + 1]; q.position.z = posm[4 * i + 2];
        q.velocity.x = vel[3 * i]; q.velocity.y = vel[3 * i + 1]; q.velocity.z = vel[3 * i + 2];
        q.mass = posm[4 * i + 3];
        q.id = (uint64_t)(i0_ + i);
    }
}

void B200LambdaCDMSimulation::power_spectrum(int grid, std::vector<float>& k, std::vector<float>& pk,
                                             std::vector<int>& modes, bool shot_noise_correction) const {
    const size_t nb = grid > 0 ? (size_t)grid / 2 : 0;
    k.assign(nb, 0.0f); pk.assign(nb, 0.0f); modes.assign(nb, 0);
    check(b200_power_spectrum_dev(ctx_, d_posm_, num_particles_, grid, box_size_, /*mass_weighted=*/1,
                                  shot_noise_correction ? 1 : 0, k.data(), pk.data(), modes.data(), stream_),
          "power spectrum");
}

bool B200LambdaCDMSimulation::needs_centred_frame(B200ForceMethod m) {
    return m == B200ForceMethod::Tree || m == B200ForceMethod::TreeFixed;
}

void B200LambdaCDMSimulation::set_force_method(B200ForceMethod m, float theta, int leaf_capacity, int max_depth) {
    // Particles this object generated sit in the frame of the method that was current then (origin-centred for
    // Tree / TreeFixed, [0, box) otherwise).  The reference tree's root cube is fixed at the origin and the periodic
    // methods wrap into [0, box), so a method that needs the other frame would silently see 7/8 of the particles
    // outside its domain: refuse.  (Particles handed in through set_particles are the caller's responsibility.)
    if (generated_frame_ != Frame::Unknown) {
        const bool want_centred = needs_centred_frame(m);
        const bool frame_matters = m == B200ForceMethod::Tree || m == B200ForceMethod::Direct ||
                                   m == B200ForceMethod::TreeFixedPeriodic;       // TreeFixed / DirectOpen fit any frame
        if (frame_matters && want_centred != (generated_frame_ == Frame::Centred))
            throw std::logic_error("B200LambdaCDMSimulation::set_force_method: the particles were generated in the "
                                   "coordinate frame of the previous force method; set the force method before "
                                   "initialize_particles / set_initial_conditions_from_power_spectrum");
    }
    flush_pending_kick();
    method_ = m; theta_ = theta; leaf_capacity_ = leaf_capacity; max_depth_ = max_depth;
    have_forces_ = false;
}

void B200LambdaCDMSimulation::compute_forces() {
    const size_t n = num_particles_;
    if (n == 0) return;
    flush_pending_kick();          // a deferred closing kick still needs the accelerations about to be replaced
    // every rank sees all sources (replicated positions), and evaluates its own targets only
    if (method_ == B200ForceMethod::Tree || method_ == B200ForceMethod::TreeFixed ||
        method_ == B200ForceMethod::TreeFixedPeriodic) {
        if (method_ == B200ForceMethod::Tree) {
            // Sharded runs build the octree by octants: rank r only the subtrees of its own octants of the root, then
            // the ranks exchange their walk tables (same tree, same forces; the replicated build was the part of a step
            // that did not shrink with the GPU count).  Particles stored in space-filling order keep the reference's
            // insertion order through the arrival array.
            const bool by_octant = world_ > 1 && world_ <= 8 && n > (size_t)leaf_capacity_;
            if (by_octant || d_arrival_) {
                check(b200_tree_build_part_dev(ctx_, d_posm_, d_arrival_, n, box_size_, leaf_capacity_, max_depth_,
                                               by_octant ? rank_ : 0, by_octant ? world_ : 1, stream_),
                      "tree build (octant part)");
                if (by_octant) check(b200_tree_forest_publish(ctx_, stream_), "tree table exchange");
            } else {
                check(b200_tree_build_dev(ctx_, d_posm_, n, box_size_, leaf_capacity_, max_depth_, stream_), "tree build");
            }
        } else
            check(b200_tree_build_fixed_dev(ctx_, d_posm_, n, leaf_capacity_, max_depth_, softening_, stream_),
                  "tree build (fixed physics)");
        check(b200_tree_set_periodic(ctx_, method_ == B200ForceMethod::TreeFixedPeriodic ? box_size_ : 0.0f),
              "tree periodic box");
        check(b200_tree_walk_dev(ctx_, i0_, n_local_, theta_, d_acc_, stream_), "tree walk");
    } else {
        const float box = (method_ == B200ForceMethod::Direct) ? box_size_ : 0.0f;   // K1/K2 are periodic
        check(b200_direct_forces_dev(ctx_, d_posm_, n, i0_, n_local_, softening_, box, d_acc_, stream_),
              "direct forces");
    }
    have_forces_ = true;
}

void B200LambdaCDMSimulation::compute_energy() {
    // launch_energy_computation (lambda_cdm_kernels.cu:492-516).  Potential energy by the same method as the
    // forces: the fixed-physics tree walks its potential (O(N log N); theta capped at 1/sqrt(3), where a
    // particle can no longer accept a cell that contains it), every other method takes the O(N^2) pair sum of
    // the direct-sum kernel -- minimum image whenever the force method is periodic.
    double e[2] = {0.0, 0.0};
    check_tree_overflow();
    flush_pending_kick();
    const bool tree_fixed = method_ == B200ForceMethod::TreeFixed || method_ == B200ForceMethod::TreeFixedPeriodic;
    if (tree_fixed && num_particles_ > 0) {
        check(b200_tree_build_fixed_dev(ctx_, d_posm_, num_particles_, leaf_capacity_, max_depth_, softening_, stream_),
              "tree build (fixed physics)");
        check(b200_tree_set_periodic(ctx_, method_ == B200ForceMethod::TreeFixedPeriodic ? box_size_ : 0.0f),
              "tree periodic box");
        check(b200_tree_energy_dev(ctx_, i0_, n_local_, d_vel_, theta_ < 0.577f ? theta_ : 0.577f, &e[0], &e[1], stream_),
              "tree energy");
    } else {
        const float box = (method_ == B200ForceMethod::Direct) ? box_size_ : 0.0f;
        check(b200_energy_dev(ctx_, d_posm_, num_particles_, i0_, n_local_, d_vel_, softening_, box, &e[0], &e[1], stream_),
              "energy");
    }
    if (world_ > 1) check(b200_allreduce_sum_f64(ctx_, e, 2), "energy all-reduce");
    kinetic_energy_ = e[0];
    potential_energy_ = e[1];
}

void B200LambdaCDMSimulation::update_scale_factor(double dt) {
    scale_factor_ += scale_factor_ * cosmology_.hubble_parameter_a(scale_factor_) * dt;   // lambda_cdm_impl.cu:261-269
}

float B200LambdaCDMSimulation::wrap_box() const {
    return (method_ == B200ForceMethod::Direct || method_ == B200ForceMethod::TreeFixedPeriodic)
               ? box_size_ : 0.0f;                                   // the periodic methods wrap into [0, box)
}

// The closing half-kick of a step is not launched by step(): it rides in the NEXT step's single
// kick-kick-drift pass (same scale factor, same dt/2 -- one launch per step instead of two, bit-identical
// velocities because the per-particle operations and their order do not change).  Anything that reads
// velocities first calls this.
void B200LambdaCDMSimulation::flush_pending_kick() const {
    if (!pending_kick_) return;
    pending_kick_ = false;
    void* my_posm = (char*)d_posm_ + i0_ * 16;
    check(b200_leapfrog_dev(ctx_, my_posm, d_vel_, d_acc_, n_local_, 1, pending_dt_kick_, pending_a_, 0.0f, wrap_box(),
                            stream_),
          "closing kick");
}

void B200LambdaCDMSimulation::step(double dt) {
    const size_t n = num_particles_;
    if (n == 0) { ++current_step_; return; }
    if (!have_forces_) compute_forces();
    const float wrap = wrap_box();
    const float dt_kick = (float)(dt * 0.5);
    // lambda_cdm_impl.cu:167-213: kick(dt/2, a) -> drift(dt) -> a update -> forces -> kick(dt/2, a_new)
    void* my_posm = (char*)d_posm_ + i0_ * 16;
    if (pending_kick_ && (pending_dt_kick_ != dt_kick || pending_a_ != scale_factor_)) flush_pending_kick();
    const int n_kicks = pending_kick_ ? 2 : 1;       // previous step's closing kick + this step's opening kick
    pending_kick_ = false;
    check(b200_leapfrog_dev(ctx_, my_posm, d_vel_, d_acc_, n_local_, n_kicks, dt_kick, scale_factor_, (float)dt, wrap,
                            stream_),
          "kick+drift");
    update_scale_factor(dt);
    if (world_ > 1) check(b200_allgather_sources_dev(ctx_, d_posm_, n, stream_), "all-gather sources");
    compute_forces();
    pending_kick_ = true;                             // closing kick with the new scale factor: deferred
    pending_dt_kick_ = dt_kick;
    pending_a_ = scale_factor_;
    ++current_step_;                                  // no host synchronisation here: read-backs synchronise
}

// The fixed-physics node table has no a-priori bound short of 2.3 N; a build that ran out of slots makes the walk
// write NaN.  Checked where the host synchronises anyway (read-backs, energies), not per step.
void B200LambdaCDMSimulation::check_tree_overflow() const {
    if (!have_forces_) return;
    if (method_ != B200ForceMethod::TreeFixed && method_ != B200ForceMethod::TreeFixedPeriodic) return;
    int overflow = 0;
    if (b200_tree_overflowed(ctx_, &overflow) == B200_OK && overflow)
        throw std::runtime_error("B200LambdaCDMSimulation: the fixed-physics octree ran out of node slots "
                                 "(pathological clustering); accelerations and everything integrated from them are NaN");
}

void B200LambdaCDMSimulation::copy_positions_to_host(float* positions) const {
    if (!num_particles_) return;
    check_tree_overflow();
    check(b200_unpack_pos3_dev(ctx_, d_posm_, num_particles_, d_tmp3_, stream_), "unpack");
    check(b200_memcpy_d2h(ctx_, positions, d_tmp3_, num_particles_ * 12, stream_), "download positions");
}

void B200LambdaCDMSimulation::copy_velocities_to_host(float* velocities) const {
    if (!n_local_) return;
    check_tree_overflow();
    flush_pending_kick();
    check(b200_memcpy_d2h(ctx_, velocities, d_vel_, n_local_ * 12, stream_), "download velocities");
}

void B200LambdaCDMSimulation::copy_forces_to_host(float* accelerations) const {
    if (!n_local_) return;
    check_tree_overflow();
    check(b200_memcpy_d2h(ctx_, accelerations, d_acc_, n_local_ * 12, stream_), "download accelerations");
}

}  // namespace physics
