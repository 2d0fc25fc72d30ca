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
    have_forces_ = false;
    current_step_ = 0;
}

void B200LambdaCDMSimulation::set_particles_spatially_ordered(const float* pos3, const float* vel3, const float* mass) {
    const size_t n = num_particles_;
    order_.assign(n, 0);
    if (n == 0) return;
    // keys need positions relative to a cube centred on the origin
    const bool centred = method_ == B200ForceMethod::Tree || method_ == B200ForceMethod::TreeFixed;
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
    for (size_t k = 0; k < n; ++k) {
        const size_t i = (size_t)order_[k];
        for (int c = 0; c < 3; ++c) { ps[3 * k + c] = pos3[3 * i + c]; vs[3 * k + c] = vel3[3 * i + c]; }
        if (mass) ms[k] = mass[i];
    }
    set_particles(ps.data(), vs.data(), mass ? ms.data() : nullptr);
}

void B200LambdaCDMSimulation::initialize_particles(uint32_t seed) {
    // generate_initial_conditions (lambda_cdm_impl.cu:26-49): uniform positions in the box, Gaussian
    // velocities with dispersion 100*sqrt(omega_m) (:153), unit masses -- seeded here.
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> uni(0.0f, box_size_);
    std::normal_distribution<float> nrm(0.0f, 100.0f * (float)std::sqrt(params_.omega_m));
    std::vector<float> pos(3 * num_particles_), vel(3 * num_particles_);
    for (size_t i = 0; i < num_particles_; ++i) {
        for (int k = 0; k < 3; ++k) pos[3 * i + k] = uni(rng);
        for (int k = 0; k < 3; ++k) vel[3 * i + k] = nrm(rng);
    }
    set_particles(pos.data(), vel.data(), nullptr);
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
    const bool centred = method_ == B200ForceMethod::Tree || method_ == B200ForceMethod::TreeFixed;
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
    have_forces_ = false;
    current_step_ = 0;
}

void B200LambdaCDMSimulation::copy_particles_to_host(std::vector<Particle>& particles) const {
    particles.resize(n_local_);
    if (!n_local_) return;
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

void B200LambdaCDMSimulation::set_force_method(B200ForceMethod m, float theta, int leaf_capacity, int max_depth) {
    method_ = m; theta_ = theta; leaf_capacity_ = leaf_capacity; max_depth_ = max_depth;
    have_forces_ = false;
}

void B200LambdaCDMSimulation::compute_forces() {
    const size_t n = num_particles_;
    if (n == 0) return;
    // every rank sees all sources (replicated positions), and evaluates its own targets only
    if (method_ == B200ForceMethod::Tree || method_ == B200ForceMethod::TreeFixed ||
        method_ == B200ForceMethod::TreeFixedPeriodic) {
        if (method_ == B200ForceMethod::Tree)
            check(b200_tree_build_dev(ctx_, d_posm_, n, box_size_, leaf_capacity_, max_depth_, stream_), "tree build");
        else
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

void B200LambdaCDMSimulation::step(double dt) {
    const size_t n = num_particles_;
    if (n == 0) { ++current_step_; return; }
    if (!have_forces_) compute_forces();
    const float wrap = (method_ == B200ForceMethod::Direct || method_ == B200ForceMethod::TreeFixedPeriodic)
                           ? box_size_ : 0.0f;                                   // the periodic methods wrap into [0, box)
    // lambda_cdm_impl.cu:167-213: kick(dt/2, a) -> drift(dt) -> a update -> forces -> kick(dt/2, a_new)
    void* my_posm = (char*)d_posm_ + i0_ * 16;
    check(b200_leapfrog_dev(ctx_, my_posm, d_vel_, d_acc_, n_local_, 1, (float)(dt * 0.5), scale_factor_, (float)dt, wrap,
                            stream_),
          "kick+drift");
    update_scale_factor(dt);
    if (world_ > 1) check(b200_allgather_sources_dev(ctx_, d_posm_, n, stream_), "all-gather sources");
    compute_forces();
    check(b200_leapfrog_dev(ctx_, my_posm, d_vel_, d_acc_, n_local_, 1, (float)(dt * 0.5), scale_factor_, 0.0f, wrap,
                            stream_),
          "closing kick");
    check(b200_ctx_sync(ctx_, stream_), "sync");
    ++current_step_;
}

void B200LambdaCDMSimulation::copy_positions_to_host(float* positions) const {
    if (!num_particles_) return;
    check(b200_unpack_pos3_dev(ctx_, d_posm_, num_particles_, d_tmp3_, stream_), "unpack");
    check(b200_memcpy_d2h(ctx_, positions, d_tmp3_, num_particles_ * 12, stream_), "download positions");
}

void B200LambdaCDMSimulation::copy_velocities_to_host(float* velocities) const {
    if (!n_local_) return;
    check(b200_memcpy_d2h(ctx_, velocities, d_vel_, n_local_ * 12, stream_), "download velocities");
}

void B200LambdaCDMSimulation::copy_forces_to_host(float* accelerations) const {
    if (!n_local_) return;
    check(b200_memcpy_d2h(ctx_, accelerations, d_acc_, n_local_ * 12, stream_), "download accelerations");
}

}  // namespace physics
