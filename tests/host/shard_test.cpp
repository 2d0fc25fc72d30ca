// shard_test.cpp -- the C++ multi-GPU path without Python: G host threads, one
// B200LambdaCDMSimulation per GPU, sharded with enable_sharding() (NCCL all-gather
// of the float4 positions through b200_allgather_sources_dev), against the same
// run on ONE GPU.  Tree: positions/velocities must be bitwise equal (the walk is a
// pure function of target and tree).  Direct: the kernel's work partition depends
// on the number of local targets, so the FP64 partial sums are added in another
// order -- agreement to 1e-6 relative.
//
// Needs >= 2 GPUs; prints SKIP and exits 0 otherwise.  Run on a multi-GPU box by
// tests/test_gpu_host_plugin.py.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "b200_simulation.hpp"
#include "b200grav.h"

static int failures = 0;
#define CHECK(cond, ...)                                               \
    do {                                                               \
        if (cond) { std::printf("PASS  "); } else { std::printf("FAIL  "); ++failures; } \
        std::printf(__VA_ARGS__); std::printf("\n");                   \
    } while (0)

using physics::B200ForceMethod;
using physics::B200LambdaCDMSimulation;

struct Result { std::vector<float> pos, vel; double a = 0; };

static void make_inputs(size_t n, std::vector<float>& pos, std::vector<float>& vel, std::vector<float>& mass) {
    std::mt19937 rng(42);
    std::uniform_real_distribution<float> u(-50.f, 50.f), um(0.5f, 1.5f);
    std::normal_distribution<float> nv(0.f, 100.f);
    pos.resize(3 * n); vel.resize(3 * n); mass.resize(n);
    for (size_t i = 0; i < 3 * n; ++i) pos[i] = u(rng);
    for (size_t i = 0; i < 3 * n; ++i) vel[i] = nv(rng);
    for (size_t i = 0; i < n; ++i) mass[i] = um(rng);
}

static void run_rank(size_t n, B200ForceMethod method, int steps, double dt, const unsigned char* id, int rank, int world,
                     Result* out, std::string* error) {
    try {
        std::vector<float> pos, vel, mass;
        make_inputs(n, pos, vel, mass);
        B200LambdaCDMSimulation sim(n, 100.0f, physics::CosmologyParams(), /*cuda_device=*/rank);
        sim.set_force_method(method);
        if (world > 1) sim.enable_sharding(id, rank, world);
        sim.set_particles(pos.data(), vel.data(), mass.data());
        for (int s = 0; s < steps; ++s) sim.step(dt);
        out->pos.resize(3 * n);
        out->vel.resize(3 * sim.get_local_count());
        sim.copy_positions_to_host(out->pos.data());
        sim.copy_velocities_to_host(out->vel.data());
        out->a = sim.get_scale_factor();
    } catch (const std::exception& e) {
        *error = e.what();
    }
}

static double max_abs_diff(const float* a, const float* b, size_t n) {
    double m = 0;
    for (size_t i = 0; i < n; ++i) m = std::fmax(m, std::fabs((double)a[i] - b[i]));
    return m;
}

int main() {
    // device count through the C ABI: contexts on devices 0 and 1 must both come up
    b200_ctx* probe = nullptr;
    int gpus = 0;
    while (gpus < 8 && b200_ctx_create(gpus, 0, &probe) == B200_OK) { b200_ctx_destroy(probe); ++gpus; }
    if (gpus < 2) { std::printf("SKIP  %d usable B200 device(s); the sharded run needs 2\n", gpus); return 0; }
    const int world = gpus >= 4 ? 4 : 2;

    struct Case { const char* name; B200ForceMethod method; size_t n; int steps; bool bitwise; };
    const Case cases[] = {
        {"tree   20001 particles (ragged shards)", B200ForceMethod::Tree, 20001, 5, true},
        {"tree   65536 particles", B200ForceMethod::Tree, 65536, 3, true},
        {"direct 16384 particles, open boundary", B200ForceMethod::DirectOpen, 16384, 5, false},
        {"direct 12345 particles, periodic (ragged)", B200ForceMethod::Direct, 12345, 5, false},
        // large enough for the 6-targets-per-thread instance (> 48 KB of shared memory: the per-device
        // opt-in must have happened on every GPU, not only on the first one a process touched)
        {"direct 262144 particles, open boundary", B200ForceMethod::DirectOpen, 262144, 1, false},
        {"tree fixed-physics 50000 particles", B200ForceMethod::TreeFixed, 50000, 3, true},
        {"tree fixed-physics periodic 30001 particles", B200ForceMethod::TreeFixedPeriodic, 30001, 3, true},
    };
    for (const Case& c : cases) {
        Result single;
        std::string err;
        run_rank(c.n, c.method, c.steps, 1e-4, nullptr, 0, 1, &single, &err);
        if (!err.empty()) { CHECK(false, "%s: single-GPU run failed: %s", c.name, err.c_str()); continue; }

        unsigned char id[B200_SHARD_ID_BYTES];
        const int rc = b200_shard_unique_id(id);
        if (rc != B200_OK) { CHECK(false, "%s: b200_shard_unique_id -> %s", c.name, b200_error_string(rc)); continue; }
        std::vector<Result> res(world);
        std::vector<std::string> errs(world);
        std::vector<std::thread> th;
        for (int r = 0; r < world; ++r)
            th.emplace_back(run_rank, c.n, c.method, c.steps, 1e-4, id, r, world, &res[r], &errs[r]);
        for (auto& t : th) t.join();
        bool ok = true;
        for (int r = 0; r < world; ++r)
            if (!errs[r].empty()) { CHECK(false, "%s: rank %d failed: %s", c.name, r, errs[r].c_str()); ok = false; }
        if (!ok) continue;

        // every rank must hold the same, complete position array
        bool replicas_equal = true;
        for (int r = 1; r < world; ++r)
            replicas_equal &= std::memcmp(res[0].pos.data(), res[r].pos.data(), 3 * c.n * sizeof(float)) == 0;
        CHECK(replicas_equal, "%s: all %d ranks hold identical positions after %d steps", c.name, world, c.steps);
        // assemble the sharded velocities
        std::vector<float> vel(3 * c.n);
        for (int r = 0; r < world; ++r) {
            size_t i0 = 0, nl = 0;
            b200_shard_range(c.n, r, world, &i0, &nl);
            std::memcpy(vel.data() + 3 * i0, res[r].vel.data(), 3 * nl * sizeof(float));
        }
        const double dp = max_abs_diff(res[0].pos.data(), single.pos.data(), 3 * c.n);
        const double dv = max_abs_diff(vel.data(), single.vel.data(), 3 * c.n);
        if (c.bitwise)
            CHECK(dp == 0.0 && dv == 0.0 && res[0].a == single.a,
                  "%s: %d-GPU run == 1-GPU run bitwise (max |dx| %.3g, |dv| %.3g)", c.name, world, dp, dv);
        else
            CHECK(dp <= 1e-4 * 100.0 * 1e-2 && dv <= 1e-3 && res[0].a == single.a,
                  "%s: %d-GPU run vs 1-GPU run: max |dx| %.3g, |dv| %.3g", c.name, world, dp, dv);
    }
    std::printf("%s (%d failure(s))\n", failures ? "FAILED" : "ALL PASS", failures);
    return failures ? 1 : 0;
}
