// engine_wiring_test.cpp -- SURVEY 8f N1 made executable: the reference's own SimulationBuilder / SimulationEngine,
// patched with integration/engine_wiring.patch (applied to a temporary copy of the reference sources by
// tests/host/Makefile), drives the B200 force path end to end:
//
//     SimulationBuilder().with_num_particles(N).with_force_computer("DirectForceComputer")...enable_gpu(0).build()->run()
//
// for 10 KDK steps on seeded particles, device-resident and through host arrays, and the final positions are held
// to the same loop run on the reference's CPU code (TreeForceComputer: its leaf pair loop with leaf_capacity > N is
// the reference's direct sum; theta 0.5 / leaf 8 its Barnes-Hut) -- positions within 1e-4 of the box.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <memory>
#include <random>
#include <sstream>
#include <vector>

#include "core/simulation_engine.hpp"
#include "forces/force_computer_factory.hpp"
#include "forces/tree_force_computer.hpp"
#include "physics/cosmology_model.hpp"

#include "b200_force_computers.hpp"

using namespace core;
using namespace forces;

static int failures = 0;
#define CHECK(cond, ...)                                                         \
    do {                                                                         \
        std::printf("%s  ", (cond) ? "PASS" : "FAIL");                           \
        std::printf(__VA_ARGS__);                                                \
        std::printf("\n");                                                       \
        if (!(cond)) ++failures;                                                 \
    } while (0)

// The particles the patched engine draws for ("ic_seed", seed): see initialize_simulation_data in the patch.
static void seeded_particles(size_t n, int seed, float box, float sigma, std::vector<float>& x, std::vector<float>& v) {
    std::mt19937 gen((unsigned)seed);
    std::uniform_real_distribution<float> pd(-0.5f * box, 0.5f * box);
    x.resize(3 * n); v.resize(3 * n);
    for (auto& q : x) q = pd(gen);
    std::normal_distribution<float> vd(0.0f, sigma);
    for (auto& q : v) q = vd(gen);
}

// KDK of lambda_cdm_impl.cu:167-213 on the host with a CPU force computer (unit masses)
static void cpu_kdk(IForceComputer& fc, std::vector<float>& x, std::vector<float>& v, size_t n, int steps, double dt,
                    double& a) {
    std::vector<float> m(n, 1.0f), f(3 * n);
    physics::CosmologyModel cosmo{physics::CosmologyParams()};
    fc.compute_forces(x.data(), m.data(), f.data(), n);
    const float hdt = (float)(dt * 0.5), fdt = (float)dt;
    for (int s = 0; s < steps; ++s) {
        float a2 = (float)(1.0f / (a * a));
        for (size_t i = 0; i < 3 * n; ++i) { v[i] += f[i] * 1.0f * hdt * a2; x[i] += v[i] * fdt; }
        a += a * cosmo.hubble_parameter_a(a) * dt;
        fc.compute_forces(x.data(), m.data(), f.data(), n);
        a2 = (float)(1.0f / (a * a));
        for (size_t i = 0; i < 3 * n; ++i) v[i] += f[i] * 1.0f * hdt * a2;
    }
}

static bool run_case(const char* fc_type, size_t n, bool device_resident, IForceComputer& cpu_ref, double vel_tol) {
    const int steps = 10, seed = 4242;
    const double dt = 1e-3;
    auto sim = SimulationBuilder()
                   .with_num_particles(n)
                   .with_box_size(100.0f)
                   .with_time_step(dt)
                   .with_max_time(1.0e9)
                   .with_max_steps(steps)
                   .with_random_particles(seed, 100.0f)
                   .with_force_computer(fc_type)
                   .with_integrator("LeapfrogIntegrator")
                   .with_cosmology_model("LambdaCDMModel")
                   .with_device_resident_state(device_resident)
                   .enable_gpu(0)
                   .build();
    if (!sim) { CHECK(false, "%s: SimulationBuilder::build() returned null", fc_type); return false; }
    const bool ok = sim->run();
    std::vector<float> xg(sim->get_positions(), sim->get_positions() + 3 * n);
    std::vector<float> vg(sim->get_velocities(), sim->get_velocities() + 3 * n);
    std::vector<float> xc, vc;
    seeded_particles(n, seed, 100.0f, 100.0f, xc, vc);
    double ac = 1.0;
    cpu_kdk(cpu_ref, xc, vc, n, steps, dt, ac);
    double mx = 0, num = 0, den = 0;
    for (size_t i = 0; i < 3 * n; ++i) {
        mx = std::fmax(mx, std::fabs((double)xg[i] - xc[i]));
        num += ((double)vg[i] - vc[i]) * ((double)vg[i] - vc[i]);
        den += (double)vc[i] * vc[i];
    }
    const double vrel = std::sqrt(num / den);
    const bool pass = ok && sim->get_statistics().current_step == (size_t)steps && sim->get_scale_factor() == ac &&
                      mx < 1e-4 * 100.0 && vrel < vel_tol && std::isfinite(mx);
    CHECK(pass, "SimulationEngine::run() with %s, %zu particles, %s: %d steps, a = %.6f (cpu %.6f), max |dx| = %.2e "
                "(gate 1e-2), velocities rel-L2 %.1e",
          fc_type, n, device_resident ? "device-resident" : "host arrays", (int)sim->get_statistics().current_step,
          sim->get_scale_factor(), ac, mx, vrel);
    return pass;
}

int main() {
    std::ostringstream quiet;
    std::streambuf* old = std::cout.rdbuf(quiet.rdbuf());        // the engine and the computers are chatty
    ForceComputerFactory::register_all_builtin_computers();      // CPU "TreeForceComputer"
    const size_t n_direct = 4096, n_tree = 8192;
    SimulationContext cctx;
    TreeForceComputer cpu_direct("cpu_direct", 0.5f, n_direct + 1, 20);     // one root leaf: the reference's CPU direct sum
    TreeForceComputer cpu_tree("cpu_tree", 0.5f, 8, 20);
    cpu_direct.initialize(cctx);
    cpu_tree.initialize(cctx);
    cpu_direct.set_box_size(100.0f);
    cpu_tree.set_box_size(100.0f);
    register_b200_force_computers(/*replace_cpu_tree=*/true);   // "DirectForceComputer", "TreeForceComputer" -> B200
    std::cout.rdbuf(old);
    run_case("DirectForceComputer", n_direct, true, cpu_direct, 1e-4);
    run_case("DirectForceComputer", n_direct, false, cpu_direct, 1e-4);
    run_case("TreeForceComputer", n_tree, true, cpu_tree, 1e-4);
    run_case("TreeForceComputer", n_tree, false, cpu_tree, 1e-4);
    std::printf("%s (%d failure%s)\n", failures ? "ENGINE WIRING FAILED" : "ENGINE WIRING OK", failures, failures == 1 ? "" : "s");
    return failures ? 1 : 0;
}
