// nbody_b200.cpp -- the reference's smoke program for its CUDA simulation class
// (examples/cuda_nbody_test.cpp: N particles, K leapfrog steps, energy every 10 steps, ms per step and
// particle-updates per second at the end) run on the B200 engine through B200LambdaCDMSimulation.
//
//   nbody_b200 [particles = 10000] [steps = 100] [method = direct|direct-open|tree|tree-fixed|tree-periodic]
//              [ics = random|zeldovich]
//
// Differences from the reference program, on purpose: the particles are seeded (the reference seeds from
// the clock), the first half-kick uses forces at the initial positions (the reference reads uninitialised
// memory), and the force method is selectable.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>

#include "b200_simulation.hpp"

int main(int argc, char** argv) {
    using namespace physics;
    try {
        const size_t n = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 10000;
        const int steps = argc > 2 ? std::atoi(argv[2]) : 100;
        const std::string method = argc > 3 ? argv[3] : "direct";
        const std::string ics = argc > 4 ? argv[4] : "random";
        B200ForceMethod m;
        if (method == "direct") m = B200ForceMethod::Direct;
        else if (method == "direct-open") m = B200ForceMethod::DirectOpen;
        else if (method == "tree") m = B200ForceMethod::Tree;
        else if (method == "tree-fixed") m = B200ForceMethod::TreeFixed;
        else if (method == "tree-periodic") m = B200ForceMethod::TreeFixedPeriodic;
        else { std::fprintf(stderr, "unknown method %s\n", method.c_str()); return 2; }

        std::printf("\n=== Lambda-CDM N-body on B200 ===\nparticles %zu, steps %d, forces %s, initial conditions %s\n",
                    n, steps, method.c_str(), ics.c_str());
        CosmologyParams params;                                  // cuda_nbody_test.cpp:26-30
        params.omega_m = 0.31; params.omega_lambda = 0.69; params.h = 0.67;
        const float box = 100.0f;
        B200LambdaCDMSimulation sim(n, box, params);
        sim.set_softening(box / 1000.0f);                        // :37
        sim.set_force_method(m);
        if (ics == "zeldovich") sim.set_initial_conditions_from_power_spectrum(12345, 49.0);
        else sim.initialize_particles(12345);

        sim.compute_energy();                                    // :44-49
        const double e0 = sim.get_total_energy();
        std::printf("initial energy %.6e (kinetic %.6e, potential %.6e), z = %.3f\n", e0, sim.get_kinetic_energy(),
                    sim.get_potential_energy(), sim.get_redshift());

        const double dt = 0.001;                                 // :53
        const auto t0 = std::chrono::high_resolution_clock::now();
        double energy_seconds = 0.0;
        for (int i = 0; i < steps; ++i) {
            sim.step(dt);
            if ((i + 1) % 10 == 0) {                             // :60-71 (the energy pass is not part of the step time)
                const auto a = std::chrono::high_resolution_clock::now();
                sim.compute_energy();
                energy_seconds += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - a).count();
                std::printf("step %4d | z = %8.3f | energy error %.3e\n", i + 1, sim.get_redshift(),
                            std::fabs((sim.get_total_energy() - e0) / e0));
            }
        }
        const double seconds =
            std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count() - energy_seconds;
        sim.compute_energy();
        std::printf("\n=== done ===\ntime per step %.4f ms\nfinal redshift %.4f\nfinal energy %.6e (kinetic %.6e, potential %.6e)\n",
                    1e3 * seconds / steps, sim.get_redshift(), sim.get_total_energy(), sim.get_kinetic_energy(),
                    sim.get_potential_energy());
        std::printf("performance: %.4e particle-updates/second\n", (double)n * steps / seconds);   // :91-93
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
