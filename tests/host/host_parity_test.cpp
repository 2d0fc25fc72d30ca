// host_parity_test.cpp -- the plugin seen from the reference's side, written the
// way the reference's own (never compiled) test reads (CONTRIBUTING.md:106-131):
// build computers through forces::ForceComputerFactory, call
// IForceComputer::compute_forces on host arrays, compare with the reference's
// own CPU TreeForceComputer on identical inputs.
//
// TEST INFRASTRUCTURE: links the reference's CPU sources (compiled from where
// they lie by tests/host/Makefile) as the checker.  Run on the GPU box by
// tests/test_gpu_host_plugin.py.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <random>
#include <sstream>
#include <vector>

#include "b200_force_computers.hpp"
#include "b200_simulation.hpp"
#include "core/simulation_context.hpp"
#include "forces/tree_force_computer.hpp"

static int failures = 0;
#define CHECK(cond, ...)                                               \
    do {                                                               \
        if (cond) { std::printf("PASS  "); } else { std::printf("FAIL  "); ++failures; } \
        std::printf(__VA_ARGS__); std::printf("\n");                   \
    } while (0)

static double rel_l2(const std::vector<float>& a, const std::vector<float>& b) {
    double num = 0, den = 0;
    for (size_t i = 0; i < a.size(); ++i) { double d = (double)a[i] - b[i]; num += d * d; den += (double)b[i] * b[i]; }
    return std::sqrt(num / (den > 0 ? den : 1));
}

int main() {
    using namespace forces;
    std::ostringstream quiet;
    auto* old = std::cout.rdbuf(quiet.rdbuf());            // the reference is chatty on stdout
    core::SimulationContext ctx;
    ForceComputerFactory::register_all_builtin_computers();                       // basic_simulation.cpp:12-13
    auto cpu_tree = ForceComputerFactory::create_force_computer("TreeForceComputer", "cpu_tree");
    cpu_tree->initialize(ctx);                                                     // downgrades itself to CPU
    TreeForceComputer cpu_direct("cpu_direct", 0.5f, /*leaf_capacity=*/size_t(1) << 40, 20);   // one root leaf
    cpu_direct.initialize(ctx);

    register_b200_force_computers();                                               // the plugin
    auto direct = ForceComputerFactory::create_direct_computer("direct");
    auto tree = ForceComputerFactory::create_tree_computer("tree");
    std::cout.rdbuf(old);
    CHECK(direct && direct->get_type() == "DirectForceComputer", "factory creates DirectForceComputer");
    CHECK(tree && tree->get_type() == "TreeForceComputer" && tree->supports_gpu(), "factory creates the B200 TreeForceComputer");
    std::cout.rdbuf(quiet.rdbuf());
    const bool ok_d = direct->initialize(ctx), ok_t = tree->initialize(ctx);
    std::cout.rdbuf(old);
    CHECK(ok_d && ok_t, "initialize() on a B200 (no CPU fallback exists)");
    if (!ok_d || !ok_t) return 2;

    {   // CONTRIBUTING.md:118-131
        std::vector<float> pos = {0, 0, 0, 1, 0, 0}, m = {1, 1}, f(6), g(6);
        direct->compute_forces(pos.data(), m.data(), f.data(), 2);
        cpu_direct.compute_forces(pos.data(), m.data(), g.data(), 2);
        CHECK(std::fabs(f[0] - 0.999850035f) < 2e-6f && std::fabs(f[3] + 0.999850035f) < 2e-6f && std::fabs(f[0] - g[0]) < 2e-6f,
              "two-body KAT: a0.x = %.9f (reference CPU %.9f)", f[0], g[0]);
        direct->compute_forces(pos.data(), m.data(), f.data(), 0);               // num_particles == 0 -> return
    }

    const size_t n = 20000;
    std::mt19937 gen(42);
    std::uniform_real_distribution<float> u(-50.0f, 50.0f);
    std::vector<float> pos(3 * n), mass(n, 1.0f), a_gpu(3 * n), a_cpu(3 * n);
    for (size_t i = 0; i < 3 * n; ++i) pos[i] = u(gen);

    direct->compute_forces(pos.data(), mass.data(), a_gpu.data(), n);
    cpu_direct.compute_forces(pos.data(), mass.data(), a_cpu.data(), n);
    double e = rel_l2(a_gpu, a_cpu);
    CHECK(e < 1e-5, "direct sum vs reference CPU leaf loop, N=%zu: rel-L2 %.2e (gate 1e-5)", n, e);

    tree->compute_forces(pos.data(), mass.data(), a_gpu.data(), n);
    cpu_tree->compute_forces(pos.data(), mass.data(), a_cpu.data(), n);
    e = rel_l2(a_gpu, a_cpu);
    CHECK(e < 1e-3, "Barnes-Hut theta=0.5 leaf=8 vs reference CPU tree: rel-L2 %.2e (gate 1e-3)", e);
    auto* bt = dynamic_cast<B200TreeForceComputer*>(tree.get());
    auto* ct = dynamic_cast<TreeForceComputer*>(cpu_tree.get());
    CHECK(bt && ct && bt->get_node_count() == ct->get_node_count() && bt->get_leaf_count() == ct->get_leaf_count() &&
              bt->get_tree_depth() == ct->get_tree_depth(),
          "tree statistics equal: %zu nodes, %zu leaves, depth %zu", bt->get_node_count(), bt->get_leaf_count(), bt->get_tree_depth());

    {   // the two phases the reference class exposes, and its statistics getters (tree_force_computer.hpp:78-103)
        std::vector<float> f2(3 * n), f1(3 * n);
        tree->compute_forces(pos.data(), mass.data(), f1.data(), n);
        bt->reset_statistics();
        bt->build_tree(pos.data(), mass.data(), n);
        bt->compute_tree_forces(pos.data(), f2.data(), n);
        bool same = true;
        for (size_t i = 0; i < 3 * n && same; ++i) same = f1[i] == f2[i];
        unsigned long long cnt[3];
        bt->count_interactions(n, cnt);
        tree->compute_forces(pos.data(), mass.data(), f1.data(), n);
        CHECK(same && bt->get_tree_traversals() == n && bt->get_force_evaluations() == n && cnt[0] > n && cnt[1] > 0 &&
                  cnt[2] > 0 && bt->get_tree_efficiency() > 0.0f,
              "build_tree + compute_tree_forces == compute_forces bitwise; %.0f node visits, %.0f cell and %.0f pair "
              "interactions per particle", (double)cnt[0] / n, (double)cnt[1] / n, (double)cnt[2] / n);
    }

    {   // ForceComputeParameters through std::any
        ForceComputeParameters p;
        p.softening_length = 0.5f;
        std::vector<float> b(3 * n);
        direct->compute_forces(pos.data(), mass.data(), b.data(), n, p);
        CHECK(rel_l2(b, a_cpu) > 1e-3, "softening_length from ForceComputeParameters is honoured");
    }

    {   // fixed-physics mode (not in the reference): against an FP64 direct sum, where the reference's
        // own tree semantics are ~0.3 away (orphans, unit-mass leaves)
        const size_t m_n = 6000;
        std::vector<float> f_fixed(3 * m_n), f_faithful(3 * m_n);
        std::vector<double> want(3 * m_n, 0.0);
        for (size_t i = 0; i < m_n; ++i)
            for (size_t j = 0; j < m_n; ++j) {
                const double dx = (double)pos[3 * j] - pos[3 * i], dy = (double)pos[3 * j + 1] - pos[3 * i + 1],
                             dz = (double)pos[3 * j + 2] - pos[3 * i + 2];
                const double r2 = dx * dx + dy * dy + dz * dz + 1e-4;
                const double f = mass[j] / (r2 * std::sqrt(r2));
                want[3 * i] += f * dx; want[3 * i + 1] += f * dy; want[3 * i + 2] += f * dz;
            }
        std::vector<float> wantf(want.begin(), want.end());
        bt->set_fixed_physics(true);
        bt->set_softening(0.01f);
        tree->compute_forces(pos.data(), mass.data(), f_fixed.data(), m_n);
        const size_t stored_nodes = bt->get_node_count();
        bt->set_fixed_physics(false);
        tree->compute_forces(pos.data(), mass.data(), f_faithful.data(), m_n);
        CHECK(rel_l2(f_fixed, wantf) < 4e-3 && rel_l2(f_faithful, wantf) > 0.1 && stored_nodes > 0,
              "fixed-physics tree vs FP64 direct sum: rel-L2 %.2e (the reference's tree semantics: %.2e)",
              rel_l2(f_fixed, wantf), rel_l2(f_faithful, wantf));
    }

    {   // KDK: IIntegrator + IForceComputer + ICosmologyModel, 5 steps, against the same loop on the CPU tree
        physics::B200LeapfrogIntegrator integ("leapfrog");
        physics::LambdaCDMModel cosmo("lcdm");
        integ.initialize(ctx);
        const size_t m_n = 4096;
        std::vector<float> x(pos.begin(), pos.begin() + 3 * m_n), v(3 * m_n, 0.0f), f(3 * m_n), ms(m_n, 1.0f);
        std::vector<float> xc = x, vc = v, fc(3 * m_n);
        double a = 1.0, ac = 1.0;
        const double dt = 1e-3;
        tree->compute_forces(x.data(), ms.data(), f.data(), m_n);
        cpu_tree->compute_forces(xc.data(), ms.data(), fc.data(), m_n);
        for (int s = 0; s < 5; ++s) {
            physics::LeapfrogStepParams p; p.scale_factor = a; p.n_kicks = 1; p.drift = true;
            integ.step(x.data(), v.data(), f.data(), m_n, dt, p);
            cosmo.update_scale_factor(a, dt);
            tree->compute_forces(x.data(), ms.data(), f.data(), m_n);
            p.scale_factor = a; p.drift = false;
            integ.step(x.data(), v.data(), f.data(), m_n, dt, p);
            // CPU restatement of lambda_cdm_kernels.cu:307-333 on the reference tree's forces
            const float hdt = (float)(dt * 0.5), fdt = (float)dt;
            float a2 = (float)(1.0f / (ac * ac));
            for (size_t i = 0; i < 3 * m_n; ++i) { vc[i] += fc[i] * 1.0f * hdt * a2; xc[i] += vc[i] * fdt; }
            ac += ac * cosmo.hubble_function(ac) * dt;
            cpu_tree->compute_forces(xc.data(), ms.data(), fc.data(), m_n);
            a2 = (float)(1.0f / (ac * ac));
            for (size_t i = 0; i < 3 * m_n; ++i) vc[i] += fc[i] * 1.0f * hdt * a2;
        }
        double mx = 0;
        for (size_t i = 0; i < 3 * m_n; ++i) mx = std::fmax(mx, std::fabs((double)x[i] - xc[i]));
        CHECK(a == ac && mx < 1e-4 * 100.0, "5 KDK steps through IIntegrator/ICosmologyModel: a = %.6f, max |dx| = %.2e (gate 1e-2)", a, mx);
    }

    {   // device-resident driver with the LambdaCDMSimulation call shape (examples/cuda_nbody_test.cpp:31-60):
        // 10 steps of dt = 1e-3 on 4096 particles, tree forces, against the same loop on the reference CPU tree
        const size_t m_n = 4096;
        physics::B200LambdaCDMSimulation sim(m_n, 100.0f);
        std::vector<float> x(pos.begin(), pos.begin() + 3 * m_n), v(3 * m_n), ms(m_n, 1.0f), fc(3 * m_n);
        std::mt19937 g2(7);
        std::normal_distribution<float> nv(0.0f, 55.0f);
        for (auto& q : v) q = nv(g2);
        sim.set_force_method(physics::B200ForceMethod::Tree);
        sim.set_particles(x.data(), v.data(), ms.data());
        std::vector<float> xc = x, vc = v;
        physics::LambdaCDMModel cosmo("lcdm");
        double ac = 1.0;
        const double dt = 1e-3;
        cpu_tree->compute_forces(xc.data(), ms.data(), fc.data(), m_n);
        for (int s = 0; s < 10; ++s) {
            sim.step(dt);
            const float hdt = (float)(dt * 0.5), fdt = (float)dt;
            float a2 = (float)(1.0f / (ac * ac));
            for (size_t i = 0; i < 3 * m_n; ++i) { vc[i] += fc[i] * 1.0f * hdt * a2; xc[i] += vc[i] * fdt; }
            cosmo.update_scale_factor(ac, dt);
            cpu_tree->compute_forces(xc.data(), ms.data(), fc.data(), m_n);
            a2 = (float)(1.0f / (ac * ac));
            for (size_t i = 0; i < 3 * m_n; ++i) vc[i] += fc[i] * 1.0f * hdt * a2;
        }
        std::vector<float> xg(3 * m_n), vg(3 * m_n);
        sim.copy_positions_to_host(xg.data());
        sim.copy_velocities_to_host(vg.data());
        double mx = 0;
        for (size_t i = 0; i < 3 * m_n; ++i) mx = std::fmax(mx, std::fabs((double)xg[i] - xc[i]));
        CHECK(sim.get_scale_factor() == ac && sim.get_current_step() == 10 && mx < 1e-4 * 100.0 && rel_l2(vg, vc) < 1e-4,
              "B200LambdaCDMSimulation: 10 device-resident KDK steps, z = %.4f, max |dx| = %.2e, vel rel-L2 %.1e",
              sim.get_redshift(), mx, rel_l2(vg, vc));
        physics::B200LambdaCDMSimulation per(2048, 100.0f);      // periodic direct path, seeded ICs
        per.initialize_particles(12345);
        per.set_softening(0.1f);
        for (int s = 0; s < 3; ++s) per.step(dt);
        std::vector<float> xp(3 * 2048);
        per.copy_positions_to_host(xp.data());
        bool inside = true;
        for (float q : xp) inside = inside && q >= 0.0f && q < 100.0f && std::isfinite(q);
        CHECK(inside && per.get_current_step() == 3, "periodic direct driver keeps particles in [0, box)");

        // Tree + initialize_particles: the generated particles must sit inside the origin-centred root cube, and
        // the forces must be the CPU tree's on those positions; a later switch to a method that works in the
        // other frame is refused
        physics::B200LambdaCDMSimulation tr(4096, 100.0f);
        tr.set_force_method(physics::B200ForceMethod::Tree);
        tr.initialize_particles(777);
        std::vector<float> xt(3 * 4096), ft(3 * 4096), fr(3 * 4096), mt(4096, 1.0f);
        tr.copy_positions_to_host(xt.data());
        bool centred = true;
        for (float q : xt) centred = centred && q >= -50.0f && q < 50.0f;
        tr.compute_forces();
        tr.copy_forces_to_host(ft.data());
        cpu_tree->compute_forces(xt.data(), mt.data(), fr.data(), 4096);
        bool refused = false;
        try { tr.set_force_method(physics::B200ForceMethod::Direct); } catch (const std::logic_error&) { refused = true; }
        CHECK(centred && rel_l2(ft, fr) < 1e-3 && refused,
              "Tree + initialize_particles: particles inside the root cube, forces vs CPU tree rel-L2 %.1e, frame switch %s",
              rel_l2(ft, fr), refused ? "refused" : "ACCEPTED");
    }

    std::cout.rdbuf(quiet.rdbuf());
    {   // particles stored in space-filling-curve order: same physics, indices mapped back
        const size_t m_n = 20000;
        std::vector<float> v0(3 * m_n, 0.0f);
        physics::B200LambdaCDMSimulation a(m_n, 100.0f), b(m_n, 100.0f);
        a.set_force_method(physics::B200ForceMethod::TreeFixed);
        b.set_force_method(physics::B200ForceMethod::TreeFixed);
        a.set_particles(pos.data(), v0.data(), mass.data());
        b.set_particles_spatially_ordered(pos.data(), v0.data(), mass.data());
        for (int s = 0; s < 3; ++s) { a.step(1e-4); b.step(1e-4); }
        std::vector<float> xa(3 * m_n), xb(3 * m_n), xb_back(3 * m_n);
        a.copy_positions_to_host(xa.data());
        b.copy_positions_to_host(xb.data());
        const std::vector<int>& ord = b.get_particle_order();
        std::vector<char> seen(m_n, 0);
        bool perm_ok = ord.size() == m_n;
        double jump = 0.0;
        for (size_t k = 0; k < m_n && perm_ok; ++k) {
            perm_ok = ord[k] >= 0 && (size_t)ord[k] < m_n && !seen[ord[k]];
            if (perm_ok) seen[ord[k]] = 1;
            for (int c = 0; c < 3; ++c) xb_back[3 * (size_t)ord[k] + c] = xb[3 * k + c];
            if (k) jump += std::fabs(xb[3 * k] - xb[3 * k - 3]) + std::fabs(xb[3 * k + 1] - xb[3 * k - 2]) +
                           std::fabs(xb[3 * k + 2] - xb[3 * k - 1]);
        }
        double dmax = 0.0;
        for (size_t i = 0; i < 3 * m_n; ++i) dmax = std::fmax(dmax, std::fabs((double)xa[i] - xb_back[i]));
        CHECK(perm_ok && dmax < 1e-4 && jump / m_n < 12.0,
              "spatially ordered storage: same trajectories (max |dx| %.2e), neighbours in storage %.1f apart (L1)",
              dmax, jump / m_n);

        // the reference-faithful tree depends on insertion order: stored along the curve, it must still be the tree of
        // the caller's index order (arrival array) -- same node count, same forces particle for particle
        physics::B200LambdaCDMSimulation c(m_n, 100.0f), d(m_n, 100.0f);
        c.set_force_method(physics::B200ForceMethod::Tree);
        d.set_force_method(physics::B200ForceMethod::Tree);
        c.set_particles(pos.data(), v0.data(), mass.data());
        d.set_particles_spatially_ordered(pos.data(), v0.data(), mass.data());
        c.compute_forces();
        d.compute_forces();
        std::vector<float> fc(3 * m_n), fd(3 * m_n), fd_back(3 * m_n);
        c.copy_forces_to_host(fc.data());
        d.copy_forces_to_host(fd.data());
        const std::vector<int>& ord2 = d.get_particle_order();
        for (size_t k = 0; k < m_n; ++k)
            for (int q = 0; q < 3; ++q) fd_back[3 * (size_t)ord2[k] + q] = fd[3 * k + q];
        CHECK(rel_l2(fd_back, fc) < 1e-6, "faithful tree on spatially ordered storage == index-order storage (rel-L2 %.1e)",
              rel_l2(fd_back, fc));
    }

    {   // device-generated initial conditions + the particle / power-spectrum accessors of the simulation class
        const size_t m_n = 32768;                                   // 32^3: every grid point
        physics::B200LambdaCDMSimulation sim(m_n, 100.0f);
        sim.set_force_method(physics::B200ForceMethod::TreeFixed);
        sim.set_initial_conditions_from_power_spectrum(12345, 49.0);
        std::vector<physics::Particle> parts;
        sim.copy_particles_to_host(parts);
        std::vector<float> x(3 * m_n);
        sim.copy_positions_to_host(x.data());
        bool same = parts.size() == m_n, inside = true;
        double disp2 = 0.0;
        for (size_t i = 0; i < m_n && same; ++i) {
            same = parts[i].position.x == x[3 * i] && parts[i].position.z == x[3 * i + 2] && parts[i].mass == 1.0f &&
                   parts[i].id == i;
            inside = inside && x[3 * i] >= -50.0f && x[3 * i] < 50.0f;
            const size_t ix = i / (32 * 32);
            const double q = (ix + 0.5) * (100.0 / 32) - 50.0;     // cell centre, origin-centred
            double d = x[3 * i] - q;
            d -= 100.0 * std::round(d / 100.0);
            disp2 += d * d;
        }
        const double rms_x = std::sqrt(disp2 / m_n);
        std::vector<float> kk, pk;
        std::vector<int> modes;
        sim.power_spectrum(32, kk, pk, modes, false);
        CHECK(same && inside && rms_x > 0.02 && rms_x < 0.4 && std::fabs(sim.get_redshift() - 49.0) < 1e-9 &&
                  pk.size() == 16 && modes[1] == 26 && pk[1] > 0.0f,
              "set_initial_conditions_from_power_spectrum: %zu particles, rms x-displacement %.3f Mpc/h, P(k_1) %.3g",
              parts.size(), rms_x, pk.size() > 1 ? pk[1] : 0.0f);
        sim.step(1e-4);
        sim.compute_energy();
        CHECK(std::isfinite(sim.get_total_energy()) && sim.get_potential_energy() < 0.0 && sim.get_kinetic_energy() > 0.0,
              "step + compute_energy on the generated particles: KE %.4g PE %.4g", sim.get_kinetic_energy(),
              sim.get_potential_energy());
    }

    direct->finalize(); direct->finalize();                                        // idempotent
    tree->finalize(); cpu_tree->finalize();
    std::cout.rdbuf(old);
    std::printf("%s (%d failure%s)\n", failures ? "HOST PARITY FAILED" : "HOST PARITY OK", failures, failures == 1 ? "" : "s");
    return failures ? 1 : 0;
}
