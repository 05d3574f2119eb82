/*
 * sim_harness.cpp — BASELINE config C1 as the reference runs it: the reference's own ECSSimulator (src/sim.cpp,
 * unmodified) ticking its own KeplerianDiskScenario headless, all eight systems in their hard-coded order
 * (sim.cpp:107-114). Built twice by oracle/Makefile from the same sources:
 *   oracle/_ref/sim_ref      every system is the reference's
 *   oracle/_ref/sim_dropin   include/systems/{barnes_hut,boundary}.hpp and their .cpp are replaced by the drop-in classes
 *                            of little-physics-engine_b200/host/systems (-> liblpe_bh.so -> GPU); src/sim.cpp,
 *                            i_scenario.hpp and everything else compile unchanged, which is the drop-in claim.
 * TEST INFRASTRUCTURE ONLY (run by tests/test_sim_tick_gpu.py). The scenario seeds its RNG from time(),
 * so after reset() the bodies' Position / Velocity / Mass are overwritten, in creation order, with the deterministic
 * Keplerian disk of csrc/workloads.cpp: both binaries then tick the same registry.
 * usage: sim_* <n_bodies> <ticks> <out.bin>   -> JSON line {ticks, ms_per_tick, ...}; out.bin = n x (x, y, vx, vy) doubles
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <memory>
#include <string>
#include <vector>

#include <entt/entt.hpp>

#define private public   // KeplerianDiskConfig::particleCount has no setter; the reference's node pool must be pre-sized (D1)
#include "scenarios/keplerian_disk.hpp"
#include "systems/barnes_hut.hpp"
#include "sim.hpp"
#undef private
#include "entities/entity_components.hpp"

typedef int (*workload_fn)(int, uint64_t, uint64_t, double, double*, double*, double*, double*, double*);

int main(int argc, char** argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 10000;
    const int ticks = argc > 2 ? std::atoi(argv[2]) : 100;
    const char* out = argc > 3 ? argv[3] : nullptr;

    // the deterministic generator, next to the product library (host code only)
    std::string self = argv[0];
    const size_t slash = self.rfind('/');
    const std::string dir = slash == std::string::npos ? "." : self.substr(0, slash);
    void* h = dlopen((dir + "/../../little-physics-engine_b200/libworkloads.so").c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) { std::printf("{\"error\": \"%s\"}\n", dlerror()); return 2; }
    auto workload = reinterpret_cast<workload_fn>(dlsym(h, "lpe_bh_workload"));
    if (!workload) return 2;

    auto scenario = std::make_unique<KeplerianDiskScenario>();
    scenario->scenarioEntityConfig.particleCount = n;
    const ScenarioSystemConfig cfg = scenario->getSystemsConfig();
    ECSSimulator& sim = ECSSimulator::getInstance();
    sim.applyConfig(cfg);
    sim.loadScenario(std::move(scenario));
    sim.reset();
    entt::registry& reg = sim.getRegistry();

    auto& ms = reg.storage<Components::Mass>();
    const size_t nb = ms.size();
    std::vector<double> x(nb), y(nb), vx(nb), vy(nb), m(nb);
    if (workload(3, nb, 11, cfg.sharedConfig.UniverseSizeMeters, x.data(), y.data(), vx.data(), vy.data(), m.data())) return 2;
    std::vector<entt::entity> ents(ms.data(), ms.data() + nb);   // packed order = creation order
    for (size_t i = 0; i < nb; ++i) {
        auto& p = reg.get<Components::Position>(ents[i]);
        auto& v = reg.get<Components::Velocity>(ents[i]);
        p.x = x[i]; p.y = y[i];
        v.x = vx[i]; v.y = vy[i];
        reg.get<Components::Mass>(ents[i]).value = m[i];
    }
#ifndef LPE_DROPIN
    for (auto& s : sim.systems)   // reference defect D1: the node pool must never grow while a build holds pointers into it
        if (auto* bh = dynamic_cast<Systems::BarnesHutSystem*>(s.get())) bh->nodePool_.resize(12 * nb + 4096);
#endif
    sim.tick();   // first tick apart: device context / buffers on one side, page faults of the pool on the other
    const auto t0 = std::chrono::steady_clock::now();
    for (int t = 1; t < ticks; ++t) sim.tick();
    const double ms_tick = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / (ticks > 1 ? ticks - 1 : 1);
    if (out) {
        FILE* f = std::fopen(out, "wb");
        if (!f) return 2;
        for (size_t i = 0; i < nb; ++i) {
            const auto& p = reg.get<Components::Position>(ents[i]);
            const auto& v = reg.get<Components::Velocity>(ents[i]);
            const double rec[4] = {p.x, p.y, v.x, v.y};
            std::fwrite(rec, sizeof(double), 4, f);
        }
        std::fclose(f);
    }
#ifdef LPE_DROPIN
    const char* which = "drop-in BarnesHutSystem + BoundarySystem (GPU), all other systems the reference's";
#else
    const char* which = "reference systems";
#endif
    std::printf("{\"bodies\": %zu, \"ticks\": %d, \"ms_per_tick\": %.6f, \"body_steps_per_s\": %.6e, \"systems\": \"%s\"}\n",
                nb, ticks, ms_tick, (double)nb / (ms_tick * 1e-3), which);
    return 0;
}
