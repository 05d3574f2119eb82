/*
 * dropin_bench.cpp — wall-clock cost of the drop-in itself: Systems::BarnesHutSystem::update(entt::registry&)
 * (host/systems/barnes_hut.{hpp,cpp} -> liblpe_bh.so -> GPU) on a real EnTT registry built the way the reference's
 * scenarios build theirs (one entity after the other, Position / Velocity / Mass / ParticlePhase emplaced in order:
 * src/scenarios/keplerian_disk.cpp:130-138). This is what a maintainer who swaps the class in pays per tick.
 *
 * Product-side harness (no oracle code): compiled against the reference's HEADERS only (include/, vendor/entt), by
 * host/Makefile, where /root/reference exists; the binary travels to the GPU box.
 * usage: dropin_bench <n_bodies> <ticks> [pagewise|per_entity] [kick|fused]     -> one JSON line
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "systems/barnes_hut.hpp"   // ours (host/ precedes the reference include dir)
#include "lpe_bh.h"

int main(int argc, char** argv) {
    const uint64_t n = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 1000000;
    const int ticks = argc > 2 ? std::atoi(argv[2]) : 10;
    const std::string mode = argc > 3 ? argv[3] : "pagewise";
    const bool fused = argc > 4 && std::string(argv[4]) == "fused";
    const double U = 1048576.0;
    std::vector<double> x(n), y(n), vx(n), vy(n), m(n);
    if (lpe_bh_workload(0, n, 42, U, x.data(), y.data(), vx.data(), vy.data(), m.data())) return 2;   // C2: uniform disk

    SharedSystemConfig sc{};
    sc.UniverseSizeMeters = U;
    sc.GravitationalSoftener = U / 16384.0;
    sc.SecondsPerTick = 1.0 / 120.0;
    sc.TimeAcceleration = 1.0;
    sc.MetersPerPixel = U / 600.0;
    Systems::BarnesHutConfig bc;
    bc.theta = 0.5;
    bc.smallMassThreshold = 0.0;   // C2's configuration

    entt::registry reg;
    auto st = reg.create();
    reg.emplace<Components::SimulatorState>(st, 1.0, 1.0);
    for (uint64_t i = 0; i < n; ++i) {
        auto e = reg.create();
        reg.emplace<Components::Position>(e, x[i], y[i]);
        reg.emplace<Components::Velocity>(e, vx[i], vy[i]);
        reg.emplace<Components::Mass>(e, m[i]);
        reg.emplace<Components::ParticlePhase>(e, Components::Phase::Gas);
    }
    Systems::BarnesHutSystem sys;
    sys.setSharedSystemConfig(sc);
    sys.setSpecificConfig(bc);
    Systems::BarnesHutDeviceOptions opt;
    opt.pagewiseStaging = mode != "per_entity";
    opt.fuseMovement = fused;
    if (const char* t = std::getenv("LPE_STAGING_THREADS")) opt.stagingThreads = std::atoi(t);   // (experiments)
    sys.setDeviceOptions(opt);
    for (int w = 0; w < 3; ++w) sys.update(reg);   // warm-up: context, buffers, page-locked staging
    const auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < ticks; ++t) sys.update(reg);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / ticks;
    // a checksum so that the work cannot be optimised away and two runs can be compared
    double sum = 0.0;
    for (auto&& [e, v] : reg.view<Components::Velocity>().each()) sum += v.x * 1e-3 + v.y * 1e-3;
    std::printf("{\"bodies\": %llu, \"ticks\": %d, \"ms_per_update\": %.6f, \"body_steps_per_s\": %.6e, "
                "\"staging\": \"%s\", \"staging_path_taken\": %d, \"fused_movement\": %s, \"h2d_bytes_per_tick\": %llu, "
                "\"d2h_bytes_per_tick\": %llu, \"velocity_checksum\": %.17g}\n",
                (unsigned long long)n, ticks, ms, (double)n / (ms * 1e-3), mode.c_str(), sys.lastStagingPath(),
                fused ? "true" : "false", (unsigned long long)(40 * n), (unsigned long long)((fused ? 32 : 16) * n), sum);
    return 0;
}
