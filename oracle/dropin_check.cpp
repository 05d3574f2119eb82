/*
 * dropin_check.cpp — proves the drop-in: an entt::registry driven through OUR Systems::BarnesHutSystem
 * (little-physics-engine_b200/host/systems/barnes_hut.{hpp,cpp} -> liblpe_bh.so -> GPU) followed by the
 * REFERENCE's own MovementSystem, compared with the same registry driven by the reference's own
 * BarnesHutSystem (inside oracle/_ref/libref_bh.so, loaded with RTLD_LOCAL so the two same-named classes
 * never meet).
 *
 * TEST INFRASTRUCTURE ONLY. Built by oracle/Makefile (target `dropin`) against the reference headers and
 * vendored EnTT under /root/reference, into oracle/_ref/; run on the GPU box by tests/test_dropin_gpu.py.
 * Usage: dropin_check <n> <seed> [keplerian|uniform|boundary] [steps]   -> one JSON line on stdout, exit 0 on parity.
 * kind "boundary" runs OUR Systems::BoundarySystem (host/systems/boundary.{hpp,cpp}) against the reference's.
 */
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "systems/barnes_hut.hpp"  // ours (host/ precedes the reference include dir)
#include "systems/boundary.hpp"    // ours as well
#include "systems/movement.hpp"    // the reference's
#include "lpe_bh.h"
#include "oracle_abi.h"

typedef int (*ref_boundary_fn)(const orc_boundary_params*, uint64_t, double*, double*, double*, double*, const uint8_t*);

// kind "boundary": our Systems::BoundarySystem on a registry vs the reference's (ref_boundary), bit for bit
static int boundary_check(void* h, uint64_t n, uint64_t seed) {
    auto ref_boundary = reinterpret_cast<ref_boundary_fn>(dlsym(h, "ref_boundary"));
    if (!ref_boundary) { std::printf("{\"error\": \"ref_boundary missing\"}\n"); return 2; }
    const double U = 1000.0;
    std::vector<double> x(n), y(n), vx(n), vy(n);
    std::vector<uint8_t> comp(n);
    uint64_t s = seed * 6364136223846793005ull + 1442695040888963407ull;
    auto u = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (double)(s >> 11) * (1.0 / 9007199254740992.0); };
    for (uint64_t i = 0; i < n; ++i) {
        x[i] = -100.0 + 1200.0 * u(); y[i] = -100.0 + 1200.0 * u();
        vx[i] = 6.0 * (u() - 0.5); vy[i] = 6.0 * (u() - 0.5);
        comp[i] = ORC_HAS_MASS | ORC_HAS_VELOCITY;
        if (i % 7 == 3) comp[i] = ORC_HAS_MASS;          // no Velocity
        if (i % 9 == 5) comp[i] |= ORC_ASLEEP;
    }
    SharedSystemConfig sc{};
    sc.UniverseSizeMeters = U;
    sc.MetersPerPixel = 0.5;
    Systems::BoundaryConfig bc;   // defaults: 15 px, 0.7, 1.0
    entt::registry reg;
    std::vector<entt::entity> ents(n);
    for (uint64_t i = 0; i < n; ++i) {
        auto e = reg.create();
        ents[i] = e;
        reg.emplace<Components::Position>(e, x[i], y[i]);
        if (comp[i] & ORC_HAS_VELOCITY) reg.emplace<Components::Velocity>(e, vx[i], vy[i]);
        if (comp[i] & ORC_ASLEEP) { Components::Sleep sl; sl.asleep = true; reg.emplace<Components::Sleep>(e, sl); }
    }
    Systems::BoundarySystem bs;
    bs.setSharedSystemConfig(sc);
    bs.setSpecificConfig(bc);
    bs.update(reg);
    orc_boundary_params bp{U, bc.marginPixels * sc.MetersPerPixel, bc.bounceDamping, bc.maxSpeed};
    if (ref_boundary(&bp, n, x.data(), y.data(), vx.data(), vy.data(), comp.data())) return 2;
    uint64_t bad = 0, moved = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const auto& q = reg.get<Components::Position>(ents[i]);
        if (std::memcmp(&q.x, &x[i], 8) || std::memcmp(&q.y, &y[i], 8)) ++bad;
        if (const auto* v = reg.try_get<Components::Velocity>(ents[i])) {
            if (std::memcmp(&v->x, &vx[i], 8) || std::memcmp(&v->y, &vy[i], 8)) ++bad;
        }
        if (x[i] == bp.margin || x[i] == U - bp.margin) ++moved;
    }
    std::printf("{\"n\": %llu, \"kind\": \"boundary\", \"clamped_x\": %llu, \"mismatches\": %llu, \"ok\": %s}\n",
                (unsigned long long)n, (unsigned long long)moved, (unsigned long long)bad, bad == 0 && moved > 0 ? "true" : "false");
    return bad == 0 && moved > 0 ? 0 : 1;
}

typedef int (*ref_run_fn)(const orc_params*, uint64_t, const double*, const double*, const double*, const double*,
                          const double*, const uint8_t*, int, uint64_t, double*, double*, double*, double*, orc_stats*);

int main(int argc, char** argv) {
    const uint64_t n = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 2000;
    const uint64_t seed = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 1;
    const std::string kind = argc > 3 ? argv[3] : "keplerian";
    const int steps = argc > 4 ? std::atoi(argv[4]) : 3;
    // "misalign": the Velocity pool gets another packed order than Position / Mass (a few components removed and put back),
    // which the page-wise staging only notices while the device already works on the tick: the tick must start over on the
    // entity-by-entity path. (The reference's result does not depend on the Velocity pool's order.)
    const bool misalign = argc > 5 && std::string(argv[5]) == "misalign";

    // locate libref_bh.so next to this binary
    std::string self = argv[0];
    const size_t slash = self.rfind('/');
    const std::string dir = slash == std::string::npos ? "." : self.substr(0, slash);
    void* h = dlopen((dir + "/libref_bh.so").c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) { std::printf("{\"error\": \"%s\"}\n", dlerror()); return 2; }
    if (kind == "boundary") return boundary_check(h, n, seed);
    auto ref_run = reinterpret_cast<ref_run_fn>(dlsym(h, "ref_bh_run"));
    if (!ref_run) { std::printf("{\"error\": \"ref_bh_run missing\"}\n"); return 2; }

    const bool kep = kind == "keplerian";
    const double U = kep ? 6e9 : 1024.0;
    std::vector<double> x(n), y(n), vx(n), vy(n), m(n);
    if (lpe_bh_workload(kep ? 3 : 0, n, seed, U, x.data(), y.data(), vx.data(), vy.data(), m.data())) return 2;

    SharedSystemConfig sc{};
    sc.UniverseSizeMeters = U;
    sc.GravitationalSoftener = kep ? 2e7 : U / 16384.0;
    sc.SecondsPerTick = 1.0 / 120.0;
    sc.TimeAcceleration = kep ? 0.8107 : 1.0;
    sc.MetersPerPixel = U / 600.0;
    Systems::BarnesHutConfig bc;  // defaults: theta 0.5, smallMassThreshold 1e3

    // ---- our system on a real registry ----
    entt::registry reg;
    auto st = reg.create();
    reg.emplace<Components::SimulatorState>(st, 1.0, 1.0);
    std::vector<entt::entity> ents(n);
    for (uint64_t i = 0; i < n; ++i) {
        auto e = reg.create();
        ents[i] = e;
        reg.emplace<Components::Position>(e, x[i], y[i]);
        reg.emplace<Components::Velocity>(e, vx[i], vy[i]);
        reg.emplace<Components::ParticlePhase>(e, Components::Phase::Gas);
        reg.emplace<Components::Mass>(e, m[i]);
    }
    if (misalign) {
        for (uint64_t i = 3; i < n; i += 97) {
            const auto v = reg.get<Components::Velocity>(ents[i]);
            reg.remove<Components::Velocity>(ents[i]);
            reg.emplace<Components::Velocity>(ents[i], v.x, v.y);
        }
    }
    Systems::BarnesHutSystem bh;
    Systems::MovementSystem mv;
    bh.setSharedSystemConfig(sc);
    bh.setSpecificConfig(bc);
    mv.setSharedSystemConfig(sc);
    for (int s = 0; s < steps; ++s) {
        bh.update(reg);
        mv.update(reg);
    }

    // ---- the reference on the same bodies ----
    orc_params p{};
    p.universe_size = U; p.softening = sc.GravitationalSoftener; p.seconds_per_tick = sc.SecondsPerTick;
    p.time_acceleration = sc.TimeAcceleration; p.base_time_acceleration = 1.0; p.time_scale = 1.0;
    p.theta = bc.theta; p.small_mass_threshold = bc.smallMassThreshold; p.G = 6.674e-11; p.run_movement = 1; p.quirk = 1;
    std::vector<double> rx(n), ry(n), rvx(n), rvy(n);
    orc_stats stt{};
    const int rc = ref_run(&p, n, x.data(), y.data(), vx.data(), vy.data(), m.data(), nullptr, steps, 0, rx.data(),
                           ry.data(), rvx.data(), rvy.data(), &stt);
    if (rc) { std::printf("{\"error\": \"ref_bh_run rc=%d\"}\n", rc); return 2; }

    // ---- compare the velocity change and the positions ----
    double num = 0.0, den = 0.0, maxdx = 0.0;
    std::vector<double> mags(n);
    for (uint64_t i = 0; i < n; ++i) mags[i] = std::hypot(rvx[i] - vx[i], rvy[i] - vy[i]);
    std::vector<double> sorted = mags;
    std::nth_element(sorted.begin(), sorted.begin() + n / 2, sorted.end());
    const double floorMag = 1e-3 * sorted[n / 2];
    double maxrel = 0.0;
    for (uint64_t i = 0; i < n; ++i) {
        const auto& v = reg.get<Components::Velocity>(ents[i]);
        const auto& q = reg.get<Components::Position>(ents[i]);
        const double ex = v.x - rvx[i], ey = v.y - rvy[i];
        num += ex * ex + ey * ey;
        den += mags[i] * mags[i];
        maxrel = std::max(maxrel, std::hypot(ex, ey) / std::max(mags[i], floorMag));
        maxdx = std::max(maxdx, std::hypot(q.x - rx[i], q.y - ry[i]) / U);
    }
    const double norm = std::sqrt(num / std::max(den, 1e-300));
    // (page-wise staging must have been taken exactly when the pools were aligned)
    const bool pathOk = bh.lastStagingPath() == (misalign ? 0 : 1);
    const bool ok = norm <= 1e-4 && maxrel <= 1e-4 && pathOk;
    std::printf("{\"n\": %llu, \"kind\": \"%s\", \"steps\": %d, \"dv_norm_rel\": %.3e, \"dv_max_rel\": %.3e, "
                "\"x_max_over_U\": %.3e, \"staging_path\": %d, \"ok\": %s}\n",
                (unsigned long long)n, kind.c_str(), steps, norm, maxrel, maxdx, bh.lastStagingPath(), ok ? "true" : "false");
    return ok ? 0 : 1;
}
