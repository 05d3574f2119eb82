/*
 * ref_harness.cpp — thin C-callable driver around the UNMODIFIED reference sources.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_abi.h). This file contains no physics: it
 * fills an entt::registry from flat arrays, calls the reference's own
 *   Systems::BarnesHutSystem::update   (/root/reference/src/systems/barnes_hut.cpp:50)
 *   Systems::MovementSystem::update    (/root/reference/src/systems/movement.cpp:13)
 * and copies the components back out. The reference .cpp files are compiled where
 * they lie under /root/reference by oracle/Makefile; nothing is copied into this repo.
 *
 * Two liberties, both forced by reference defects:
 *  - `#define private public` around barnes_hut.hpp, solely to pre-size nodePool_
 *    (defect D1: allocateNode()'s resize invalidates live pointers,
 *    barnes_hut.cpp:38-48) and to read the pool back for tree-parity dumps.
 *  - built with -DENTT_ID_TYPE=std::uint64_t so more than 2^20-1 entities fit
 *    (entt.hpp:11856); the reference compiles and behaves identically with it.
 */
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#define private public
#include "core/debug.hpp"          // DebugStats' counters are private statics
#include "systems/barnes_hut.hpp"
#undef private
#include "systems/movement.hpp"
#include "systems/boundary.hpp"
#include "entities/entity_components.hpp"
#include "entities/sim_components.hpp"

#include "core/debug.hpp"
#include "oracle_abi.h"

namespace {

using Clock = std::chrono::steady_clock;

struct World {
    entt::registry reg;
    std::vector<entt::entity> ents;
};

SharedSystemConfig makeShared(const orc_params& p) {
    SharedSystemConfig c{};
    c.UniverseSizeMeters = p.universe_size;
    c.TimeAcceleration = p.time_acceleration;
    c.MetersPerPixel = 1.0;
    c.SecondsPerTick = p.seconds_per_tick;
    c.GravitationalSoftener = p.softening;
    c.DragCoeff = 0.0;
    c.ParticleDensity = 0.0;
    c.GridSize = 1;
    c.CellSizePixels = 1.0;
    return c;
}

void fill(World& w, const orc_params& p, uint64_t n, const double* x, const double* y,
          const double* vx, const double* vy, const double* m, const uint8_t* comp) {
    // sim.cpp:81-101 creates the SimulatorState entity before the scenario's entities.
    auto st = w.reg.create();
    w.reg.emplace<Components::SimulatorState>(st, p.base_time_acceleration, p.time_scale);
    w.ents.resize(n);
    for (uint64_t i = 0; i < n; ++i) {
        const uint8_t c = comp ? comp[i] : uint8_t(ORC_HAS_MASS | ORC_HAS_VELOCITY);
        auto e = w.reg.create();
        w.ents[i] = e;
        // same emplace order as keplerian_disk.cpp:130-138 (Position, Velocity, Phase, Mass)
        w.reg.emplace<Components::Position>(e, x[i], y[i]);
        if (c & ORC_HAS_VELOCITY) w.reg.emplace<Components::Velocity>(e, vx ? vx[i] : 0.0, vy ? vy[i] : 0.0);
        if (c & ORC_LIQUID) w.reg.emplace<Components::ParticlePhase>(e, Components::Phase::Liquid);
        if (c & ORC_HAS_MASS) w.reg.emplace<Components::Mass>(e, m[i]);
        if (c & ORC_BOUNDARY) w.reg.emplace<Components::Boundary>(e, true);
    }
}

uint64_t autoPool(uint64_t n, uint64_t pool_nodes) {
    return pool_nodes ? pool_nodes : 12 * n + 4096;
}

void treeStats(const Systems::BarnesHutSystem& sys, double U, orc_stats* st) {
    if (!st) return;
    st->pool_nodes = sys.nextNodeIndex_;
    uint64_t nonempty = 0, internal = 0;
    int maxDepth = 0;
    for (size_t k = 0; k < sys.nextNodeIndex_; ++k) {
        const auto& nd = sys.nodePool_[k];
        if (nd.totalMass != 0.0) {
            ++nonempty;
            int depth = (int)std::lround(std::log2(U / nd.boundarySize));
            if (depth > maxDepth) maxDepth = depth;
        }
        if (!nd.isLeaf) ++internal;
    }
    st->nonempty_nodes = nonempty;
    st->internal_nodes = internal;
    st->max_depth = maxDepth;
}

}  // namespace

extern "C" {

/* Run nsteps of {BarnesHutSystem::update; [MovementSystem::update]} on the given bodies. */
int ref_bh_run(const orc_params* p, uint64_t n, const double* x, const double* y, const double* vx,
               const double* vy, const double* m, const uint8_t* comp, int nsteps, uint64_t pool_nodes,
               double* ox, double* oy, double* ovx, double* ovy, orc_stats* st) {
    if (!p || !x || !y || !m) return 1;
    if (st) std::memset(st, 0, sizeof(*st));
    World w;
    fill(w, *p, n, x, y, vx, vy, m, comp);

    Systems::BarnesHutSystem bh;
    Systems::MovementSystem mv;
    const SharedSystemConfig sc = makeShared(*p);
    bh.setSharedSystemConfig(sc);
    mv.setSharedSystemConfig(sc);
    Systems::BarnesHutConfig bc;
    bc.theta = p->theta;
    bc.smallMassThreshold = p->small_mass_threshold;
    bh.setSpecificConfig(bc);
    const uint64_t pool = autoPool(n, pool_nodes);
    bh.nodePool_.resize(pool);

    DebugStats::max_force = 0.0;   // the reference's own process globals (core/debug.hpp:24-41)
    DebugStats::total_force = 0.0;
    DebugStats::force_count = 0;
    auto t0 = Clock::now();
    for (int s = 0; s < nsteps; ++s) {
        bh.update(w.reg);
        if (p->run_movement) mv.update(w.reg);
    }
    auto t1 = Clock::now();

    treeStats(bh, p->universe_size, st);
    if (st) {
        st->force_max = DebugStats::max_force;
        st->force_sum = DebugStats::total_force;
        st->force_count = (uint64_t)DebugStats::force_count;
        st->total_seconds = std::chrono::duration<double>(t1 - t0).count();
        st->pool_overflow = bh.nodePool_.size() != pool ? 1 : 0;
    }
    for (uint64_t i = 0; i < n; ++i) {
        const auto e = w.ents[i];
        const auto& pos = w.reg.get<Components::Position>(e);
        if (ox) ox[i] = pos.x;
        if (oy) oy[i] = pos.y;
        if (const auto* v = w.reg.try_get<Components::Velocity>(e)) {
            if (ovx) ovx[i] = v->x;
            if (ovy) ovy[i] = v->y;
        } else {
            if (ovx) ovx[i] = vx ? vx[i] : 0.0;
            if (ovy) ovy[i] = vy ? vy[i] : 0.0;
        }
    }
    return (st && st->pool_overflow) ? 2 : 0;
}

/* Position of each body in the iteration of view<Position,Mass>(exclude<Boundary>)
 * (barnes_hut.cpp:117), i.e. the order in which buildTree inserts; 0xFFFFFFFF if not in the view. */
int ref_bh_view_rank(uint64_t n, const uint8_t* comp, uint32_t* rank_out) {
    orc_params p{};
    p.base_time_acceleration = p.time_scale = 1.0;
    std::vector<double> z(n, 0.0), one(n, 1.0);
    World w;
    fill(w, p, n, z.data(), z.data(), z.data(), z.data(), one.data(), comp);
    for (uint64_t i = 0; i < n; ++i) rank_out[i] = 0xFFFFFFFFu;
    // entity index = creation index + 1 (state entity first)
    uint32_t r = 0;
    auto view = w.reg.view<Components::Position, Components::Mass>(entt::exclude<Components::Boundary>);
    for (auto e : view) {
        const uint64_t idx = (uint64_t)entt::to_entity(e) - 1;
        rank_out[idx] = r++;
    }
    return 0;
}

/* Build the reference tree once and dump every non-empty node in pool (allocation) order. */
int ref_bh_tree(const orc_params* p, uint64_t n, const double* x, const double* y, const double* m,
                const uint8_t* comp, uint64_t pool_nodes, orc_node* out, uint64_t cap, uint64_t* count,
                orc_stats* st) {
    if (!p || !x || !y || !m) return 1;
    if (st) std::memset(st, 0, sizeof(*st));
    World w;
    fill(w, *p, n, x, y, nullptr, nullptr, m, comp);
    Systems::BarnesHutSystem bh;
    bh.setSharedSystemConfig(makeShared(*p));
    Systems::BarnesHutConfig bc;
    bc.theta = p->theta;
    bc.smallMassThreshold = p->small_mass_threshold;
    bh.setSpecificConfig(bc);
    const uint64_t pool = autoPool(n, pool_nodes);
    bh.nodePool_.resize(pool);
    auto t0 = Clock::now();
    bh.buildTree(w.reg);
    auto t1 = Clock::now();
    treeStats(bh, p->universe_size, st);
    if (st) {
        st->build_seconds = std::chrono::duration<double>(t1 - t0).count();
        st->pool_overflow = bh.nodePool_.size() != pool ? 1 : 0;
    }
    uint64_t k = 0;
    for (size_t i = 0; i < bh.nextNodeIndex_; ++i) {
        const auto& nd = bh.nodePool_[i];
        if (nd.totalMass == 0.0) continue;
        if (out && k < cap) {
            orc_node& o = out[k];
            o.mass = nd.totalMass;
            o.comx = nd.centerOfMassX;
            o.comy = nd.centerOfMassY;
            o.bx = nd.boundaryX;
            o.by = nd.boundaryY;
            o.bsize = nd.boundarySize;
            o.is_leaf = nd.isLeaf ? 1 : 0;
            o.all_small = nd.allSmall ? 1 : 0;
            // for an internal node singleParticle still names its first occupant (never reset by subdivide)
            o.single = (nd.singleParticle != entt::null) ? (int64_t)entt::to_entity(nd.singleParticle) - 1 : -1;
        }
        ++k;
    }
    if (count) *count = k;
    return (st && st->pool_overflow) ? 2 : 0;
}

const char* ref_bh_describe(void) {
    return "reference sean-peters-au/little-physics-engine: src/systems/barnes_hut.cpp + movement.cpp, "
           "compiled unmodified (g++ -O2, ENTT_ID_TYPE=uint64), single thread";
}

/* The reference's own BoundarySystem::update (src/systems/boundary.cpp) on a registry filled from the arrays. */
int ref_boundary(const orc_boundary_params* p, uint64_t n, double* x, double* y, double* vx, double* vy,
                 const uint8_t* comp) {
    if (!p || (n && (!x || !y || !vx || !vy))) return 1;
    entt::registry reg;
    std::vector<entt::entity> ents(n);
    for (uint64_t i = 0; i < n; ++i) {
        const unsigned cm = comp ? comp[i] : (ORC_HAS_MASS | ORC_HAS_VELOCITY);
        const auto e = reg.create();
        ents[i] = e;
        reg.emplace<Components::Position>(e, x[i], y[i]);
        if (cm & ORC_HAS_VELOCITY) reg.emplace<Components::Velocity>(e, vx[i], vy[i]);
        if (cm & ORC_ASLEEP) {
            Components::Sleep sl;
            sl.asleep = true;
            reg.emplace<Components::Sleep>(e, sl);
        }
    }
    Systems::BoundarySystem bs;
    SharedSystemConfig sc{};
    sc.UniverseSizeMeters = p->universe_size;
    sc.MetersPerPixel = 1.0;              // margin is handed over in metres
    bs.setSharedSystemConfig(sc);
    Systems::BoundaryConfig bc;
    bc.marginPixels = p->margin;
    bc.bounceDamping = p->bounce_damping;
    bc.maxSpeed = p->max_speed;
    bs.setSpecificConfig(bc);
    bs.update(reg);
    for (uint64_t i = 0; i < n; ++i) {
        const auto& pos = reg.get<Components::Position>(ents[i]);
        x[i] = pos.x;
        y[i] = pos.y;
        if (auto* v = reg.try_get<Components::Velocity>(ents[i])) {
            vx[i] = v->x;
            vy[i] = v->y;
        }
    }
    return 0;
}

}  // extern "C"
