/**
 * @file barnes_hut.hpp — drop-in replacement of the reference's Systems::BarnesHutSystem.
 *
 * Same class name, base class, public methods and config struct as the reference header
 * (include/systems/barnes_hut.hpp:31-74 in sean-peters-au/little-physics-engine), so src/sim.cpp
 * (construction at :111, dynamic_cast configuration at :66-67 and :137-138) and
 * include/scenarios/i_scenario.hpp:15,34 compile unchanged when this directory precedes the reference's include
 * directory. The quadtree, the force traversal and the velocity kick run in liblpe_bh.so on a B200
 * (C ABI: include/lpe_bh.h); this class only stages the registry's components, as the reference's own GPU system does
 * for its fluid particles (src/systems/fluid.cpp:250-302).
 *
 * There is no CPU fallback: if no CUDA device can be opened, update() logs to std::cerr and returns,
 * the reference's own error style (barnes_hut.cpp:76-79).
 *
 * Staging: Position / Velocity are {double x, y} records and Mass one double, kept by EnTT in
 * packed pools of 1024-element pages (entt.hpp:60, storage.raw() :17392, packed entity array storage.data() :16319).
 * When the three pools hold the same entities in the same order — true whenever the components were emplaced entity
 * by entity and nothing was removed, e.g. keplerian_disk.cpp:130-138 — and no entity is a Boundary, the pages are
 * copied as they are into page-locked buffers (a few worker threads, one memcpy per page and component) and the
 * new velocities are copied back the same way. Otherwise the reference's own views are walked entity by entity.
 */
#pragma once

#include <entt/entt.hpp>
#include <cstdint>
#include <memory>

#include "systems/i_system.hpp"
#include "entities/entity_components.hpp"
#include "entities/sim_components.hpp"

struct lpe_bh_ctx;

namespace Systems {

/** Same fields, meaning and defaults as the reference (barnes_hut.hpp:31-46). */
struct BarnesHutConfig {
    double theta = 0.5;
    double smallMassThreshold = 1e3;
};

/** Device-side options; the defaults reproduce the reference's observable behaviour exactly. */
struct BarnesHutDeviceOptions {
    int device = 0;
    bool referenceQuirk = true;   ///< first-occupant double count of the reference tree
    bool strictFp64 = false;      ///< every interaction in fp64 in the reference's expression order
    /** Also apply MovementSystem's drift on the device to every Position + Velocity non-Boundary, non-Liquid entity
     *  (movement.cpp:20-33), massless ones included; MovementSystem must then be skipped by the caller. The drift uses
     *  the velocity right after the kick, i.e. before any system that the tick would run between the two. */
    bool fuseMovement = false;
    /** Collect what the reference feeds DebugStats::updateForce with at barnes_hut.cpp:278 (max / sum / count of
     *  G*M*m/distSq over the accepted nodes) with a device reduction; read it with lastForceStats(). Uses the counting
     *  kernels, which are slower. The reference's own counters are private statics with prints compiled out
     *  (core/debug.hpp:6,76), so they are not written. */
    bool collectForceStats = false;
    bool pagewiseStaging = true;  ///< take the pool-page fast path when the pools allow it
    int stagingThreads = 4;       ///< worker threads of the page-wise copies (0 = copy on the calling thread)
};

class BarnesHutSystem : public ConfigurableSystem<BarnesHutConfig> {
public:
    BarnesHutSystem();
    ~BarnesHutSystem() override;
    BarnesHutSystem(const BarnesHutSystem&) = delete;
    BarnesHutSystem& operator=(const BarnesHutSystem&) = delete;

    /** Applies one gravity kick to every (Position, Velocity, Mass) non-Boundary entity. */
    void update(entt::registry& registry) override;

    void setDeviceOptions(const BarnesHutDeviceOptions& o) { options_ = o; }
    const BarnesHutDeviceOptions& getDeviceOptions() const { return options_; }
    /** How the last update staged the registry: 1 = pool pages, 0 = entity by entity, -1 = no update yet. */
    int lastStagingPath() const { return lastPath_; }
    struct ForceStats { double maxForce = 0.0, totalForce = 0.0; unsigned long long count = 0; };
    /** DebugStats::updateForce's three numbers for the last update (collectForceStats). */
    const ForceStats& lastForceStats() const { return forceStats_; }

private:
    struct Staging;   // page-locked buffers + worker threads (barnes_hut.cpp)
    bool ensureContext();
    bool stagePagewise(entt::registry& registry, std::size_t& n);
    std::size_t stagePerEntity(entt::registry& registry);

    BarnesHutDeviceOptions options_;
    lpe_bh_ctx* ctx_ = nullptr;
    bool contextFailed_ = false;
    int lastPath_ = -1;
    ForceStats forceStats_;
    std::unique_ptr<Staging> st_;
};

}  // namespace Systems
