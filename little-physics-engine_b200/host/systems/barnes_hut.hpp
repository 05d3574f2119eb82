/**
 * @file barnes_hut.hpp — drop-in replacement for the reference's include/systems/barnes_hut.hpp.
 *
 * Same seam as the reference: class Systems::BarnesHutSystem deriving
 * ConfigurableSystem<BarnesHutConfig> with `void update(entt::registry&) override`
 * (reference include/systems/barnes_hut.hpp:58-74, include/systems/i_system.hpp:35), the same
 * BarnesHutConfig fields and defaults (barnes_hut.hpp:31-46), so src/sim.cpp:66-67,111,137-138 and
 * include/scenarios/i_scenario.hpp:15,34 compile unchanged when this directory precedes the
 * reference's include/ on the include path and barnes_hut.cpp here replaces src/systems/barnes_hut.cpp.
 *
 * What differs is private: instead of a host quadtree (nodePool_) the system owns a device context of
 * liblpe_bh.so (include/lpe_bh.h). update() stages Position/Velocity/Mass of the reference's own views
 * into flat arrays, runs one Barnes-Hut step on the GPU and writes the kicked velocities back — the only
 * observable effect of the reference's update() (barnes_hut.cpp:285-286). There is no CPU fallback: if no
 * CUDA device can be opened, update() reports it on std::cerr and returns, like the reference's own
 * error path (barnes_hut.cpp:76-79).
 */
#pragma once

#include <entt/entt.hpp>
#include <cstdint>
#include <vector>

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
    bool fuseMovement = false;    ///< also apply MovementSystem's drift on the device (then skip MovementSystem)
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

private:
    bool ensureContext();

    BarnesHutDeviceOptions options_;
    lpe_bh_ctx* ctx_ = nullptr;
    bool contextFailed_ = false;
    // staging (host, grown on demand, reused across ticks)
    std::vector<entt::entity> entities_;
    std::vector<double> x_, y_, vx_, vy_, m_;
    std::vector<std::uint8_t> comp_;
    std::vector<std::uint32_t> rank_;
};

}  // namespace Systems
