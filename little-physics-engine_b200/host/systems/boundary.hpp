/**
 * @file boundary.hpp — drop-in replacement for the reference's include/systems/boundary.hpp.
 *
 * Same seam: class Systems::BoundarySystem deriving ConfigurableSystem<BoundaryConfig> with
 * `void update(entt::registry&) override` (reference include/systems/boundary.hpp:38-65), the same BoundaryConfig
 * fields and defaults (boundary.hpp:27-36). update() stages the reference's own view<Position, Velocity>
 * (src/systems/boundary.cpp:22) — with the Sleep test of boundary.cpp:29-31 as a component bit — into flat arrays,
 * runs lpe_bh_boundary on the device and writes Position and Velocity back in place: the observable effect of the
 * reference's update(), bit for bit. No CPU fallback: without a CUDA device update() reports it and returns.
 *
 * Through host buffers this pass is PCIe-bound and only there for completeness of the ECS seam; its purpose is
 * the resident tick (lpe_bh_boundary + lpe_bh_step on one context), where it is one launch.
 */
#pragma once

#include <entt/entt.hpp>
#include <cstdint>
#include <vector>

#include "systems/i_system.hpp"
#include "entities/entity_components.hpp"

struct lpe_bh_ctx;

namespace Systems {

/** Same fields, meaning and defaults as the reference (boundary.hpp:27-36). */
struct BoundaryConfig {
    double marginPixels = 15.0;
    double bounceDamping = 0.7;
    double maxSpeed = 1.0;
};

class BoundarySystem : public ConfigurableSystem<BoundaryConfig> {
public:
    BoundarySystem();
    ~BoundarySystem() override;
    BoundarySystem(const BoundarySystem&) = delete;
    BoundarySystem& operator=(const BoundarySystem&) = delete;

    /** Clamps and bounces every (Position, Velocity) entity that is not asleep. */
    void update(entt::registry& registry) override;

    void setDevice(int device) { device_ = device; }

private:
    int device_ = 0;
    lpe_bh_ctx* ctx_ = nullptr;
    bool contextFailed_ = false;
    std::vector<entt::entity> entities_;
    std::vector<double> x_, y_, vx_, vy_, m_;
    std::vector<std::uint8_t> comp_;
};

}  // namespace Systems
