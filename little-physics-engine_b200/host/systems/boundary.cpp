/**
 * @file boundary.cpp — ECS side of the drop-in BoundarySystem (see boundary.hpp here).
 * Control flow of the reference's update() (src/systems/boundary.cpp:13-69): derive the margin in metres (:16),
 * walk view<Position, Velocity> (:22), skip sleepers (:29-31), clamp / bounce / cap — the last part is one call
 * into liblpe_bh.so.
 */
#include "systems/boundary.hpp"

#include <iostream>

#include "core/profile.hpp"
#include "lpe_bh.h"

namespace Systems {

BoundarySystem::BoundarySystem() = default;

BoundarySystem::~BoundarySystem() {
    if (ctx_) lpe_bh_destroy(ctx_);
}

void BoundarySystem::update(entt::registry& registry) {
    PROFILE_SCOPE("BoundarySystem");

    if (!ctx_) {
        if (contextFailed_) return;
        if (lpe_bh_create(device_, &ctx_) != 0) {
            std::cerr << "[Boundary] Warning: cannot open CUDA device " << device_ << ": "
                      << lpe_bh_last_error(nullptr) << ". Skipping update.\n";
            contextFailed_ = true;
            ctx_ = nullptr;
            return;
        }
    }

    auto view = registry.view<Components::Position, Components::Velocity>();   // boundary.cpp:22
    entities_.clear();
    x_.clear(); y_.clear(); vx_.clear(); vy_.clear(); comp_.clear();
    for (auto&& [entity, pos, vel] : view.each()) {
        std::uint8_t comp = LPE_HAS_VELOCITY;
        if (auto* sleep = registry.try_get<Components::Sleep>(entity); sleep && sleep->asleep) comp |= LPE_ASLEEP;
        entities_.push_back(entity);
        x_.push_back(pos.x); y_.push_back(pos.y);
        vx_.push_back(vel.x); vy_.push_back(vel.y);
        comp_.push_back(comp);
    }
    if (entities_.empty()) return;
    m_.assign(entities_.size(), 0.0);   // masses play no part in this system

    lpe_bh_boundary_params bp{};
    bp.universe_size = sysConfig.UniverseSizeMeters;                            // boundary.cpp:17
    bp.margin = specificConfig.marginPixels * sysConfig.MetersPerPixel;         // boundary.cpp:16
    bp.bounce_damping = specificConfig.bounceDamping;                           // boundary.cpp:18
    bp.max_speed = specificConfig.maxSpeed;                                     // boundary.cpp:19

    const std::uint64_t n = entities_.size();
    if (lpe_bh_upload(ctx_, n, x_.data(), y_.data(), vx_.data(), vy_.data(), m_.data(), nullptr, comp_.data()) != 0 ||
        lpe_bh_boundary(ctx_, &bp) != 0 ||
        lpe_bh_download(ctx_, x_.data(), y_.data(), vx_.data(), vy_.data()) != 0) {
        std::cerr << "[Boundary] Warning: device pass failed: " << lpe_bh_last_error(ctx_) << ". Skipping update.\n";
        return;
    }

    // the reference mutates Position / Velocity in place through the view references (boundary.cpp:36-66)
    for (std::size_t i = 0; i < entities_.size(); ++i) {
        auto& pos = registry.get<Components::Position>(entities_[i]);
        auto& vel = registry.get<Components::Velocity>(entities_[i]);
        pos.x = x_[i]; pos.y = y_[i];
        vel.x = vx_[i]; vel.y = vy_[i];
    }
}

}  // namespace Systems
