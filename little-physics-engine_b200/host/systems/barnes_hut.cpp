/**
 * @file barnes_hut.cpp — ECS side of the drop-in BarnesHutSystem (see barnes_hut.hpp here).
 *
 * Mirrors the control flow of the reference's update() (src/systems/barnes_hut.cpp:50-99):
 *   early exit when no mass reaches smallMassThreshold (:55-71), SimulatorState lookup with the same
 *   warning (:75-80), then "build tree + force loop" — which here is one call into liblpe_bh.so.
 */
#include "systems/barnes_hut.hpp"

#include <iostream>

#include "core/constants.hpp"
#include "core/profile.hpp"
#include "lpe_bh.h"

namespace Systems {

BarnesHutSystem::BarnesHutSystem() = default;

BarnesHutSystem::~BarnesHutSystem() {
    if (ctx_) lpe_bh_destroy(ctx_);
}

bool BarnesHutSystem::ensureContext() {
    if (ctx_) return true;
    if (contextFailed_) return false;
    if (lpe_bh_create(options_.device, &ctx_) != 0) {
        std::cerr << "[BarnesHut] Warning: cannot open CUDA device " << options_.device << ": "
                  << lpe_bh_last_error(nullptr) << ". Skipping update.\n";
        contextFailed_ = true;
        ctx_ = nullptr;
        return false;
    }
    return true;
}

void BarnesHutSystem::update(entt::registry& registry) {
    PROFILE_SCOPE("BarnesHutSystem");

    // Early exit, reference barnes_hut.cpp:55-71: skip when every non-boundary mass is below the threshold.
    if (specificConfig.smallMassThreshold > 0.0) {
        bool shouldSkip = true;
        auto massCheckView = registry.view<Components::Mass>(entt::exclude<Components::Boundary>);
        for (auto entity : massCheckView) {
            if (massCheckView.get<Components::Mass>(entity).value >= specificConfig.smallMassThreshold) {
                shouldSkip = false;
                break;
            }
        }
        if (shouldSkip) return;
    }

    // reference barnes_hut.cpp:75-80
    auto stateView = registry.view<Components::SimulatorState>();
    if (stateView.empty()) {
        std::cerr << "[BarnesHut] Warning: No SimulatorState found. Skipping update.\n";
        return;
    }
    const auto& simState = stateView.get<Components::SimulatorState>(stateView.front());

    if (!ensureContext()) return;

    // Stage the bodies in the iteration order of buildTree's own view (barnes_hut.cpp:117): that order IS the
    // insertion order, which decides each cell's first occupant. Every target of the force
    // loop (view<Position,Velocity,Mass>, :89) is also in this view, so one pass stages sources and targets.
    auto insertView = registry.view<Components::Position, Components::Mass>(entt::exclude<Components::Boundary>);
    entities_.clear();
    x_.clear(); y_.clear(); vx_.clear(); vy_.clear(); m_.clear(); comp_.clear(); rank_.clear();
    for (auto entity : insertView) {
        const auto& pos = insertView.get<Components::Position>(entity);
        const auto& mass = insertView.get<Components::Mass>(entity);
        std::uint8_t comp = LPE_HAS_MASS;
        double vx = 0.0, vy = 0.0;
        if (const auto* vel = registry.try_get<Components::Velocity>(entity)) {
            comp |= LPE_HAS_VELOCITY;
            vx = vel->x;
            vy = vel->y;
        }
        if (options_.fuseMovement) {
            if (const auto* ph = registry.try_get<Components::ParticlePhase>(entity))
                if (ph->phase == Components::Phase::Liquid) comp |= LPE_LIQUID;   // movement.cpp:25-29
        }
        rank_.push_back(static_cast<std::uint32_t>(entities_.size()));
        entities_.push_back(entity);
        x_.push_back(pos.x); y_.push_back(pos.y);
        vx_.push_back(vx); vy_.push_back(vy);
        m_.push_back(mass.value);
        comp_.push_back(comp);
    }
    if (entities_.empty()) return;

    lpe_bh_params p{};
    p.universe_size = sysConfig.UniverseSizeMeters;
    p.softening = sysConfig.GravitationalSoftener;
    p.theta = specificConfig.theta;
    p.small_mass_threshold = specificConfig.smallMassThreshold;
    p.G = SimulatorConstants::RealG;
    p.dt_kick = sysConfig.SecondsPerTick * simState.baseTimeAcceleration * simState.timeScale;  // barnes_hut.cpp:284
    p.dt_drift = sysConfig.SecondsPerTick * sysConfig.TimeAcceleration;                           // movement.cpp:17
    p.quirk_mode = options_.referenceQuirk ? 1 : 0;
    p.precision = options_.strictFp64 ? LPE_PREC_STRICT : LPE_PREC_FAST;
    p.do_drift = options_.fuseMovement ? 1 : 0;
    p.max_depth = 0;

    if (lpe_bh_update_host(ctx_, &p, entities_.size(), x_.data(), y_.data(), vx_.data(), vy_.data(), m_.data(),
                           rank_.data(), comp_.data()) != 0) {
        std::cerr << "[BarnesHut] Warning: device step failed: " << lpe_bh_last_error(ctx_) << ". Skipping update.\n";
        return;
    }

    // The reference mutates Velocity in place through the view reference (barnes_hut.cpp:285-286): no signals.
    for (std::size_t i = 0; i < entities_.size(); ++i) {
        if (!(comp_[i] & LPE_HAS_VELOCITY)) continue;
        auto& vel = registry.get<Components::Velocity>(entities_[i]);
        vel.x = vx_[i];
        vel.y = vy_[i];
        if (options_.fuseMovement && !(comp_[i] & LPE_LIQUID)) {
            auto& pos = registry.get<Components::Position>(entities_[i]);
            pos.x = x_[i];
            pos.y = y_[i];
        }
    }
}

}  // namespace Systems
