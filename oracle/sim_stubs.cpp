// Stubs that let the reference's src/sim.cpp link headless on Linux; TEST INFRASTRUCTURE ONLY.
//  * FluidSystem is Apple-Metal code (src/systems/fluid/fluid.cpp); the real update() returns at once when the registry
//    holds no liquid particle (fluid.cpp:966-972), which is the case in the Keplerian scenario: a no-op stands in.
//  * ContactSolver::solveContactConstraints is ARM-NEON code (src/systems/rigid/contact_solver.cpp:11,207-254); it is
//    only reached with solid bodies in contact, none exist here.
#include "systems/fluid/fluid.hpp"
#include "systems/rigid/contact_solver.hpp"

namespace Systems {
FluidSystem::FluidSystem() {}
FluidSystem::~FluidSystem() {}
void FluidSystem::update(entt::registry&) {}
}  // namespace Systems

namespace RigidBodyCollision {
void ContactSolver::solveContactConstraints(entt::registry&, ContactManager&, const ContactSolverConfig&) {}
}  // namespace RigidBodyCollision
