/*
 * oracle_abi.h — shared plain-C types for the two CPU checkers under oracle/.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load these libraries, and only as the checker / baseline.
 *
 *   oracle/_ref/libref_bh.so   the reference's own barnes_hut.cpp + movement.cpp,
 *                              compiled unmodified from /root/reference (ref_harness.cpp)
 *   oracle/liboracle_bh.so     plain-C restatement of the same algorithm (bh_oracle.c)
 *
 * Both take bodies as flat arrays in *creation order* (index i = i-th entity
 * created) with a per-body component mask, mirroring what an EnTT registry
 * would hold (reference: include/entities/entity_components.hpp:21-29,111-116).
 */
#ifndef LPE_ORACLE_ABI_H
#define LPE_ORACLE_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* component mask bits (Position is always present) */
#define ORC_HAS_MASS     1u  /* Components::Mass */
#define ORC_HAS_VELOCITY 2u  /* Components::Velocity */
#define ORC_BOUNDARY     4u  /* Components::Boundary  (excluded from every view) */
#define ORC_LIQUID       8u  /* ParticlePhase == Liquid (skipped by MovementSystem) */
#define ORC_ASLEEP      16u  /* Components::Sleep with asleep == true (skipped by BoundarySystem, boundary.cpp:29-31) */

/* BoundarySystem (reference include/systems/boundary.hpp:27-36, src/systems/boundary.cpp:16-19) */
typedef struct {
    double universe_size;    /* SharedSystemConfig::UniverseSizeMeters */
    double margin;           /* BoundaryConfig::marginPixels * SharedSystemConfig::MetersPerPixel, in metres */
    double bounce_damping;   /* BoundaryConfig::bounceDamping */
    double max_speed;        /* BoundaryConfig::maxSpeed */
} orc_boundary_params;

typedef struct {
    double universe_size;          /* SharedSystemConfig::UniverseSizeMeters   (shared_system_config.hpp:11) */
    double softening;              /* SharedSystemConfig::GravitationalSoftener (shared_system_config.hpp:15) */
    double seconds_per_tick;       /* SharedSystemConfig::SecondsPerTick */
    double time_acceleration;      /* SharedSystemConfig::TimeAcceleration  (drift dt, movement.cpp:17) */
    double base_time_acceleration; /* SimulatorState (sim_components.hpp:4-11)  (kick dt, barnes_hut.cpp:284) */
    double time_scale;
    double theta;                  /* BarnesHutConfig::theta (barnes_hut.hpp:36) */
    double small_mass_threshold;   /* BarnesHutConfig::smallMassThreshold (barnes_hut.hpp:45) */
    double G;                      /* port only; the compiled reference uses RealG = 6.674e-11 (constants.cpp:8) */
    int32_t run_movement;          /* 1: also run MovementSystem::update after BarnesHutSystem::update */
    int32_t quirk;                 /* port only: 1 = reference first-occupant double count, 0 = textbook tree */
} orc_params;

/* One dumped tree node (only nodes with totalMass != 0 are dumped). */
typedef struct {
    double mass, comx, comy;
    double bx, by, bsize;
    int64_t single;     /* creation index of singleParticle: the occupant of a leaf, the FIRST occupant of an internal node */
    int32_t is_leaf;
    int32_t all_small;
} orc_node;

typedef struct {
    uint64_t pool_nodes;      /* nodes allocated by the build (1 + 4*internal) */
    uint64_t nonempty_nodes;
    uint64_t internal_nodes;
    uint64_t accepted;        /* port only: accepted interactions summed over targets */
    uint64_t visited;         /* port only: non-empty nodes visited summed over targets */
    int32_t  max_depth;
    int32_t  pool_overflow;   /* ref only: 1 if the node pool had to grow (defect D1 => result invalid) */
    double   build_seconds;   /* port only (the reference does not split build from force) */
    double   force_seconds;
    double   total_seconds;
    /* DebugStats::updateForce (include/core/debug.hpp:37-41), called once per accepted node at barnes_hut.cpp:278:
     * max / sum / count of force = G*M*m/distSq over the run. ref: the reference's own process globals. */
    double   force_max, force_sum;
    uint64_t force_count;
} orc_stats;

/* BoundarySystem::update on flat arrays, in place: bodies with ORC_HAS_VELOCITY and without ORC_ASLEEP are clamped
 * to [margin, U - margin] and bounced (boundary.cpp:21-66). Same signature in both libraries. */
int orc_boundary(const orc_boundary_params* p, uint64_t n, double* x, double* y, double* vx, double* vy,
                 const uint8_t* comp);
int ref_boundary(const orc_boundary_params* p, uint64_t n, double* x, double* y, double* vx, double* vy,
                 const uint8_t* comp);

#ifdef __cplusplus
}
#endif
#endif
