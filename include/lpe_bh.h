/*
 * lpe_bh.h — C ABI of the B200-native Barnes-Hut step (liblpe_bh.so).
 *
 * This is the drop-in boundary for ONE hot path of sean-peters-au/little-physics-engine:
 *   Systems::BarnesHutSystem::update(entt::registry&)   reference src/systems/barnes_hut.cpp:50-99
 *   Systems::MovementSystem::update(entt::registry&)    reference src/systems/movement.cpp:13-39
 * The reference has no FFI layer of its own; the C++ class in
 * little-physics-engine_b200/host/systems/barnes_hut.hpp keeps the reference's class/ISystem
 * contract and calls these entry points. Plain pointers and sizes only; host buffers stay
 * caller-owned. Every entry point returns 0 on success, non-zero on error with a message
 * available from lpe_bh_last_error(). There is NO CPU fallback: without a CUDA device
 * lpe_bh_create fails.
 *
 * Bodies are flat arrays in entity-creation order (index i = i-th entity), with a per-body
 * component mask mirroring the registry (entity_components.hpp:21-29,111-116).
 */
#ifndef LPE_BH_H
#define LPE_BH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LPE_BH_ABI_VERSION 2

/* component mask bits (Position is implied) */
#define LPE_HAS_MASS     1u /* Components::Mass      -> source if not Boundary (barnes_hut.cpp:117) */
#define LPE_HAS_VELOCITY 2u /* Components::Velocity  -> target if it also has Mass (barnes_hut.cpp:89) */
#define LPE_BOUNDARY     4u /* Components::Boundary  -> excluded from every view */
#define LPE_LIQUID       8u /* ParticlePhase::Liquid -> not moved by MovementSystem (movement.cpp:25-29) */
#define LPE_ASLEEP      16u /* Components::Sleep{asleep = true} -> skipped by BoundarySystem (boundary.cpp:29-31) */

/* lpe_bh_params.precision */
#define LPE_PREC_FAST   0 /* fp64 state + fp64 differences, fp32 interaction math, exact fp64 re-test of borderline theta decisions */
#define LPE_PREC_STRICT 1 /* every interaction in fp64, in the reference's expression order */

/* lpe_bh_params.key_order */
#define LPE_KEYS_AUTO    0 /* Hilbert for FAST precision (compact warps: measured 8-12 % faster traversal), Morton for STRICT */
#define LPE_KEYS_MORTON  1 /* Z-order: child digit = (x >= mid) + 2*(y >= mid), the reference's nw,ne,sw,se recursion order */
#define LPE_KEYS_HILBERT 2 /* Hilbert index of the same depth-D cell */

typedef struct lpe_bh_ctx lpe_bh_ctx;

typedef struct {
    double universe_size;        /* SharedSystemConfig::UniverseSizeMeters: root square is [0,U)^2 (barnes_hut.cpp:110-112) */
    double softening;            /* SharedSystemConfig::GravitationalSoftener (barnes_hut.cpp:261) */
    double theta;                /* BarnesHutConfig::theta (barnes_hut.hpp:36) */
    double small_mass_threshold; /* BarnesHutConfig::smallMassThreshold, 0 disables (barnes_hut.hpp:45) */
    double G;                    /* SimulatorConstants::RealG = 6.674e-11 (constants.cpp:8) */
    double dt_kick;              /* SecondsPerTick*baseTimeAcceleration*timeScale (barnes_hut.cpp:284) */
    double dt_drift;             /* SecondsPerTick*TimeAcceleration (movement.cpp:17); used when do_drift != 0 */
    int32_t quirk_mode;          /* 1 (reference): every internal node counts its first occupant twice (barnes_hut.cpp:157-177); 0: textbook tree */
    int32_t precision;           /* LPE_PREC_* */
    int32_t do_drift;            /* 0: BarnesHutSystem only (velocity kick); 1: MovementSystem fused into the same step */
    int32_t max_depth;           /* 0: automatic (softening bound, SURVEY.md Q4, capped at 30); else forced key depth 1..30 */
    int32_t key_order;           /* LPE_KEYS_*: space-filling curve of the sort keys (same cells, same tree, same results) */
    int32_t reserved;
} lpe_bh_params;

typedef struct {
    uint64_t n_bodies;       /* bodies uploaded */
    uint64_t n_in_tree;      /* sources inside [0,U)^2 */
    uint64_t n_terminals;    /* distinct depth-D cells (leaves + aggregated cells at the depth bound) */
    uint64_t n_nodes;        /* nodes of the path-compressed tree in pre-order (terminals + branching cells) */
    uint64_t interactions;   /* accepted node interactions of the last step (only counted when stats are enabled) */
    uint64_t visits;         /* node visits (per lane) of the last step (only when stats are enabled) */
    uint64_t warp_visits;    /* list entries / node visits per warp, summed over warps (only when stats are enabled) */
    uint64_t overflow_chunks;/* 32-body chunks the two-phase kernel handed to the depth-first kernel in the last step */
    uint64_t t2_kinds[8];    /* two-phase diagnostics (stats only): A-clean, A-dirty, O-dirty, M->all accept, M->all open, M->split, rounds, frontier nodes */
    int32_t  depth;          /* key depth D used by the last step */
    int32_t  sort_passes;
    int32_t  hilbert;        /* 1 if the last step sorted by Hilbert index, 0 for Morton code */
    int32_t  pad_;
    float ms_keygen, ms_sort, ms_build, ms_traverse, ms_total; /* last step, CUDA events; only when timing is enabled */
    float pad2_;
    /* DebugStats::updateForce (reference include/core/debug.hpp:37-41, called per accepted node at barnes_hut.cpp:278):
     * max and sum of force = G*M*m/distSq over the accepted interactions of the last step (count = interactions); only
     * when stats are enabled */
    double force_max, force_sum;
} lpe_bh_stats;

/* tree dump for parity tests; every pointer may be NULL. Arrays are sized by the caller from lpe_bh_get_stats. */
typedef struct {
    uint64_t* sorted_keys;   /* [n_bodies] sort keys (Morton code or Hilbert index, see stats.hilbert) in sorted order (bit 2D set = not in tree) */
    uint32_t* sorted_index;  /* [n_bodies] creation index of the body at each sorted position */
    int32_t*  node_level;    /* [n_nodes] level of a branching cell; -1 single-body leaf; -2 aggregated cell at the depth bound */
    uint64_t* node_key;      /* [n_nodes] sort key (depth D) of the first body of the node */
    uint32_t* node_skip;     /* [n_nodes] pre-order index of the first node after this node's subtree */
    uint32_t* node_first;    /* [n_nodes] creation index of the node's first occupant (minimum insertion rank) */
    uint32_t* node_count;    /* [n_nodes] bodies under the node */
    double*   node_mass;     /* [n_nodes] node mass as the traversal sees it (includes the quirk when enabled) */
    double*   node_comx;     /* [n_nodes] */
    double*   node_comy;     /* [n_nodes] */
} lpe_bh_tree_dump;

/* device-resident views for zero-copy interop (torch / NCCL plumbing). The pointers are valid until the next
 * upload, STEP or destroy: every step re-orders the state into this step's key order in a second set of buffers and
 * swaps the two sets. Slot i of body / vel holds the body whose creation index is orig[i] when key_ordered != 0;
 * right after an upload (key_ordered == 0) slot i is body i and orig is not meaningful. */
typedef struct {
    void* body;       /* {double x, y, m; uint32 rank, comp}[n], 32 B per body */
    void* vel;        /* double2[n] */
    void* orig;       /* uint32[n]: creation index of the body in each slot (key_ordered != 0) */
    void* xchg_send;  /* double4[xchg_chunk]  this rank's packed slice (x,y,vx,vy), see lpe_bh_set_shard */
    void* xchg_recv;  /* double4[xchg_chunk * nranks] */
    uint64_t n;
    uint64_t xchg_chunk; /* elements per rank in the exchange buffers */
    int32_t key_ordered; /* 0: creation order (no step since the upload); 1: key order, see orig */
    int32_t pad_;
} lpe_bh_device_view;

const char* lpe_bh_version(void);
int  lpe_bh_device_count(void);

int  lpe_bh_create(int device, lpe_bh_ctx** out);
void lpe_bh_destroy(lpe_bh_ctx* ctx);
const char* lpe_bh_last_error(const lpe_bh_ctx* ctx); /* ctx may be NULL: last creation error */

/* Run all work of this context on an existing CUDA stream (cudaStream_t as void*); NULL restores the context's own stream. */
int  lpe_bh_set_stream(lpe_bh_ctx* ctx, void* cuda_stream);
/* flags: bit0 = per-phase CUDA-event timing, bit1 = count interactions/visits (slower; parity tests only),
 *        bit2 = FAST precision uses the depth-first kernel instead of the two-phase kernel (A/B testing),
 *        bit3 = the two-phase kernel hands every chunk to its overflow path (tests of that path),
 *        bit4 = plain launches: never replay a captured CUDA graph of the step (A/B testing; see lpe_bh_graph_replays),
 *        bit5 = the two-phase kernel walks per warp only, without the far field shared by the warps of a CTA (A/B testing) */
int  lpe_bh_set_instrumentation(lpe_bh_ctx* ctx, int flags);

/* Stage bodies into device SoA buffers. rank[i] = position of body i in the iteration of
 * view<Position,Mass>(exclude<Boundary>) (the reference's insertion order); NULL = EnTT's default, newest
 * entity first. comp NULL = every body has Mass and Velocity. vx/vy NULL = zero. Asynchronous on the context's
 * stream when the host arrays are pinned. At most 2^28 bodies per context (32-bit record slots; about what fits
 * in 180 GB of HBM at ~450 B/body). */
int  lpe_bh_upload(lpe_bh_ctx* ctx, uint64_t n, const double* x, const double* y, const double* vx,
                   const double* vy, const double* m, const uint32_t* rank, const uint8_t* comp);
/* make the bodies ON THE DEVICE instead of uploading them — kind 4 = the Keplerian-disk scenario's
 * entity law (reference src/scenarios/keplerian_disk.cpp:45-146) with one independent random stream per body, so the
 * same (kind, n, seed) always gives the same bodies and lpe_bh_workload(4, ...) restates them on the host to libm
 * rounding. Bodies are in creation order with EnTT's default insertion ranks, all with Mass and Velocity. Asynchronous. */
int  lpe_bh_generate(lpe_bh_ctx* ctx, int kind, uint64_t n, uint64_t seed, double universe_size);
/* Positions only (e.g. after other ECS systems moved bodies); n must match the last upload. */
int  lpe_bh_upload_positions(lpe_bh_ctx* ctx, const double* x, const double* y);
int  lpe_bh_upload_velocities(lpe_bh_ctx* ctx, const double* vx, const double* vy);

/* nsteps x { keygen, sort, tree build + aggregation, traversal, kick [, drift] } on device-resident state. Asynchronous. */
int  lpe_bh_step(lpe_bh_ctx* ctx, const lpe_bh_params* p, int nsteps);

/* Copy state back (synchronises). Any pointer may be NULL. */
int  lpe_bh_download(lpe_bh_ctx* ctx, double* x, double* y, double* vx, double* vy);
int  lpe_bh_synchronize(lpe_bh_ctx* ctx);

/* One-call host round trip used by the ECS drop-in: upload, one step, download. x,y,vx,vy are updated in place. */
int  lpe_bh_update_host(lpe_bh_ctx* ctx, const lpe_bh_params* p, uint64_t n, double* x, double* y, double* vx,
                        double* vy, const double* m, const uint32_t* rank, const uint8_t* comp);

/* The same tick for callers that keep {x, y} records (EnTT's Position / Velocity pools: reference
 * include/math/vector_math.hpp:46-49,120-123): pos and vel are 2n doubles each, updated in place. */
int  lpe_bh_update_host_aos(lpe_bh_ctx* ctx, const lpe_bh_params* p, uint64_t n, double* pos, double* vel,
                            const double* m, const uint32_t* rank, const uint8_t* comp);

/* The same tick in three calls for a caller that first has to gather its components (the ECS drop-in): each call
 * queues the device work its array unlocks and returns, so the next array is gathered while the GPU runs.
 *   lpe_bh_tick_begin   positions ({x,y} records) [+ component masks]  -> keys, sort          (asynchronous)
 *   lpe_bh_tick_mass    masses [+ insertion ranks]                       -> gather, tree build and, in FAST precision,
 *                       the tree walk, which does not need the velocities (asynchronous)
 *   lpe_bh_tick_finish  velocities ({vx,vy} records, updated in place)   -> kick [+ drift: pos updated in place when
 *                       do_drift] (STRICT precision: tree walk + kick); synchronises
 * Calling lpe_bh_tick_begin again before lpe_bh_tick_finish abandons the tick that was under way.
 * Host arrays should be page-locked (lpe_bh_alloc_pinned: the first two calls then replay captured CUDA graphs) and must
 * stay untouched until lpe_bh_tick_finish returns. */
int  lpe_bh_tick_begin(lpe_bh_ctx* ctx, const lpe_bh_params* p, uint64_t n, const double* pos, const uint8_t* comp);
int  lpe_bh_tick_mass(lpe_bh_ctx* ctx, const double* m, const uint32_t* rank);
int  lpe_bh_tick_finish(lpe_bh_ctx* ctx, double* pos, double* vel);

int  lpe_bh_get_stats(lpe_bh_ctx* ctx, lpe_bh_stats* out);          /* synchronises */
int  lpe_bh_dump_tree(lpe_bh_ctx* ctx, lpe_bh_tree_dump* out);      /* tree of the last step; synchronises */
/* per-body accepted / visited counts of the last step in creation order (instrumentation bit1 must be on) */
int  lpe_bh_get_counts(lpe_bh_ctx* ctx, uint32_t* accepted, uint32_t* visited);

/* Largest mass among the resident bodies that have Mass and are not Boundary: the answer to the reference's two
 * per-tick host scans over Mass (barnes_hut.cpp:55-71 early exit when it is below smallMassThreshold; gravity.cpp:41-49
 * uniform field off when it reaches 1e10), reduced on the device at upload. Synchronises. */
int  lpe_bh_max_source_mass(lpe_bh_ctx* ctx, double* max_mass);

/* Direct O(N^2) sum with the same force law, for the accuracy cross-check: accelerations of bodies
 * [first, first+count) in creation order, fp64, from device-resident state. Synchronises. */
int  lpe_bh_direct_accel(lpe_bh_ctx* ctx, const lpe_bh_params* p, uint64_t first, uint64_t count, double* ax,
                         double* ay);

/* ---- multi-GPU (one context per process/GPU; the collective itself is the caller's, e.g. NCCL allgather) ----
 * Every rank holds all bodies and builds the same tree; rank r traverses and integrates the sorted-order blocks
 * b with b % nranks == r (blocks of LPE_SHARD_BLOCK key-consecutive bodies), packs their new (x,y,vx,vy) into
 * xchg_send, and after the caller's allgather into xchg_recv, lpe_bh_step_finish scatters every rank's slice
 * back into the state arrays. nranks == 1 restores the single-GPU path. */
#define LPE_SHARD_BLOCK 2048u
int  lpe_bh_set_shard(lpe_bh_ctx* ctx, int rank, int nranks);
int  lpe_bh_step_begin(lpe_bh_ctx* ctx, const lpe_bh_params* p);   /* build + own-slice traversal -> xchg_send */
int  lpe_bh_step_finish(lpe_bh_ctx* ctx);                          /* xchg_recv -> state */
int  lpe_bh_get_device_view(lpe_bh_ctx* ctx, lpe_bh_device_view* out);
/* host-staged exchange (tests, or a transport without device pointers): copy this rank's packed slice out /
 * another rank's slice in; each is 4*xchg_chunk doubles. Synchronises. */
int  lpe_bh_xchg_read_send(lpe_bh_ctx* ctx, double* host);
int  lpe_bh_xchg_write_recv(lpe_bh_ctx* ctx, int src_rank, const double* host);
/* ---- BoundarySystem as a device pass over the resident state -------------------------------
 * Replaces Systems::BoundarySystem::update (reference src/systems/boundary.cpp:13-69, config
 * include/systems/boundary.hpp:27-36) for resident runs, where it is the step right before BarnesHutSystem each
 * tick (sim.cpp system order): bodies with LPE_HAS_VELOCITY and without LPE_ASLEEP are clamped to
 * [margin, U - margin], the velocity component is reflected with damping, and after a bounce the speed is capped
 * at max_speed. fp64, operation for operation as the reference (bit-exact against it). Asynchronous on the
 * context's stream. */
typedef struct lpe_bh_boundary_params {
    double universe_size;    /* SharedSystemConfig::UniverseSizeMeters */
    double margin;           /* BoundaryConfig::marginPixels * SharedSystemConfig::MetersPerPixel (metres) */
    double bounce_damping;   /* BoundaryConfig::bounceDamping (default 0.7) */
    double max_speed;        /* BoundaryConfig::maxSpeed (default 1.0) */
} lpe_bh_boundary_params;
int  lpe_bh_boundary(lpe_bh_ctx* ctx, const lpe_bh_boundary_params* p);

/* Direct exchange over peer memory (NVLink / NVSwitch), 2..8 ranks of one node: once every rank's receive buffer is
 * known to this context, the traversal kernel itself stores each new (x,y,vx,vy) into ALL ranks' receive buffers
 * (the exchange overlaps the force computation) and the caller only needs a barrier between lpe_bh_step_begin and
 * lpe_bh_step_finish instead of the allgather. One process per GPU: lpe_bh_xchg_export gives a 64-byte CUDA IPC
 * handle of this rank's receive buffer (and registers the own buffer); ship it to the other ranks by any channel
 * and lpe_bh_xchg_import it there. Same process (tests): lpe_bh_xchg_set_peer with the raw device pointer from
 * lpe_bh_get_device_view. The receive buffer holds two generations (step parity), so one barrier per step is
 * enough. A reallocation (different n) drops the peer table: exchange handles again. */
int  lpe_bh_xchg_export(lpe_bh_ctx* ctx, void* handle64);
int  lpe_bh_xchg_import(lpe_bh_ctx* ctx, int rank, const void* handle64);
int  lpe_bh_xchg_set_peer(lpe_bh_ctx* ctx, int rank, void* recv_device_ptr);
int  lpe_bh_xchg_p2p_ready(const lpe_bh_ctx* ctx);
/* forget every peer buffer (closes the IPC mappings): back to the collective exchange, e.g. when not every rank
 * could open every handle */
int  lpe_bh_xchg_reset(lpe_bh_ctx* ctx);
/* pure host helper (no GPU): which rank owns sorted position i, and where it sits in that rank's packed slice */
int  lpe_bh_shard_owner(uint64_t sorted_pos, int nranks, int* rank_out, uint64_t* slot_out);
uint64_t lpe_bh_shard_chunk(uint64_t n_bodies, int nranks);        /* elements per rank in the exchange buffers */

/* ---- multi-GPU, domain-decomposed (contiguous key ranges per GPU, per-GPU sort + build, one exchange
 * of tree nodes) ------------------------------------------------------------------------------------------------
 * Nothing is replicated: rank r owns the bodies whose sort key lies in [K_r, K_r+1) (splitters = arbitrary keys),
 * sorts and builds only those, publishes the roots of its part of the tree to every rank and stores into rank d's
 * record array the child blocks of the cells that a body of d's key range could open (a conservative box test: the
 * locally essential tree). Every rank then builds the few cells that straddle a splitter from the published roots
 * — same child order, same sums as the single-GPU build — and traverses for its own bodies: per-body accept / open
 * decisions are those of the single-GPU tree (= the reference's). Bodies that leave a key range are stored straight
 * into their new owner's state arrays. All traffic is plain stores to NVLink peer memory inside the step's own
 * kernels plus two in-stream flag barriers per step; there is no collective call. FAST precision only.
 *
 * Setup (every rank, same capacity / import_blocks / nranks):
 *   lpe_bh_dd_init -> exchange windows (same process: lpe_bh_dd_window + lpe_bh_dd_set_peer; one process per GPU:
 *   lpe_bh_dd_export + lpe_bh_dd_import, handles shipped by any channel) -> lpe_bh_dd_upload (the WHOLE input on every
 *   rank; each keeps its share) -> a host barrier across the ranks -> lpe_bh_dd_step in lockstep.
 * lpe_bh_dd_phase runs one of the step's three phases without relying on concurrently running peers (tests that play
 * several ranks on one GPU: phase p on every context, synchronise all, next phase). */
typedef struct {
    uint64_t capacity;        /* body slots of this rank */
    uint64_t n_live;          /* bodies this rank owns after the last step */
    uint64_t n_in_tree, n_terminals, n_cells;   /* of the rank's own part of the tree */
    uint64_t n_roots;         /* roots published by all ranks (leaves of the top of the tree) */
    uint64_t exported_blocks[8]; /* child blocks (128 B records + 128 B fp64 side records) stored into each rank */
    uint64_t interactions;    /* accepted interactions of the rank's targets (stats only) */
    uint64_t work_cost;       /* list entries evaluated by the rank's traversal */
    uint64_t overflow_chunks;
    uint32_t import_blocks;   /* capacity of one sender's import region */
    uint32_t fault;           /* device fault bits since the last check */
    int32_t  rank, nranks, depth;
    int32_t  export_rounds;   /* generations of the exporter's breadth-first walk (its critical path) */
    /* last step, CUDA events, only when timing is enabled: keys + migration | wait for every rank's migrants |
     * inbox keys + sort | build | flags + publish + export | wait for every rank's export | top of the tree | traversal */
    float ms_keygen, ms_wait_a, ms_sort, ms_build, ms_export, ms_wait_b, ms_top, ms_traverse, ms_total, pad2_;
} lpe_bh_dd_stats;

int  lpe_bh_dd_init(lpe_bh_ctx* ctx, int rank, int nranks, uint64_t capacity, uint32_t import_blocks /* 0 = default */);
int  lpe_bh_dd_export(lpe_bh_ctx* ctx, void* handle64);                    /* CUDA IPC handle of this rank's window */
int  lpe_bh_dd_import(lpe_bh_ctx* ctx, int rank, const void* handle64);
void* lpe_bh_dd_window(lpe_bh_ctx* ctx);                                   /* same-process form: raw device pointer */
int  lpe_bh_dd_set_peer(lpe_bh_ctx* ctx, int rank, void* window, int peer_device /* -1: same device */);
int  lpe_bh_dd_ready(const lpe_bh_ctx* ctx);
int  lpe_bh_dd_upload(lpe_bh_ctx* ctx, const lpe_bh_params* p, uint64_t n, const double* x, const double* y,
                      const double* vx, const double* vy, const double* m, const uint32_t* rank, const uint8_t* comp);
int  lpe_bh_dd_step(lpe_bh_ctx* ctx, const lpe_bh_params* p, int nsteps);  /* all ranks in lockstep; asynchronous */
int  lpe_bh_dd_phase(lpe_bh_ctx* ctx, const lpe_bh_params* p, int phase);  /* 0, 1, 2 */
/* this rank's bodies with their creation indices; arrays of >= capacity elements, any may be NULL; synchronises */
int  lpe_bh_dd_download(lpe_bh_ctx* ctx, uint64_t* n_out, uint32_t* index, double* x, double* y, double* vx,
                        double* vy, uint32_t* accepted);
int  lpe_bh_dd_get_stats(lpe_bh_ctx* ctx, lpe_bh_dd_stats* out);
/* load-balance input: per 32-body chunk of this rank's sorted bodies, the depth-30 key of its first body and the list
 * entries its warp evaluated in the last traversal; arrays of >= capacity / 32 + 1 elements; synchronises */
int  lpe_bh_dd_chunk_costs(lpe_bh_ctx* ctx, uint64_t* n_chunks, uint64_t* first_key30, uint32_t* cost);
/* splitters as depth-30 keys, nranks + 1 values (first 0, last 2^60); set: same values on every rank before the same step */
int  lpe_bh_dd_get_splitters(lpe_bh_ctx* ctx, uint64_t* split30);
int  lpe_bh_dd_set_splitters(lpe_bh_ctx* ctx, const uint64_t* split30);
/* pure host helpers (no GPU): the sort key of cell (ix, iy) at a level and its inverse */
uint64_t lpe_bh_cell_key(uint32_t ix, uint32_t iy, int level, int hilbert);
void lpe_bh_key_cell(uint64_t key, int level, int hilbert, uint32_t* ix, uint32_t* iy);

/* cumulative number of this library's kernels launched by the context (bench.py's gpu_launches; kernels inside a replayed
 * CUDA graph count like plain launches) */
uint64_t lpe_bh_launch_count(const lpe_bh_ctx* ctx);
/* Whole steps are captured as CUDA graphs and replayed (lpe_bh_step, lpe_bh_update_host(_aos) and lpe_bh_tick_* with
 * page-locked host buffers): one submission per step instead of ~30. A step is captured the second time the same
 * (parameters, body count, buffers) come up; timing instrumentation, sharded and decomposed contexts, and the
 * environment variable LPE_BH_GRAPHS=0 keep plain launches. Returns how many steps were replays so far. */
uint64_t lpe_bh_graph_replays(const lpe_bh_ctx* ctx);
/* FP32 FMA peak of the device by a register-resident FMA loop, TFLOP/s (2 flops per FMA): the roofline denominator
 * of the traversal, which is FP32-pipe bound, not HBM or tensor bound. Synchronises. */
int  lpe_bh_fma_peak(lpe_bh_ctx* ctx, double* tflops);

/* page-locked host memory for staging buffers (full PCIe rate for upload/download) */
void* lpe_bh_alloc_pinned(uint64_t bytes);
void  lpe_bh_free_pinned(void* p);

/* ---- deterministic synthetic workloads (host, std::mt19937_64, u=(g()>>11)*2^-53) ----
 * kind: 0 = uniform disk (C2), 1 = Plummer sphere projected (C3), 2 = two-galaxy collision (C4),
 *       3 = Keplerian disk with the law of reference src/scenarios/keplerian_disk.cpp:78-147 (C1 stand-in)
 *       4 = the same law, counter-based: one random stream per body (what lpe_bh_generate makes on the device) */
int  lpe_bh_workload(int kind, uint64_t n, uint64_t seed, double universe_size, double* x, double* y,
                     double* vx, double* vy, double* m);

#ifdef __cplusplus
}
#endif
#endif
