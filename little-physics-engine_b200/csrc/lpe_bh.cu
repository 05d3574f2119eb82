// lpe_bh.cu — host side of the C ABI declared in include/lpe_bh.h: device buffers, staging, and the launch
// sequence of one Barnes-Hut step. No CPU fallback: every entry point needs a CUDA device.
//
// One step, no host synchronisation (counts live in a device `Scal` block):
//   keygen -> radix sort (u64 key, u32 index: histogram, bases, one look-back scatter per pass) ->
//   [side stream: gather into key order] | head-flag scan that writes the terminals -> witnesses ->
//   mask-popcount scan -> level offsets -> topology -> aggregation level by level -> traverse + kick (+ drift),
//   then the depth-first kernel for the (normally zero) chunks whose frontier overflowed.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/lpe_bh.h"
#include "bh_common.cuh"
#include "bh_sort.cuh"
#include "bh_build.cuh"
#include "bh_traverse.cuh"
#include "bh_traverse2.cuh"
#include "bh_dd.cuh"
#include "kepler_gen.h"

using namespace lpe;

// Record slots and child codes are 32-bit (4 slots per cell, one flag bit): 2^28 bodies per context keeps every index
// in range. At ~450 B/body that is also about what fits in 180 GB of HBM next to the sort buffers.
static constexpr uint64_t LPE_MAX_BODIES = 1ull << 28;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

}  // namespace

struct lpe_bh_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    int instr = 0;
    bool force_dfs = false;  // instrumentation bit2: use the depth-first kernel in FAST mode too
    bool force_overflow = false;  // instrumentation bit3: the two-phase kernel hands EVERY chunk to the overflow path

    uint64_t n = 0, cap = 0;
    uint64_t launches = 0;
    int shard_rank = 0, shard_n = 1;
    uint64_t xchg_chunk = 0;

    // state
    Body* body = nullptr;
    double2* vel = nullptr;
    // every step's gather re-orders the state into key order (second set of buffers); orig[slot] is then the creation
    // index of the body in that slot (orig_valid = false: the state is in creation order, right after an upload)
    Body* body2 = nullptr;
    double4* stage4 = nullptr;    // {x, y, vx, vy} records in creation order on their way to the host (k_finish_tick, k_stage_state)
    double2* vel2 = nullptr;
    unsigned int *orig = nullptr, *orig2 = nullptr;
    bool orig_valid = false;
    unsigned int* rank_in = nullptr;   // staging of the caller's rank / component arrays
    unsigned char* comp_in = nullptr;
    // staging
    double* tmp = nullptr;  // 5*cap doubles
    // lpe_bh_update_host: uploads run on a second stream, the step waits for each array only where it first reads it
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t evc[4] = {nullptr, nullptr, nullptr, nullptr};
    // the gather into key order runs beside the terminal / witness / scan kernels (they only need the sorted keys)
    cudaStream_t side_stream = nullptr;
    cudaEvent_t evs[2] = {nullptr, nullptr};
    bool pend_mass = false, pend_vel = false, pend_rank = false, pend_vel_aos = false;
    bool tick_has_comp = false;
    bool capturing = false;       // the calls being made are recorded into a CUDA graph (run_graphed)
    bool tracing = false;
    cudaEvent_t trace_ev[7] = {};
    bool defer_kick = false;      // host tick, FAST precision: the traversal stores velocity changes, k_finish_tick applies them
    int tick_stage = 0;           // lpe_bh_tick_begin / _mass / _finish: which call comes next
    StepConst tick_k{};
    lpe_bh_params tick_p{};
    // sort
    unsigned long long* keys[2] = {nullptr, nullptr};
    unsigned int* vals[2] = {nullptr, nullptr};
    unsigned int* totals = nullptr;            // per step: digit histograms / bases [8][512], tile counters [8], fault flag
    unsigned long long* lbstatus = nullptr;    // look-back status words of the sort passes (epoch-tagged, never cleared)
    unsigned int epoch = 0;                    // host mirror of *epoch_dev (0 = the status words have to be cleared first)
    unsigned int* epoch_dev = nullptr;
    unsigned int* fault_host = nullptr;        // pinned copy of the sort's fault flag
    int sorted_sel = 0;
    unsigned int *selfslot = nullptr, *ovf_list = nullptr;
    // scans
    unsigned int* P = nullptr;
    // terminals
    unsigned long long* tkey = nullptr;
    unsigned int *tfirst = nullptr, *mask = nullptr, *wstart = nullptr;
    signed char* delta = nullptr;
    // nodes (pre-order index), cells (ordinal), child blocks
    uint64_t node_cap = 0;
    unsigned int *child = nullptr, *levelMeta = nullptr;
    uint2* levelList = nullptr;
    Agg* agg = nullptr;
    TravRec* rec = nullptr;
    // stats
    unsigned int *cntAcc = nullptr, *cntVis = nullptr;
    Scal* scal = nullptr;
    // exchange
    double4 *xchg_send = nullptr, *xchg_recv = nullptr;   // recv holds two generations (step parity) of nranks slices
    double4* peer_recv[LPE_MAX_P2P] = {};                 // every rank's xchg_recv (own pointer, or opened through CUDA IPC)
    void* peer_opened[LPE_MAX_P2P] = {};                  // IPC mappings to close
    int xchg_parity = 0;
    int sms = 148;
    // domain decomposition (lpe_bh_dd.inl)
    bool dd = false;
    int dd_rank = 0, dd_R = 1;
    unsigned int dd_icap = 0;                 // import blocks per sender
    char* dd_win = nullptr;                   // own window (peer-visible): header, root tables, state x2, records
    size_t dd_win_bytes = 0;
    char* dd_peer_win[LPE_MAX_P2P] = {};      // every rank's window (own pointer, raw peer pointer, or an IPC mapping)
    void* dd_peer_opened[LPE_MAX_P2P] = {};   // IPC mappings to close
    int dd_cur = 0;                           // which of the two state buffer sets of the window is current
    unsigned long long dd_epoch = 0;          // step number: the value the barrier flags carry (host mirror of *dd_step_dev)
    unsigned long long* dd_step_dev = nullptr;
    unsigned long long dd_split30[LPE_MAX_P2P + 1] = {};   // splitters as depth-30 keys
    int dd_hilbert = -1;                      // key order the splitters were made for
    int dd_dom_depth = -1;                    // depth the device copy of the domain table was built for
    uint64_t dd_n_total = 0;                  // bodies of the whole input
    double dd_U = 0.0;
    DDDomain* dd_dom = nullptr;
    std::vector<char> dd_dom_host;
    DDRoot* dd_myroots = nullptr;             // roots of this rank's inner quadrants (k_dd_roots -> k_dd_export)
    unsigned int* dd_queue = nullptr;         // [dest][icap] breadth-first queues of the exporter
    unsigned long long* dd_oob = nullptr;     // ordered-encoded box of the out-of-tree targets seen by phase A
    unsigned int* dd_pushed = nullptr;        // round flags of the exporter (behind dd_oob)
    double* dd_payload = nullptr;             // 2 x 6 doubles: this step's mail of the two barrier points
    double4* dd_xrec = nullptr;
    unsigned int* dd_chunk_cost = nullptr;
    DDTop dd_top{};
    cudaEvent_t dd_ev[10] = {};
    // timing
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    lpe_bh_stats last{};
    StepConst last_c{};
    bool have_step = false;
    std::vector<void*> allocs;
    // whole steps as CUDA graphs (resident steps and host ticks): see run_graphed
    struct GraphEntry {
        std::string key;
        cudaGraphExec_t exec = nullptr;
        // what queueing the step does to the context on the host side, re-applied on every replay
        bool swapped = false;
        bool orig_valid = false;
        int sorted_sel = 0;
        uint64_t launches = 0;
        unsigned int epochs = 0;
        bool dd_flip = false;
        unsigned long long dd_steps = 0;
        lpe_bh_stats last{};
        StepConst last_c{};
    };
    std::vector<GraphEntry> graphs;
    std::vector<std::string> graph_seen;     // keys met once: a step is captured the second time its key comes up
    int use_graphs = -1;                     // -1: not decided yet (LPE_BH_GRAPHS=0 turns them off)
    uint64_t graph_replays = 0;
};

namespace {

#define CU_TRY(ctx, expr)                                                                       \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                    \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

// Every entry point runs on the context's device and puts the caller's current device back afterwards (the ECS
// drop-in lives inside a host application that may use other devices).
struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DevGuard(const DevGuard&) = delete;
    DevGuard& operator=(const DevGuard&) = delete;
};

int fail(lpe_bh_ctx* c, const std::string& m) {
    if (c) c->err = m; else g_create_error = m;
    return 1;
}

template <class T>
int dalloc(lpe_bh_ctx* c, T*& p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 16);
    if (e != cudaSuccess) {
        c->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
        return 1;
    }
    c->allocs.push_back(q);
    p = static_cast<T*>(q);
    return 0;
}

void drop_graphs(lpe_bh_ctx* c) {   // (they hold the addresses of the buffers they were captured with)
    for (auto& g : c->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    c->graphs.clear();
    c->graph_seen.clear();
}

void free_all(lpe_bh_ctx* c) {
    drop_graphs(c);
    for (void* p : c->allocs) cudaFree(p);
    c->allocs.clear();
    c->cap = 0;
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

void close_peers(lpe_bh_ctx* c);

void dd_release(lpe_bh_ctx* c);
int dd_check_fault(lpe_bh_ctx* c);

// dd: the context becomes one rank of a domain-decomposed run with n slots; the state arrays and the records then live
// in the rank's peer-visible window (lpe_bh_dd.inl) instead of being allocated here
int ensure_capacity(lpe_bh_ctx* c, uint64_t n, bool dd = false) {
    if (!dd && !c->dd && n <= c->cap && c->cap != 0) return 0;
    cudaStreamSynchronize(c->stream);
    if (c->side_stream) cudaStreamSynchronize(c->side_stream);   // (a step that failed half-way may not have joined them)
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    free_all(c);
    dd_release(c);
    const uint64_t cap = n < 1024 ? 1024 : n;
    const uint64_t ncap = 2 * cap + 8;
    const uint64_t recSlots = 4 * (cap + 8 + (dd ? DD_TOPCAP : 0));
    const int sortTiles = cdiv((long long)cap, SORT_TILE);
    const int scanTiles = cdiv((long long)cap + 1, SCAN_TILE);
    int rc = 0;
    rc |= dalloc(c, c->rank_in, cap) | dalloc(c, c->comp_in, cap) | dalloc(c, c->tmp, 5 * cap);
    if (!dd)
        rc |= dalloc(c, c->body, cap) | dalloc(c, c->vel, cap) | dalloc(c, c->body2, cap) | dalloc(c, c->vel2, cap) | dalloc(c, c->stage4, cap) |
              dalloc(c, c->orig, cap) | dalloc(c, c->orig2, cap) | dalloc(c, c->rec, recSlots);
    rc |= dalloc(c, c->keys[0], cap) | dalloc(c, c->keys[1], cap) | dalloc(c, c->vals[0], cap) |
          dalloc(c, c->vals[1], cap) | dalloc(c, c->lbstatus, (size_t)sortTiles * (256 * (SORT_MAX_PASSES - 1) + 512) + 2 * ((size_t)scanTiles + 2)) | dalloc(c, c->totals, 512 * 8 + 16) | dalloc(c, c->epoch_dev, 4);
    rc |= dalloc(c, c->selfslot, cap) |
          dalloc(c, c->ovf_list, (size_t)cdiv((long long)cap, LPE_SHARD_BLOCK) * (LPE_SHARD_BLOCK / 32) + 8);
    rc |= dalloc(c, c->P, cap + 2);
    rc |= dalloc(c, c->tkey, cap + 2) | dalloc(c, c->tfirst, cap + 2) | dalloc(c, c->mask, cap + 2) |
          dalloc(c, c->wstart, cap + 2) | dalloc(c, c->delta, cap + 2);
    rc |= dalloc(c, c->child, 4 * (cap + 8)) | dalloc(c, c->levelList, cap + 8) |
          dalloc(c, c->levelMeta, 3 * 32) | dalloc(c, c->agg, ncap);
    rc |= dalloc(c, c->cntAcc, cap) | dalloc(c, c->cntVis, cap) | dalloc(c, c->scal, 1);
    if (rc) {
        free_all(c);
        return 1;
    }
    c->cap = cap;
    c->node_cap = ncap;
    c->epoch = 0u;   // new status words and a new device epoch: cleared before the next step (epoch_prepare)
    c->xchg_send = c->xchg_recv = nullptr;
    c->xchg_chunk = 0;
    close_peers(c);
    c->orig_valid = false;
    return 0;
}

void close_peers(lpe_bh_ctx* c) {
    for (int r = 0; r < LPE_MAX_P2P; ++r) {
        if (c->peer_opened[r]) cudaIpcCloseMemHandle(c->peer_opened[r]);
        c->peer_opened[r] = nullptr;
        c->peer_recv[r] = nullptr;
    }
    c->xchg_parity = 0;
}
bool p2p_ready(const lpe_bh_ctx* c) {
    if (c->shard_n <= 1 || c->shard_n > LPE_MAX_P2P || !c->xchg_recv) return false;
    for (int r = 0; r < c->shard_n; ++r)
        if (!c->peer_recv[r]) return false;
    return true;
}

int ensure_xchg(lpe_bh_ctx* c) {
    const uint64_t chunk = lpe_bh_shard_chunk(c->n, c->shard_n);
    if (c->xchg_send && c->xchg_chunk == chunk) return 0;
    // (old exchange buffers, if any, stay in the allocation list and are released with the context)
    if (dalloc(c, c->xchg_send, chunk) || dalloc(c, c->xchg_recv, 2 * chunk * (uint64_t)c->shard_n)) return 1;
    c->xchg_chunk = chunk;
    close_peers(c);   // the receive buffer moved: peers must exchange handles again
    return 0;
}

// host arrays are always in creation order; `orig` (null = identity) maps a state slot to its creation index
__global__ void k_pack2(int n, const double* __restrict__ a, const double* __restrict__ b, double2* __restrict__ out,
                        const unsigned int* __restrict__ orig) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned int s = orig ? orig[i] : (unsigned int)i;
        out[i] = make_double2(a ? a[s] : 0.0, b ? b[s] : 0.0);
    }
}
__global__ void k_unpack2(int n, const double2* __restrict__ in, double* __restrict__ a, double* __restrict__ b,
                          const unsigned int* __restrict__ orig) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double2 v = in[i];
        const unsigned int s = orig ? orig[i] : (unsigned int)i;
        a[s] = v.x;
        b[s] = v.y;
    }
}
__global__ void k_unpermute_u32(int n, const unsigned int* __restrict__ in, unsigned int* __restrict__ out,
                                const unsigned int* __restrict__ orig) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[orig[i]] = in[i];
}
// x, y, m (+ optional rank / component arrays) -> one 32-byte Body per entity
__global__ void k_pack_body(int n, const double* __restrict__ x, const double* __restrict__ y,
                            const double* __restrict__ m, const unsigned int* __restrict__ rank,
                            const unsigned char* __restrict__ comp, Body* __restrict__ body, Scal* __restrict__ s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    Body b;
    b.m = 0.0; b.comp = 0u;
    if (i < n) {
        b.x = x[i]; b.y = y[i]; b.m = m[i];
        // EnTT iterates the leading pool back to front: newest entity is inserted first
        b.rank = rank ? rank[i] : (unsigned int)(n - 1 - i);
        b.comp = comp ? (unsigned int)comp[i] : (unsigned int)(LPE_HAS_MASS | LPE_HAS_VELOCITY);
        body[i] = b;
    }
    block_max_mass(b.m, b.comp, s);
}
// the two halves of k_pack_body, for the pipelined host path: positions + components first, masses + ranks later
__global__ void k_pack_pos(int n, const double* __restrict__ x, const double* __restrict__ y,
                           const unsigned char* __restrict__ comp, Body* __restrict__ body) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    *reinterpret_cast<double2*>(&body[i].x) = make_double2(x[i], y[i]);
    body[i].comp = comp ? (unsigned int)comp[i] : (unsigned int)(LPE_HAS_MASS | LPE_HAS_VELOCITY);
}
__global__ void k_pack_mass(int n, const double* __restrict__ m, const unsigned int* __restrict__ rank,
                            Body* __restrict__ body, Scal* __restrict__ s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double mi = 0.0;
    unsigned int cm = 0u;
    if (i < n) {
        mi = m[i];
        cm = body[i].comp;   // written by k_pack_pos before the sort
        body[i].m = mi;
        body[i].rank = rank ? rank[i] : (unsigned int)(n - 1 - i);
    }
    block_max_mass(mi, cm, s);
}
// the scenario's bodies made on the device, one thread per body (kepler_gen.h)
__global__ void __launch_bounds__(256)
k_generate_keplerian(int n, unsigned long long seed, double U, Body* __restrict__ body, double2* __restrict__ vel, Scal* __restrict__ s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    Body b;
    b.m = 0.0; b.comp = 0u;
    if (i < n) {
        double vx, vy;
        lpe_keplerian_body((uint64_t)i, seed, U, &b.x, &b.y, &vx, &vy, &b.m);
        b.rank = (unsigned int)(n - 1 - i);    // EnTT's view order: newest entity first
        b.comp = (unsigned int)(LPE_HAS_MASS | LPE_HAS_VELOCITY);
        body[i] = b;
        vel[i] = make_double2(vx, vy);
    }
    block_max_mass(b.m, b.comp, s);
}
// MovementSystem::update as its own pass (STRICT precision: the traversal reads leaf bodies from the state, so the
// positions must not move under it). Same expression as the fused drift of the traversal kernels.
__global__ void __launch_bounds__(256) k_drift(int n, double dtD, Body* __restrict__ body, const double2* __restrict__ vel,
                                               const Scal* __restrict__ live) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (live && (unsigned int)i >= live->n_live)) return;
    const unsigned int cm = body[i].comp;
    const bool mover = (cm & 2u) && !(cm & 4u) && !(cm & 8u);   // movement.cpp:20-29
    if (!mover) return;
    double2 p = *reinterpret_cast<const double2*>(&body[i].x);
    const double2 v = vel[i];
    p.x = drift_step(p.x, v.x, dtD);                                            // movement.cpp:32-33
    p.y = drift_step(p.y, v.y, dtD);
    *reinterpret_cast<double2*>(&body[i].x) = p;
}
// Host tick, FAST precision: the tree walk does not need the velocities, so it runs while they are still coming over
// PCIe; its epilogue leaves {x, y, dvx, dvy} per body as one 32-byte record at the body's CREATION index (a random place
// as far as key order is concerned: one full-sector store per body, hidden inside the compute-bound walk). This pass is
// then the kick and the drift in creation order, every access streaming: uploaded velocity + change -> new velocity,
// drifted position, written over the staging arrays the host's copies are made from (slot o is read and written by the
// same thread). At 16 M bodies 0.2 ms, where kicking in key order and scattering four 8-byte values per body into the
// host-shaped arrays took 2.9 ms (partial sectors: read-modify-write in DRAM).
template <bool AOS>
__global__ void __launch_bounds__(256)
k_finish_tick(int n, int do_drift, double dtD, const double4* __restrict__ stage, const unsigned char* __restrict__ comp,
              double* va, double* vb, double* xa, double* xb) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    const double4 r = stage[o];
    double2 v = AOS ? reinterpret_cast<const double2*>(va)[o] : make_double2(va[o], vb[o]);
    const unsigned int cm = comp ? (unsigned int)comp[o] : (unsigned int)(LPE_HAS_MASS | LPE_HAS_VELOCITY);
    if ((cm & 1u) && (cm & 2u) && !(cm & 4u)) {                          // a target, barnes_hut.cpp:89
        v.x = __dadd_rn(v.x, r.z);                                       // r.z = 0 + acc * dt (kick_step), barnes_hut.cpp:284-286
        v.y = __dadd_rn(v.y, r.w);
    }
    double2 p = make_double2(r.x, r.y);
    if (do_drift && (cm & 2u) && !(cm & 4u) && !(cm & 8u)) {             // a mover, movement.cpp:20-29
        p.x = drift_step(p.x, v.x, dtD);                                 // movement.cpp:32-33
        p.y = drift_step(p.y, v.y, dtD);
    }
    if (AOS) {
        reinterpret_cast<double2*>(va)[o] = v;
        reinterpret_cast<double2*>(xa)[o] = p;
    } else {
        va[o] = v.x; vb[o] = v.y;
        xa[o] = p.x; xb[o] = p.y;
    }
}
// ... and the resident state (key order) catches up from the same arrays on the side stream, beside the downloads
template <bool AOS>
__global__ void __launch_bounds__(256)
k_refresh_state(int n, int do_drift, const unsigned int* __restrict__ orig, Body* __restrict__ body, double2* __restrict__ vel,
                const double* __restrict__ va, const double* __restrict__ vb, const double* __restrict__ xa,
                const double* __restrict__ xb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int o = orig ? orig[i] : (unsigned int)i;
    vel[i] = AOS ? reinterpret_cast<const double2*>(va)[o] : make_double2(va[o], vb[o]);
    if (do_drift)
        *reinterpret_cast<double2*>(&body[i].x) = AOS ? reinterpret_cast<const double2*>(xa)[o] : make_double2(xa[o], xb[o]);
}
// the resident state as {x, y, vx, vy} records in creation order (lpe_bh_download: same reason as above)
__global__ void __launch_bounds__(256)
k_stage_state(int n, const Body* __restrict__ body, const double2* __restrict__ vel, const unsigned int* __restrict__ orig,
              double4* __restrict__ stage) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 p = *reinterpret_cast<const double2*>(&body[i].x);
    const double2 v = vel[i];
    stage[orig ? orig[i] : (unsigned int)i] = make_double4(p.x, p.y, v.x, v.y);
}
// records -> the arrays the host wants: four arrays of doubles (x, y, vx, vy; any may be null), or with AOS two arrays of
// {x, y} records (pos = xa, vel = va)
template <bool AOS>
__global__ void __launch_bounds__(256)
k_unstage(int n, const double4* __restrict__ stage, double* __restrict__ xa, double* __restrict__ xb,
          double* __restrict__ va, double* __restrict__ vb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 r = stage[i];
    if (AOS) {
        if (xa) reinterpret_cast<double2*>(xa)[i] = make_double2(r.x, r.y);
        if (va) reinterpret_cast<double2*>(va)[i] = make_double2(r.z, r.w);
    } else {
        if (xa) xa[i] = r.x;
        if (xb) xb[i] = r.y;
        if (va) va[i] = r.z;
        if (vb) vb[i] = r.w;
    }
}
// array-of-structs forms for the ECS drop-in: EnTT keeps Position / Velocity as {double x, y} records, so its pool pages
// can be copied as they are (lpe_bh_update_host_aos)
__global__ void k_pack_pos_aos(int n, const double2* __restrict__ pos, const unsigned char* __restrict__ comp, Body* __restrict__ body) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    *reinterpret_cast<double2*>(&body[i].x) = pos[i];
    body[i].comp = comp ? (unsigned int)comp[i] : (unsigned int)(LPE_HAS_MASS | LPE_HAS_VELOCITY);
}
__global__ void k_pack_vel_aos(int n, const double2* __restrict__ v, double2* __restrict__ out, const unsigned int* __restrict__ orig) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v[orig ? orig[i] : (unsigned int)i];
}
__global__ void k_unpack_vel_aos(int n, const double2* __restrict__ in, double2* __restrict__ out, const unsigned int* __restrict__ orig) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[orig ? orig[i] : (unsigned int)i] = in[i];
}
__global__ void k_get_pos_aos(int n, const Body* __restrict__ body, double2* __restrict__ out, const unsigned int* __restrict__ orig) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[orig ? orig[i] : (unsigned int)i] = *reinterpret_cast<const double2*>(&body[i].x);
}
__global__ void k_set_pos(int n, const double* __restrict__ x, const double* __restrict__ y, Body* __restrict__ body,
                          const unsigned int* __restrict__ orig) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned int s = orig ? orig[i] : (unsigned int)i;
        *reinterpret_cast<double2*>(&body[i].x) = make_double2(x[s], y[s]);
    }
}
__global__ void k_get_pos(int n, const Body* __restrict__ body, double* __restrict__ x, double* __restrict__ y,
                          const unsigned int* __restrict__ orig) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double2 v = *reinterpret_cast<const double2*>(&body[i].x);
        const unsigned int s = orig ? orig[i] : (unsigned int)i;
        x[s] = v.x;
        y[s] = v.y;
    }
}

// sharded mode: every rank's packed slice -> state arrays (sidx null: the state is in key order, slot = position)
__global__ void k_xchg_scatter(int n, int nranks, unsigned long long chunk, const unsigned int* __restrict__ sidx,
                               const double4* __restrict__ recv, Body* __restrict__ body,
                               double2* __restrict__ vel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int gblock = (unsigned int)i / LPE_SHARD_BLOCK;
    const unsigned int r = gblock % (unsigned int)nranks, lblock = gblock / (unsigned int)nranks;
    const unsigned long long slot = (unsigned long long)lblock * LPE_SHARD_BLOCK + ((unsigned int)i % LPE_SHARD_BLOCK);
    const double4 v = recv[(unsigned long long)r * chunk + slot];
    const unsigned int b = sidx ? sidx[i] : (unsigned int)i;
    *reinterpret_cast<double2*>(&body[b].x) = make_double2(v.x, v.y);
    vel[b] = make_double2(v.z, v.w);
}

// BoundarySystem::update (reference src/systems/boundary.cpp:26-68) over the resident state, one body per thread.
// Round-to-nearest intrinsics keep nvcc from contracting a*b+c into an FMA the reference (g++ -O2, x86-64) never uses.
__global__ void __launch_bounds__(256)
k_boundary(int n, Body* __restrict__ body, double2* __restrict__ vel, double marginM, double universeSizeM,
           double bounceDamping, double maxSpeed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int cm = body[i].comp;
    if (!(cm & LPE_HAS_VELOCITY) || (cm & LPE_ASLEEP)) return;   // view<Position, Velocity>, asleep skipped
    double2 p = *reinterpret_cast<const double2*>(&body[i].x);
    double2 v = vel[i];
    const double hiEdge = __dsub_rn(universeSizeM, marginM);
    bool bounced = false;
    if (p.x < marginM) {
        p.x = marginM; v.x = __dmul_rn(fabs(v.x), bounceDamping); bounced = true;
    } else if (p.x > hiEdge) {
        p.x = hiEdge; v.x = __dmul_rn(-fabs(v.x), bounceDamping); bounced = true;
    }
    if (p.y < marginM) {
        p.y = marginM; v.y = __dmul_rn(fabs(v.y), bounceDamping); bounced = true;
    } else if (p.y > hiEdge) {
        p.y = hiEdge; v.y = __dmul_rn(-fabs(v.y), bounceDamping); bounced = true;
    }
    if (!bounced) return;
    const double speed = __dsqrt_rn(__dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y)));
    if (speed > maxSpeed) {
        v.x = __dmul_rn(__ddiv_rn(v.x, speed), maxSpeed);
        v.y = __dmul_rn(__ddiv_rn(v.y, speed), maxSpeed);
    }
    *reinterpret_cast<double2*>(&body[i].x) = p;
    vel[i] = v;
}

// Direct O(N^2) sum in fp64 with the reference's force law (barnes_hut.cpp:257-282), tiled through shared memory.
__global__ void __launch_bounds__(256)
k_direct(int n, const Body* __restrict__ body, double U, double eps2, double G, int first, int count,
         double* __restrict__ ax, double* __restrict__ ay, const unsigned int* __restrict__ orig) {
    __shared__ double sx[256], sy[256], sm[256];
    // targets are the creation indices [first, first + count). With a re-ordered state (orig != null) every slot is
    // looked at and the ones whose body falls into the range are computed.
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    int i = first + k, out = k;
    bool want = k < count;
    if (orig) {
        i = k;
        out = (k < n) ? (int)orig[k] - first : -1;
        want = k < n && out >= 0 && out < count;
    }
    double2 p = make_double2(0.0, 0.0);
    if (want) p = make_double2(body[i].x, body[i].y);
    double accx = 0.0, accy = 0.0;
    for (int base = 0; base < n; base += 256) {
        const int j = base + threadIdx.x;
        double2 q = make_double2(0.0, 0.0);
        double m = 0.0;
        if (j < n) {
            const Body bj = body[j];
            const unsigned int cm = bj.comp;
            q = make_double2(bj.x, bj.y);
            const bool src = (cm & 1u) && !(cm & 4u) && q.x >= 0.0 && q.x < U && q.y >= 0.0 && q.y < U;
            m = src ? bj.m : 0.0;
        }
        __syncthreads();
        sx[threadIdx.x] = q.x; sy[threadIdx.x] = q.y; sm[threadIdx.x] = m;
        __syncthreads();
        const int lim = min(256, n - base);
        for (int t = 0; t < lim; ++t) {
            if (base + t == i) continue;
            const double dx = sx[t] - p.x, dy = sy[t] - p.y;
            const double d2 = dx * dx + dy * dy + eps2;
            const double f = G * sm[t] / (d2 * sqrt(d2));
            accx += dx * f;
            accy += dy * f;
        }
    }
    if (want) {
        ax[out] = accx;
        ay[out] = accy;
    }
}

// 16 independent FMA chains per thread, register resident
__global__ void __launch_bounds__(256) k_fma_peak(int iters, float a, float* __restrict__ sink) {
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = (float)(threadIdx.x + k) * 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = fmaf(v[k], a, 0.5f);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += v[k];
    if (s == 123.456f) sink[threadIdx.x] = s;
}

// The look-back of the sort gives up after a bounded wait and raises a flag instead of hanging the device; the flag
// rides along with every synchronising call (no extra round trip).
int fetch_fault(lpe_bh_ctx* c) {
    if (!c->totals || !c->have_step || !c->fault_host) return 0;
    CU_TRY(c, cudaMemcpyAsync(c->fault_host, c->totals + 512 * SORT_MAX_PASSES + SORT_MAX_PASSES, sizeof(unsigned int),
                              cudaMemcpyDeviceToHost, c->stream));
    return 0;
}
int check_fault(lpe_bh_ctx* c) {
    if (c->fault_host && *c->fault_host) {
        const unsigned int f = *c->fault_host;
        *c->fault_host = 0;
        if (f & 2u) return fail(c, "checked build: radix sort scatter index out of range");
        return fail(c, "radix sort look-back timed out (internal error): results of the last step are invalid");
    }
#ifdef LPE_CHECKED
    if (c->scal && c->have_step) {
        unsigned int f = 0;
        if (cudaMemcpy(&f, &c->scal->check_fault, sizeof(f), cudaMemcpyDeviceToHost) == cudaSuccess && f) {
            char buf[96];
            std::snprintf(buf, sizeof(buf), "checked build: index out of range, check bits 0x%x", f);
            return fail(c, buf);
        }
    }
#endif
    return 0;
}

int choose_depth(const lpe_bh_params& p) {
    if (p.max_depth > 0) return p.max_depth > LPE_MAX_DEPTH ? LPE_MAX_DEPTH : p.max_depth;
    if (!(p.softening > 0.0) || !(p.theta > 0.0)) return LPE_MAX_DEPTH;
    // a cell with s < theta*eps is accepted whatever the distance (s^2/(d^2+eps^2) < theta^2), so the
    // reference never descends below the first level where that holds. A small safety margin keeps the fp64
    // comparison in the reference (barnes_hut.cpp:269) strictly on the accepting side.
    const double lim = p.theta * p.softening * (1.0 - 1e-9);
    int D = 1;
    while (D < LPE_MAX_DEPTH && std::ldexp(p.universe_size, -D) >= lim) ++D;
    return D;
}

int make_const(lpe_bh_ctx* c, const lpe_bh_params& p, StepConst& k) {
    if (!(p.universe_size > 0.0) || !std::isfinite(p.universe_size)) return fail(c, "universe_size must be positive and finite");
    if (!(p.theta >= 0.0)) return fail(c, "theta must be >= 0");
    if (p.precision != LPE_PREC_FAST && p.precision != LPE_PREC_STRICT) return fail(c, "unknown precision");
    std::memset(&k, 0, sizeof(k));
    k.U = p.universe_size;
    int e = 0;
    std::frexp(p.universe_size, &e);  // U = f * 2^e, f in [0.5,1)
    k.S = std::ldexp(1.0, e);
    k.invS = std::ldexp(1.0, -e);
    k.eps = p.softening;
    const double es = p.softening * k.invS;
    k.eps2s = es * es;
    k.eps2f = (float)k.eps2s;
    k.theta = p.theta;
    k.theta2 = p.theta * p.theta;
    k.thr = p.small_mass_threshold;
    k.G = p.G;
    k.dtK = p.dt_kick;
    k.dtD = p.dt_drift;
    k.D = choose_depth(p);
    k.h = std::ldexp(p.universe_size, -k.D);
    k.invh = 1.0 / k.h;
    k.quirk = p.quirk_mode ? 1 : 0;
    k.do_drift = p.do_drift ? 1 : 0;
    k.n = (int)c->n;
    k.shard_rank = c->shard_rank;
    k.shard_n = c->shard_n;
    // the body's-own-leaf bookkeeping is only needed when a self interaction would not vanish by itself (eps == 0)
    // or when interactions are counted
    k.need_self = ((c->instr & 2) || !((float)k.eps2s > 0.0f)) ? 1 : 0;
    k.test_overflow = c->force_overflow ? 1 : 0;
    if (p.key_order < 0 || p.key_order > 2) return fail(c, "unknown key_order");
    k.hilbert = (p.key_order == LPE_KEYS_HILBERT || (p.key_order == LPE_KEYS_AUTO && p.precision == LPE_PREC_FAST)) ? 1 : 0;
    k.k32 = (k.D <= 16) ? 1 : 0;
    k.dd = 0;
    k.blockBase = 1u;
    k.bodyCap = (unsigned int)c->cap;
    k.nodeCap = (unsigned int)c->node_cap;
    k.recSlots = 4u * ((unsigned int)c->cap + 8u);
    return 0;
}

// ---- one step, in the pieces the single-GPU step and the domain-decomposed phases are assembled from ------------------
struct SortPlan { int passes, topBits; };
SortPlan sort_plan(const StepConst& k) {
    SortPlan sp;
    if (k.k32) {
        // 32-bit containers: 8-bit digits over the 2D bits of the cell index; "not in the tree" rides in the payload and
        // joins the top pass as digit bit 8 (always 512 bins there)
        sp.passes = std::max(1, (2 * k.D + 7) / 8);
        sp.topBits = 9;
        return sp;
    }
    // 8-bit digits; a key of 8p+1 bits gets a 9-bit top digit instead of one more pass
    const int keyBits = 2 * k.D + 1;
    sp.passes = (keyBits + 7) / 8;
    sp.topBits = keyBits - 8 * (sp.passes - 1);
    if (sp.passes > 1 && sp.topBits == 1) { --sp.passes; sp.topBits = 9; }
    if (sp.topBits < 8) sp.topBits = 8;
    return sp;
}

// Look-back status words carry the step's epoch, so the arrays are cleared only when the context's buffers are new
// (c->epoch == 0) or the epoch is about to wrap. Called before a step is queued (and before a step is CAPTURED into a
// CUDA graph: the clearing must not become part of the graph).
int epoch_prepare(lpe_bh_ctx* c) {
    if (c->epoch != 0u && c->epoch < (1u << 30) - 1u) return 0;
    cudaStream_t st = c->stream;
    CU_TRY(c, cudaMemsetAsync(c->lbstatus, 0, sizeof(unsigned long long) *
                                  ((size_t)cdiv((long long)c->cap, SORT_TILE) * (256 * (SORT_MAX_PASSES - 1) + 512) +
                                   2 * ((size_t)cdiv((long long)c->cap + 1, SCAN_TILE) + 2)), st));
    k_epoch_set<<<1, 1, 0, st>>>(c->epoch_dev, 1u);   // (cleared words carry epoch 0: never "ready")
    c->epoch = 1u;
    return 0;
}

int step_prologue(lpe_bh_ctx* c, int n) {
    cudaStream_t st = c->stream;
    // (the first word, the mass scale, belongs to the upload)
    CU_TRY(c, cudaMemsetAsync(reinterpret_cast<char*>(c->scal) + 8, 0, sizeof(Scal) - 8, st));
    CU_TRY(c, cudaMemsetAsync(c->totals, 0, sizeof(unsigned int) * (512 * SORT_MAX_PASSES + 16), st));
    CU_TRY(c, cudaMemsetAsync(c->mask, 0, sizeof(unsigned int) * ((size_t)n + 1), st));
    CU_TRY(c, cudaMemsetAsync(c->child, 0xFF, sizeof(unsigned int) * 4 * ((size_t)n + 1), st));
    CU_TRY(c, cudaMemsetAsync(c->levelMeta, 0, sizeof(unsigned int) * 3 * 32, st));
    return 0;
}

// keys[0] / vals[0] -> sorted keys / payload in keys[sorted_sel] / vals[sorted_sel]
int step_sort(lpe_bh_ctx* c, const StepConst& k, int n, const unsigned int* n_dev = nullptr) {
    cudaStream_t st = c->stream;
    const SortPlan plan = sort_plan(k);
    const int passes = plan.passes, topBits = plan.topBits;
    // (status words are laid out for the smaller tile, so they also hold the 4096-key tiles of the 32-bit path)
    const int statusTiles = cdiv(n, SORT_TILE);
    const int sortTiles = k.k32 ? cdiv(n, 2 * SORT_TILE) : statusTiles;
    int sel = 0;
    if (epoch_prepare(c)) return 1;
    ++c->epoch;   // (advanced on the device by k_sort_bases)
    unsigned int* hist = c->totals;
    unsigned int* tileCounter = c->totals + 512 * SORT_MAX_PASSES;
    unsigned int* fault = tileCounter + SORT_MAX_PASSES;
    const int lastBins = 1 << topBits;
    const int histBlocks = std::min(sortTiles, 148 * 8);
    auto run = [&](auto keyTag) {
        using KeyT = decltype(keyTag);
        constexpr int THREADS = sizeof(KeyT) == 4 ? 2 * SORT_THREADS : SORT_THREADS;   // tile = THREADS x SORT_ITEMS keys
        KeyT* kb[2] = {reinterpret_cast<KeyT*>(c->keys[0]), reinterpret_cast<KeyT*>(c->keys[1])};
        k_sort_hist<KeyT><<<histBlocks, 256, sizeof(unsigned int) * SORT_HIST_STRIDE * passes, st>>>(kb[0], c->vals[0], n, passes, lastBins, hist, n_dev);
        k_sort_bases<<<passes, 512, 0, st>>>(hist, c->epoch_dev);
        for (int ps = 0; ps < passes; ++ps) {
            const int shift = 8 * ps;
            const int top = ps == passes - 1;
            unsigned long long* status = c->lbstatus + (size_t)ps * 256 * statusTiles;
            const unsigned int* base = hist + 512 * ps;
            // (ranking by ballots while a pass is about one wave of tiles, by MATCH.ANY beyond: bh_sort.cuh)
            const bool ballot = sizeof(KeyT) == 4 && n <= SORT_BALLOT_MAX;
            auto pass = [&](auto binsTag) {
                constexpr int BINS = decltype(binsTag)::value;
                constexpr size_t smem = sort_smem_bytes<BINS, KeyT, THREADS, SORT_ITEMS>();
                auto go = [&](auto kern) {
                    // (more than 48 KB of shared memory has to be asked for, per kernel and per device: a host-side call, no stream work)
                    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    kern<<<sortTiles, THREADS, smem, st>>>(kb[sel], c->vals[sel], kb[sel ^ 1], c->vals[sel ^ 1], n, shift, base, status,
                                                          c->epoch_dev, tileCounter + ps, fault, n_dev);
                };
                if (ballot) go(k_sort_onesweep<BINS, KeyT, THREADS, SORT_ITEMS, true>);
                else go(k_sort_onesweep<BINS, KeyT, THREADS, SORT_ITEMS, false>);
            };
            if (top && lastBins == 512) pass(std::integral_constant<int, 512>{});
            else pass(std::integral_constant<int, 256>{});
            sel ^= 1;
        }
    };
    if (k.k32) run((unsigned int)0); else run((unsigned long long)0);
    c->sorted_sel = sel;
    c->launches += 2 + (uint64_t)passes;
    c->last.sort_passes = passes;
    return 0;
}

// gather into key order (side stream) | terminals -> witnesses -> ordinals -> topology -> aggregation, deepest level first
int step_build(lpe_bh_ctx* c, const StepConst& k, int n, const unsigned int* n_dev = nullptr) {
    cudaStream_t st = c->stream;
    const int g256 = cdiv(n, 256);
    const unsigned long long* skeys = c->keys[c->sorted_sel];
    const unsigned int* sidx = c->vals[c->sorted_sel];
    // fork: the gather (and, in the host tick, the late mass / rank pack before it) runs on the side stream while
    // this stream goes on with the kernels that only need the sorted keys; joined before k_topology
    cudaStream_t sg = c->side_stream;
    CU_TRY(c, cudaEventRecord(c->evs[0], st));
    CU_TRY(c, cudaStreamWaitEvent(sg, c->evs[0], 0));
    if (c->pend_mass) {   // host path: masses and ranks were uploaded behind the key generation and the sort
        CU_TRY(c, cudaStreamWaitEvent(sg, c->evc[2], 0));
        k_pack_mass<<<g256, 256, 0, sg>>>(n, c->tmp + 2 * c->cap, c->pend_rank ? c->rank_in : nullptr, c->body, c->scal);
        c->pend_mass = false;
    }
    // (host tick: the velocities are still on their way and are packed straight into key order before the kick)
    k_gather<<<g256, 256, 0, sg>>>(n, k.need_self, sidx, c->body, c->pend_vel ? nullptr : c->vel,
                                   c->orig_valid ? c->orig : nullptr, c->body2, c->vel2, c->orig2, c->selfslot, n_dev,
                                   (unsigned int)c->cap, c->scal);
    CU_TRY(c, cudaEventRecord(c->evs[1], sg));
    // from here on the state IS in key order
    std::swap(c->body, c->body2);
    std::swap(c->vel, c->vel2);
    std::swap(c->orig, c->orig2);
    c->orig_valid = true;
    // scans are single-pass (look-back over one status word per tile), each fused with its consumer
    const int scanTiles = cdiv((long long)n + 1, SCAN_TILE);
    unsigned int* scanTicket = c->totals + 512 * SORT_MAX_PASSES + SORT_MAX_PASSES + 1;
    unsigned int* sortFault = c->totals + 512 * SORT_MAX_PASSES + SORT_MAX_PASSES;
    unsigned long long* scanStatus = c->lbstatus + (size_t)cdiv((long long)c->cap, SORT_TILE) * (256 * (SORT_MAX_PASSES - 1) + 512);
    k_scan_chained<<<scanTiles, SCAN_THREADS, 0, st>>>(HeadFlag{skeys, c->scal, k.k32}, TerminalSink{skeys, k.k32, c->tkey, c->tfirst, c->scal, n},
                                                       n, scanStatus, c->epoch_dev, scanTicket, sortFault);
    unsigned int *levelCount = c->levelMeta, *levelBase = c->levelMeta + 32, *levelCursor = c->levelMeta + 64;
    k_witness<<<g256, 256, 0, st>>>(k.D, c->tkey, c->delta, c->mask, c->wstart, levelCount, c->scal);
    k_scan_chained<<<scanTiles, SCAN_THREADS, 0, st>>>(MaskPop{c->mask, c->scal}, StoreSink{c->P}, n,
                                                       scanStatus + (size_t)cdiv((long long)c->cap + 1, SCAN_TILE) + 1, c->epoch_dev,
                                                       scanTicket + 1, sortFault);
    Topo topo{c->wstart, c->child, c->agg, c->levelList, levelCount, levelBase, levelCursor, c->tfirst, c->body, c->selfslot, c->rec};
    CU_TRY(c, cudaStreamWaitEvent(st, c->evs[1], 0));   // join: bodies are in key order
    k_topology<<<g256, 256, 0, st>>>(k, c->tkey, c->delta, c->mask, c->P, topo, c->scal);
    NodeOut no{c->agg, c->rec, c->selfslot, c->body};
    // branching cells, deepest level first; the handful of cells of levels <= 4 share one single-block launch
    const int sms = c->sms;
    const int Ltop = k.D - 1 < 4 ? k.D - 1 : 4;
    int levelLaunches = 0;
    for (int L = k.D - 1; L > Ltop; --L) {
        // a level holds at most min(4^L, n/2) cells
        long long maxCells = (L < 15) ? (1ll << (2 * L)) : (long long)n;
        if (maxCells > n) maxCells = n;
        int grid = cdiv(maxCells * 4, 256);   // four lanes per cell
        if (grid > sms * 4) grid = sms * 4;   // (64 registers: four blocks are resident per SM; 7 / 8 / 16 per SM measured 1-2 % slower)
        k_agg_level<<<grid, 256, 0, st>>>(k, L, c->levelList, levelBase, levelCount, c->child, no, c->scal);
        ++levelLaunches;
    }
    k_agg_top<<<1, 1024, 0, st>>>(k, Ltop, c->levelList, levelBase, levelCount, c->child, no, c->scal);
    // gather, 2 scans, witness, topology, one launch per level above Ltop, agg_top
    c->launches += 1 + 2 + 1 + 1 + (uint64_t)levelLaunches + 1;
    return 0;
}

int step_traverse(lpe_bh_ctx* c, const StepConst& k, const lpe_bh_params& p, int n, bool sharded_begin) {
    cudaStream_t st = c->stream;
    const bool stats = c->instr & 2;
    const int g256 = cdiv(n, 256);
    const int sms = c->sms;
    TravArgs ta{};
    ta.rec = c->rec; ta.agg = c->agg;
    ta.body = c->body; ta.vel = c->vel;
    ta.xchg_send = c->xchg_send; ta.cntAcc = c->cntAcc; ta.cntVis = c->cntVis; ta.s = c->scal;
    ta.npeer = 0;
    ta.xrec = c->dd_xrec; ta.localLo = 0u; ta.localHi = 0xFFFFFFFFu; ta.chunk_cost = nullptr;
    ta.stage_out = (c->defer_kick && !k.dd) ? c->stage4 : nullptr;
    ta.orig = c->orig_valid ? c->orig : nullptr;
    if (k.dd) {
        ta.localLo = 4u * k.blockBase;
        ta.localHi = 4u * (k.blockBase + (unsigned int)c->cap + 8u);
        ta.chunk_cost = c->dd_chunk_cost;
    }
    if (sharded_begin && p2p_ready(c)) {
        // this step's generation of every rank's receive buffer, at this rank's slice
        c->xchg_parity ^= 1;
        const size_t gen = (size_t)c->xchg_parity * c->xchg_chunk * (size_t)c->shard_n;
        for (int r = 0; r < c->shard_n; ++r) ta.peer[r] = c->peer_recv[r] + gen + (size_t)c->shard_rank * c->xchg_chunk;
        ta.npeer = c->shard_n;
    }
    const unsigned int nblocks = (unsigned int)cdiv(n, LPE_SHARD_BLOCK);
    const unsigned int own = (nblocks + (unsigned int)k.shard_n - 1u - (unsigned int)k.shard_rank) / (unsigned int)k.shard_n;
    ta.n_chunks_local = own * (LPE_SHARD_BLOCK / 32u);
    ta.selfslot = c->selfslot; ta.chunk_list = nullptr;
    const int maxGridCtas = sms * 8;
    int travLaunches = 1;
    if (p.precision == LPE_PREC_FAST && !c->force_dfs) {
        // two-phase kernel, then the depth-first kernel for the (normally zero) chunks whose frontier overflowed
        const size_t smem = sizeof(T2Warp) * T2_WARPS;
        const bool selfT = k.need_self != 0;
        static_assert(sizeof(T2Warp) * T2_WARPS + sizeof(T3Cta) <= 48 * 1024, "per-CTA work areas fit the default dynamic shared memory limit");
        int grid = cdiv(ta.n_chunks_local, T2_WARPS);
        if (grid > sms * T2_MIN_CTAS) grid = sms * T2_MIN_CTAS;
        if (grid < 1) grid = 1;
        // the far field shared by the four warps of a CTA (bh_traverse2.cuh) unless instrumentation bit 5 asks for the
        // warp-only walk (A/B testing)
        const bool cta = !(c->instr & 32);
        auto launch = [&](auto modeTag, auto ctaTag) {
            constexpr int MODE = decltype(modeTag)::value;
            constexpr bool CTA = decltype(ctaTag)::value;
            const size_t sm = smem + (CTA ? sizeof(T3Cta) : 0);
            if (stats) k_traverse2<true, true, MODE, CTA><<<grid, T2_THREADS, sm, st>>>(k, ta, c->ovf_list);
            else if (selfT) k_traverse2<false, true, MODE, CTA><<<grid, T2_THREADS, sm, st>>>(k, ta, c->ovf_list);
            else k_traverse2<false, false, MODE, CTA><<<grid, T2_THREADS, sm, st>>>(k, ta, c->ovf_list);
        };
        auto launchMode = [&](auto ctaTag) {
            if (k.dd) launch(std::integral_constant<int, T2_DD>{}, ctaTag);
            else if (ta.stage_out) launch(std::integral_constant<int, T2_STAGED>{}, ctaTag);
            else launch(std::integral_constant<int, T2_RESIDENT>{}, ctaTag);
        };
        if (cta) launchMode(std::true_type{}); else launchMode(std::false_type{});
        ta.chunk_list = c->ovf_list;
        if (stats) k_traverse<0, true><<<sms, TRAV_THREADS, 0, st>>>(k, ta);
        else k_traverse<0, false><<<sms, TRAV_THREADS, 0, st>>>(k, ta);
        travLaunches = 2;
    } else {
        int grid = cdiv(ta.n_chunks_local, TRAV_THREADS / 32);
        if (grid > maxGridCtas) grid = maxGridCtas;
        if (grid < 1) grid = 1;
        if (p.precision == LPE_PREC_FAST) {
            if (stats) k_traverse<0, true><<<grid, TRAV_THREADS, 0, st>>>(k, ta);
            else k_traverse<0, false><<<grid, TRAV_THREADS, 0, st>>>(k, ta);
        } else {
            if (stats) k_traverse<1, true><<<grid, TRAV_THREADS, 0, st>>>(k, ta);
            else k_traverse<1, false><<<grid, TRAV_THREADS, 0, st>>>(k, ta);
            if (k.do_drift && c->shard_n == 1) {   // STRICT: the drift is its own pass (see k_drift)
                k_drift<<<g256, 256, 0, st>>>(n, k.dtD, c->body, c->vel, k.dd ? c->scal : nullptr);
                ++travLaunches;
            }
        }
    }
    c->launches += (uint64_t)travLaunches;
    return 0;
}

int run_step(lpe_bh_ctx* c, const lpe_bh_params& p, bool sharded_begin) {
    if (c->dd) return fail(c, "context is in domain-decomposed mode: use lpe_bh_dd_step / lpe_bh_dd_phase");
    StepConst k;
    if (make_const(c, p, k)) return 1;
    const int n = (int)c->n;
    if (n == 0) return 0;
    if (c->shard_n > 1 && ensure_xchg(c)) return 1;
    cudaStream_t st = c->stream;
    const bool timing = c->instr & 1;
    if (timing) cudaEventRecord(c->ev[0], st);
    if (step_prologue(c, n)) return 1;
    k_keygen<<<cdiv(n, 256), 256, 0, st>>>(k, c->body, c->keys[0], c->vals[0], c->scal);
    c->launches += 1;
    if (timing) cudaEventRecord(c->ev[1], st);
    if (step_sort(c, k, n)) return 1;
    if (timing) cudaEventRecord(c->ev[2], st);
    if (step_build(c, k, n)) return 1;
    if (timing) cudaEventRecord(c->ev[3], st);
    // host path: the velocities were uploaded behind the build. FAST precision does not need them for the walk: the kick
    // is deferred to k_finish_tick (the caller launches it once they have arrived). STRICT sums in the reference's
    // order starting from the velocity itself, so it waits for them here.
    c->defer_kick = c->pend_vel && p.precision == LPE_PREC_FAST && c->shard_n == 1;
    if (c->pend_vel && !c->defer_kick) {
        CU_TRY(c, cudaStreamWaitEvent(st, c->evc[3], 0));
        if (c->pend_vel_aos)
            k_pack_vel_aos<<<cdiv(n, 256), 256, 0, st>>>(n, reinterpret_cast<const double2*>(c->tmp + 3 * c->cap), c->vel,
                                                         c->orig_valid ? c->orig : nullptr);
        else
            k_pack2<<<cdiv(n, 256), 256, 0, st>>>(n, c->tmp + 3 * c->cap, c->tmp + 4 * c->cap, c->vel, c->orig_valid ? c->orig : nullptr);
    }
    c->pend_vel = false;
    c->pend_vel_aos = false;
    if (step_traverse(c, k, p, n, sharded_begin)) return 1;
    if (timing) cudaEventRecord(c->ev[4], st);
    CU_TRY(c, cudaGetLastError());
    c->last_c = k;
    c->have_step = true;
    c->last.depth = k.D;
    c->last.hilbert = k.hilbert;
    return 0;
}


// ---- whole steps as CUDA graphs ---------------------------------------------------------------------------------
// A step is ~30 small launches and memsets with no host decision in between (every count lives in device memory, the
// look-back epoch too), so it can be captured once and replayed: one submission instead of thirty. That matters at
// 1 M bodies, where the build is bound by launch latency, and most in the host tick, where kernel launches queue up
// behind the PCIe traffic of the uploads (measured: the build of a 1 M-body tick took 0.33 ms against 0.26 resident, and
// 1.1 ms in the three-call tick of the ECS drop-in). `body` queues the work on c->stream (and on the copy / side streams,
// forked from and joined to it with events) and updates the context's host-side state; the key says everything the
// queued work depends on. A key is captured the second time it comes up (a single step is not worth an instantiation);
// the host-side effects of the step are recorded with the graph and re-applied on every replay.
struct HostState {
    Body *body, *body2; double2 *vel, *vel2; unsigned int *orig, *orig2;
    bool orig_valid, have_step, pend_mass, pend_vel, pend_rank, pend_vel_aos, defer_kick;
    int sorted_sel; uint64_t launches; unsigned int epoch; lpe_bh_stats last; StepConst last_c;
    int dd_cur; unsigned long long dd_epoch;
};
HostState host_state(const lpe_bh_ctx* c) {
    return HostState{c->body, c->body2, c->vel, c->vel2, c->orig, c->orig2, c->orig_valid, c->have_step, c->pend_mass, c->pend_vel,
                     c->pend_rank, c->pend_vel_aos, c->defer_kick, c->sorted_sel, c->launches, c->epoch, c->last, c->last_c,
                     c->dd_cur, c->dd_epoch};
}
void restore_host_state(lpe_bh_ctx* c, const HostState& h) {
    c->body = h.body; c->body2 = h.body2; c->vel = h.vel; c->vel2 = h.vel2; c->orig = h.orig; c->orig2 = h.orig2;
    c->orig_valid = h.orig_valid; c->have_step = h.have_step; c->pend_mass = h.pend_mass; c->pend_vel = h.pend_vel;
    c->pend_rank = h.pend_rank; c->pend_vel_aos = h.pend_vel_aos; c->defer_kick = h.defer_kick;
    c->sorted_sel = h.sorted_sel; c->launches = h.launches; c->epoch = h.epoch; c->last = h.last; c->last_c = h.last_c;
    c->dd_cur = h.dd_cur; c->dd_epoch = h.dd_epoch;
}
template <class T>
void key_add(std::string& k, const T& v) { k.append(reinterpret_cast<const char*>(&v), sizeof(T)); }
bool graphs_on(lpe_bh_ctx* c, bool dd_step = false) {
    if (c->use_graphs < 0) {
        const char* e = getenv("LPE_BH_GRAPHS");
        c->use_graphs = (e && e[0] == '0') ? 0 : 1;
    }
    // (timing events between the phases and the sharded / decomposed modes keep the plain launches)
    return c->use_graphs == 1 && !(c->instr & (1 | 16)) && c->shard_n == 1 && (dd_step || !c->dd);
}
// page-locked host memory? (asynchronous copies of pageable memory cannot be captured)
bool is_pinned(const void* p) {
    if (!p) return true;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

template <class F>
int run_graphed(lpe_bh_ctx* c, const std::string& key, F&& body, bool dd_step = false) {
    if (!graphs_on(c, dd_step)) return body();
    if (epoch_prepare(c)) return 1;   // never part of a graph
    for (auto& g : c->graphs) {
        if (g.key != key) continue;
        CU_TRY(c, cudaGraphLaunch(g.exec, c->stream));
        if (g.swapped) { std::swap(c->body, c->body2); std::swap(c->vel, c->vel2); std::swap(c->orig, c->orig2); }
        c->orig_valid = g.orig_valid;
        c->sorted_sel = g.sorted_sel;
        c->launches += g.launches;
        c->epoch += g.epochs;
        if (g.dd_flip) c->dd_cur ^= 1;
        c->dd_epoch += g.dd_steps;
        c->last = g.last;
        c->last_c = g.last_c;
        c->have_step = true;
        c->pend_mass = c->pend_vel = c->pend_rank = c->pend_vel_aos = c->defer_kick = false;
        ++c->graph_replays;
        return 0;
    }
    if (std::find(c->graph_seen.begin(), c->graph_seen.end(), key) == c->graph_seen.end()) {
        if (c->graph_seen.size() >= 16) c->graph_seen.erase(c->graph_seen.begin());
        c->graph_seen.push_back(key);
        return body();
    }
    const HostState before = host_state(c);
    if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        c->use_graphs = 0;
        return body();
    }
    c->capturing = true;
    const int rc = body();
    c->capturing = false;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    if (!rc && e == cudaSuccess && graph) e = cudaGraphInstantiate(&exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (rc || e != cudaSuccess || !exec) {
        // nothing was queued: put the host state back, give graphs up for this context, run the step the plain way
        cudaGetLastError();
        restore_host_state(c, before);
        c->use_graphs = 0;
        c->err.clear();
        return body();
    }
    if (cudaGraphLaunch(exec, c->stream) != cudaSuccess) {
        cudaGetLastError();
        cudaGraphExecDestroy(exec);
        restore_host_state(c, before);
        c->use_graphs = 0;
        return body();
    }
    if (c->graphs.size() >= 8) {
        if (c->graphs.front().exec) cudaGraphExecDestroy(c->graphs.front().exec);
        c->graphs.erase(c->graphs.begin());
    }
    lpe_bh_ctx::GraphEntry g;
    g.key = key;
    g.exec = exec;
    g.swapped = c->body != before.body;
    g.orig_valid = c->orig_valid;
    g.sorted_sel = c->sorted_sel;
    g.launches = c->launches - before.launches;
    g.epochs = c->epoch - before.epoch;
    g.dd_flip = c->dd_cur != before.dd_cur;
    g.dd_steps = c->dd_epoch - before.dd_epoch;
    g.last = c->last;
    g.last_c = c->last_c;
    c->graphs.push_back(std::move(g));
    return 0;
}

// Tail of a host tick whose kick was deferred (FAST precision): kick + drift in creation order over the staging arrays
// (main stream), the downloads behind it, and the resident key-ordered state catching up on the side stream beside the
// downloads. SoA: hx / hy / hvx / hvy are four host arrays; AoS: hx = {x, y} records, hvx = {vx, vy} records.
int queue_deferred_finish(lpe_bh_ctx* c, bool aos, bool has_comp, bool want_pos, double* hx, double* hy, double* hvx, double* hvy) {
    cudaStream_t st = c->stream, sg = c->side_stream;
    const uint64_t n = c->n;
    const size_t cap = c->cap, bytes = sizeof(double) * n;
    const int g = cdiv((long long)n, 256);
    double* t = c->tmp;
    const int do_drift = c->last_c.do_drift;
    const unsigned char* comp = has_comp ? c->comp_in : nullptr;
    const unsigned int* orig = c->orig_valid ? c->orig : nullptr;
    c->defer_kick = false;
    if (aos) k_finish_tick<true><<<g, 256, 0, st>>>((int)n, do_drift, c->last_c.dtD, c->stage4, comp, t + 3 * cap, nullptr, t, nullptr);
    else k_finish_tick<false><<<g, 256, 0, st>>>((int)n, do_drift, c->last_c.dtD, c->stage4, comp, t + 3 * cap, t + 4 * cap, t, t + cap);
    CU_TRY(c, cudaEventRecord(c->evs[0], st));
    if (c->tracing) cudaEventRecord(c->trace_ev[5], st);
    CU_TRY(c, cudaStreamWaitEvent(sg, c->evs[0], 0));
    if (aos) k_refresh_state<true><<<g, 256, 0, sg>>>((int)n, do_drift, orig, c->body, c->vel, t + 3 * cap, nullptr, t, nullptr);
    else k_refresh_state<false><<<g, 256, 0, sg>>>((int)n, do_drift, orig, c->body, c->vel, t + 3 * cap, t + 4 * cap, t, t + cap);
    CU_TRY(c, cudaEventRecord(c->evs[1], sg));
    c->launches += 2;
    if (aos) {
        CU_TRY(c, cudaMemcpyAsync(hvx, t + 3 * cap, 2 * bytes, cudaMemcpyDeviceToHost, st));
        if (want_pos) CU_TRY(c, cudaMemcpyAsync(hx, t, 2 * bytes, cudaMemcpyDeviceToHost, st));
    } else {
        CU_TRY(c, cudaMemcpyAsync(hvx, t + 3 * cap, bytes, cudaMemcpyDeviceToHost, st));
        CU_TRY(c, cudaMemcpyAsync(hvy, t + 4 * cap, bytes, cudaMemcpyDeviceToHost, st));
        if (want_pos) {
            CU_TRY(c, cudaMemcpyAsync(hx, t, bytes, cudaMemcpyDeviceToHost, st));
            CU_TRY(c, cudaMemcpyAsync(hy, t + cap, bytes, cudaMemcpyDeviceToHost, st));
        }
    }
    if (c->tracing) cudaEventRecord(c->trace_ev[6], st);
    CU_TRY(c, cudaStreamWaitEvent(st, c->evs[1], 0));   // join: the context's stream is done when the state has caught up, too
    // (the sort's fault flag comes down with the results)
    CU_TRY(c, cudaMemcpyAsync(c->fault_host, c->totals + 512 * SORT_MAX_PASSES + SORT_MAX_PASSES, sizeof(unsigned int),
                              cudaMemcpyDeviceToHost, st));
    return 0;
}
// ... and the host's wait for it
int wait_deferred_finish(lpe_bh_ctx* c) {
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->tracing) {
        c->tracing = false;
        float e[7] = {};
        for (int z = 1; z < 7; ++z) cudaEventElapsedTime(&e[z], c->trace_ev[0], c->trace_ev[z]);
        fprintf(stderr, "tick trace, ms after the first upload was queued: x,y up %.2f | m up %.2f | vx,vy up %.2f | traversal done %.2f | "
                        "kick + drift done %.2f | downloads done %.2f\n", e[1], e[2], e[3], e[4], e[5], e[6]);
        if (c->instr & 1) {
            float k[5] = {};
            for (int z = 0; z < 5; ++z) cudaEventElapsedTime(&k[z], c->trace_ev[0], c->ev[z]);
            fprintf(stderr, "            step events: begin %.3f | keys %.3f | sorted %.3f | built %.3f | traversed %.3f\n", k[0], k[1], k[2], k[3], k[4]);
        }
    }
    CU_TRY(c, cudaGetLastError());
    return check_fault(c);
}

// One host tick in one call (lpe_bh_update_host: four + one host arrays; lpe_bh_update_host_aos: {x, y} records).
// The arrays go up on the copy stream in the order the step first reads them (positions and components -> keys;
// masses and ranks -> gather; velocities -> kick) and the step waits per array.
int host_tick(lpe_bh_ctx* c, const lpe_bh_params* p, uint64_t n, bool aos, double* hx, double* hy, double* hvx, double* hvy,
              const double* m, const uint32_t* rank, const uint8_t* comp) {
    if (ensure_capacity(c, n)) return 1;
    c->n = n;
    c->have_step = false;
    c->orig_valid = false;
    cudaStream_t st = c->stream, cs = c->copy_stream;
    const size_t bytes = sizeof(double) * n;
    double* t = c->tmp;
    const size_t cap = c->cap;
    const int g = cdiv((long long)n, 256);
    // LPE_TICK_TRACE=1: timeline of every tick on stderr (where the PCIe legs and the kernels overlap)
    const bool trace = getenv("LPE_TICK_TRACE") != nullptr;
    cudaEvent_t* tr = c->trace_ev;
    if (trace && !tr[0]) for (int z = 0; z < 7; ++z) cudaEventCreate(&tr[z]);
    c->tracing = trace;
    auto queue = [&]() -> int {
        CU_TRY(c, cudaEventRecord(c->evc[0], st));           // the staging buffers are free once earlier work is done
        CU_TRY(c, cudaStreamWaitEvent(cs, c->evc[0], 0));
        if (trace) cudaEventRecord(tr[0], cs);
        if (aos) {
            CU_TRY(c, cudaMemcpyAsync(t, hx, 2 * bytes, cudaMemcpyHostToDevice, cs));
        } else {
            CU_TRY(c, cudaMemcpyAsync(t, hx, bytes, cudaMemcpyHostToDevice, cs));
            CU_TRY(c, cudaMemcpyAsync(t + cap, hy, bytes, cudaMemcpyHostToDevice, cs));
        }
        if (comp) CU_TRY(c, cudaMemcpyAsync(c->comp_in, comp, n, cudaMemcpyHostToDevice, cs));
        CU_TRY(c, cudaEventRecord(c->evc[1], cs));
        if (trace) cudaEventRecord(tr[1], cs);
        CU_TRY(c, cudaMemcpyAsync(t + 2 * cap, m, bytes, cudaMemcpyHostToDevice, cs));
        if (rank) CU_TRY(c, cudaMemcpyAsync(c->rank_in, rank, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, cs));
        CU_TRY(c, cudaEventRecord(c->evc[2], cs));
        if (trace) cudaEventRecord(tr[2], cs);
        if (aos) {
            CU_TRY(c, cudaMemcpyAsync(t + 3 * cap, hvx, 2 * bytes, cudaMemcpyHostToDevice, cs));
        } else {
            CU_TRY(c, cudaMemcpyAsync(t + 3 * cap, hvx, bytes, cudaMemcpyHostToDevice, cs));
            CU_TRY(c, cudaMemcpyAsync(t + 4 * cap, hvy, bytes, cudaMemcpyHostToDevice, cs));
        }
        CU_TRY(c, cudaEventRecord(c->evc[3], cs));
        if (trace) cudaEventRecord(tr[3], cs);
        CU_TRY(c, cudaStreamWaitEvent(st, c->evc[1], 0));
        CU_TRY(c, cudaMemsetAsync(c->scal, 0, 8, st));   // the mass scale is rebuilt by k_pack_mass (side stream, after the sort)
        if (aos) k_pack_pos_aos<<<g, 256, 0, st>>>((int)n, reinterpret_cast<const double2*>(t), comp ? c->comp_in : nullptr, c->body);
        else k_pack_pos<<<g, 256, 0, st>>>((int)n, t, t + cap, comp ? c->comp_in : nullptr, c->body);
        c->launches += 1;
        c->pend_mass = true;
        c->pend_rank = rank != nullptr;
        c->pend_vel = true;
        c->pend_vel_aos = aos;
        if (run_step(c, *p, false)) return 1;
        if (c->defer_kick) {
            if (trace) cudaEventRecord(tr[4], st);
            CU_TRY(c, cudaStreamWaitEvent(st, c->evc[3], 0));
            return queue_deferred_finish(c, aos, comp != nullptr, p->do_drift != 0, hx, hy, hvx, hvy);
        }
        return 0;
    };
    // FAST precision: the whole tick, copies included, is one CUDA graph when the host arrays are page-locked
    const bool deferred = p->precision == LPE_PREC_FAST;
    int rc;
    if (deferred && !trace && graphs_on(c) && is_pinned(hx) && is_pinned(hy) && is_pinned(hvx) && is_pinned(hvy) && is_pinned(m) &&
        is_pinned(rank) && is_pinned(comp)) {
        std::string key(aos ? "tickA" : "tickS");
        key_add(key, *p); key_add(key, n); key_add(key, c->cap); key_add(key, c->instr); key_add(key, c->force_dfs);
        key_add(key, c->force_overflow); key_add(key, c->body); key_add(key, c->stream);
        key_add(key, hx); key_add(key, hy); key_add(key, hvx); key_add(key, hvy); key_add(key, m); key_add(key, rank); key_add(key, comp);
        rc = run_graphed(c, key, queue);
    } else {
        rc = queue();
    }
    if (rc) {   // a failed step must not leave waits dangling for the next one
        c->pend_mass = c->pend_vel = c->pend_vel_aos = c->defer_kick = false;
        cudaStreamSynchronize(cs);
        return 1;
    }
    if (deferred) return wait_deferred_finish(c);
    // STRICT precision kicked inside the traversal (it sums in the reference's order, starting from the velocity).
    // BarnesHutSystem only changes Velocity (barnes_hut.cpp:285-286); positions move only when the drift is fused
    if (!aos) return lpe_bh_download(c, p->do_drift ? hx : nullptr, p->do_drift ? hy : nullptr, hvx, hvy);
    const unsigned int* orig = c->orig_valid ? c->orig : nullptr;
    if (p->do_drift) {
        k_get_pos_aos<<<g, 256, 0, st>>>((int)n, c->body, reinterpret_cast<double2*>(t), orig);
        CU_TRY(c, cudaMemcpyAsync(hx, t, 2 * bytes, cudaMemcpyDeviceToHost, st));
    }
    k_unpack_vel_aos<<<g, 256, 0, st>>>((int)n, c->vel, reinterpret_cast<double2*>(t + 3 * cap), orig);
    CU_TRY(c, cudaMemcpyAsync(hvx, t + 3 * cap, 2 * bytes, cudaMemcpyDeviceToHost, st));
    if (fetch_fault(c)) return 1;
    CU_TRY(c, cudaStreamSynchronize(st));
    CU_TRY(c, cudaGetLastError());
    return check_fault(c);
}

}  // namespace

extern "C" {

#ifdef LPE_CHECKED
const char* lpe_bh_version(void) { return "lpe_bh 0.2 (sm_100a, ABI 2, CHECKED build: bounds checks on)"; }
#else
const char* lpe_bh_version(void) { return "lpe_bh 0.2 (sm_100a, ABI 2)"; }
#endif

int lpe_bh_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int lpe_bh_create(int device, lpe_bh_ctx** out) {
    if (!out) return fail(nullptr, "out is NULL");
    *out = nullptr;
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0)
        return fail(nullptr, std::string("no CUDA device (this library has no CPU path): ") + cudaGetErrorString(e));
    if (device < 0 || device >= cnt) return fail(nullptr, "device index out of range");
    DevGuard _dg(device);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(nullptr, cudaGetErrorString(e));
    lpe_bh_ctx* c = new lpe_bh_ctx();
    c->device = device;
    bool ok = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (auto& ev : c->evs) ok = ok && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
    for (auto& ev : c->ev) ok = ok && cudaEventCreate(&ev) == cudaSuccess;
    for (auto& ev : c->evc) ok = ok && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->fault_host, sizeof(unsigned int)) == cudaSuccess;
    if (!ok) {
        const std::string msg = std::string("cannot create streams / events / pinned flag: ") + cudaGetErrorString(cudaGetLastError());
        c->stream = c->own_stream;
        lpe_bh_destroy(c);
        return fail(nullptr, msg);
    }
    *c->fault_host = 0;
    c->stream = c->own_stream;
    cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, device);
    if (c->sms <= 0) c->sms = 148;
    *out = c;
    return 0;
}

void lpe_bh_destroy(lpe_bh_ctx* c) {
    if (!c) return;
    DevGuard _dg(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    dd_release(c);
    for (auto& ev : c->dd_ev) if (ev) cudaEventDestroy(ev);
    if (c->side_stream) cudaStreamSynchronize(c->side_stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    close_peers(c);
    free_all(c);
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->evc) if (ev) cudaEventDestroy(ev);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (auto& ev : c->evs) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->trace_ev) if (ev) cudaEventDestroy(ev);
    if (c->side_stream) cudaStreamDestroy(c->side_stream);
    if (c->fault_host) cudaFreeHost(c->fault_host);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

const char* lpe_bh_last_error(const lpe_bh_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int lpe_bh_set_stream(lpe_bh_ctx* c, void* s) {
    if (!c) return 1;
    cudaStreamSynchronize(c->stream);
    c->stream = s ? static_cast<cudaStream_t>(s) : c->own_stream;
    return 0;
}

int lpe_bh_set_instrumentation(lpe_bh_ctx* c, int flags) {
    if (!c) return 1;
    c->instr = flags;
    c->force_dfs = (flags & 4) != 0;
    c->force_overflow = (flags & 8) != 0;
    return 0;
}

int lpe_bh_upload(lpe_bh_ctx* c, uint64_t n, const double* x, const double* y, const double* vx, const double* vy,
                  const double* m, const uint32_t* rank, const uint8_t* comp) {
    if (!c) return 1;
    if (n > LPE_MAX_BODIES) return fail(c, "too many bodies for one context (limit 2^28)");
    if (n && (!x || !y || !m)) return fail(c, "x, y and m are required");
    DevGuard _dg(c->device);
    if (ensure_capacity(c, n)) return 1;
    c->n = n;
    c->have_step = false;
    c->orig_valid = false;
    if (n == 0) return 0;
    cudaStream_t st = c->stream;
    const size_t bytes = sizeof(double) * n;
    const int g = cdiv((long long)n, 256);
    double *t0 = c->tmp, *t1 = c->tmp + c->cap, *t2 = c->tmp + 2 * c->cap;
    CU_TRY(c, cudaMemcpyAsync(t0, x, bytes, cudaMemcpyHostToDevice, st));
    CU_TRY(c, cudaMemcpyAsync(t1, y, bytes, cudaMemcpyHostToDevice, st));
    CU_TRY(c, cudaMemcpyAsync(t2, m, bytes, cudaMemcpyHostToDevice, st));
    if (rank) CU_TRY(c, cudaMemcpyAsync(c->rank_in, rank, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    if (comp) CU_TRY(c, cudaMemcpyAsync(c->comp_in, comp, n, cudaMemcpyHostToDevice, st));
    CU_TRY(c, cudaMemsetAsync(c->scal, 0, sizeof(Scal), st));
    k_pack_body<<<g, 256, 0, st>>>((int)n, t0, t1, t2, rank ? c->rank_in : nullptr, comp ? c->comp_in : nullptr, c->body, c->scal);
    if (vx) CU_TRY(c, cudaMemcpyAsync(t0, vx, bytes, cudaMemcpyHostToDevice, st));
    if (vy) CU_TRY(c, cudaMemcpyAsync(t1, vy, bytes, cudaMemcpyHostToDevice, st));
    k_pack2<<<g, 256, 0, st>>>((int)n, vx ? t0 : nullptr, vy ? t1 : nullptr, c->vel, nullptr);
    CU_TRY(c, cudaGetLastError());
    return 0;
}

int lpe_bh_generate(lpe_bh_ctx* c, int kind, uint64_t n, uint64_t seed, double universe_size) {
    if (!c) return 1;
    if (kind != 4) return fail(c, "lpe_bh_generate: only kind 4 (counter-based Keplerian disk) is made on the device");
    if (n > LPE_MAX_BODIES) return fail(c, "too many bodies for one context (limit 2^28)");
    if (!(universe_size > 0.0)) return fail(c, "universe_size must be positive");
    DevGuard _dg(c->device);
    if (ensure_capacity(c, n)) return 1;
    c->n = n;
    c->have_step = false;
    c->orig_valid = false;
    if (n == 0) return 0;
    CU_TRY(c, cudaMemsetAsync(c->scal, 0, sizeof(Scal), c->stream));
    k_generate_keplerian<<<cdiv((long long)n, 256), 256, 0, c->stream>>>((int)n, seed, universe_size, c->body, c->vel, c->scal);
    CU_TRY(c, cudaGetLastError());
    c->launches += 1;
    return 0;
}

int lpe_bh_upload_positions(lpe_bh_ctx* c, const double* x, const double* y) {
    if (!c || !x || !y) return 1;
    if (c->n == 0) return 0;
    DevGuard _dg(c->device);
    const size_t bytes = sizeof(double) * c->n;
    double *t0 = c->tmp, *t1 = c->tmp + c->cap;
    CU_TRY(c, cudaMemcpyAsync(t0, x, bytes, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(t1, y, bytes, cudaMemcpyHostToDevice, c->stream));
    k_set_pos<<<cdiv((long long)c->n, 256), 256, 0, c->stream>>>((int)c->n, t0, t1, c->body, c->orig_valid ? c->orig : nullptr);
    CU_TRY(c, cudaGetLastError());
    return 0;
}

int lpe_bh_upload_velocities(lpe_bh_ctx* c, const double* vx, const double* vy) {
    if (!c || !vx || !vy) return 1;
    if (c->n == 0) return 0;
    DevGuard _dg(c->device);
    const size_t bytes = sizeof(double) * c->n;
    double *t2 = c->tmp + 2 * c->cap, *t3 = c->tmp + 3 * c->cap;
    CU_TRY(c, cudaMemcpyAsync(t2, vx, bytes, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(t3, vy, bytes, cudaMemcpyHostToDevice, c->stream));
    k_pack2<<<cdiv((long long)c->n, 256), 256, 0, c->stream>>>((int)c->n, t2, t3, c->vel, c->orig_valid ? c->orig : nullptr);
    CU_TRY(c, cudaGetLastError());
    return 0;
}

int lpe_bh_step(lpe_bh_ctx* c, const lpe_bh_params* p, int nsteps) {
    if (!c || !p) return 1;
    if (c->shard_n > 1) return fail(c, "sharded context: use lpe_bh_step_begin / lpe_bh_step_finish");
    DevGuard _dg(c->device);
    for (int s = 0; s < nsteps; ++s) {
        std::string key("step");
        key_add(key, *p); key_add(key, c->n); key_add(key, c->cap); key_add(key, c->instr); key_add(key, c->force_dfs);
        key_add(key, c->force_overflow); key_add(key, c->body); key_add(key, c->orig_valid); key_add(key, c->stream);
        if (run_graphed(c, key, [&]() { return run_step(c, *p, false); })) return 1;
    }
    return 0;
}

int lpe_bh_download(lpe_bh_ctx* c, double* x, double* y, double* vx, double* vy) {
    if (!c) return 1;
    if (c->dd) return fail(c, "context is in domain-decomposed mode: use lpe_bh_dd_download");
    DevGuard _dg(c->device);
    const uint64_t n = c->n;
    if (n) {
        cudaStream_t st = c->stream;
        const size_t bytes = sizeof(double) * n;
        const int g = cdiv((long long)n, 256);
        double *t0 = c->tmp, *t1 = c->tmp + c->cap, *t2 = c->tmp + 2 * c->cap, *t3 = c->tmp + 3 * c->cap;
        // one 32-byte record per body into creation order, then split into the four arrays (see k_finish_tick)
        double4* stage = c->stage4;
        k_stage_state<<<g, 256, 0, st>>>((int)n, c->body, c->vel, c->orig_valid ? c->orig : nullptr, stage);
        k_unstage<false><<<g, 256, 0, st>>>((int)n, stage, x ? t0 : nullptr, y ? t1 : nullptr, vx ? t2 : nullptr, vy ? t3 : nullptr);
        if (x) CU_TRY(c, cudaMemcpyAsync(x, t0, bytes, cudaMemcpyDeviceToHost, st));
        if (y) CU_TRY(c, cudaMemcpyAsync(y, t1, bytes, cudaMemcpyDeviceToHost, st));
        if (vx) CU_TRY(c, cudaMemcpyAsync(vx, t2, bytes, cudaMemcpyDeviceToHost, st));
        if (vy) CU_TRY(c, cudaMemcpyAsync(vy, t3, bytes, cudaMemcpyDeviceToHost, st));
    }
    if (fetch_fault(c)) return 1;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    CU_TRY(c, cudaGetLastError());
    return check_fault(c);
}

int lpe_bh_synchronize(lpe_bh_ctx* c) {
    if (!c) return 1;
    DevGuard _dg(c->device);
    if (fetch_fault(c)) return 1;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    CU_TRY(c, cudaGetLastError());
    if (c->dd && dd_check_fault(c)) return 1;
    return check_fault(c);
}

int lpe_bh_update_host(lpe_bh_ctx* c, const lpe_bh_params* p, uint64_t n, double* x, double* y, double* vx, double* vy,
                       const double* m, const uint32_t* rank, const uint8_t* comp) {
    if (!c || !p) return 1;
    if (c->shard_n > 1) return fail(c, "sharded context: use lpe_bh_step_begin / lpe_bh_step_finish");
    if (!vx || !vy || n == 0) {
        if (lpe_bh_upload(c, n, x, y, vx, vy, m, rank, comp)) return 1;
        if (lpe_bh_step(c, p, 1)) return 1;
        return lpe_bh_download(c, p->do_drift ? x : nullptr, p->do_drift ? y : nullptr, vx, vy);
    }
    if (c->dd) return fail(c, "context is in domain-decomposed mode");
    if (n > LPE_MAX_BODIES) return fail(c, "too many bodies for one context (limit 2^28)");
    if (!x || !y || !m) return fail(c, "x, y and m are required");
    DevGuard _dg(c->device);
    return host_tick(c, p, n, false, x, y, vx, vy, m, rank, comp);
}

// The same tick for callers whose components are {x, y} records (EnTT pools): pos / vel are 2n doubles each.
int lpe_bh_update_host_aos(lpe_bh_ctx* c, const lpe_bh_params* p, uint64_t n, double* pos, double* vel, const double* m,
                           const uint32_t* rank, const uint8_t* comp) {
    if (!c || !p) return 1;
    if (c->dd) return fail(c, "context is in domain-decomposed mode");
    if (c->shard_n > 1) return fail(c, "sharded context: use lpe_bh_step_begin / lpe_bh_step_finish");
    if (n == 0) return 0;
    if (n > LPE_MAX_BODIES) return fail(c, "too many bodies for one context (limit 2^28)");
    if (!pos || !vel || !m) return fail(c, "pos, vel and m are required");
    DevGuard _dg(c->device);
    return host_tick(c, p, n, true, pos, nullptr, vel, nullptr, m, rank, comp);
}

// ---- the same tick in three calls, so that a caller that has to GATHER its components first (the ECS drop-in) can
// overlap that with the device: positions -> keys + sort are queued; masses -> the build; velocities -> kick, result.
int lpe_bh_tick_begin(lpe_bh_ctx* c, const lpe_bh_params* p, uint64_t n, const double* pos, const uint8_t* comp) {
    if (!c || !p) return 1;
    if (c->dd) return fail(c, "context is in domain-decomposed mode");
    if (c->shard_n > 1) return fail(c, "sharded context: use lpe_bh_step_begin / lpe_bh_step_finish");
    if (n == 0 || n > LPE_MAX_BODIES) return fail(c, "tick: 1 .. 2^28 bodies");
    if (!pos) return fail(c, "pos is required");
    DevGuard _dg(c->device);
    if (ensure_capacity(c, n)) return 1;
    c->n = n;
    c->have_step = false;
    c->orig_valid = false;
    c->tick_stage = 0;
    c->defer_kick = false;
    c->tick_has_comp = comp != nullptr;
    if (make_const(c, *p, c->tick_k)) return 1;
    c->tick_p = *p;
    cudaStream_t st = c->stream;
    const int g = cdiv((long long)n, 256);
    c->tracing = getenv("LPE_TICK_TRACE") != nullptr;
    if (c->tracing) {
        if (!c->trace_ev[0]) for (int z = 0; z < 7; ++z) cudaEventCreate(&c->trace_ev[z]);
        cudaEventRecord(c->trace_ev[0], st);
        for (int z = 1; z < 5; ++z) cudaEventRecord(c->trace_ev[z], st);   // (placeholders: the copy-stream marks of the one-call tick)
        c->instr |= 1;
    }
    // Each of the three calls is one CUDA graph (copies included) when the caller's staging buffers are page-locked and
    // stay where they are from tick to tick, as the ECS drop-in's do: see run_graphed.
    auto queue = [&]() -> int {
        if (c->instr & 1) cudaEventRecord(c->ev[0], st);
        CU_TRY(c, cudaMemcpyAsync(c->tmp, pos, 16 * n, cudaMemcpyHostToDevice, st));
        if (comp) CU_TRY(c, cudaMemcpyAsync(c->comp_in, comp, n, cudaMemcpyHostToDevice, st));
        CU_TRY(c, cudaMemsetAsync(c->scal, 0, 8, st));   // the mass scale is rebuilt by lpe_bh_tick_mass
        k_pack_pos_aos<<<g, 256, 0, st>>>((int)n, reinterpret_cast<const double2*>(c->tmp), comp ? c->comp_in : nullptr, c->body);
        if (step_prologue(c, (int)n)) return 1;
        k_keygen<<<g, 256, 0, st>>>(c->tick_k, c->body, c->keys[0], c->vals[0], c->scal);
        c->launches += 2;
        if (c->instr & 1) cudaEventRecord(c->ev[1], st);
        if (step_sort(c, c->tick_k, (int)n)) return 1;
        if (c->instr & 1) cudaEventRecord(c->ev[2], st);
        CU_TRY(c, cudaGetLastError());
        return 0;
    };
    int rc;
    if (graphs_on(c) && is_pinned(pos) && is_pinned(comp)) {
        std::string key("tick1");
        key_add(key, *p); key_add(key, n); key_add(key, c->cap); key_add(key, c->instr); key_add(key, c->body); key_add(key, c->stream);
        key_add(key, pos); key_add(key, comp);
        rc = run_graphed(c, key, queue);
        c->have_step = false;   // (a replay marks the step complete; this is a third of one)
    } else {
        rc = queue();
    }
    if (rc) return 1;
    c->tick_stage = 1;
    return 0;
}

int lpe_bh_tick_mass(lpe_bh_ctx* c, const double* m, const uint32_t* rank) {
    if (!c) return 1;
    if (c->tick_stage != 1) return fail(c, "lpe_bh_tick_mass: call lpe_bh_tick_begin first");
    if (!m) return fail(c, "m is required");
    DevGuard _dg(c->device);
    cudaStream_t st = c->stream, cs = c->copy_stream;
    const uint64_t n = c->n;
    // the masses go up on the copy stream, beside the key generation and the sort queued by lpe_bh_tick_begin; the build
    // packs them on its side stream right before the gather (pend_mass), the kernels that only need keys do not wait
    auto queue = [&]() -> int {
        if (c->capturing) {   // the copy stream joins the capture (a plain call leaves it free-running: its buffer is idle)
            CU_TRY(c, cudaEventRecord(c->evc[0], st));
            CU_TRY(c, cudaStreamWaitEvent(cs, c->evc[0], 0));
        }
        CU_TRY(c, cudaMemcpyAsync(c->tmp + 2 * c->cap, m, 8 * n, cudaMemcpyHostToDevice, cs));
        if (rank) CU_TRY(c, cudaMemcpyAsync(c->rank_in, rank, 4 * n, cudaMemcpyHostToDevice, cs));
        CU_TRY(c, cudaEventRecord(c->evc[2], cs));
        c->pend_mass = true;
        c->pend_rank = rank != nullptr;
        c->pend_vel = true;    // the gather leaves the velocities alone: they arrive with lpe_bh_tick_finish
        const int rc = step_build(c, c->tick_k, (int)n);
        c->pend_vel = false;
        if (rc) return 1;
        if (c->instr & 1) cudaEventRecord(c->ev[3], st);
        // FAST precision walks the tree without the velocities (the kick is deferred, see k_finish_tick): the traversal is
        // queued here, behind the build, and runs while the caller is still staging its Velocity pool
        c->defer_kick = c->tick_p.precision == LPE_PREC_FAST;
        if (c->defer_kick) {
            if (step_traverse(c, c->tick_k, c->tick_p, (int)n, false)) { c->defer_kick = false; return 1; }
            if (c->instr & 1) cudaEventRecord(c->ev[4], st);
        }
        CU_TRY(c, cudaGetLastError());
        c->launches += 1;
        return 0;
    };
    int rc;
    if (graphs_on(c) && is_pinned(m) && is_pinned(rank)) {
        std::string key("tick2");
        key_add(key, c->tick_p); key_add(key, n); key_add(key, c->cap); key_add(key, c->instr); key_add(key, c->body); key_add(key, c->stream);
        key_add(key, m); key_add(key, rank);
        key_add(key, c->force_dfs); key_add(key, c->force_overflow);
        rc = run_graphed(c, key, queue);
        c->have_step = false;
    } else {
        rc = queue();
    }
    if (rc) { c->pend_mass = c->pend_vel = c->defer_kick = false; cudaStreamSynchronize(cs); return 1; }
    c->defer_kick = c->tick_p.precision == LPE_PREC_FAST;   // (a replay resets the flag with the other pending-work flags)
    c->tick_stage = 2;
    return 0;
}

int lpe_bh_tick_finish(lpe_bh_ctx* c, double* pos, double* vel) {
    if (!c) return 1;
    if (c->tick_stage != 2) return fail(c, "lpe_bh_tick_finish: call lpe_bh_tick_begin and lpe_bh_tick_mass first");
    if (!vel) return fail(c, "vel is required");
    c->tick_stage = 0;
    DevGuard _dg(c->device);
    cudaStream_t st = c->stream, cs = c->copy_stream;
    const uint64_t n = c->n;
    const int g = cdiv((long long)n, 256);
    double* t = c->tmp;
    const size_t cap = c->cap;
    const bool want_pos = c->tick_p.do_drift && pos;
    if (c->defer_kick) {
        // FAST precision: the traversal was queued by lpe_bh_tick_mass and stores velocity changes (k_finish_tick); the
        // velocities go up on the copy stream beside it. Five plain operations: not worth a graph, and a graph on the
        // context's stream could not start its copy before the traversal has finished.
        CU_TRY(c, cudaMemcpyAsync(t + 3 * cap, vel, 16 * n, cudaMemcpyHostToDevice, cs));
        CU_TRY(c, cudaEventRecord(c->evc[3], cs));
        c->last_c = c->tick_k;
        c->have_step = true;
        c->last.depth = c->tick_k.D;
        c->last.hilbert = c->tick_k.hilbert;
        CU_TRY(c, cudaStreamWaitEvent(st, c->evc[3], 0));
        if (queue_deferred_finish(c, true, c->tick_has_comp, want_pos, pos, nullptr, vel, nullptr)) {
            c->defer_kick = false;
            cudaStreamSynchronize(cs);
            return 1;
        }
        return wait_deferred_finish(c);
    }
    // STRICT precision sums in the reference's order starting from the velocity itself: upload, pack, walk, download
    const unsigned int* orig = c->orig_valid ? c->orig : nullptr;
    c->last_c = c->tick_k;
    c->have_step = true;
    c->last.depth = c->tick_k.D;
    c->last.hilbert = c->tick_k.hilbert;
    CU_TRY(c, cudaMemcpyAsync(t + 3 * cap, vel, 16 * n, cudaMemcpyHostToDevice, st));
    k_pack_vel_aos<<<g, 256, 0, st>>>((int)n, reinterpret_cast<const double2*>(t + 3 * cap), c->vel, orig);
    if (step_traverse(c, c->tick_k, c->tick_p, (int)n, false)) return 1;
    if (c->instr & 1) cudaEventRecord(c->ev[4], st);
    if (want_pos) {
        k_get_pos_aos<<<g, 256, 0, st>>>((int)n, c->body, reinterpret_cast<double2*>(t), orig);
        CU_TRY(c, cudaMemcpyAsync(pos, t, 16 * n, cudaMemcpyDeviceToHost, st));
    }
    k_unpack_vel_aos<<<g, 256, 0, st>>>((int)n, c->vel, reinterpret_cast<double2*>(t + 3 * cap), orig);
    CU_TRY(c, cudaMemcpyAsync(vel, t + 3 * cap, 16 * n, cudaMemcpyDeviceToHost, st));
    c->launches += 3;
    if (fetch_fault(c)) return 1;
    CU_TRY(c, cudaStreamSynchronize(st));
    CU_TRY(c, cudaGetLastError());
    return check_fault(c);
}

int lpe_bh_get_stats(lpe_bh_ctx* c, lpe_bh_stats* out) {
    if (!c || !out) return 1;
    DevGuard _dg(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    lpe_bh_stats s = c->last;
    s.n_bodies = c->n;
    if (c->have_step && c->n) {
        Scal h;
        CU_TRY(c, cudaMemcpy(&h, c->scal, sizeof(h), cudaMemcpyDeviceToHost));
        s.n_in_tree = h.n_in;
        s.n_terminals = h.n_term;
        s.n_nodes = (uint64_t)h.n_term + h.n_internal;
        s.interactions = h.interactions;
        s.visits = h.visits;
        s.warp_visits = h.warp_visits;
        s.overflow_chunks = h.ovf_count;
        s.force_sum = h.force_sum;
        { double fm; std::memcpy(&fm, &h.force_max_bits, sizeof(fm)); s.force_max = fm; }
        for (int z = 0; z < 8; ++z) s.t2_kinds[z] = h.t2[z];
        s.ms_keygen = s.ms_sort = s.ms_build = s.ms_traverse = s.ms_total = 0.f;   // (only with timing instrumentation)
        if (c->instr & 1) {
            cudaEventElapsedTime(&s.ms_keygen, c->ev[0], c->ev[1]);
            cudaEventElapsedTime(&s.ms_sort, c->ev[1], c->ev[2]);
            cudaEventElapsedTime(&s.ms_build, c->ev[2], c->ev[3]);
            cudaEventElapsedTime(&s.ms_traverse, c->ev[3], c->ev[4]);
            cudaEventElapsedTime(&s.ms_total, c->ev[0], c->ev[4]);
        }
    }
    *out = s;
    return 0;
}

int lpe_bh_dump_tree(lpe_bh_ctx* c, lpe_bh_tree_dump* o) {
    if (!c || !o) return 1;
    if (!c->have_step) return fail(c, "no step has been run since the last upload");
    if (c->dd) return fail(c, "lpe_bh_dump_tree: not available for a domain-decomposed rank");
    DevGuard _dg(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    const size_t n = c->n;
    Scal h;
    CU_TRY(c, cudaMemcpy(&h, c->scal, sizeof(h), cudaMemcpyDeviceToHost));
    const size_t nt = h.n_term, nn = (size_t)h.n_term + h.n_internal;
    if (o->sorted_keys) {
        if (c->last_c.k32) {   // 32-bit containers + "outside" flag in the payload -> the documented 64-bit form
            std::vector<unsigned int> k32(n), v32(n);
            CU_TRY(c, cudaMemcpy(k32.data(), c->keys[c->sorted_sel], 4 * n, cudaMemcpyDeviceToHost));
            CU_TRY(c, cudaMemcpy(v32.data(), c->vals[c->sorted_sel], 4 * n, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < n; ++i)
                o->sorted_keys[i] = (v32[i] & LPE_VAL_OUT) ? ((1ull << (2 * c->last_c.D)) | k32[i]) : (uint64_t)k32[i];
        } else {
            CU_TRY(c, cudaMemcpy(o->sorted_keys, c->keys[c->sorted_sel], 8 * n, cudaMemcpyDeviceToHost));
        }
    }
    std::vector<unsigned int> sidx(n);
    // creation index of the body at each sorted position (after a step the state is in sorted order itself)
    CU_TRY(c, cudaMemcpy(sidx.data(), c->orig, 4 * n, cudaMemcpyDeviceToHost));
    if (o->sorted_index) std::memcpy(o->sorted_index, sidx.data(), 4 * n);
    if (nn == 0) return 0;
    // The device keeps no per-node topology array: a node's level, first terminal and extent follow from the
    // terminals' level masks and their prefix sums (bh_build.cuh), which is what this test-only call rebuilds.
    std::vector<Agg> ag(nn);
    std::vector<unsigned long long> tk(nt);
    std::vector<unsigned int> tf(nt + 1), mk(nt), P(nt + 1);
    std::vector<Body> sb(n);
    CU_TRY(c, cudaMemcpy(ag.data(), c->agg, sizeof(Agg) * nn, cudaMemcpyDeviceToHost));
    CU_TRY(c, cudaMemcpy(tk.data(), c->tkey, 8 * nt, cudaMemcpyDeviceToHost));
    CU_TRY(c, cudaMemcpy(tf.data(), c->tfirst, 4 * (nt + 1), cudaMemcpyDeviceToHost));
    CU_TRY(c, cudaMemcpy(mk.data(), c->mask, 4 * nt, cudaMemcpyDeviceToHost));
    CU_TRY(c, cudaMemcpy(P.data(), c->P, 4 * (nt + 1), cudaMemcpyDeviceToHost));
    // the bodies as the tree saw them: the step's drift has moved the key-ordered state since, but the buffer the gather
    // read from (now body2) still holds the pre-step bodies, and the sort's payload maps sorted position -> old slot
    {
        std::vector<Body> old(n);
        std::vector<unsigned int> perm(n);
        CU_TRY(c, cudaMemcpy(old.data(), c->body2, sizeof(Body) * n, cudaMemcpyDeviceToHost));
        CU_TRY(c, cudaMemcpy(perm.data(), c->vals[c->sorted_sel], 4 * n, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; ++i) sb[i] = old[perm[i] & ~LPE_VAL_OUT];
    }
    const int D = c->last_c.D;
    // first node that starts at terminal t (its shallowest cell, or the terminal itself); nn past the last terminal
    auto nodeAt = [&](size_t t) -> size_t { return t >= nt ? nn : t + P[t]; };
    for (size_t t = 0; t < nt; ++t) {
        const size_t base = t + P[t];
        unsigned int rest = mk[t];
        size_t j = 0;
        auto emit = [&](size_t i, int level, size_t tEnd, const Agg& a) {
            double M, cx, cy;
            node_centre(a, level, c->last_c.quirk, M, cx, cy);
            if (o->node_level) o->node_level[i] = level;
            if (o->node_key) o->node_key[i] = tk[t];
            if (o->node_skip) o->node_skip[i] = (uint32_t)nodeAt(tEnd);
            if (o->node_first) o->node_first[i] = sidx[a.fidx];
            if (o->node_count) o->node_count[i] = tf[tEnd] - tf[t];
            if (o->node_mass) o->node_mass[i] = M;
            if (o->node_comx) o->node_comx[i] = cx;
            if (o->node_comy) o->node_comy[i] = cy;
        };
        while (rest) {   // the branching cells that start at t, shallow to deep
            const int L = __builtin_ctz(rest);
            rest &= rest - 1;
            const int shift = 2 * (D - L);
            size_t tEnd = t + 1;   // one past the last terminal that shares the cell's level-L prefix
            {
                size_t lo = t, hi = nt;   // tk[lo] shares it, tk[hi] (or the end) does not
                while (hi - lo > 1) {
                    const size_t mid = lo + (hi - lo) / 2;
                    if ((tk[mid] >> shift) == (tk[t] >> shift)) lo = mid; else hi = mid;
                }
                tEnd = lo + 1;
            }
            emit(base + j, L, tEnd, ag[base + j]);
            ++j;
        }
        const size_t i = base + j;   // the terminal's own node
        const unsigned int first = tf[t], last = tf[t + 1];
        if (last - first == 1u) {    // single-body leaves keep no aggregate on the device: rebuild it from the body
            const Body& s0 = sb[first];
            Agg a{};
            a.m = s0.m; a.sx = s0.m * s0.x; a.sy = s0.m * s0.y; a.mf = s0.m; a.xf = s0.x; a.yf = s0.y;
            a.frank = s0.rank; a.fidx = first; a.ordinal = 0; a.small = 0;
            emit(i, -1, t + 1, a);
        } else {
            emit(i, -2, t + 1, ag[i]);
        }
    }
    return 0;
}

int lpe_bh_get_counts(lpe_bh_ctx* c, uint32_t* accepted, uint32_t* visited) {
    if (!c) return 1;
    if (!(c->instr & 2)) return fail(c, "enable instrumentation bit1 before the step");
    DevGuard _dg(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    for (int which = 0; which < 2; ++which) {
        uint32_t* dst = which ? visited : accepted;
        const unsigned int* src = which ? c->cntVis : c->cntAcc;
        if (!dst || c->n == 0) continue;
        if (c->orig_valid) {   // counters are per state slot: back to creation order
            unsigned int* t = reinterpret_cast<unsigned int*>(c->tmp);
            k_unpermute_u32<<<cdiv((long long)c->n, 256), 256, 0, c->stream>>>((int)c->n, src, t, c->orig);
            CU_TRY(c, cudaStreamSynchronize(c->stream));
            src = t;
        }
        CU_TRY(c, cudaMemcpy(dst, src, 4 * c->n, cudaMemcpyDeviceToHost));
    }
    return 0;
}

int lpe_bh_direct_accel(lpe_bh_ctx* c, const lpe_bh_params* p, uint64_t first, uint64_t count, double* ax, double* ay) {
    if (!c || !p || !ax || !ay) return 1;
    if (first + count > c->n) return fail(c, "target range out of bounds");
    if (count == 0) return 0;
    DevGuard _dg(c->device);
    double *dax = c->tmp, *day = c->tmp + c->cap;
    k_direct<<<cdiv((long long)(c->orig_valid ? c->n : count), 256), 256, 0, c->stream>>>(
        (int)c->n, c->body, p->universe_size, p->softening * p->softening, p->G, (int)first, (int)count, dax, day,
        c->orig_valid ? c->orig : nullptr);
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(ax, dax, 8 * count, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(ay, day, 8 * count, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// the two per-tick host scans over Mass — BarnesHutSystem's early exit "no mass reaches
// smallMassThreshold" (barnes_hut.cpp:55-71) and BasicGravitySystem's "some mass >= 1e10 disables the uniform field"
// (gravity.cpp:41-49) — both only ask for the LARGEST mass among the non-Boundary bodies that have one. The upload
// kernels reduce exactly that on the device (Scal::max_mass_bits, it also fixes the traversal's mass unit), so for
// resident bodies the answer is one 8-byte read. Synchronises.
int lpe_bh_max_source_mass(lpe_bh_ctx* c, double* max_mass) {
    if (!c || !max_mass) return 1;
    *max_mass = 0.0;
    if (!c->scal || c->n == 0) return 0;
    DevGuard _dg(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    unsigned long long bits = 0;
    CU_TRY(c, cudaMemcpy(&bits, &c->scal->max_mass_bits, sizeof(bits), cudaMemcpyDeviceToHost));
    std::memcpy(max_mass, &bits, sizeof(bits));
    return 0;
}

// ---------------------------------------------------------------------------------------------- multi-GPU
uint64_t lpe_bh_shard_chunk(uint64_t n, int nranks) {
    if (nranks < 1) nranks = 1;
    const uint64_t nblocks = (n + LPE_SHARD_BLOCK - 1) / LPE_SHARD_BLOCK;
    const uint64_t per = (nblocks + (uint64_t)nranks - 1) / (uint64_t)nranks;
    return per * LPE_SHARD_BLOCK;
}

int lpe_bh_shard_owner(uint64_t pos, int nranks, int* rank_out, uint64_t* slot_out) {
    if (nranks < 1) return 1;
    const uint64_t gblock = pos / LPE_SHARD_BLOCK;
    if (rank_out) *rank_out = (int)(gblock % (uint64_t)nranks);
    if (slot_out) *slot_out = (gblock / (uint64_t)nranks) * LPE_SHARD_BLOCK + pos % LPE_SHARD_BLOCK;
    return 0;
}

int lpe_bh_set_shard(lpe_bh_ctx* c, int rank, int nranks) {
    if (!c) return 1;
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(c, "bad shard rank / nranks");
    c->shard_rank = rank;
    c->shard_n = nranks;
    return 0;
}

int lpe_bh_step_begin(lpe_bh_ctx* c, const lpe_bh_params* p) {
    if (!c || !p) return 1;
    DevGuard _dg(c->device);
    return run_step(c, *p, true);
}

int lpe_bh_step_finish(lpe_bh_ctx* c) {
    if (!c) return 1;
    if (c->shard_n <= 1 || c->n == 0) return 0;
    if (!c->have_step) return fail(c, "lpe_bh_step_begin has not run");
    DevGuard _dg(c->device);
    k_xchg_scatter<<<cdiv((long long)c->n, 256), 256, 0, c->stream>>>((int)c->n, c->shard_n, c->xchg_chunk,
                                                                       nullptr,
                                                                       c->xchg_recv + (p2p_ready(c) ? (size_t)c->xchg_parity * c->xchg_chunk * (size_t)c->shard_n : 0),
                                                                       c->body, c->vel);
    CU_TRY(c, cudaGetLastError());
    return 0;
}

int lpe_bh_xchg_read_send(lpe_bh_ctx* c, double* host) {
    if (!c || !host) return 1;
    if (c->shard_n <= 1 || !c->xchg_send) return fail(c, "context is not sharded");
    DevGuard _dg(c->device);
    CU_TRY(c, cudaMemcpyAsync(host, c->xchg_send, sizeof(double4) * c->xchg_chunk, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int lpe_bh_xchg_write_recv(lpe_bh_ctx* c, int src, const double* host) {
    if (!c || !host) return 1;
    if (c->shard_n <= 1 || !c->xchg_recv) return fail(c, "context is not sharded");
    if (src < 0 || src >= c->shard_n) return fail(c, "source rank out of range");
    DevGuard _dg(c->device);
    CU_TRY(c, cudaMemcpyAsync(c->xchg_recv + (size_t)src * c->xchg_chunk, host, sizeof(double4) * c->xchg_chunk,
                              cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int lpe_bh_boundary(lpe_bh_ctx* c, const lpe_bh_boundary_params* p) {
    if (!c || !p) return 1;
    if (!(p->universe_size > 0.0) || !(p->margin >= 0.0)) return fail(c, "boundary: universe_size must be > 0 and margin >= 0");
    if (c->n == 0) return 0;
    DevGuard _dg(c->device);
    // the pass is per body and order-free: it runs on the state in whatever order it currently is
    k_boundary<<<cdiv((long long)c->n, 256), 256, 0, c->stream>>>((int)c->n, c->body, c->vel, p->margin, p->universe_size,
                                                                   p->bounce_damping, p->max_speed);
    CU_TRY(c, cudaGetLastError());
    c->launches += 1;
    return 0;
}

// ---- direct exchange over peer memory (NVLink / NVSwitch): handles travel through the caller's own channel ----
int lpe_bh_xchg_export(lpe_bh_ctx* c, void* handle64) {
    if (!c || !handle64) return 1;
    if (c->shard_n <= 1 || c->n == 0) return fail(c, "context is not sharded (set_shard + upload first)");
    if (c->shard_n > LPE_MAX_P2P) return fail(c, "direct exchange supports at most 8 ranks");
    DevGuard _dg(c->device);
    if (ensure_xchg(c)) return 1;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
    cudaIpcMemHandle_t hnd;
    CU_TRY(c, cudaIpcGetMemHandle(&hnd, c->xchg_recv));
    std::memcpy(handle64, &hnd, sizeof(hnd));
    c->peer_recv[c->shard_rank] = c->xchg_recv;
    return 0;
}

int lpe_bh_xchg_import(lpe_bh_ctx* c, int rank, const void* handle64) {
    if (!c || !handle64) return 1;
    if (c->shard_n <= 1 || c->shard_n > LPE_MAX_P2P) return fail(c, "direct exchange needs 2..8 ranks");
    if (rank < 0 || rank >= c->shard_n || rank == c->shard_rank) return fail(c, "bad peer rank");
    DevGuard _dg(c->device);
    cudaIpcMemHandle_t hnd;
    std::memcpy(&hnd, handle64, sizeof(hnd));
    void* ptr = nullptr;
    CU_TRY(c, cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
    if (c->peer_opened[rank]) cudaIpcCloseMemHandle(c->peer_opened[rank]);
    c->peer_opened[rank] = ptr;
    c->peer_recv[rank] = static_cast<double4*>(ptr);
    return 0;
}

int lpe_bh_xchg_set_peer(lpe_bh_ctx* c, int rank, void* recv_device_ptr) {
    if (!c) return 1;
    if (c->shard_n <= 1 || c->shard_n > LPE_MAX_P2P) return fail(c, "direct exchange needs 2..8 ranks");
    if (rank < 0 || rank >= c->shard_n) return fail(c, "bad peer rank");
    c->peer_recv[rank] = static_cast<double4*>(recv_device_ptr);
    return 0;
}

int lpe_bh_xchg_p2p_ready(const lpe_bh_ctx* c) { return (c && p2p_ready(c)) ? 1 : 0; }

int lpe_bh_xchg_reset(lpe_bh_ctx* c) {
    if (!c) return 1;
    DevGuard _dg(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    close_peers(c);
    return 0;
}

uint64_t lpe_bh_launch_count(const lpe_bh_ctx* c) { return c ? c->launches : 0; }
uint64_t lpe_bh_graph_replays(const lpe_bh_ctx* c) { return c ? c->graph_replays : 0; }
#ifdef SORT_TRACE
void lpe_bh_debug_sort_trace(unsigned long long* out8, int reset) {   // (variant builds only)
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out8, lpe::g_sort_trace, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {}; cudaMemcpyToSymbol(lpe::g_sort_trace, z, sizeof(z)); }
}
#endif

int lpe_bh_fma_peak(lpe_bh_ctx* c, double* tflops) {
    if (!c || !tflops) return 1;
    DevGuard _dg(c->device);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    float* sink = nullptr;
    CU_TRY(c, cudaMalloc(&sink, sizeof(float) * 1024));
    const int iters = 1 << 14, blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, c->stream);
        k_fma_peak<<<blocks, threads, 0, c->stream>>>(iters, 1.0001f, sink);
        cudaEventRecord(e1, c->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    CU_TRY(c, cudaGetLastError());
    *tflops = best;
    return 0;
}

void* lpe_bh_alloc_pinned(uint64_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void lpe_bh_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

int lpe_bh_get_device_view(lpe_bh_ctx* c, lpe_bh_device_view* o) {
    if (!c || !o) return 1;
    if (c->shard_n > 1 && c->n && ensure_xchg(c)) return 1;
    o->body = c->body;
    o->vel = c->vel;
    o->orig = c->orig;
    o->key_ordered = c->orig_valid ? 1 : 0;
    o->pad_ = 0;
    o->xchg_send = c->xchg_send;
    o->xchg_recv = c->xchg_recv;
    o->n = c->n;
    o->xchg_chunk = c->xchg_chunk;
    return 0;
}

}  // extern "C"

#include "lpe_bh_dd.inl"
