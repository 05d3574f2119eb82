// bh_dd.cuh — domain-decomposed Barnes-Hut step across the GPUs of one NVLink / NVSwitch domain.
//
// The reference builds ONE tree per step (barnes_hut.cpp:101-131). Rebuilding it in full on every GPU caps the scaling
// (round 1: 3.5 of the 5.0 ms 8-GPU step at 16 M bodies were replicated sort + build), so here every phase is sharded:
//
//   ownership   rank r owns the bodies whose sort key (Hilbert / Morton index of the depth-D cell) lies in
//               [K_r, K_r+1); the splitters K are arbitrary keys (cost-balanced), the bodies live only on their owner.
//   phase A     keys of the own bodies; a body whose key left the range is stored straight into its new owner's
//               state arrays (peer memory, slots handed out from the end of the array by an atomic counter).
//   phase B     sort + build of the OWN bodies only. Every cell whose key interval lies inside the rank's range is
//               complete and identical to the same cell of the single-GPU tree. The cells that straddle a splitter
//               (at most D per splitter) form the TOP of the tree: each rank publishes the roots of its maximal inner
//               quadrants (<= 6 per level) to every rank, and exports to rank d the child blocks of those of its cells
//               that some body of d's key range could OPEN (a conservative box test against d's quadrants: the
//               locally essential tree) — plain stores into d's record array over NVLink.
//   phase C     every rank builds the top of the tree from the published roots (same child order and the same
//               (c0 + c1) + (c2 + c3) sums as the single-GPU build), then traverses for its own bodies: the records a
//               warp can reach are exactly the ones it would reach in the single-GPU tree, with the same contents and
//               the same sibling order, so every accept / open decision is the reference's.
//   Two in-stream barriers per step (after A and after B): one flag per peer in peer memory, no collective.
#pragma once
#include "bh_common.cuh"
#include "bh_build.cuh"
#include "bh_sort.cuh"
#include "bh_traverse.cuh"

namespace lpe {

constexpr int DD_MAXQ = 6 * LPE_MAX_DEPTH + 8;          // quadrants of one rank's key range (<= 3 per level and side)
constexpr int DD_MAXROOTS = DD_MAXQ;                    // published roots per rank
constexpr int DD_TOPROOTS = DD_MAXROOTS * LPE_MAX_P2P;  // leaves of the top tree
constexpr unsigned int DD_POISON_BLOCK = 1u;            // child block of a cell that was not exported: NaN leaves
constexpr unsigned int DD_TOPB = 2u;                    // top cell with ordinal q has child block DD_TOPB + q
constexpr unsigned int DD_TOPCAP = DD_TOPB + DD_TOPROOTS;   // first local block (blockBase of the local build)

struct DDQuad {                 // one maximal aligned quadrant of a rank's key range
    unsigned long long key;     // first depth-D key
    int level;                  // quadtree level of the quadrant (0 = the universe)
    int pad;
    double x0, y0, x1, y1;      // [x0, x1) x [y0, y1): every in-tree body of the quadrant lies inside (keygen's own bounds)
};
constexpr int DD_BVH_LEAVES = 256;                      // >= DD_MAXQ, power of two
struct DDDomain {               // identical on every rank; rebuilt on the host when the splitters or the depth change
    int nq[LPE_MAX_P2P];
    double box[LPE_MAX_P2P][4];                 // bounding box of the rank's quadrants: x0, y0, x1, y1
    DDQuad q[LPE_MAX_P2P][DD_MAXQ];
    // The same quadrants as a binary tree of boxes (heap order: node k has children 2k and 2k+1, leaf j is node
    // DD_BVH_LEAVES + j), in scaled units (x / S) as floats rounded OUTWARD; x0 > x1 marks an empty node. The quadrants
    // are in key order, so neighbours in the list are neighbours in space and the inner boxes are tight.
    float4 bvh[LPE_MAX_P2P][2 * DD_BVH_LEAVES];
};
struct DDSplit {                // depth-D splitters: rank r owns keys in [k[r], k[r+1]); k[R] = all ones
    unsigned long long k[LPE_MAX_P2P + 1];
    int me, R;
};

struct __align__(16) DDRoot {   // root of one non-empty inner quadrant, as published to every rank
    Agg agg;                    // exact sums (a single-body leaf: the body itself)
    unsigned long long key;     // first depth-D key of the quadrant
    int level;                  // level of the node: >= 0 branching cell, -1 single-body leaf, -2 aggregated terminal
    unsigned int cblock;        // child block of the cell ON THE RECEIVING RANK (DD_POISON_BLOCK if it was not exported)
    unsigned int leafpos;       // single-body leaf: sorted position of the body on its owner, else LPE_NONE
    unsigned int owner;
    unsigned int pad[2];
};
static_assert(sizeof(DDRoot) == 96, "DDRoot layout");

struct __align__(16) DDMail {   // one barrier point, one sender: flag = epoch of the step, plus a small payload
    unsigned long long flag;
    unsigned long long pad;
    double payload[6];
};
struct __align__(16) DDHeader { // head of every rank's window (peer-visible)
    unsigned int inbox_count;   // migrants stored into this rank's tail slots in the current step (peers atomicAdd)
    unsigned int fault;         // bit0 inbox overflow, bit1 import overflow, bit2 root table overflow, bit3 barrier timeout,
                                // bit4 migrant arrived outside the rank's range, bit5 frontier overflow
    unsigned int n_live;        // live bodies (slots [0, n_live) of the current state buffers)
    unsigned int pad;
    unsigned int root_count[LPE_MAX_P2P];   // roots published by each sender
    DDMail mail[2][LPE_MAX_P2P];
};

struct DDPeers {                // every rank's window pieces (own entries included), current buffer parity
    DDHeader* hdr[LPE_MAX_P2P];
    Body* body[LPE_MAX_P2P];
    double2* vel[LPE_MAX_P2P];
    unsigned int* orig[LPE_MAX_P2P];
    DDRoot* roots[LPE_MAX_P2P];     // [sender][DD_MAXROOTS]
    TravRec* rec[LPE_MAX_P2P];
    double4* xrec[LPE_MAX_P2P];
};

// ---- sys-scope flag access for the barriers -----------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Signal: lane s stores this rank's payload and then the step's epoch into rank s's mailbox (own mailbox included).
// Everything this rank stored into peer memory in earlier kernels of the stream is complete by then (kernel boundary
// + fence), so a rank that sees the flag sees the data.
// (the step number lives in device memory, advanced by k_dd_next_step: a captured CUDA graph of the step replays unchanged)
__global__ void k_dd_next_step(unsigned long long* __restrict__ step) { *step += 1ull; }
__global__ void k_dd_signal(int point, const unsigned long long* __restrict__ step, int me, int R, DDPeers peers, const double* __restrict__ payload) {
    const int s = threadIdx.x;
    if (s >= R) return;
    const unsigned long long epoch = *step;
    DDMail* m = &peers.hdr[s]->mail[point][me];
    if (payload)
        for (int k = 0; k < 6; ++k) m->payload[k] = payload[k];
    __threadfence_system();
    st_release_sys(&m->flag, epoch);
}
// Wait: lane s spins until rank s's flag in the own mailbox carries this step's epoch. Bounded: a dead peer becomes an
// error flag, not a hung device.
__global__ void k_dd_wait(int point, const unsigned long long* __restrict__ step, int R, DDHeader* hdr) {
    const int s = threadIdx.x;
    if (s >= R) return;
    const unsigned long long epoch = *step;
    const unsigned long long* f = &hdr->mail[point][s].flag;
    unsigned long long spins = 0;
    while (ld_acquire_sys(f) < epoch) {
        __nanosleep(200);
        if (++spins > (1ull << 27)) {   // about half a minute: a peer that is merely slow (host stall) must not trip it
            atomicOr(&hdr->fault, 8u);
            break;
        }
    }
}

__device__ __forceinline__ unsigned long long ordered_bits(double v) {   // monotone map double -> u64
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double ordered_value(unsigned long long o) {
    const unsigned long long b = (o >> 63) ? (o & 0x7FFFFFFFFFFFFFFFull) : ~o;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, sizeof(d));
    return d;
#endif
}

__device__ __forceinline__ int dd_owner(const DDSplit& sp, unsigned long long key) {
    int r = 0;
#pragma unroll
    for (int k = 1; k < LPE_MAX_P2P; ++k) r += (k < sp.R && key >= sp.k[k]) ? 1 : 0;
    return r;
}

// sort key of a body exactly as k_keygen computes it; `inside` = the body is a source inside [0,U)^2
__device__ __forceinline__ unsigned long long dd_body_key(const StepConst& c, const unsigned char* lut, double2 p,
                                                          unsigned int cm, bool& inside) {
    const bool src = (cm & 1u) && !(cm & 4u);
    inside = src && p.x >= 0.0 && p.x < c.U && p.y >= 0.0 && p.y < c.U;   // barnes_hut.cpp:117-124
    if (!inside) return 1ull << (2 * c.D);
    const unsigned int kmax = (1u << c.D) - 1u;
    const unsigned int ix = cell_index(p.x, c.h, c.invh, kmax);
    const unsigned int iy = cell_index(p.y, c.h, c.invh, kmax);
    if (!c.hilbert) return spread_bits32(ix) | (spread_bits32(iy) << 1);
    return lut ? hilbert_index_lut(lut, ix, iy, c.D) : hilbert_index(ix, iy, c.D);
}
__device__ __forceinline__ unsigned long long dd_dead_key(int D) { return (1ull << (2 * D)) | 1ull; }   // sorts last

// ---- phase A: keys of the own bodies; leavers go straight into their new owner's state arrays ----------------------
// Slots [0, n_live) hold the rank's bodies. A body whose key left the rank's range is stored right behind the live
// bodies of its new owner (slot = the owner's n_live + a ticket from the owner's counter): the owner then sorts
// n_live + arrivals slots, whatever the capacity. The slots behind n_live belong to the peers during this phase and are
// not touched here; their keys are made by k_dd_keygen_inbox after the barrier.
__global__ void __launch_bounds__(256)
k_dd_keygen(StepConst c, DDSplit sp, const Body* __restrict__ body, const double2* __restrict__ vel,
            const unsigned int* __restrict__ orig, void* __restrict__ keys, unsigned int* __restrict__ vals,
            Scal* __restrict__ s, DDHeader* __restrict__ hdr, DDPeers peers, unsigned long long* __restrict__ oob) {
    __shared__ unsigned char lut[64];
    if (threadIdx.x < 64) lut[threadIdx.x] = hilbert_lut_entry(threadIdx.x);
    __syncthreads();
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int n_prev = hdr->n_live;
    if (blockIdx.x * blockDim.x >= n_prev) return;
    unsigned int in = 0, live = 0;
    if (i < n_prev) {
        const Body b = body[i];
        bool inside;
        unsigned long long key = dd_body_key(c, lut, make_double2(b.x, b.y), b.comp, inside);
        const int owner = dd_owner(sp, key);
        const bool target = (b.comp & 1u) && (b.comp & 2u) && !(b.comp & 4u);
        if (!inside && target) {   // a target outside the tree: the last rank's domain must cover it
            atomicMin(&oob[0], ordered_bits(b.x)); atomicMin(&oob[1], ordered_bits(b.y));
            atomicMax(&oob[2], ordered_bits(b.x)); atomicMax(&oob[3], ordered_bits(b.y));
        }
        if (owner == sp.me) {
            live = 1; in = inside ? 1u : 0u;
        } else {
            DDHeader* oh = peers.hdr[owner];
            const unsigned int slot = oh->n_live + atomicAdd(&oh->inbox_count, 1u);
            if (slot < (unsigned int)c.n) {   // (the bounds check of the migration: a full owner raises fault bit 0)
                peers.body[owner][slot] = b;
                peers.vel[owner][slot] = vel[i];
                peers.orig[owner][slot] = orig[i];
            } else {
                atomicOr(&hdr->fault, 1u);
            }
            key = dd_dead_key(c.D);   // sorts behind everything: the slot is free again after the gather
        }
        store_key(keys, vals, c.k32, c.D, i, key);
    }
    const unsigned int cin = __syncthreads_count(in);
    const unsigned int clive = __syncthreads_count(live);
    if (threadIdx.x == 0) {
        if (cin) atomicAdd(&s->n_in, cin);
        if (clive) atomicAdd(&s->n_live, clive);
    }
}

// ---- phase B, first kernel: keys of the bodies that arrived behind the live ones; how many slots the sort covers -----
__global__ void __launch_bounds__(256)
k_dd_keygen_inbox(StepConst c, DDSplit sp, const Body* __restrict__ body, void* __restrict__ keys,
                  unsigned int* __restrict__ vals, Scal* __restrict__ s, DDHeader* __restrict__ hdr) {
    __shared__ unsigned char lut[64];
    if (threadIdx.x < 64) lut[threadIdx.x] = hilbert_lut_entry(threadIdx.x);
    __syncthreads();
    unsigned int cnt = hdr->inbox_count;
    const unsigned int n_prev = hdr->n_live;
    const unsigned int room = (unsigned int)c.n - n_prev;
    if (cnt > room) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&hdr->fault, 1u);
        cnt = room;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) s->n_sort = n_prev + cnt;
    unsigned int in = 0, live = 0;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < cnt; k += gridDim.x * blockDim.x) {
        const unsigned int slot = n_prev + k;
        const Body b = body[slot];
        bool inside;
        unsigned long long key = dd_body_key(c, lut, make_double2(b.x, b.y), b.comp, inside);
        if (dd_owner(sp, key) != sp.me) {   // cannot happen while all ranks use the same splitters
            atomicOr(&hdr->fault, 16u);
            key = dd_dead_key(c.D);
        } else {
            ++live;
            in += inside ? 1u : 0u;
        }
        store_key(keys, vals, c.k32, c.D, slot, key);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        in += __shfl_xor_sync(0xFFFFFFFFu, in, o);
        live += __shfl_xor_sync(0xFFFFFFFFu, live, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (in) atomicAdd(&s->n_in, in);
        if (live) atomicAdd(&s->n_live, live);
    }
}

// ---- phase B: which of the own cells can a body of rank d open? -------------------------------------------------------
// Conservative: the reference opens a cell for a body at p when !(size^2 / (|com - p|^2 + eps^2) < theta^2)
// (barnes_hut.cpp:261-269). For every p inside a box, |com - p|^2 >= dmin^2 (distance from com to the box), so if
// size^2 < theta^2 * (dmin^2 + eps^2) * (1 - 1e-9) no body of the box opens the cell and its children are never
// visited from there.
__device__ __forceinline__ bool dd_box_may_open(double cx, double cy, double sizeSq, double eps2, double theta2,
                                                double x0, double y0, double x1, double y1) {
    const double dx = fmax(fmax(x0 - cx, cx - x1), 0.0), dy = fmax(fmax(y0 - cy, cy - y1), 0.0);
    const double d2 = dx * dx + dy * dy + eps2;
    return !(sizeSq < theta2 * d2 * (1.0 - 1e-9));
}

// fp64 side record of a non-local slot: {centre x, centre y, mass, level} as the traversal's exact test wants them
__device__ __forceinline__ double4 dd_xrec(const Agg& a, int level, int quirk) {
    double M, cx, cy;
    node_centre(a, level, quirk, M, cx, cy);
    return make_double4(cx, cy, M, (double)level);
}

__device__ __forceinline__ unsigned int dd_lower_bound(const unsigned long long* __restrict__ a, unsigned int n, unsigned long long v) {
    unsigned int lo = 0, hi = n;
    while (lo < hi) {
        const unsigned int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---- phase B: the roots of the rank's non-empty inner quadrants (one block) ---------------------------------------------
struct DDRootsIn {
    const unsigned long long* tkey;
    const unsigned int* tfirst;
    const unsigned int* mask;
    const unsigned int* P;
    const Agg* agg;
    const Body* body;
};
// myroots[i].cblock = ORDINAL of the root cell (LPE_NONE for a leaf / aggregated terminal); the exporter turns it into
// the child block index each destination will see
__global__ void __launch_bounds__(256)
k_dd_roots(StepConst c, DDSplit sp, const DDDomain* __restrict__ dom, DDRootsIn a, DDRoot* __restrict__ myroots,
           Scal* __restrict__ s, DDHeader* __restrict__ hdr) {
    __shared__ unsigned int sh[9];
    const int j = threadIdx.x;
    const int nq = dom->nq[sp.me];
    const unsigned int n_term = s->n_term;
    bool nonempty = false;
    unsigned int t0 = 0, t1 = 0;
    DDQuad q{};
    if (j < nq) {
        q = dom->q[sp.me][j];
        const unsigned long long kend = q.key + (1ull << (2 * (c.D - q.level)));
        t0 = dd_lower_bound(a.tkey, n_term, q.key);
        t1 = dd_lower_bound(a.tkey, n_term, kend);
        nonempty = t1 > t0;
    }
    unsigned int total;
    const unsigned int pos = block_exclusive_scan_256(nonempty ? 1u : 0u, sh, &total);
    if (nonempty) {
        DDRoot r;
        r.key = q.key; r.owner = (unsigned int)sp.me; r.leafpos = LPE_NONE; r.pad[0] = r.pad[1] = 0u; r.cblock = LPE_NONE;
        if (t1 - t0 == 1u) {   // one terminal: a single-body leaf or an aggregated depth-D cell
            const unsigned int first = a.tfirst[t0], last = a.tfirst[t0 + 1];
            if (last - first == 1u) {
                r.level = -1; r.agg = body_agg(a.body[first], first, c.thr); r.leafpos = first;
            } else {
                r.level = -2; r.agg = a.agg[t0 + a.P[t0] + (unsigned int)__popc(a.mask[t0])];   // the terminal's own node
            }
        } else {               // the lowest cell that holds every body of the quadrant
            const int L = lca_level(a.tkey[t0], a.tkey[t1 - 1], c.D);
            const unsigned int ordinal = a.P[t0] + (unsigned int)__popc(a.mask[t0] & ((1u << L) - 1u));
            r.level = L; r.agg = a.agg[t0 + ordinal]; r.cblock = ordinal;
        }
        if (pos < (unsigned int)DD_MAXROOTS) myroots[pos] = r;
        else atomicOr(&hdr->fault, 4u);
    }
    if (j == 0) {
        s->dd_myroots = total < (unsigned int)DD_MAXROOTS ? total : (unsigned int)DD_MAXROOTS;
        hdr->n_live = s->n_live;     // the state is compact again: next step's phase A reads slots [0, n_live)
        hdr->inbox_count = 0u;       // consumed; the peers touch it again only after the barrier that follows
    }
}

// ---- phase B: publish the roots to rank d and export the child blocks rank d can reach ----------------------------------
// Breadth first from the roots: a cell's child block goes to rank d only if some body of d's key range could open the
// cell AND all its ancestors — exactly the records d's traversal can reach, never the rest of this rank's tree (the
// work is proportional to the exported part, a few thousand cells, not to the rank's million cells). The queue lives in
// global memory and doubles as the numbering: the cell at queue position e is exported as block e of this rank's
// import region on d, so a parent knows its children's block numbers the moment it pushes them.
// The reachability test below the roots is the traversal's own group classification (bh_traverse2.cuh, phase 1) against
// d's quadrants instead of a warp's bounding box: fp32 from the record alone, box edges rounded outward, the same
// safety margins — if it says "every body of the box accepts", every warp of rank d (whose box lies inside) says so
// too and never asks for the children. The quadrants are searched through a small tree of boxes (DDDomain::bvh).
// One thread-block CLUSTER per destination: the generations of the breadth-first walk are separated by cluster
// barriers, four lanes work on one exported cell (one per child record).
struct DDExportArgs {
    const DDRoot* myroots;
    const unsigned int* child;      // [4 * ordinal + digit]
    const TravRec* rec;             // own records (local blocks)
    const Agg* agg;
    const Body* body;
    unsigned int* queue;            // [dest][icap] ordinals of the exported cells
    unsigned int* pushed;           // [dest][DD_EXPORT_ROUNDS] "somebody pushed in this round"
    unsigned int icap, importBase;
};
constexpr int DD_EXPORT_ROUNDS = 2 * LPE_MAX_DEPTH + 4;   // (a round handles at least one generation of the walk)
constexpr int DD_EXPORT_THREADS = 1024;
constexpr int DD_EXPORT_CLUSTER = 8;
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __launch_bounds__(DD_EXPORT_THREADS)
k_dd_export(StepConst c, DDSplit sp, const DDDomain* __restrict__ dom, DDExportArgs a, DDPeers peers, Scal* __restrict__ s,
            DDHeader* __restrict__ hdr) {
    __shared__ float4 bvh[2 * DD_BVH_LEAVES];
    __shared__ float4 oboxf;
    __shared__ int haveOob;
    const int d = blockIdx.y;
    const int crank = blockIdx.x;                 // rank of this block inside its cluster
    const int tid = threadIdx.x;
    const unsigned int nroots = s->dd_myroots;
    const bool remote = d != sp.me;
    if (!remote && crank != 0) return;            // the own table is written by one block; nobody waits for the others
    if (tid == 0) {
        // targets outside the universe live on the last rank: union of every rank's box of such bodies
        double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
        bool any = false;
        if (d == sp.R - 1)
            for (int r = 0; r < sp.R; ++r) {
                const double* pl = hdr->mail[0][r].payload;
                if (pl[0] <= pl[2]) { any = true; x0 = fmin(x0, pl[0]); y0 = fmin(y0, pl[1]); x1 = fmax(x1, pl[2]); y1 = fmax(y1, pl[3]); }
            }
        oboxf = make_float4(__double2float_rd(x0 * c.invS), __double2float_rd(y0 * c.invS),
                            __double2float_ru(x1 * c.invS), __double2float_ru(y1 * c.invS));
        haveOob = any ? 1 : 0;
    }
    for (int k = tid; k < 2 * DD_BVH_LEAVES; k += blockDim.x) bvh[k] = dom->bvh[d][k];
    __syncthreads();
    const float eps2f = c.eps2f;
    auto boxAccepts = [&](const TravRec& R, const float4 bx) -> bool {   // every point of the box accepts the node
        const float ax0 = (R.c.x - bx.x) + R.c.z, ax1 = (R.c.x - bx.z) + R.c.z;
        const float ay0 = (R.c.y - bx.y) + R.c.w, ay1 = (R.c.y - bx.w) + R.c.w;
        const float dxmin = fmaxf(fmaxf(-ax0, ax1), 0.f), dymin = fmaxf(fmaxf(-ay0, ay1), 0.f);
        const float d2min = fmaf(dxmin, dxmin, fmaf(dymin, dymin, eps2f));
        return d2min * (1.0f - T2_MARGIN) >= R.open_t * (1.0f + OPEN_BAND);
    };
    auto mayOpen32 = [&](const TravRec& R) -> bool {
        if (haveOob && !boxAccepts(R, oboxf)) return true;
        // depth-first over the heap-ordered tree without a stack: leaving node k = climb while k is a right child,
        // then step to the sibling
        unsigned int k = 1u;
        while (k) {
            const float4 bx = bvh[k];
            if (bx.x > bx.z || boxAccepts(R, bx)) {   // empty, or the whole group of quadrants accepts
                k >>= __ffs(~k) - 1;
                k = (k > 1u) ? k + 1u : 0u;
            } else if (k >= (unsigned int)DD_BVH_LEAVES) {
                return true;                          // a quadrant the node may be opened from
            } else {
                k <<= 1;
            }
        }
        return false;
    };
    unsigned int* queue = a.queue + (size_t)d * a.icap;
    unsigned int* tailNext = &s->exp_count[d];     // zeroed with the step's scalars
    const unsigned int regionBase = a.importBase + (unsigned int)sp.me * a.icap;   // this rank's import region on every rank
    // a cell that may be opened from d: next queue position = its export index; returns the child block index d will see
    auto pushCell = [&](unsigned int ordinal) -> unsigned int {
        const unsigned int e = atomicAdd(tailNext, 1u);
        if (e >= a.icap) {
            atomicOr(&hdr->fault, 2u);
            return DD_POISON_BLOCK;
        }
        queue[e] = ordinal;
        return regionBase + e;
    };
    // ---- the roots: published to d whatever they are; root CELLS that d may open start the queue ----
    if (crank == 0) {
        for (unsigned int i = tid; i < nroots; i += blockDim.x) {
            DDRoot r = a.myroots[i];
            const unsigned int ordinal = r.cblock;
            r.cblock = 0u;
            if (ordinal != LPE_NONE) {
                if (!remote) r.cblock = c.blockBase + ordinal;
                else {   // (a root has no record on this rank: make the one the top builders will make)
                    const TravRec R = make_record(c, r.agg, r.level, LPE_NONE, 0u, mass_scale_inv(s->max_mass_bits));
                    r.cblock = (R.open_t >= 0.f && mayOpen32(R)) ? pushCell(ordinal) : DD_POISON_BLOCK;
                }
            }
            peers.roots[d][(size_t)sp.me * DD_MAXROOTS + i] = r;
        }
        if (tid == 0) peers.hdr[d]->root_count[sp.me] = nroots;
    }
    if (!remote) return;
    cluster_sync_all();
    // ---- breadth first, one generation per round; a quad of lanes per exported cell ----
    // Queue position e is always handled by quad (e mod quads), so it does not matter that the lanes read the tail at
    // slightly different moments (a lane that sees entries pushed in the current round just handles them early).
    // What must be uniform is the decision to stop: round g's pushers raise pushed[g], which is only read after the
    // barrier that ends the round. One cluster barrier per generation.
    unsigned int head = 0;
    const int r4 = tid & 3;
    const unsigned int quadsPerBlock = blockDim.x >> 2;
    const unsigned int quads = quadsPerBlock * DD_EXPORT_CLUSTER;
    const unsigned int myQuad = (unsigned int)crank * quadsPerBlock + ((unsigned int)tid >> 2);
    unsigned int* pushed = a.pushed + (size_t)d * DD_EXPORT_ROUNDS;   // zeroed with the step's scratch
    for (int round = 0; round < DD_EXPORT_ROUNDS - 1; ++round) {
        unsigned int tail = __ldcg(tailNext);
        if (tail > a.icap) tail = a.icap;
        bool any = false;
        for (unsigned int e = head + ((myQuad + quads - head % quads) % quads); e < tail; e += quads) {
            const unsigned int q = __ldcg(queue + e);
            const uint4* src = reinterpret_cast<const uint4*>(a.rec + lpe_idx(4u * (c.blockBase + q) + r4, c.recSlots, 15, s));
            uint4 v0 = src[0], v1 = src[1];
            TravRec R;
            R.c = make_float4(__uint_as_float(v0.x), __uint_as_float(v0.y), __uint_as_float(v0.z), __uint_as_float(v0.w));
            R.gm = __uint_as_float(v1.x); R.open_t = __uint_as_float(v1.y); R.node = v1.z; R.cblock = v1.w;
            if (R.cblock != 0u) {   // a cell: its own child block as numbered on d, if d can open it at all
                const bool open = R.open_t >= 0.f && mayOpen32(R);   // (open_t = -2: skipped by the small-mass rule, never opened)
                const unsigned int blk = open ? pushCell((R.cblock >> 2) - c.blockBase) : DD_POISON_BLOCK;
                any = any || open;
                v1.w = (blk << 2) | (R.cblock & 3u);
            }
            uint4* o = reinterpret_cast<uint4*>(peers.rec[d] + lpe_idx(4u * (regionBase + e) + r4, c.recSlots, 14, s));
            __stcs(o, v0);
            __stcs(o + 1, v1);
        }
        if (any) pushed[round] = 1u;
        head = tail;
        cluster_sync_all();   // this round's pushes (queue entries, the tail, the flag) are visible to the whole cluster
        if (__ldcg(pushed + round) == 0u) {
            if (tid == 0 && crank == 0) atomicMax(&s->dd_rounds, (unsigned int)round + 1u);
            break;
        }
    }
}

// the fp64 side records of the exported blocks (exact centre / mass / level of every child: guard-band re-tests and
// STRICT precision on the receiving rank) — one thread per exported record, off the breadth-first critical path
__global__ void __launch_bounds__(256)
k_dd_export_x(StepConst c, int me, DDExportArgs a, DDPeers peers, const Scal* __restrict__ s) {
    const int d = blockIdx.y;
    if (d == me) return;
    const unsigned int count = min(s->exp_count[d], a.icap);
    const unsigned int* queue = a.queue + (size_t)d * a.icap;
    const unsigned int regionBase = a.importBase + (unsigned int)me * a.icap;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4u * count; i += gridDim.x * blockDim.x) {
        const unsigned int e = i >> 2, r4 = i & 3u;
        const unsigned int q = queue[e];
        const uint4 cd4 = *reinterpret_cast<const uint4*>(a.child + (size_t)q * 4);
        const unsigned int cds[4] = {cd4.x, cd4.y, cd4.z, cd4.w};
        unsigned int code = LPE_NONE, seen = 0;   // the r-th valid child in digit order sits in slot r
#pragma unroll
        for (int dg = 0; dg < 4; ++dg)
            if (cds[dg] != LPE_NONE) {
                if (seen == r4) code = cds[dg];
                ++seen;
            }
        double4 X = make_double4(0.0, 0.0, 0.0, -1.0);
        if (code != LPE_NONE) {
            if (code & LPE_LEAF_FLAG) {
                const Body b = a.body[code & ~LPE_LEAF_FLAG];
                X = make_double4(b.x, b.y, b.m, -1.0);
            } else {
                const Agg ag = a.agg[code];
                X = dd_xrec(ag, agg_level(ag), c.quirk);
            }
        }
        double2* ox = reinterpret_cast<double2*>(peers.xrec[d] + 4u * ((size_t)regionBase + e) + r4);
        __stcs(ox, make_double2(X.x, X.y));
        __stcs(ox + 1, make_double2(X.z, X.w));
    }
}

// ---- phase C: the top of the tree from every rank's published roots (one block) ------------------------------------------
// Same construction as the single-GPU topology (bh_build.cuh) with the roots as "terminals": adjacent roots i, i+1
// witness the branching cell at their common level, cells that start at root a are numbered shallow to deep, a cell's
// first child is the next deeper cell starting there (or the root), every other child starts after a witness. The sums
// are (c0 + c1) + (c2 + c3) by child digit, as in aggregate_cell_quad, so every shared cell gets bit for bit the
// aggregate the single-GPU build gives it.
struct DDTop {
    unsigned int* mask;     // [DD_TOPROOTS]
    unsigned int* P;        // [DD_TOPROOTS + 1]
    unsigned int* wstart;   // [DD_TOPROOTS]
    unsigned int* child;    // [4 * DD_TOPROOTS]
    int* cellLevel;         // [DD_TOPROOTS]
    Agg* agg;               // [DD_TOPROOTS]
    signed char* delta;     // [DD_TOPROOTS]
};
#define DD_CELL_FLAG 0x40000000u
constexpr int DD_TOP_SMEM_ROOTS = 640;   // up to this many roots the whole construction runs out of shared memory
constexpr size_t DD_TOP_SMEM_BYTES = (size_t)DD_TOP_SMEM_ROOTS * (sizeof(DDRoot) + sizeof(Agg) + 16 + 4 * 4 + 4);
__global__ void __launch_bounds__(1024)
k_dd_top(StepConst c, int me, int R, const DDHeader* __restrict__ hdr, const DDRoot* __restrict__ roots, DDTop t,
         TravRec* __restrict__ rec, double4* __restrict__ xrec, unsigned int* __restrict__ selfslot, Scal* __restrict__ s) {
    extern __shared__ __align__(16) unsigned char dsm[];
    __shared__ unsigned int base[LPE_MAX_P2P + 1];
    __shared__ unsigned long long key[DD_TOPROOTS];
    __shared__ unsigned int levelsMask, ncellsSh;
    const int tid = threadIdx.x;
    if (tid == 0) {
        unsigned int run = 0;
        for (int r = 0; r < R; ++r) { base[r] = run; run += hdr->root_count[r]; }
        for (int r = R; r <= LPE_MAX_P2P; ++r) base[r] = run;
        levelsMask = 0u;
    }
    __syncthreads();
    const int N = (int)base[R];
    // work arrays: shared memory when the roots fit (the usual case: a few dozen roots per rank), else global scratch.
    // Generic pointers, one code path.
    const bool inShared = N <= DD_TOP_SMEM_ROOTS;
    DDRoot* sroots = reinterpret_cast<DDRoot*>(dsm);
    Agg* cagg = inShared ? reinterpret_cast<Agg*>(dsm + (size_t)DD_TOP_SMEM_ROOTS * sizeof(DDRoot)) : t.agg;
    unsigned int* child = inShared ? reinterpret_cast<unsigned int*>(dsm + (size_t)DD_TOP_SMEM_ROOTS * (sizeof(DDRoot) + sizeof(Agg))) : t.child;
    unsigned int* mask = inShared ? child + 4 * DD_TOP_SMEM_ROOTS : t.mask;
    unsigned int* P = inShared ? mask + DD_TOP_SMEM_ROOTS : t.P;          // (P[N] lives in ncellsSh)
    unsigned int* wstart = inShared ? P + DD_TOP_SMEM_ROOTS : t.wstart;
    int* cellLevel = inShared ? reinterpret_cast<int*>(wstart + DD_TOP_SMEM_ROOTS) : t.cellLevel;
    signed char* delta = inShared ? reinterpret_cast<signed char*>(cellLevel + DD_TOP_SMEM_ROOTS) : t.delta;
    auto globalRoot = [&](int i) -> const DDRoot& {
        int r = 0;
        while (r + 1 < R && (unsigned int)i >= base[r + 1]) ++r;
        return roots[(size_t)r * DD_MAXROOTS + ((unsigned int)i - base[r])];
    };
    const double msi = mass_scale_inv(s->max_mass_bits);
    if (tid == 0) s->dd_nroots = (unsigned int)N;
    if (N == 0) return;
    if (inShared) {   // the published roots, rank after rank = in key order
        const int words = (int)(sizeof(DDRoot) / 16);
        for (int k = tid; k < N * words; k += blockDim.x) {
            const int i = k / words, w = k % words;
            reinterpret_cast<uint4*>(sroots + i)[w] = reinterpret_cast<const uint4*>(&globalRoot(i))[w];
        }
        __syncthreads();
    }
    auto rootAt = [&](int i) -> const DDRoot& { return inShared ? sroots[i] : globalRoot(i); };
    auto writeChild = [&](unsigned int slot, const Agg& a, int level, unsigned int cblockIndex, const DDRoot* root) {
        slot = lpe_idx(slot, 4u * DD_TOPCAP, 16, s);
        rec[slot] = make_record(c, a, level, LPE_NONE, cblockIndex, msi);
        xrec[slot] = dd_xrec(a, level, c.quirk);   // (node_centre is inlined in both: the divisions are shared)
        if (root && root->owner == (unsigned int)me && root->leafpos != LPE_NONE && c.need_self) selfslot[root->leafpos] = slot;
    };
    if (N == 1) {   // one root: it is the root of the tree
        if (tid == 0) {
            const DDRoot& r0 = rootAt(0);
            writeChild(0u, r0.agg, r0.level, r0.cblock, &r0);
            rec[1] = rec[2] = rec[3] = invalid_record();
        }
        return;
    }
    for (int i = tid; i < N; i += blockDim.x) { key[i] = rootAt(i).key; mask[i] = 0u; }
    __syncthreads();
    for (int i = tid; i < N - 1; i += blockDim.x) {
        const int L = lca_level(key[i], key[i + 1], c.D);
        const int shift = 2 * (c.D - L);
        int a = i;
        while (a > 0 && (key[a - 1] >> shift) == (key[i] >> shift)) --a;
        delta[i] = (signed char)L;
        wstart[i] = (unsigned int)a;
        atomicOr(&mask[a], 1u << L);
        atomicOr(&levelsMask, 1u << L);
    }
    __syncthreads();
    if (tid < 32) {   // exclusive scan of popc(mask) by one warp
        unsigned int run = 0;
        for (int b0 = 0; b0 < N; b0 += 32) {
            const int i = b0 + tid;
            const unsigned int v = (i < N) ? (unsigned int)__popc(mask[i]) : 0u;
            unsigned int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int u = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (tid >= o) inc += u;
            }
            if (i < N) P[i] = run + inc - v;
            run += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
        if (tid == 0) ncellsSh = run;
    }
    __syncthreads();
    const int ncells = (int)ncellsSh;   // <= N - 1
    for (int k = tid; k < 4 * ncells; k += blockDim.x) child[k] = LPE_NONE;
    __syncthreads();
    for (int i = tid; i < N; i += blockDim.x) {
        const unsigned int mk = mask[i], Pi = P[i];
        unsigned int rest = mk, j = 0;
        while (rest) {   // cells that start at root i, shallow to deep
            const int L = __ffs(rest) - 1;
            rest &= rest - 1;
            const unsigned int ord = Pi + j;
            cellLevel[ord] = L;
            const unsigned int digit = (unsigned int)(key[i] >> (2 * (c.D - L - 1))) & 3u;
            child[(size_t)ord * 4 + digit] = rest ? (DD_CELL_FLAG | (ord + 1u)) : (unsigned int)i;
            ++j;
        }
        if (i < N - 1) {   // as the witness of the cell at level delta[i]: the child that starts at root i + 1
            const int L = (int)delta[i];
            const unsigned int a0 = wstart[i];
            const unsigned int q = P[a0] + (unsigned int)__popc(mask[a0] & ((1u << L) - 1u));
            const unsigned int code = mask[i + 1] ? (DD_CELL_FLAG | P[i + 1]) : (unsigned int)(i + 1);
            const unsigned int digit = (unsigned int)(key[i + 1] >> (2 * (c.D - L - 1))) & 3u;
            child[(size_t)q * 4 + digit] = code;
        }
    }
    __syncthreads();
    // aggregates, deepest level first (sums only: no division on this serial path)
    auto childAgg = [&](unsigned int code, int& level, unsigned int& cb, const DDRoot*& root) -> Agg {
        Agg a;
        a.m = 0.0; a.sx = 0.0; a.sy = 0.0; a.mf = 0.0; a.xf = 0.0; a.yf = 0.0;
        a.frank = 0xFFFFFFFFu; a.fidx = 0; a.ordinal = 0; a.small = 1u;
        level = -3; cb = 0u; root = nullptr;
        if (code != LPE_NONE) {
            if (code & DD_CELL_FLAG) {
                const unsigned int o2 = code & ~DD_CELL_FLAG;
                a = cagg[o2]; level = cellLevel[o2]; cb = DD_TOPB + o2;
            } else {
                const DDRoot& r0 = rootAt((int)code);
                a = r0.agg; level = r0.level; cb = r0.cblock; root = &r0;
            }
        }
        return a;
    };
    unsigned int levels = levelsMask;
    while (levels) {
        const int L = 31 - __clz(levels);
        levels &= ~(1u << L);
        for (int ord = tid; ord < ncells; ord += blockDim.x) {
            if (cellLevel[ord] != L) continue;
            Agg ch[4];
            unsigned int nvalid = 0;
#pragma unroll
            for (int dg = 0; dg < 4; ++dg) {
                int lv; unsigned int cb; const DDRoot* rt;
                ch[dg] = childAgg(child[(size_t)ord * 4 + dg], lv, cb, rt);
                nvalid += lv != -3 ? 1u : 0u;
            }
            // (c0 + c1) + (c2 + c3); first occupant = minimum insertion rank
            Agg r;
            r.m = (ch[0].m + ch[1].m) + (ch[2].m + ch[3].m);
            r.sx = (ch[0].sx + ch[1].sx) + (ch[2].sx + ch[3].sx);
            r.sy = (ch[0].sy + ch[1].sy) + (ch[2].sy + ch[3].sy);
            int best = 0;
#pragma unroll
            for (int dg = 1; dg < 4; ++dg)
                if (ch[dg].frank < ch[best].frank) best = dg;
            r.mf = ch[best].mf; r.xf = ch[best].xf; r.yf = ch[best].yf; r.frank = ch[best].frank; r.fidx = ch[best].fidx;
            r.ordinal = (unsigned int)ord;
            r.small = (ch[0].small & ch[1].small & ch[2].small & ch[3].small & 1u) | ((nvalid - 1u) << 1) | ((unsigned int)(L + 2) << 8);
            cagg[ord] = r;
        }
        __syncthreads();
    }
    // the records: one thread per (cell, child digit); the valid children of a cell fill its block in digit order
    for (int k = tid; k < 4 * ncells; k += blockDim.x) {
        const int ord = k >> 2, dg = k & 3;
        unsigned int before = 0, nvalid = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool v = child[(size_t)ord * 4 + j] != LPE_NONE;
            nvalid += v ? 1u : 0u;
            before += (v && j < dg) ? 1u : 0u;
        }
        int lv; unsigned int cb; const DDRoot* rt;
        const Agg a = childAgg(child[(size_t)ord * 4 + dg], lv, cb, rt);
        const unsigned int blockSlot = 4u * (DD_TOPB + (unsigned int)ord);
        if (lv != -3) writeChild(blockSlot + before, a, lv, cb, rt);
        else rec[blockSlot + nvalid + ((unsigned int)dg - before)] = invalid_record();   // the unused slots behind them
    }
    if (tid == 0) {   // the root: the shallowest cell that starts at root 0
        writeChild(0u, cagg[0], cellLevel[0], DD_TOPB, nullptr);
        rec[1] = rec[2] = rec[3] = invalid_record();
    }
}

// poison block: four leaves with a NaN mass — a traversal that ever opened a cell whose children were not exported
// would produce NaN velocities instead of a silently wrong answer
__global__ void k_dd_poison(TravRec* __restrict__ rec) {
    if (threadIdx.x < 4) {
        TravRec r = invalid_record();
        r.gm = __int_as_float(0x7fc00000);
        rec[4u * DD_POISON_BLOCK + threadIdx.x] = r;
    }
}

// ---- upload: every rank sees the whole input and keeps the bodies of its own key range ------------------------------
struct DDSelIn {
    const double *x, *y, *vx, *vy, *m;
    const unsigned int* rank;
    const unsigned char* comp;
    unsigned int first;         // creation index of element 0 of this chunk
    unsigned int ntotal;        // bodies of the whole input (default rank = ntotal - 1 - creation index)
};
__device__ __forceinline__ unsigned int dd_sel_comp(const DDSelIn& in, int i) {
    return in.comp ? (unsigned int)in.comp[i] : (unsigned int)(1u | 2u);
}
struct DDSelLoad {
    StepConst c;            // depth LPE_MAX_DEPTH
    DDSplit sp;             // depth-30 splitters
    DDSelIn in;
    __device__ __forceinline__ unsigned int operator()(int i) const {
        bool inside;
        const unsigned long long key = dd_body_key(c, nullptr, make_double2(in.x[i], in.y[i]), dd_sel_comp(in, i), inside);
        return dd_owner(sp, key) == sp.me ? 1u : 0u;
    }
};
struct DDSelSink {
    DDSelIn in;
    Body* body;
    double2* vel;
    unsigned int* orig;
    unsigned int base, cap;
    unsigned int* total;
    int n;
    __device__ __forceinline__ void operator()(int i, unsigned int excl, unsigned int v) const {
        if (i == n) { *total = excl; return; }
        if (!v) return;
        const unsigned int slot = base + excl;
        if (slot >= cap) return;
        Body b;
        b.x = in.x[i]; b.y = in.y[i]; b.m = in.m[i];
        const unsigned int ci = in.first + (unsigned int)i;
        b.rank = in.rank ? in.rank[i] : (in.ntotal - 1u - ci);
        b.comp = dd_sel_comp(in, i);
        body[slot] = b;
        vel[slot] = make_double2(in.vx ? in.vx[i] : 0.0, in.vy ? in.vy[i] : 0.0);
        orig[slot] = ci;
    }
};
// depth-30 keys of a strided sample (splitter estimate) and the largest source mass of the whole input
__global__ void __launch_bounds__(256)
k_dd_sample_keys(StepConst c, int n, const double* __restrict__ x, const double* __restrict__ y,
                 const unsigned char* __restrict__ comp, unsigned long long* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool inside;
    keys[i] = dd_body_key(c, nullptr, make_double2(x[i], y[i]), comp ? (unsigned int)comp[i] : 3u, inside);
}
__global__ void __launch_bounds__(256)
k_dd_max_mass(int n, const double* __restrict__ m, const unsigned char* __restrict__ comp, Scal* __restrict__ s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    block_max_mass(i < n ? m[i] : 0.0, i < n ? (comp ? (unsigned int)comp[i] : 3u) : 0u, s);
}

}  // namespace lpe
