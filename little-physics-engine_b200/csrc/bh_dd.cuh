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
struct DDDomain {               // identical on every rank; rebuilt on the host when the splitters or the depth change
    int nq[LPE_MAX_P2P];
    double box[LPE_MAX_P2P][4];                 // bounding box of the rank's quadrants: x0, y0, x1, y1
    DDQuad q[LPE_MAX_P2P][DD_MAXQ];
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
__global__ void k_dd_signal(int point, unsigned long long epoch, int me, int R, DDPeers peers, const double* __restrict__ payload) {
    const int s = threadIdx.x;
    if (s >= R) return;
    DDMail* m = &peers.hdr[s]->mail[point][me];
    if (payload)
        for (int k = 0; k < 6; ++k) m->payload[k] = payload[k];
    __threadfence_system();
    st_release_sys(&m->flag, epoch);
}
// Wait: lane s spins until rank s's flag in the own mailbox carries this step's epoch. Bounded: a dead peer becomes an
// error flag, not a hung device.
__global__ void k_dd_wait(int point, unsigned long long epoch, int R, DDHeader* hdr) {
    const int s = threadIdx.x;
    if (s >= R) return;
    const unsigned long long* f = &hdr->mail[point][s].flag;
    unsigned long long spins = 0;
    while (ld_acquire_sys(f) < epoch) {
        __nanosleep(200);
        if (++spins > (1ull << 24)) {   // a few seconds
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

// ---- phase A: keys of the own bodies; leavers go straight into their new owner's tail slots -------------------------
// c.n = S (slots). Slots [0, n_live) hold bodies; the tail belongs to the peers in this phase (they store migrants
// there) and is NOT read here: its keys are written by k_dd_keygen_inbox after the barrier.
__global__ void __launch_bounds__(256)
k_dd_keygen(StepConst c, DDSplit sp, const Body* __restrict__ body, const double2* __restrict__ vel,
            const unsigned int* __restrict__ orig, unsigned long long* __restrict__ keys, unsigned int* __restrict__ vals,
            Scal* __restrict__ s, DDHeader* __restrict__ hdr, DDPeers peers, unsigned long long* __restrict__ oob, int icap) {
    __shared__ unsigned char lut[64];
    if (threadIdx.x < 64) lut[threadIdx.x] = hilbert_lut_entry(threadIdx.x);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int n_prev = hdr->n_live;
    unsigned int in = 0, live = 0;
    if (i < c.n) {
        unsigned long long key = dd_dead_key(c.D);
        if ((unsigned int)i < n_prev) {
            const Body b = body[i];
            bool inside;
            key = dd_body_key(c, lut, make_double2(b.x, b.y), b.comp, inside);
            const int owner = dd_owner(sp, key);
            const bool target = (b.comp & 1u) && (b.comp & 2u) && !(b.comp & 4u);
            if (!inside && target) {   // a target outside the tree: the last rank's domain must cover it
                atomicMin(&oob[0], ordered_bits(b.x)); atomicMin(&oob[1], ordered_bits(b.y));
                atomicMax(&oob[2], ordered_bits(b.x)); atomicMax(&oob[3], ordered_bits(b.y));
            }
            if (owner == sp.me) {
                live = 1; in = inside ? 1u : 0u;
            } else {
                const unsigned int k = atomicAdd(&peers.hdr[owner]->inbox_count, 1u);
                if (k < (unsigned int)icap) {
                    const unsigned int slot = (unsigned int)c.n - 1u - k;
                    peers.body[owner][slot] = b;
                    peers.vel[owner][slot] = vel[i];
                    peers.orig[owner][slot] = orig[i];
                } else {
                    atomicOr(&hdr->fault, 1u);
                }
                key = dd_dead_key(c.D);
            }
        }
        keys[i] = key;
        vals[i] = (unsigned int)i;
    }
    const unsigned int cin = __syncthreads_count(in);
    const unsigned int clive = __syncthreads_count(live);
    if (threadIdx.x == 0) {
        if (cin) atomicAdd(&s->n_in, cin);
        if (clive) atomicAdd(&s->n_live, clive);
    }
}

// ---- phase B, first kernel: keys of the bodies that arrived in the tail slots ---------------------------------------
__global__ void __launch_bounds__(256)
k_dd_keygen_inbox(StepConst c, DDSplit sp, const Body* __restrict__ body, unsigned long long* __restrict__ keys,
                  unsigned int* __restrict__ vals, Scal* __restrict__ s, DDHeader* __restrict__ hdr) {
    __shared__ unsigned char lut[64];
    if (threadIdx.x < 64) lut[threadIdx.x] = hilbert_lut_entry(threadIdx.x);
    __syncthreads();
    unsigned int cnt = hdr->inbox_count;
    const unsigned int room = (unsigned int)c.n - hdr->n_live;
    if (cnt > room) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&hdr->fault, 1u);
        cnt = room;
    }
    unsigned int in = 0, live = 0;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < cnt; k += gridDim.x * blockDim.x) {
        const unsigned int slot = (unsigned int)c.n - 1u - k;
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
        keys[slot] = key;   // (vals[slot] = slot was written by phase A)
    }
    (void)vals;
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
// visited from there. A flagged cell gets an export index per destination (any unique index will do: the layout of the
// imported blocks does not change any result).
__device__ __forceinline__ bool dd_box_may_open(double cx, double cy, double sizeSq, double eps2, double theta2,
                                                double x0, double y0, double x1, double y1) {
    const double dx = fmax(fmax(x0 - cx, cx - x1), 0.0), dy = fmax(fmax(y0 - cy, cy - y1), 0.0);
    const double d2 = dx * dx + dy * dy + eps2;
    return !(sizeSq < theta2 * d2 * (1.0 - 1e-9));
}

struct DDExport {
    const uint2* levelList;         // every local cell: {pre-order index, ordinal}
    const NodeMeta* meta;
    const Agg* agg;
    const unsigned long long* tkey;
    unsigned int* eidx;             // [dest][cellCap] export index of the cell's child block, LPE_NONE = not exported
    uint4* list;                    // {ordinal, dest, export index, -}
    unsigned int cellCap;
    unsigned int icap;              // import blocks per sender on every rank
    unsigned int listCap;
};

__global__ void __launch_bounds__(256)
k_dd_export_flags(StepConst c, DDSplit sp, const DDDomain* __restrict__ dom, DDExport e, Scal* __restrict__ s,
                  DDHeader* __restrict__ hdr) {
    const unsigned int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= s->n_internal) return;
    const uint2 pq = e.levelList[idx];
    const NodeMeta mt = e.meta[pq.x];
    const int L = mt.level;
    const int shift = 2 * (c.D - L);
    // a cell is complete (and a cell of the global tree) iff its whole key interval lies inside this rank's range
    const unsigned long long kstart = (e.tkey[mt.start] >> shift) << shift;
    const unsigned long long kend = kstart + (1ull << shift);
    const bool inner = kstart >= sp.k[sp.me] && kend <= sp.k[sp.me + 1] && kend != 0ull;
    double cx = 0.0, cy = 0.0, sizeSq = 0.0;
    if (inner) {
        double M;
        node_centre(e.agg[pq.x], L, c.quirk, M, cx, cy);
        const double size = ldexp(c.U, -L);
        sizeSq = size * size;
    }
    const double eps2 = c.eps * c.eps;
    for (int d = 0; d < sp.R; ++d) {
        if (d == sp.me) continue;
        bool open = false;
        if (inner) {
            const int nq = dom->nq[d];
            if (nq > 0 && dd_box_may_open(cx, cy, sizeSq, eps2, c.theta2, dom->box[d][0], dom->box[d][1], dom->box[d][2], dom->box[d][3])) {
                for (int k = 0; k < nq && !open; ++k) {
                    const DDQuad& q = dom->q[d][k];
                    open = dd_box_may_open(cx, cy, sizeSq, eps2, c.theta2, q.x0, q.y0, q.x1, q.y1);
                }
            }
            if (!open && d == sp.R - 1) {
                // targets outside the universe live on the last rank: union of every rank's box of such bodies
                double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
                bool any = false;
                for (int r = 0; r < sp.R; ++r) {
                    const double* pl = hdr->mail[0][r].payload;
                    if (pl[0] <= pl[2]) { any = true; x0 = fmin(x0, pl[0]); y0 = fmin(y0, pl[1]); x1 = fmax(x1, pl[2]); y1 = fmax(y1, pl[3]); }
                }
                if (any) open = dd_box_may_open(cx, cy, sizeSq, eps2, c.theta2, x0, y0, x1, y1);
            }
        }
        unsigned int ex = LPE_NONE;
        if (open) {
            ex = atomicAdd(&s->exp_count[d], 1u);
            if (ex >= e.icap) {
                atomicOr(&hdr->fault, 2u);
                ex = LPE_NONE;
            } else {
                const unsigned int li = atomicAdd(&s->exp_list_count, 1u);
                if (li < e.listCap) e.list[li] = make_uint4(pq.y, (unsigned int)d, ex, 0u);
            }
        }
        e.eidx[(size_t)d * e.cellCap + pq.y] = ex;
    }
}

// fp64 side record of a non-local slot: {centre x, centre y, mass, level} as the traversal's exact test wants them
__device__ __forceinline__ double4 dd_xrec(const Agg& a, int level, int quirk) {
    double M, cx, cy;
    node_centre(a, level, quirk, M, cx, cy);
    return make_double4(cx, cy, M, (double)level);
}

// ---- phase B: roots of the non-empty inner quadrants -> every rank's root table ----------------------------------------
struct DDPublish {
    const unsigned long long* tkey;
    const unsigned int* tfirst;
    const unsigned int* tnode;
    const unsigned int* mask;
    const unsigned int* P;
    const Agg* agg;
    const Body* body;
    const unsigned int* eidx;
    unsigned int cellCap;
    unsigned int icap;
    unsigned int importBase;   // first import block on every rank; sender s owns [importBase + s * icap, + icap)
};
__device__ __forceinline__ unsigned int dd_lower_bound(const unsigned long long* __restrict__ a, unsigned int n, unsigned long long v) {
    unsigned int lo = 0, hi = n;
    while (lo < hi) {
        const unsigned int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__global__ void __launch_bounds__(256)
k_dd_publish(StepConst c, DDSplit sp, const DDDomain* __restrict__ dom, DDPublish a, DDPeers peers, Scal* __restrict__ s,
             DDHeader* __restrict__ hdr) {
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
        r.key = q.key; r.owner = (unsigned int)sp.me; r.leafpos = LPE_NONE; r.pad[0] = r.pad[1] = 0u;
        unsigned int ordinal = LPE_NONE;
        if (t1 - t0 == 1u) {   // one terminal: a single-body leaf or an aggregated depth-D cell
            const unsigned int first = a.tfirst[t0], last = a.tfirst[t0 + 1];
            if (last - first == 1u) {
                r.level = -1; r.agg = body_agg(a.body[first], first, c.thr); r.leafpos = first;
            } else {
                r.level = -2; r.agg = a.agg[a.tnode[t0]];
            }
        } else {               // the lowest cell that holds every body of the quadrant
            const int L = lca_level(a.tkey[t0], a.tkey[t1 - 1], c.D);
            ordinal = a.P[t0] + (unsigned int)__popc(a.mask[t0] & ((1u << L) - 1u));
            r.level = L; r.agg = a.agg[t0 + ordinal];
        }
        if (pos < (unsigned int)DD_MAXROOTS) {
            for (int d = 0; d < sp.R; ++d) {
                r.cblock = 0u;
                if (ordinal != LPE_NONE) {
                    if (d == sp.me) r.cblock = c.blockBase + ordinal;
                    else {
                        const unsigned int ex = a.eidx[(size_t)d * a.cellCap + ordinal];
                        r.cblock = (ex == LPE_NONE) ? DD_POISON_BLOCK : a.importBase + (unsigned int)sp.me * a.icap + ex;
                    }
                }
                peers.roots[d][(size_t)sp.me * DD_MAXROOTS + pos] = r;
            }
        } else {
            atomicOr(&hdr->fault, 4u);
        }
    }
    if (j == 0) {
        const unsigned int cnt = total < (unsigned int)DD_MAXROOTS ? total : (unsigned int)DD_MAXROOTS;
        for (int d = 0; d < sp.R; ++d) peers.hdr[d]->root_count[sp.me] = cnt;
        hdr->n_live = s->n_live;     // the state is compact again: next step's phase A reads slots [0, n_live)
        hdr->inbox_count = 0u;       // consumed; the peers touch it again only after the barrier that follows
    }
}

// ---- phase B: the flagged child blocks -> the destination's record array (peer stores) ----------------------------
struct DDWrite {
    const uint4* list;
    const unsigned int* eidx;
    const unsigned int* child;      // [4 * ordinal + digit]
    const TravRec* rec;             // own records (local blocks)
    const NodeMeta* meta;
    const Agg* agg;
    const Body* body;
    unsigned int cellCap, icap, importBase, listCap;
};
__global__ void __launch_bounds__(256)
k_dd_export_write(StepConst c, int me, DDWrite w, DDPeers peers, const Scal* __restrict__ s) {
    unsigned int count = s->exp_list_count;
    if (count > w.listCap) count = w.listCap;
    const unsigned int quads = (gridDim.x * blockDim.x) >> 2;
    const int r = threadIdx.x & 3;
    for (unsigned int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 2; i < count; i += quads) {
        const uint4 e = w.list[i];
        const unsigned int q = e.x, d = e.y, ex = e.z;
        TravRec R = w.rec[4u * (c.blockBase + q) + r];
        // the r-th valid child in digit order sits in slot r
        unsigned int code = LPE_NONE, seen = 0;
#pragma unroll
        for (int dg = 0; dg < 4; ++dg) {
            const unsigned int cd = w.child[(size_t)q * 4 + dg];
            if (cd != LPE_NONE) {
                if (seen == (unsigned int)r) code = cd;
                ++seen;
            }
        }
        double4 X = make_double4(0.0, 0.0, 0.0, -1.0);
        if (code != LPE_NONE) {
            if (code & LPE_LEAF_FLAG) {
                const Body b = w.body[code & ~LPE_LEAF_FLAG];
                X = make_double4(b.x, b.y, b.m, -1.0);
            } else {
                const int level = w.meta[code].level;
                X = dd_xrec(w.agg[code], level, c.quirk);
                if (level >= 0) {   // a cell: its own child block, as numbered on the destination
                    const unsigned int qc = (R.cblock >> 2) - c.blockBase;
                    const unsigned int exc = w.eidx[(size_t)d * w.cellCap + qc];
                    const unsigned int blk = (exc == LPE_NONE) ? DD_POISON_BLOCK : w.importBase + (unsigned int)me * w.icap + exc;
                    R.cblock = (blk << 2) | (R.cblock & 3u);
                }
            }
        }
        const size_t dst = 4u * ((size_t)w.importBase + (size_t)me * w.icap + ex) + r;
        uint4* o = reinterpret_cast<uint4*>(peers.rec[d] + dst);
        const uint4* src = reinterpret_cast<const uint4*>(&R);
        __stcs(o, src[0]);
        __stcs(o + 1, src[1]);
        double2* ox = reinterpret_cast<double2*>(peers.xrec[d] + dst);
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
__global__ void __launch_bounds__(1024)
k_dd_top(StepConst c, int me, int R, const DDHeader* __restrict__ hdr, const DDRoot* __restrict__ roots, DDTop t,
         TravRec* __restrict__ rec, double4* __restrict__ xrec, unsigned int* __restrict__ selfslot, Scal* __restrict__ s) {
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
    auto rootAt = [&](int i) -> const DDRoot& {
        int r = 0;
        while (r + 1 < R && (unsigned int)i >= base[r + 1]) ++r;
        return roots[(size_t)r * DD_MAXROOTS + ((unsigned int)i - base[r])];
    };
    const double msi = mass_scale_inv(s->max_mass_bits);
    if (tid == 0) s->dd_nroots = (unsigned int)N;
    if (N == 0) return;
    auto writeChild = [&](unsigned int slot, const Agg& a, int level, unsigned int cblockIndex, const DDRoot* root) {
        rec[slot] = make_record(c, a, level, 1u, cblockIndex, msi);
        xrec[slot] = dd_xrec(a, level, c.quirk);
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
    for (int i = tid; i < N; i += blockDim.x) { key[i] = rootAt(i).key; t.mask[i] = 0u; }
    __syncthreads();
    for (int i = tid; i < N - 1; i += blockDim.x) {
        const int L = lca_level(key[i], key[i + 1], c.D);
        const int shift = 2 * (c.D - L);
        int a = i;
        while (a > 0 && (key[a - 1] >> shift) == (key[i] >> shift)) --a;
        t.delta[i] = (signed char)L;
        t.wstart[i] = (unsigned int)a;
        atomicOr(&t.mask[a], 1u << L);
        atomicOr(&levelsMask, 1u << L);
    }
    __syncthreads();
    if (tid < 32) {   // exclusive scan of popc(mask) by one warp
        unsigned int run = 0;
        for (int b0 = 0; b0 < N; b0 += 32) {
            const int i = b0 + tid;
            const unsigned int v = (i < N) ? (unsigned int)__popc(t.mask[i]) : 0u;
            unsigned int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int u = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (tid >= o) inc += u;
            }
            if (i < N) t.P[i] = run + inc - v;
            run += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
        if (tid == 0) { t.P[N] = run; ncellsSh = run; }
    }
    __syncthreads();
    const int ncells = (int)ncellsSh;
    for (int k = tid; k < 4 * ncells; k += blockDim.x) t.child[k] = LPE_NONE;
    __syncthreads();
    for (int i = tid; i < N; i += blockDim.x) {
        const unsigned int mk = t.mask[i], Pi = t.P[i];
        unsigned int rest = mk, j = 0;
        while (rest) {   // cells that start at root i, shallow to deep
            const int L = __ffs(rest) - 1;
            rest &= rest - 1;
            const unsigned int ord = Pi + j;
            t.cellLevel[ord] = L;
            const unsigned int digit = (unsigned int)(key[i] >> (2 * (c.D - L - 1))) & 3u;
            t.child[(size_t)ord * 4 + digit] = rest ? (DD_CELL_FLAG | (ord + 1u)) : (unsigned int)i;
            ++j;
        }
        if (i < N - 1) {   // as the witness of the cell at level delta[i]: the child that starts at root i + 1
            const int L = (int)t.delta[i];
            const unsigned int a0 = t.wstart[i];
            const unsigned int q = t.P[a0] + (unsigned int)__popc(t.mask[a0] & ((1u << L) - 1u));
            const unsigned int code = t.mask[i + 1] ? (DD_CELL_FLAG | t.P[i + 1]) : (unsigned int)(i + 1);
            const unsigned int digit = (unsigned int)(key[i + 1] >> (2 * (c.D - L - 1))) & 3u;
            t.child[(size_t)q * 4 + digit] = code;
        }
    }
    __syncthreads();
    unsigned int levels = levelsMask;
    while (levels) {   // deepest level first
        const int L = 31 - __clz(levels);
        levels &= ~(1u << L);
        for (int ord = tid; ord < ncells; ord += blockDim.x) {
            if (t.cellLevel[ord] != L) continue;
            Agg ch[4];
            int lvl[4];
            unsigned int cb[4];
            const DDRoot* rt[4];
            unsigned int nvalid = 0;
#pragma unroll
            for (int dg = 0; dg < 4; ++dg) {
                const unsigned int code = t.child[(size_t)ord * 4 + dg];
                Agg a;
                a.m = 0.0; a.sx = 0.0; a.sy = 0.0; a.mf = 0.0; a.xf = 0.0; a.yf = 0.0;
                a.frank = 0xFFFFFFFFu; a.fidx = 0; a.count = 0; a.small = 1u;
                lvl[dg] = -3; cb[dg] = 0u; rt[dg] = nullptr;
                if (code != LPE_NONE) {
                    ++nvalid;
                    if (code & DD_CELL_FLAG) {
                        const unsigned int o2 = code & ~DD_CELL_FLAG;
                        a = t.agg[o2]; lvl[dg] = t.cellLevel[o2]; cb[dg] = DD_TOPB + o2;
                    } else {
                        const DDRoot& r0 = rootAt((int)code);
                        a = r0.agg; lvl[dg] = r0.level; cb[dg] = r0.cblock; rt[dg] = &r0;
                    }
                }
                ch[dg] = a;
            }
            unsigned int slot = 4u * (DD_TOPB + (unsigned int)ord);
#pragma unroll
            for (int dg = 0; dg < 4; ++dg)
                if (lvl[dg] != -3) writeChild(slot++, ch[dg], lvl[dg], cb[dg], rt[dg]);
            for (; slot < 4u * (DD_TOPB + (unsigned int)ord) + 4u; ++slot) rec[slot] = invalid_record();
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
            r.count = ch[0].count + ch[1].count + ch[2].count + ch[3].count;
            r.small = (ch[0].small & ch[1].small & ch[2].small & ch[3].small & 1u) | ((nvalid - 1u) << 1);
            t.agg[ord] = r;
        }
        __syncthreads();
    }
    if (tid == 0) {   // the root: the shallowest cell that starts at root 0
        writeChild(0u, t.agg[0], t.cellLevel[0], DD_TOPB, nullptr);
        rec[1] = rec[2] = rec[3] = invalid_record();
    }
}

// poison block: four leaves with a NaN mass — a traversal that ever opened a cell whose children were not exported
// would produce NaN velocities instead of a silently wrong answer
__global__ void k_dd_poison(TravRec* __restrict__ rec) {
    if (threadIdx.x < 4) {
        TravRec r = invalid_record();
        r.gm = __int_as_float(0x7fc00000);
        r.skip = 1u;
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
