// bh_build.cuh — sort keys (Hilbert or Morton index of the depth-D cell), terminals, path-compressed quadtree
// topology in pre-order, and the bottom-up mass / centre-of-mass / first-occupant aggregation.
//
// What the reference builds by recursive insertion (barnes_hut.cpp:101-238) is rebuilt here from sorted
// keys. Equivalences used (all verified against the compiled reference):
//   Q10  child digit = (x >= mid) + 2*(y >= mid)  =>  reference DFS order == Morton order of the cells; any other
//        hierarchical curve (Hilbert) gives the same cells and the same tree, only another sibling order
//   Q3   a chain of single-child cells carries one (M, COM); only its lowest cell's size matters
//        => keep only branching cells (>= 2 non-empty children): a path-compressed tree
//   Q4   cells smaller than theta*eps are always accepted => stop at depth D (terminals may aggregate)
//   Q2   every internal cell counts its first occupant (minimum insertion rank) twice
//
// Topology without a level loop and without searches (one galloping search per witness aside): for sorted
// terminals t (distinct depth-D cells), delta[t] = LCA level of t and t+1. Every t is a "witness" of the
// branching cell at level delta[t] that contains t and t+1; that cell is identified by (a, L) with a = its first
// terminal (kept in wstart[t]). A 32-bit level mask per terminal collects the levels of the cells that start
// there (atomicOr), and with P = exclusive scan of popc(mask):
//     ordinal(a, L)    = P[a] + popc(mask[a] & ((1<<L)-1))
//     preorder(a, L)   = a + ordinal(a, L);   preorder(terminal t) = t + P[t] + popc(mask[t])
//     first child of (a, L)  = the next deeper cell starting at a (preorder + 1), else terminal a
//     other children         = for every witness t of the cell: the shallowest cell starting at t+1
//                              ((t+1) + P[t+1]), else terminal t+1
// Aggregation is level-synchronous, deepest level first (one launch per level, cells of a level listed by a
// block-aggregated counting pass in k_topology): a QUAD of lanes sums a cell's <= 4 children in the fixed order
// (c0 + c1) + (c2 + c3) and writes their traversal records side by side into the cell's CHILD BLOCK (one 128-byte
// line). The traversal always visits all children of an opened cell, so every byte it fetches is used.
#pragma once
#include "bh_common.cuh"

namespace lpe {

// ---- 1. sort keys ---------------------------------------------------------------------------------------
// Cell index along one axis at depth D, matching the reference's comparisons exactly: the reference descends with
// `x < bx + 0.5*bs` tests on boundaries k*h that are exact in fp64, so the cell is the k with
// k*h <= x < (k+1)*h; the quotient is only a first guess.
__device__ __forceinline__ unsigned int cell_index(double x, double h, double invh, unsigned int kmax) {
    long long k = (long long)(x * invh);
    if (k < 0) k = 0;
    if (k > (long long)kmax) k = kmax;
    while (k > 0 && __dmul_rn((double)k, h) > x) --k;
    while (k < (long long)kmax && __dmul_rn((double)(k + 1), h) <= x) ++k;
    return (unsigned int)k;
}

// Hilbert index of cell (x, y) at depth D: like the Morton code it is hierarchical (two bits per level, the bodies of
// every cell stay contiguous, so the same tree falls out of the sorted keys), but consecutive indices are always
// edge-adjacent cells: 32 consecutive bodies form a compact blob instead of straddling a Z-curve jump, which is what
// the traversal's per-warp grouping wants.
__host__ __device__ __forceinline__ unsigned long long hilbert_index(unsigned int x, unsigned int y, int D) {
    unsigned long long d = 0;
    const unsigned int n1 = (1u << D) - 1u;
    for (int b = D - 1; b >= 0; --b) {
        const unsigned int s = 1u << b;
        const unsigned int rx = (x >> b) & 1u, ry = (y >> b) & 1u;
        d |= (unsigned long long)((3u * rx) ^ ry) << (2 * b);
        if (ry == 0u) {
            if (rx == 1u) { x = n1 - x; y = n1 - y; }
            const unsigned int t = x; x = y; y = t;
        }
        (void)s;
    }
    return d;
}

// The same curve two levels per step: the walk above is a 4-state machine (state = swap | invert << 1 applied to the
// remaining low bits), so a 64-entry table indexed by state and a nibble (2 bits of x, 2 bits of y) gives 4 index
// bits and the next state. The table is filled from hilbert_index's own single-level rule, see hilbert_lut_entry.
__device__ __forceinline__ unsigned int hilbert_level(unsigned int state, unsigned int xb, unsigned int yb,
                                                      unsigned int& digit) {
    const unsigned int sw = state & 1u, inv = state >> 1;
    const unsigned int rx = (sw ? yb : xb) ^ inv, ry = (sw ? xb : yb) ^ inv;
    digit = (3u * rx) ^ ry;
    unsigned int nsw = sw, ninv = inv;
    if (ry == 0u) {            // hilbert_index: (flip both if rx) then swap
        if (rx == 1u) ninv ^= 1u;
        nsw ^= 1u;
    }
    return nsw | (ninv << 1);
}
__device__ __forceinline__ unsigned char hilbert_lut_entry(unsigned int idx) {   // idx = state << 4 | y2 << 2 | x2
    const unsigned int state = idx >> 4, x2 = idx & 3u, y2 = (idx >> 2) & 3u;
    unsigned int d1, d0;
    const unsigned int s1 = hilbert_level(state, x2 >> 1, y2 >> 1, d1);
    const unsigned int s0 = hilbert_level(s1, x2 & 1u, y2 & 1u, d0);
    return (unsigned char)((s0 << 4) | (d1 << 2) | d0);
}
__device__ __forceinline__ unsigned long long hilbert_index_lut(const unsigned char* lut, unsigned int x, unsigned int y, int D) {
    unsigned long long d = 0;
    unsigned int state = 0;
    int b = D;
    if (b & 1) {               // odd depth: one single level first
        --b;
        unsigned int dg;
        state = hilbert_level(0u, (x >> b) & 1u, (y >> b) & 1u, dg);
        d = dg;
    }
    for (b -= 2; b >= 0; b -= 2) {
        const unsigned int nib = ((x >> b) & 3u) | (((y >> b) & 3u) << 2);
        const unsigned int e = lut[(state << 4) | nib];
        d = (d << 4) | (e & 15u);
        state = e >> 4;
    }
    return d;
}

__global__ void __launch_bounds__(256)
k_keygen(StepConst c, const Body* __restrict__ body, void* __restrict__ keys,
         unsigned int* __restrict__ vals, Scal* __restrict__ s) {
    __shared__ unsigned char lut[64];
    if (threadIdx.x < 64) lut[threadIdx.x] = hilbert_lut_entry(threadIdx.x);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int in = 0;
    if (i < c.n) {
        // positions and components only: masses and ranks may still be on their way over PCIe (lpe_bh_update_host)
        const double2 p = *reinterpret_cast<const double2*>(&body[i].x);
        const unsigned int cm = body[i].comp;
        // buildTree's view and bounds test, barnes_hut.cpp:117-124
        const bool src = (cm & 1u) && !(cm & 4u);
        const bool inside = src && p.x >= 0.0 && p.x < c.U && p.y >= 0.0 && p.y < c.U;
        unsigned long long key = 1ull << (2 * c.D);  // not in the tree: sorts after every cell
        if (inside) {
            const unsigned int kmax = (1u << c.D) - 1u;
            const unsigned int ix = cell_index(p.x, c.h, c.invh, kmax);
            const unsigned int iy = cell_index(p.y, c.h, c.invh, kmax);
            key = c.hilbert ? hilbert_index_lut(lut, ix, iy, c.D) : (spread_bits32(ix) | (spread_bits32(iy) << 1));
            in = 1;
        }
        store_key(keys, vals, c.k32, c.D, i, key);
    }
    const unsigned int cnt = __syncthreads_count(in);
    if (threadIdx.x == 0 && cnt) atomicAdd(&s->n_in, cnt);
}

// ---- 2. gather: the state itself goes into key order (one 32-byte sector read and one written per body) ---------
// Bodies move little per step, so next step's gather reads almost in place, the kick / drift writes of the traversal
// are coalesced, and every later kernel reads bodies by sorted position. orig[] carries the creation index of each
// slot (null on input = the state is still in creation order). velIn null: velocities are still on their way over
// PCIe (host tick) and are packed straight into key order later.
__global__ void __launch_bounds__(256)
k_gather(int n, int need_self, const unsigned int* __restrict__ sidx, const Body* __restrict__ bodyIn,
         const double2* __restrict__ velIn, const unsigned int* __restrict__ origIn, Body* __restrict__ bodyOut,
         double2* __restrict__ velOut, unsigned int* __restrict__ origOut, unsigned int* __restrict__ selfslot,
         const unsigned int* __restrict__ n_dev, unsigned int cap, Scal* __restrict__ chk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dev) n = (int)*n_dev;
    if (i >= n) return;
    const unsigned int b = sidx[i] & ~LPE_VAL_OUT;
#ifdef LPE_CHECKED
    if (b >= cap) { atomicOr(&chk->check_fault, 1u << 1); return; }
#endif
    bodyOut[i] = bodyIn[b];
    if (velIn) velOut[i] = velIn[b];
    origOut[i] = origIn ? origIn[b] : b;
    if (need_self) selfslot[i] = LPE_NONE;   // set for bodies that end up alone in their depth-D cell
}

// Largest source mass -> Scal::max_mass_bits (fixes the power-of-two mass unit of the traversal records). Runs with
// the upload, not with the step: masses only change through an upload.
__device__ __forceinline__ void block_max_mass(double m, unsigned int comp, Scal* __restrict__ s) {
    unsigned long long mx = 0;
    if ((comp & 1u) && !(comp & 4u) && m > 0.0 && m < __longlong_as_double(0x7FF0000000000000ll))
        mx = (unsigned long long)__double_as_longlong(m);
    __shared__ unsigned long long shm[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xFFFFFFFFu, mx, o);
        mx = t > mx ? t : mx;
    }
    if ((threadIdx.x & 31) == 0) shm[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (unsigned int w = 1; w < (blockDim.x >> 5); ++w) mx = shm[w] > mx ? shm[w] : mx;
        if (mx) atomicMax(&s->max_mass_bits, mx);
    }
}

// scan loader: 1 where sorted position i starts a new depth-D cell
struct HeadFlag {
    const void* keys;
    const Scal* s;
    int k32;
    __device__ __forceinline__ unsigned int operator()(int i) const {
        if ((unsigned int)i >= s->n_in) return 0u;
        return (i == 0 || load_key(keys, k32, i) != load_key(keys, k32, i - 1)) ? 1u : 0u;
    }
};

// sinks of the single-pass scans
// ---- 3. terminals: the head-flag scan writes the terminal arrays itself ----------------------------------------
struct TerminalSink {
    const void* keys;
    int k32;
    unsigned long long* tkey;
    unsigned int* tfirst;
    Scal* s;
    int n;
    __device__ __forceinline__ void operator()(int i, unsigned int excl, unsigned int head) const {
        const unsigned int n_in = s->n_in;
        if ((unsigned int)i < n_in) {
            if (head) {
                tkey[excl] = load_key(keys, k32, i);
                tfirst[excl] = (unsigned int)i;
            }
            if ((unsigned int)i == n_in - 1u) tfirst[excl + head] = n_in;
        }
        if (i == n) {
            s->n_term = excl;   // grand total of the head flags
            if (n_in == 0u) tfirst[0] = 0u;
        }
    }
};
struct StoreSink {      // plain exclusive prefix, out[0..n]
    unsigned int* out;
    __device__ __forceinline__ void operator()(int i, unsigned int excl, unsigned int) const { out[i] = excl; }
};

// ---- 4. witnesses: delta[], the per-terminal level masks, and the number of branching cells per level ------
// wstart[t] = first terminal of the cell that t witnesses (the only galloping search of the whole build).
__global__ void __launch_bounds__(256)
k_witness(int D, const unsigned long long* __restrict__ tkey, signed char* __restrict__ delta,
          unsigned int* __restrict__ mask, unsigned int* __restrict__ wstart, unsigned int* __restrict__ levelCount,
          const Scal* __restrict__ s) {
    __shared__ unsigned int cnt[32];
    if (threadIdx.x < 32) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_term = (int)s->n_term;
    if (t < n_term) {
        if (t == n_term - 1) {
            delta[t] = -1;
        } else {
            const unsigned long long kt = tkey[t];
            const int L = lca_level(kt, tkey[t + 1], D);
            delta[t] = (signed char)L;
            const int shift = 2 * (D - L);
            const int a = cell_first(tkey, t, shift);
            LPE_CHECK_NR(a >= 0 && a <= t && L >= 0 && L < D, 2, const_cast<Scal*>(s));
            wstart[t] = (unsigned int)a;
            atomicOr(&mask[a], 1u << L);
            // the witness that sits in the cell's first non-empty child counts the cell (once per cell)
            if ((tkey[a] >> (shift - 2)) == (kt >> (shift - 2))) atomicAdd(&cnt[L], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < 32 && cnt[threadIdx.x]) atomicAdd(&levelCount[threadIdx.x], cnt[threadIdx.x]);
}

struct MaskPop {
    const unsigned int* mask;
    const Scal* s;
    __device__ __forceinline__ unsigned int operator()(int t) const {
        return ((unsigned int)t < s->n_term) ? (unsigned int)__popc(mask[t]) : 0u;
    }
};

// ---- 5. topology: pre-order indices, skip pointers, child slots, per-level cell lists ------------------------
struct Topo {
    const unsigned int* wstart; // [terminal] first terminal of the cell that t witnesses
    unsigned int* child;      // [4 * ordinal + digit] a cell's child: pre-order index, or LPE_LEAF_FLAG | sorted body
                              // position for a single-body leaf, or LPE_NONE
    Agg* agg;                 // [preorder] written here only for aggregated terminals (>= 2 bodies in a depth-D cell)
    uint2* levelList;         // cells grouped by level: levelList[levelBase[L] + i] = {pre-order index, cell ordinal}
    const unsigned int* levelCount;   // cells per level (k_witness)
    unsigned int* levelBase;          // exclusive scan of levelCount: every block makes its own copy, block 0 stores it for the aggregation
    unsigned int* levelCursor;        // (zeroed with the step's scratch)
    const unsigned int* tfirst;
    const Body* body;         // state in key order
    unsigned int* selfslot;   // [sorted body] record slot of its own leaf (only written here for a one-terminal tree)
    TravRec* rec;
};

// Power-of-two mass unit just above the largest source mass: node masses then fit fp32 comfortably
// (<= 2N units) whatever the caller's units are (1e36 kg in the Keplerian scenario).
__device__ __forceinline__ double mass_scale_inv(unsigned long long max_mass_bits) {
    if (max_mass_bits == 0ull) return 1.0;
    const int e = (int)((max_mass_bits >> 52) & 0x7FFull) - 1023;
    return ldexp(1.0, -(e + 1));
}

// aggregate of one body
__device__ __forceinline__ Agg body_agg(const Body& sb, unsigned int pos, double thr) {
    Agg a;
    a.m = sb.m; a.sx = sb.m * sb.x; a.sy = sb.m * sb.y;
    a.mf = sb.m; a.xf = sb.x; a.yf = sb.y;
    a.frank = sb.rank; a.fidx = pos; a.ordinal = 0u; a.small = (sb.m >= thr) ? 0u : 1u;
    return a;
}

// Traversal record of a node from its aggregate.
__device__ __forceinline__ TravRec make_record(const StepConst& c, const Agg& a, int level, unsigned int node,
                                               unsigned int cblockIndex, double massScaleInv) {
    double M, cx, cy;
    node_centre(a, level, c.quirk, M, cx, cy);
    const double cxs = cx * c.invS, cys = cy * c.invS;
    TravRec r;
    const float hx = (float)cxs, hy = (float)cys;
    r.c = make_float4(hx, hy, (float)(cxs - (double)hx), (float)(cys - (double)hy));
    // allSmall cells are skipped by the traversal but still feed their ancestors (barnes_hut.cpp:253, Q7):
    // a zero mass with "never open" is exactly that.
    const bool skipSmall = (c.thr > 0.0) && (a.small & 1u);
    r.gm = skipSmall ? 0.0f : (float)(M * massScaleInv);
    r.open_t = skipSmall ? -2.0f : -1.0f;
    if (level >= 0 && !skipSmall) {
        const double s = ldexp(c.U, -level) * c.invS;
        r.open_t = (float)((s * s) / c.theta2);
    }
    r.node = node;
    r.cblock = (level >= 0) ? ((cblockIndex << 2) | ((a.small >> 1) & 3u)) : 0u;
    return r;
}

__device__ __forceinline__ TravRec invalid_record() {
    TravRec r;
    r.c = make_float4(0.f, 0.f, 0.f, 0.f);
    r.gm = 0.f;
    r.open_t = -1.f;
    r.node = LPE_NONE;
    r.cblock = 0u;
    return r;
}

// No searches here: a cell's FIRST child is the next deeper cell that starts at the same terminal (pre-order index
// + 1) or the terminal itself; every OTHER child starts right after a witness of the cell (terminal t + 1 where
// delta[t] = level of the cell) and is the shallowest cell starting there, or that terminal. Skip pointers are filled
// bottom-up by the aggregation (a cell ends where its last child ends).
__global__ void __launch_bounds__(256)
k_topology(StepConst c, const unsigned long long* __restrict__ tkey, const signed char* __restrict__ delta,
           const unsigned int* __restrict__ mask, const unsigned int* __restrict__ P, Topo o, Scal* __restrict__ s) {
    __shared__ unsigned int cnt[32], base[32], lbase[32];
    if (threadIdx.x < 32) {
        cnt[threadIdx.x] = 0;
        // start of each level's list = exclusive scan of the 32 level counts (one warp; no launch of its own)
        const unsigned int v = o.levelCount[threadIdx.x];
        unsigned int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int u = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if ((int)threadIdx.x >= d) inc += u;
        }
        lbase[threadIdx.x] = inc - v;
        if (blockIdx.x == 0) o.levelBase[threadIdx.x] = inc - v;
    }
    __syncthreads();
    const int D = c.D;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_term = (int)s->n_term;
    if (t == 0) s->n_internal = P[n_term];
    const bool live = t < n_term;
    unsigned int mk = 0, Pt = 0;
    if (live) {
        mk = mask[t];
        Pt = P[t];
        const unsigned long long kt = tkey[t];
        const unsigned int ncell = (unsigned int)__popc(mk);
        // the terminal itself: a single-body leaf, or the sum of the bodies that share its depth-D cell
        const unsigned int idx = (unsigned int)t + Pt + ncell;
        const unsigned int first = o.tfirst[t], last = o.tfirst[t + 1];
        LPE_CHECK(idx < c.nodeCap && first < last && last <= c.bodyCap && Pt + ncell <= c.bodyCap, 3, s);
        const bool single = (last - first) == 1u;
        Agg a;
        if (single) {
            if (n_term == 1) a = body_agg(o.body[first], first, c.thr);
        } else {
            a.m = 0.0; a.sx = 0.0; a.sy = 0.0; a.mf = 0.0; a.xf = 0.0; a.yf = 0.0;
            a.frank = 0xFFFFFFFFu; a.fidx = first; a.ordinal = 0u; a.small = 1u;   // (level bits 0 = aggregated terminal)
            // Bodies that share a depth-D cell are summed in insertion-rank order, not in sorted-position order (which
            // for equal keys is an accident of the previous permutation): the sums are then the same however the
            // bodies reached this GPU (single GPU, or migrated between the ranks of a domain-decomposed run).
            const bool canonical = (last - first) <= 64u;
            unsigned int prev = 0u;
            for (unsigned int k = first; k < last; ++k) {
                unsigned int i = k;
                if (canonical) {   // next larger rank (ranks are unique)
                    unsigned int best = 0xFFFFFFFFu;
                    for (unsigned int j = first; j < last; ++j) {
                        const unsigned int r = o.body[j].rank;
                        if ((k == first || r > prev) && r < best) { best = r; i = j; }
                    }
                    prev = best;
                }
                const Body sb = o.body[i];
                a.m += sb.m; a.sx += sb.m * sb.x; a.sy += sb.m * sb.y;
                if (sb.rank < a.frank) { a.frank = sb.rank; a.fidx = i; a.mf = sb.m; a.xf = sb.x; a.yf = sb.y; }
                if (sb.m >= c.thr) a.small = 0u;
            }
            o.agg[idx] = a;
        }
        const unsigned int tcode = single ? (LPE_LEAF_FLAG | first) : idx;
        if (n_term == 1 && !c.dd) {   // a tree of one terminal: it is the root
            const double msi = mass_scale_inv(s->max_mass_bits);
            o.rec[0] = make_record(c, a, single ? -1 : -2, single ? (LPE_LEAF_FLAG | first) : idx, 0u, msi);
            o.rec[1] = o.rec[2] = o.rec[3] = invalid_record();
            if (single && c.need_self) o.selfslot[first] = 0u;
        }
        // branching cells whose first terminal is t, shallow to deep: ordinal Pt + i, pre-order t + Pt + i
        unsigned int rest = mk, i = 0;
        while (rest) {
            const int L = __ffs(rest) - 1;
            rest &= rest - 1;
            const unsigned int cidx = (unsigned int)t + Pt + i;
            // first child: the next deeper cell starting here, else the terminal
            const unsigned int digit = (unsigned int)(kt >> (2 * (D - L - 1))) & 3u;
            o.child[(size_t)(Pt + i) * 4 + digit] = rest ? cidx + 1u : tcode;
            atomicAdd(&cnt[L], 1u);
            ++i;
        }
        // as the witness of the cell at level delta[t]: the child that starts at terminal t + 1
        if (t < n_term - 1) {
            const int L = (int)delta[t];
            const unsigned int a0 = o.wstart[t];
            const unsigned int q = P[a0] + (unsigned int)__popc(mask[a0] & ((1u << L) - 1u));
            LPE_CHECK(a0 <= (unsigned int)t && q < c.bodyCap, 4, s);
            const unsigned int mk1 = mask[t + 1];
            unsigned int code = (unsigned int)(t + 1) + P[t + 1];   // shallowest cell starting at t+1, or that terminal's node
            if (mk1 == 0u) {
                const unsigned int f1 = o.tfirst[t + 1];
                if (o.tfirst[t + 2] - f1 == 1u) code = LPE_LEAF_FLAG | f1;
            }
            const unsigned int digit = (unsigned int)(tkey[t + 1] >> (2 * (D - L - 1))) & 3u;
            o.child[(size_t)q * 4 + digit] = code;
        }
    }
    // per-level lists: one global reservation per level per block, then block-local slots
    __syncthreads();
    if (threadIdx.x < 32) {
        const unsigned int cc = cnt[threadIdx.x];
        base[threadIdx.x] = cc ? atomicAdd(&o.levelCursor[threadIdx.x], cc) : 0u;
        cnt[threadIdx.x] = 0;
    }
    __syncthreads();
    unsigned int rest = mk, i = 0;
    while (rest) {
        const int L = __ffs(rest) - 1;
        rest &= rest - 1;
        const unsigned int local = atomicAdd(&cnt[L], 1u);
        o.levelList[lbase[L] + base[L] + local] = make_uint2((unsigned int)t + Pt + i, Pt + i);
        ++i;
    }
}

// ---- 6. aggregation and traversal records ---------------------------------------------------------------------
struct NodeOut {
    Agg* agg;               // [preorder] (cells and aggregated terminals; single-body leaves have none)
    TravRec* rec;           // [4 * block + slot]; block 0 = {root}, block q+1 = children of the cell with ordinal q
    unsigned int* selfslot; // [sorted body] record slot of the body's own single-body leaf, LPE_NONE otherwise
    const Body* body;       // state in key order
};

// One branching cell, handled by a QUAD of lanes: lane q owns child slot q. The four child reads are independent
// (4x the memory parallelism of one thread per cell), the four 32-byte records of the child block are written by four
// neighbouring lanes (one 128-byte line), and the sums are combined with two shuffle steps: (c0 + c1) + (c2 + c3),
// deterministic. `p` is uniform across the quad; lanes of a quad must call this together.
__device__ __forceinline__ void aggregate_cell_quad(const StepConst& c, const NodeOut& o, unsigned int p,
                                                    unsigned int qd, int cellLevel,
                                                    unsigned int ci, double msi, int q,
                                                    unsigned int quadShift, bool live, const Scal* chk) {
    // every lane of the warp runs this (the ballot / shuffles use the full mask); quads past the end of the list
    // carry live = false and neither read children nor store anything. The level list carries the cell's ordinal
    // next to its pre-order index, so the child codes are fetched without a detour through the cell's own meta.
    // (ci = the lane's child code, child[4 * ordinal + q], fetched by the caller one round ahead; LPE_NONE when !live)
    const bool valid = ci != LPE_NONE;
    Agg a;
    a.m = 0.0; a.sx = 0.0; a.sy = 0.0; a.mf = 0.0; a.xf = 0.0; a.yf = 0.0;
    a.frank = 0xFFFFFFFFu; a.fidx = 0; a.ordinal = 0; a.small = 1u;
    int level = -1;
    unsigned int cbi = 0, leafpos = LPE_NONE;
    if (valid) {
        if (ci & LPE_LEAF_FLAG) {
            // a single-body leaf: read the body itself (one sector), no aggregate was ever stored for it
            leafpos = lpe_idx(ci & ~LPE_LEAF_FLAG, c.bodyCap, 6, chk);
            a = body_agg(o.body[leafpos], leafpos, c.thr);
        } else {
            // a deeper cell or an aggregated terminal: its aggregate says what it is (level) and where its children are
            a = o.agg[lpe_idx(ci, c.nodeCap, 7, chk)];
            level = agg_level(a);
            cbi = (level >= 0) ? a.ordinal + c.blockBase : 0u;
        }
    }
    const unsigned int vmask = (__ballot_sync(0xFFFFFFFFu, valid) >> quadShift) & 0xFu;
    const unsigned int below = (1u << q) - 1u;
    const unsigned int nvalid = __popc(vmask);
    const unsigned int r = valid ? __popc(vmask & below) : nvalid + __popc(~vmask & below & 0xFu);
    const unsigned int slot = lpe_idx(4u * (qd + c.blockBase) + r, c.recSlots, 8, chk);
    p = lpe_idx(p, c.nodeCap, 9, chk);
    if (valid) {
        o.rec[slot] = make_record(c, a, level, (leafpos != LPE_NONE) ? (LPE_LEAF_FLAG | leafpos) : ci, cbi, msi);
        if (leafpos != LPE_NONE && c.need_self) o.selfslot[leafpos] = slot;
    } else if (live) {
        o.rec[slot] = invalid_record();
    }
    // quad reduction
    a.small &= 1u;
#pragma unroll
    for (int sft = 1; sft <= 2; sft <<= 1) {
        const double om = __shfl_xor_sync(0xFFFFFFFFu, a.m, sft), osx = __shfl_xor_sync(0xFFFFFFFFu, a.sx, sft);
        const double osy = __shfl_xor_sync(0xFFFFFFFFu, a.sy, sft), omf = __shfl_xor_sync(0xFFFFFFFFu, a.mf, sft);
        const double oxf = __shfl_xor_sync(0xFFFFFFFFu, a.xf, sft), oyf = __shfl_xor_sync(0xFFFFFFFFu, a.yf, sft);
        const unsigned int ofr = __shfl_xor_sync(0xFFFFFFFFu, a.frank, sft), ofi = __shfl_xor_sync(0xFFFFFFFFu, a.fidx, sft);
        const unsigned int osm = __shfl_xor_sync(0xFFFFFFFFu, a.small, sft);
        // the lower lane of each pair adds (own + other) so that the order is the same on both sides
        const bool lower = (q & sft) == 0;
        a.m = lower ? a.m + om : om + a.m;
        a.sx = lower ? a.sx + osx : osx + a.sx;
        a.sy = lower ? a.sy + osy : osy + a.sy;
        if (ofr < a.frank) { a.frank = ofr; a.fidx = ofi; a.mf = omf; a.xf = oxf; a.yf = oyf; }
        a.small &= osm;
    }
    if (q == 0 && live) {
        // children - 1 and the cell's own level and ordinal: read back when its parent makes this cell's record
        a.small = (a.small & 1u) | ((nvalid - 1u) << 1) | ((unsigned int)(cellLevel + 2) << 8);
        a.ordinal = qd;
        o.agg[p] = a;
        if (p == 0 && !c.dd) {   // the root has no parent to write its record
            o.rec[0] = make_record(c, a, cellLevel, 0u, c.blockBase, msi);
            o.rec[1] = o.rec[2] = o.rec[3] = invalid_record();
        }
    }
}

// all branching cells of one level (children are at deeper levels: finished by earlier launches); 4 lanes per cell
__global__ void __launch_bounds__(256)
k_agg_level(StepConst c, int L, const uint2* __restrict__ levelList, const unsigned int* __restrict__ levelBase,
            const unsigned int* __restrict__ levelCount, const unsigned int* __restrict__ child, NodeOut o,
            const Scal* __restrict__ s) {
    const unsigned int count = levelCount[L], base = levelBase[L];
    const double msi = mass_scale_inv(s->max_mass_bits);
    const int q = threadIdx.x & 3;
    const unsigned int quadShift = (threadIdx.x & 31) & ~3u;
    const unsigned int quads = (gridDim.x * blockDim.x) >> 2;
    // whole warps stay in the loop together (the ballot inside needs all 32 lanes). A cell costs a chain of four dependent
    // loads (list entry -> child code -> child aggregate -> ...): the first two links of the NEXT round's cell are fetched
    // before this round's cell is worked on, so consecutive rounds overlap instead of queueing up their latencies.
    const unsigned int rounds = (count + quads - 1) / quads;
    unsigned int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    auto fetch = [&](unsigned int idx, uint2& e, unsigned int& ci) {
        const bool lv = idx < count;
        e = levelList[base + (lv ? idx : 0u)];
        ci = lv ? child[(size_t)lpe_idx(e.y, c.bodyCap, 5, s) * 4 + q] : LPE_NONE;
    };
    uint2 e;
    unsigned int ci;
    fetch(i, e, ci);
    for (unsigned int k = 0; k < rounds; ++k, i += quads) {
        const bool live = i < count;
        uint2 en = e;
        unsigned int cin = LPE_NONE;
        if (k + 1 < rounds) fetch(i + quads, en, cin);
        if (__any_sync(0xFFFFFFFFu, live)) aggregate_cell_quad(c, o, e.x, e.y, L, ci, msi, q, quadShift, live, s);
        e = en;
        ci = cin;
    }
}

// the few cells of levels Ltop..0 in one block (a level has at most 4^L cells), one __syncthreads per level
__global__ void __launch_bounds__(1024)
k_agg_top(StepConst c, int Ltop, const uint2* __restrict__ levelList, const unsigned int* __restrict__ levelBase,
          const unsigned int* __restrict__ levelCount, const unsigned int* __restrict__ child, NodeOut o,
          const Scal* __restrict__ s) {
    const double msi = mass_scale_inv(s->max_mass_bits);
    const int q = threadIdx.x & 3;
    const unsigned int quadShift = (threadIdx.x & 31) & ~3u;
    const unsigned int quads = blockDim.x >> 2;
    for (int L = Ltop; L >= 0; --L) {
        const unsigned int count = levelCount[L], base = levelBase[L];
        const unsigned int rounds = (count + quads - 1) / quads;
        unsigned int i = threadIdx.x >> 2;
        for (unsigned int k = 0; k < rounds; ++k, i += quads) {
            const bool live = i < count;
            const uint2 e = levelList[base + (live ? i : 0u)];
            const unsigned int ci = live ? child[(size_t)lpe_idx(e.y, c.bodyCap, 5, s) * 4 + q] : LPE_NONE;
            if (__any_sync(0xFFFFFFFFu, live)) aggregate_cell_quad(c, o, e.x, e.y, L, ci, msi, q, quadShift, live, s);
        }
        __syncthreads();
    }
}

}  // namespace lpe
