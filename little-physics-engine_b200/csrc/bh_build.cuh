// bh_build.cuh — Morton keys, terminals, path-compressed quadtree topology in pre-order, and the
// bottom-up mass / centre-of-mass / first-occupant aggregation.
//
// What the reference builds by recursive insertion (barnes_hut.cpp:101-238) is rebuilt here from sorted
// keys. Equivalences used (all verified against the compiled reference):
//   Q10  child digit = (x >= mid) + 2*(y >= mid)  =>  reference DFS order == Morton order of the cells
//   Q3   a chain of single-child cells carries one (M, COM); only its lowest cell's size matters
//        => keep only branching cells (>= 2 non-empty children): a path-compressed tree
//   Q4   cells smaller than theta*eps are always accepted => stop at depth D (terminals may aggregate)
//   Q2   every internal cell counts its first occupant (minimum insertion rank) twice
//
// Topology without a level loop: for sorted terminals t (distinct depth-D cells), delta[t] = LCA level of
// t and t+1. Every t is a "witness" of the branching cell at level delta[t] that contains t and t+1; that
// cell is identified by (a, L) with a = its first terminal. A 32-bit level mask per terminal collects the
// levels of the cells that start there (atomicOr), and
//     preorder(a, L) = a + P[a] + popc(mask[a] & ((1<<L)-1)),   P = exclusive scan of popc(mask)
//     preorder(leaf t) = t + P[t] + popc(mask[t])
//     skip(a..b)       = (b+1) + P[b+1]
//     parent level     = max(delta[a-1], delta[b])
#pragma once
#include "bh_common.cuh"

namespace lpe {

// ---- 1. Morton keys -------------------------------------------------------------------------------------
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

__global__ void __launch_bounds__(256)
k_keygen(StepConst c, const double2* __restrict__ pos, const double* __restrict__ mass,
         const unsigned char* __restrict__ comp, unsigned long long* __restrict__ keys,
         unsigned int* __restrict__ vals, Scal* __restrict__ s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int in = 0;
    unsigned long long mbits = 0;
    if (i < c.n) {
        const unsigned char cm = comp[i];
        const double2 p = pos[i];
        // buildTree's view and bounds test, barnes_hut.cpp:117-124
        const bool src = (cm & 1u) && !(cm & 4u);
        const bool inside = src && p.x >= 0.0 && p.x < c.U && p.y >= 0.0 && p.y < c.U;
        unsigned long long key = 1ull << (2 * c.D);  // not in the tree: sorts after every cell
        if (inside) {
            const unsigned int kmax = (1u << c.D) - 1u;
            const unsigned int ix = cell_index(p.x, c.h, c.invh, kmax);
            const unsigned int iy = cell_index(p.y, c.h, c.invh, kmax);
            key = spread_bits32(ix) | (spread_bits32(iy) << 1);
            in = 1;
            const double m = mass[i];
            if (m > 0.0) mbits = (unsigned long long)__double_as_longlong(m);
        }
        keys[i] = key;
        vals[i] = (unsigned int)i;
    }
    // block reduce: count and max
    const unsigned int cnt = __syncthreads_count(in);
    __shared__ unsigned long long shm[8];
    unsigned long long mx = mbits;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xFFFFFFFFu, mx, o);
        mx = t > mx ? t : mx;
    }
    if ((threadIdx.x & 31) == 0) shm[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) mx = shm[w] > mx ? shm[w] : mx;
        if (cnt) atomicAdd(&s->n_in, cnt);
        if (mx) atomicMax(&s->max_mass_bits, mx);
    }
}

// ---- 2. gather into Morton order --------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_gather(int n, const unsigned int* __restrict__ sidx, const double2* __restrict__ pos,
         const double* __restrict__ mass, const unsigned int* __restrict__ rank, double2* __restrict__ spos,
         double* __restrict__ smass, unsigned int* __restrict__ srank) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int b = sidx[i];
    spos[i] = pos[b];
    smass[i] = mass[b];
    srank[i] = rank[b];
}

// scan loader: 1 where sorted position i starts a new depth-D cell
struct HeadFlag {
    const unsigned long long* keys;
    const Scal* s;
    __device__ __forceinline__ unsigned int operator()(int i) const {
        if ((unsigned int)i >= s->n_in) return 0u;
        return (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
    }
};

// ---- 3. terminals -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_terminals(int n, const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ headExcl,
            unsigned long long* __restrict__ tkey, unsigned int* __restrict__ tfirst, Scal* __restrict__ s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int n_in = s->n_in;
    if (i == 0) {
        s->n_term = headExcl[n];  // grand total of the head flags
        if (n_in == 0) tfirst[0] = 0;
    }
    if ((unsigned int)i >= n_in) return;
    const bool head = (i == 0) || keys[i] != keys[i - 1];
    const unsigned int t = headExcl[i];
    if (head) {
        tkey[t] = keys[i];
        tfirst[t] = (unsigned int)i;
    }
    if ((unsigned int)i == n_in - 1) tfirst[t + (head ? 1u : 0u)] = n_in;
}

// ---- 4. witnesses: delta[] and the per-terminal level masks -------------------------------------------------
__global__ void __launch_bounds__(256)
k_witness(int D, const unsigned long long* __restrict__ tkey, signed char* __restrict__ delta,
          unsigned int* __restrict__ mask, const Scal* __restrict__ s) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_term = (int)s->n_term;
    if (t >= n_term) return;
    if (t == n_term - 1) {
        delta[t] = -1;
        return;
    }
    const int L = lca_level(tkey[t], tkey[t + 1], D);
    delta[t] = (signed char)L;
    const int a = cell_first(tkey, t, 2 * (D - L));
    atomicOr(&mask[a], 1u << L);
}

struct MaskPop {
    const unsigned int* mask;
    const Scal* s;
    __device__ __forceinline__ unsigned int operator()(int t) const {
        return ((unsigned int)t < s->n_term) ? (unsigned int)__popc(mask[t]) : 0u;
    }
};

// ---- 5. topology: pre-order indices, skip pointers, parents, child slots ------------------------------------
__global__ void __launch_bounds__(256)
k_topology(int D, const unsigned long long* __restrict__ tkey, const signed char* __restrict__ delta,
           const unsigned int* __restrict__ mask, const unsigned int* __restrict__ P,
           unsigned int* __restrict__ tnode, unsigned int* __restrict__ parent, unsigned int* __restrict__ child,
           NodeB* __restrict__ nodeB, signed char* __restrict__ nlevel, unsigned int* __restrict__ nodeStart,
           Scal* __restrict__ s) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_term = (int)s->n_term;
    if (t == 0) s->n_internal = P[n_term];
    if (t >= n_term) return;
    const unsigned int mk = mask[t];
    const unsigned int Pt = P[t];
    const unsigned long long kt = tkey[t];
    const int dl = (t > 0) ? (int)delta[t - 1] : -1;

    // the terminal itself
    {
        const unsigned int idx = (unsigned int)t + Pt + (unsigned int)__popc(mk);
        tnode[t] = idx;
        const int dr = (int)delta[t];  // -1 for the last terminal
        const int Lp = max(dl, dr);
        unsigned int par = LPE_NONE;
        if (Lp >= 0) {
            const int ap = (dl < Lp) ? t : cell_first(tkey, t, 2 * (D - Lp));
            par = (unsigned int)ap + P[ap] + (unsigned int)__popc(mask[ap] & ((1u << Lp) - 1u));
            const unsigned int digit = (unsigned int)(kt >> (2 * (D - Lp - 1))) & 3u;
            child[(size_t)par * 4 + digit] = idx;
        }
        parent[idx] = par;
        nodeB[idx].skip = idx + 1;
        nodeStart[idx] = (unsigned int)t;
    }
    // branching cells whose first terminal is t
    unsigned int rest = mk;
    while (rest) {
        const int L = __ffs(rest) - 1;
        rest &= rest - 1;
        const unsigned int idx = (unsigned int)t + Pt + (unsigned int)__popc(mk & ((1u << L) - 1u));
        const int b = cell_last(tkey, t, 2 * (D - L), n_term);
        const int dr = (int)delta[b];
        const int Lp = max(dl, dr);
        unsigned int par = LPE_NONE;
        if (Lp >= 0) {
            const int ap = (dl < Lp) ? t : cell_first(tkey, t, 2 * (D - Lp));
            par = (unsigned int)ap + P[ap] + (unsigned int)__popc(mask[ap] & ((1u << Lp) - 1u));
            const unsigned int digit = (unsigned int)(kt >> (2 * (D - Lp - 1))) & 3u;
            child[(size_t)par * 4 + digit] = idx;
        }
        parent[idx] = par;
        nodeB[idx].skip = (unsigned int)(b + 1) + P[b + 1];
        nlevel[idx] = (signed char)L;
        nodeStart[idx] = (unsigned int)t;
    }
}

// Power-of-two mass unit just above the largest source mass: node masses then fit fp32 comfortably
// (<= 2N units) whatever the caller's units are (1e36 kg in the Keplerian scenario).
__device__ __forceinline__ double mass_scale_inv(unsigned long long max_mass_bits) {
    if (max_mass_bits == 0ull) return 1.0;
    const int e = (int)((max_mass_bits >> 52) & 0x7FFull) - 1023;
    return ldexp(1.0, -(e + 1));
}

// ---- 6. aggregation: leaves write their record, the last child to arrive sums its siblings ------------------
struct NodeOut {
    double2* nodeA;       // fp64 centre (scaled): exact re-test of borderline theta decisions, STRICT mode, dumps
    float4* nodeC;        // two-float centre (scaled): the traversal's inner loop
    NodeB* nodeB;
    double* nodeM;
    signed char* nlevel;  // level of a branching cell, -1 leaf, -2 aggregated terminal
};

constexpr float OPEN_BAND = 4e-6f;  // relative half-width of the fp32 guard band around s^2/theta^2

__device__ __forceinline__ void finalize_node(const StepConst& c, const NodeOut& o, unsigned int idx, const Agg& a,
                                              int level, double massScaleInv, const double2* __restrict__ spos,
                                              const double* __restrict__ smass) {
    double M = a.m, sx = a.sx, sy = a.sy;
    double cx, cy;
    if (level == -1) {
        // single-body leaf: exactly the body (barnes_hut.cpp:144-153)
        const double2 p = spos[a.fidx];
        cx = p.x;
        cy = p.y;
    } else {
        if (c.quirk) {
            // first occupant counted twice in every internal cell (barnes_hut.cpp:157-177)
            const double mf = smass[a.fidx];
            const double2 pf = spos[a.fidx];
            M += mf;
            sx += mf * pf.x;
            sy += mf * pf.y;
        }
        cx = sx / M;
        cy = sy / M;
    }
    const double cxs = cx * c.invS, cys = cy * c.invS;
    o.nodeA[idx] = make_double2(cxs, cys);
    const float hx = (float)cxs, hy = (float)cys;
    o.nodeC[idx] = make_float4(hx, hy, (float)(cxs - (double)hx), (float)(cys - (double)hy));
    o.nodeM[idx] = M;
    // allSmall cells are skipped by the traversal but still feed their ancestors (barnes_hut.cpp:253, Q7):
    // a zero mass with "never open" is exactly that.
    const bool skipSmall = (c.thr > 0.0) && a.small;
    NodeB nb;
    nb.skip = o.nodeB[idx].skip;  // written by k_topology
    nb.gm = skipSmall ? 0.0f : (float)(M * massScaleInv);
    nb.open_lo = nb.open_hi = skipSmall ? -2.0f : -1.0f;
    if (level >= 0 && !skipSmall) {
        const double s = ldexp(c.U, -level) * c.invS;
        const float t = (float)((s * s) / c.theta2);
        nb.open_lo = t * (1.0f - OPEN_BAND);
        nb.open_hi = t * (1.0f + OPEN_BAND);
    }
    o.nodeB[idx] = nb;
    o.nlevel[idx] = (signed char)level;
}

__global__ void __launch_bounds__(256)
k_aggregate(StepConst c, const unsigned int* __restrict__ tfirst, const unsigned int* __restrict__ tnode,
            const double2* __restrict__ spos, const double* __restrict__ smass,
            const unsigned int* __restrict__ srank, const unsigned int* __restrict__ parent,
            const unsigned int* __restrict__ child, unsigned int* __restrict__ arrived, Agg* __restrict__ agg,
            NodeOut o, unsigned int* __restrict__ selfnode, const Scal* __restrict__ s) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_term = (int)s->n_term;
    if (t >= n_term) return;
    const double massScaleInv = mass_scale_inv(s->max_mass_bits);
    const unsigned int first = tfirst[t], last = tfirst[t + 1];
    Agg a;
    a.m = 0.0; a.sx = 0.0; a.sy = 0.0; a.frank = 0xFFFFFFFFu; a.fidx = first; a.count = last - first; a.small = 1u;
    a.pad[0] = a.pad[1] = 0u;
    for (unsigned int i = first; i < last; ++i) {
        const double m = smass[i];
        const double2 p = spos[i];
        a.m += m;
        a.sx += m * p.x;
        a.sy += m * p.y;
        const unsigned int r = srank[i];
        if (r < a.frank) { a.frank = r; a.fidx = i; }
        if (m >= c.thr) a.small = 0u;
    }
    unsigned int idx = tnode[t];
    const bool single = (last - first) == 1;
    finalize_node(c, o, idx, a, single ? -1 : -2, massScaleInv, spos, smass);
    for (unsigned int i = first; i < last; ++i) selfnode[i] = single ? idx : LPE_NONE;
    agg[idx] = a;

    // walk up: the last child to arrive at a cell owns it
    unsigned int p = parent[idx];
    for (int hop = 0; hop <= LPE_MAX_DEPTH + 1 && p != LPE_NONE; ++hop) {  // a branching chain is at most D cells long
        const uint4 ch = reinterpret_cast<const uint4*>(child)[p];
        const unsigned int need = (ch.x != LPE_NONE) + (ch.y != LPE_NONE) + (ch.z != LPE_NONE) + (ch.w != LPE_NONE);
        __threadfence();
        const unsigned int old = atomicAdd(&arrived[p], 1u);
        if (old + 1u < need) return;
        __threadfence();
        Agg b;
        b.m = 0.0; b.sx = 0.0; b.sy = 0.0; b.frank = 0xFFFFFFFFu; b.fidx = 0; b.count = 0; b.small = 1u;
        b.pad[0] = b.pad[1] = 0u;
        const unsigned int cs[4] = {ch.x, ch.y, ch.z, ch.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (cs[q] == LPE_NONE) continue;
            // written by another SM just before its atomicAdd: read through L2, not a stale L1 line
            const double2* ad = reinterpret_cast<const double2*>(&agg[cs[q]]);
            const double2 w0 = __ldcg(ad);
            const double2 w1 = __ldcg(ad + 1);
            const uint4 w2 = __ldcg(reinterpret_cast<const uint4*>(ad + 2));
            const unsigned long long fr = (unsigned long long)__double_as_longlong(w1.y);
            const unsigned int frank = (unsigned int)(fr & 0xFFFFFFFFull), fidx = (unsigned int)(fr >> 32);
            b.m += w0.x; b.sx += w0.y; b.sy += w1.x;
            if (frank < b.frank) { b.frank = frank; b.fidx = fidx; }
            b.count += w2.x;
            b.small &= w2.y;
        }
        finalize_node(c, o, p, b, (int)o.nlevel[p], massScaleInv, spos, smass);
        agg[p] = b;
        p = parent[p];
    }
}

}  // namespace lpe
