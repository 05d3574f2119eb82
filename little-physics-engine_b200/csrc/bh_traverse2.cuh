// bh_traverse2.cuh — two-phase traversal (FAST precision): the production force kernel.
//
// Same result as the depth-first walk in bh_traverse.cuh — every lane (body) takes the reference's own accept/open
// decision for every node (barnes_hut.cpp:266-269) — but the work is organised so that the expensive part runs with
// all 32 lanes busy and no per-node control flow. One warp owns 32 key-consecutive targets and a FIFO ring of
// {record slot, lane mask} entries in shared memory; the mask holds the targets that actually REACH the node (a
// target that accepted an ancestor must never see it). Per round:
//
//   phase 1 (cooperative, one NODE per lane): 32 ring entries are popped, their records gathered, and every node is
//     classified against the bounding box of the warp's targets, conservatively:
//       A  every target accepts it   (d2_min >= s^2/theta^2 + margin, or it is a leaf / terminal) -> accept list,
//                                     with its mask
//       O  every target opens it     (d2_max <= s^2/theta^2 - margin) -> its children inherit its mask
//       M  mixed / too close to call -> mixed list of this round
//   phase 2a (one BODY per lane, the round's mixed nodes, two per iteration with packed fp32): each lane that reaches
//     the node takes the per-lane theta test (fp32 against the two edges of a guard band; if any lane lands inside
//     a band the round's loop is redone with the reference's fp64 expression deciding those cases) and accumulates
//     if it accepted. The ballot of the lanes that opened becomes the mask of the node's children; children nobody
//     opened are dropped with their whole subtree.
//   push: the children of O and M nodes are appended to the ring (positions from three ballots, one per bit of the
//     child count).
//   phase 2b (one BODY per lane, when >= 64 entries have gathered or the ring is empty): the accept list — no test at
//     all, two entries per iteration: two-float difference, rsqrt, three multiplies, two FMAs per interaction, the
//     mask only selects a zero mass.
//
// A warp whose ring would overflow hands its chunk to the depth-first kernel (second launch).
//
// In front of the per-warp walk, the four warps of a CTA take the far field of their 128 bodies together (template flag
// CTA, see T3Cta below): what every body of the block accepts or opens is classified once, not four times.
#pragma once
#include <type_traits>
#include "bh_common.cuh"
#include "bh_traverse.cuh"

namespace lpe {

#ifndef T2_MIN_CTAS
#define T2_MIN_CTAS 7
#endif
constexpr int T2_THREADS = 128;
constexpr int T2_WARPS = T2_THREADS / 32;
constexpr int T2_CAP = 256;                 // ring entries per warp (power of two)
constexpr int T2_ABUF = 98;                 // accept-list buffer (flushed at >= 64; + 1 pad entry, even size)
constexpr float T2_MARGIN = 2e-5f;          // relative safety margin of the group classification

struct __align__(16) APair {
    float xh[2], yh[2];     // centre, high parts
    float xl[2], yl[2];     // centre, low parts
    float g[2];             // mass
    unsigned int mask[2];   // lanes that reach the entry
};
struct __align__(16) MPair {
    float xh[2], yh[2], xl[2], yl[2];
    float g[2], lo[2];      // mass, lower edge of the opening threshold's guard band
    float hi[2];            // upper edge
    unsigned int mask[2];   // lanes that reach the node
};

struct __align__(16) T2Warp {
    uint2 q[T2_CAP];                        // FIFO ring of nodes to classify: {record slot, lanes (targets) that reach it}
    // accept list: entries 2k and 2k+1 share one APair, each component of the two side by side, so that one
    // LDS.128 fetches two packed-fp32 operands (three loads per pair of entries)
    APair ap[T2_ABUF / 2];
    unsigned int aslot[T2_ABUF];            //   record slot (self test / stats only)
    // mixed nodes of the current round, same layout (evaluated in pairs as well)
    MPair mp[16];
    unsigned int mslot[32];                 //   record slot
    unsigned int momask[32];                //   result: lanes that reached AND opened it
};

// ---- packed FP32 (sm_100: FADD2 / FMUL2 / FFMA2 do two fp32 operations in ONE issue slot) ----------------------
// The traversal is instruction-issue bound, not FMA-pipe bound (ncu: issue slots 75 % busy, FMA pipe 37 %), so the
// list evaluation works on TWO entries at a time with every add / multiply / fma packed across the pair.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
    return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ float lo2(f32x2_t v) { return __uint_as_float((unsigned int)v); }
__device__ __forceinline__ float hi2(f32x2_t v) { return __uint_as_float((unsigned int)(v >> 32)); }
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct LanePos2 {   // the lane's negated two-float position, each component duplicated into a register pair
    f32x2_t nphx, nphy, nplx, nply, eps2;
};

// two consecutive accept-list entries (m even) for one lane; the sums go to two partial accumulators per axis
template <bool STATS, bool SELF>
__device__ __forceinline__ void t2_accept_pair(const T2Warp& W, unsigned int m, unsigned int lanebit, unsigned int self,
                                               const LanePos2& P, f32x2_t& AX2, f32x2_t& AY2, unsigned int& nacc,
                                               double& fsum, float& fmaxq) {
    const ulonglong2* ap = reinterpret_cast<const ulonglong2*>(&W.ap[m >> 1]);
    const ulonglong2 vh = ap[0], vl = ap[1];
    const uint4 vg = *reinterpret_cast<const uint4*>(ap + 2);
    const f32x2_t xh = vh.x, yh = vh.y, xl = vl.x, yl = vl.y;
    const float2 g = make_float2(__uint_as_float(vg.x), __uint_as_float(vg.y));
    const uint2 mk = make_uint2(vg.z, vg.w);
    const bool r0 = (mk.x & lanebit) != 0u, r1 = (mk.y & lanebit) != 0u;
    // a lane that accepted an ancestor of an entry gets nothing from it
    const float g0 = r0 ? g.x : 0.f, g1 = r1 ? g.y : 0.f;
    bool self0 = false, self1 = false;
    if (SELF) {
        const uint2 as = *reinterpret_cast<const uint2*>(&W.aslot[m]);
        self0 = (as.x & 0x7FFFFFFFu) == self;
        self1 = (as.y & 0x7FFFFFFFu) == self;
    }
    const f32x2_t dx = add2(add2(xh, P.nphx), add2(xl, P.nplx));
    const f32x2_t dy = add2(add2(yh, P.nphy), add2(yl, P.nply));
    const f32x2_t d2 = fma2(dx, dx, fma2(dy, dy, P.eps2));
    const f32x2_t rinv = pack2(rsqrt_approx(lo2(d2)), rsqrt_approx(hi2(d2)));
    if (STATS) {   // SELF is on in every stats build
        const uint2 as = *reinterpret_cast<const uint2*>(&W.aslot[m]);
        const bool c0 = r0 && (as.x & 0x7FFFFFFFu) != self && !(as.x >> 31);
        const bool c1 = r1 && (as.y & 0x7FFFFFFFu) != self && !(as.y >> 31);
        nacc += (c0 ? 1u : 0u) + (c1 ? 1u : 0u);
        // the reference's force = G*M*m/distSq of the interaction, here in scaled units without the target's mass
        const float q0 = c0 ? g.x * lo2(rinv) * lo2(rinv) : 0.f, q1 = c1 ? g.y * hi2(rinv) * hi2(rinv) : 0.f;
        fsum += (double)q0 + (double)q1;
        fmaxq = fmaxf(fmaxq, fmaxf(q0, q1));
    }
    f32x2_t f = mul2(mul2(pack2(g0, g1), rinv), mul2(rinv, rinv));
    // without softening a body's own leaf has d2 = 0 (0 * inf): drop it explicitly (barnes_hut.cpp:272)
    if (SELF) f = pack2(self0 ? 0.f : lo2(f), self1 ? 0.f : hi2(f));
    AX2 = fma2(dx, f, AX2);
    AY2 = fma2(dy, f, AY2);
}

// ---- far field shared by the four warps of a CTA (template flag CTA) ----------------------------------------------
// The four warps of a CTA own 128 key-consecutive bodies and walk almost the same nodes: everything far from the whole
// block is classified four times. With CTA set, the block's bodies are taken together first: one breadth-first walk by
// all 128 threads (one node per thread) against the bounding box of the 128 bodies. A node every body of the block
// accepts goes to a shared accept list that each warp evaluates for its own lanes — without a lane mask, every body
// reaches it —, a node every body opens passes its children on, and only the rest (mixed for the block) seeds the four
// per-warp walks above, each of which then re-classifies against its own, smaller box. Per-body decisions are the
// same: both group tests are conservative.
constexpr int T3_QCAP = 256;                // block-level queue of record slots (power of two)
#ifndef T3_ABUF_N
#define T3_ABUF_N 130
#endif
#ifndef T3_FLUSH_N
#define T3_FLUSH_N 64
#endif
constexpr int T3_ABUF = T3_ABUF_N;          // shared accept list (one round adds at most T3_ABUF - 2 - fill entries)
constexpr unsigned int T3_FLUSH = T3_FLUSH_N; // ... evaluated by the four warps when it holds this many
constexpr int T3_SEEDS = 160;               // mixed nodes handed to the warps (more: the block goes to the depth-first kernel)
struct __align__(16) T3Cta {
    APair ap[T3_ABUF / 2];
    unsigned int aslot[T3_ABUF];
    unsigned int q[T3_QCAP];
    unsigned int seeds[T3_SEEDS];
    float box[T2_WARPS][4];
    uint4 wc[T2_WARPS];                     // per warp and round: entries for the accept list, seeds, children
    unsigned int head, tail, nA, nSeeds, ovf, block, nodes, pad;
};

// two consecutive entries of the block's shared accept list: every lane reaches them, no mask
template <bool STATS, bool SELF>
__device__ __forceinline__ void t3_accept_pair(const T3Cta& C, unsigned int m, unsigned int self, const LanePos2& P,
                                               f32x2_t& AX2, f32x2_t& AY2, unsigned int& nacc, double& fsum, float& fmaxq) {
    const ulonglong2* ap = reinterpret_cast<const ulonglong2*>(&C.ap[m >> 1]);
    const ulonglong2 vh = ap[0], vl = ap[1];
    const uint2 vg = *reinterpret_cast<const uint2*>(ap + 2);
    const f32x2_t xh = vh.x, yh = vh.y, xl = vl.x, yl = vl.y;
    const float2 g = make_float2(__uint_as_float(vg.x), __uint_as_float(vg.y));
    bool self0 = false, self1 = false;
    if (SELF) {
        const uint2 as = *reinterpret_cast<const uint2*>(&C.aslot[m]);
        self0 = (as.x & 0x7FFFFFFFu) == self;
        self1 = (as.y & 0x7FFFFFFFu) == self;
    }
    const f32x2_t dx = add2(add2(xh, P.nphx), add2(xl, P.nplx));
    const f32x2_t dy = add2(add2(yh, P.nphy), add2(yl, P.nply));
    const f32x2_t d2 = fma2(dx, dx, fma2(dy, dy, P.eps2));
    const f32x2_t rinv = pack2(rsqrt_approx(lo2(d2)), rsqrt_approx(hi2(d2)));
    if (STATS) {
        const uint2 as = *reinterpret_cast<const uint2*>(&C.aslot[m]);
        const bool c0 = as.x != LPE_NONE && (as.x & 0x7FFFFFFFu) != self && !(as.x >> 31);
        const bool c1 = as.y != LPE_NONE && (as.y & 0x7FFFFFFFu) != self && !(as.y >> 31);
        nacc += (c0 ? 1u : 0u) + (c1 ? 1u : 0u);
        const float q0 = c0 ? g.x * lo2(rinv) * lo2(rinv) : 0.f, q1 = c1 ? g.y * hi2(rinv) * hi2(rinv) : 0.f;
        fsum += (double)q0 + (double)q1;
        fmaxq = fmaxf(fmaxq, fmaxf(q0, q1));
    }
    f32x2_t f = mul2(mul2(pack2(g.x, g.y), rinv), mul2(rinv, rinv));
    if (SELF) f = pack2(self0 ? 0.f : lo2(f), self1 ? 0.f : hi2(f));
    AX2 = fma2(dx, f, AX2);
    AY2 = fma2(dy, f, AY2);
}

// MODE: what the kernel serves besides the resident single-GPU step (T2_RESIDENT) — a domain-decomposed rank (T2_DD: body
// count and tree size known only on the device, per-chunk cost recorded for the load balancer) or a host tick whose kick
// is deferred (T2_STAGED: the epilogue stores {x, y, dvx, dvy} at the body's creation index instead of kicking). A
// template parameter so that the resident kernel carries none of it: either costs it 1.5 % (registers, measured).
constexpr int T2_RESIDENT = 0, T2_DD = 1, T2_STAGED = 2;
template <bool STATS, bool SELF, int MODE, bool CTA>
__global__ void __launch_bounds__(T2_THREADS, T2_MIN_CTAS)
k_traverse2(const __grid_constant__ StepConst c, const __grid_constant__ TravArgs a, unsigned int* __restrict__ ovf_list) {
    constexpr bool DD = MODE == T2_DD;
    constexpr bool STAGED = MODE == T2_STAGED;
    extern __shared__ __align__(16) unsigned char t2_smem[];
    T2Warp& W = reinterpret_cast<T2Warp*>(t2_smem)[threadIdx.x >> 5];
    T3Cta& C = *reinterpret_cast<T3Cta*>(t2_smem + sizeof(T2Warp) * T2_WARPS);   // (only there when CTA)
    const int lane = threadIdx.x & 31;
    const unsigned int lt = (1u << lane) - 1u;
    const unsigned int lanebit = 1u << lane;
    const unsigned int n_nodes = DD ? a.s->dd_nroots : a.s->n_term + a.s->n_internal;
    const double massScale = 1.0 / mass_scale_inv(a.s->max_mass_bits);
    const float eps2f = c.eps2f;
    const float INF = __int_as_float(0x7f800000);
    const float FMAXV = 3.0e38f;
    constexpr unsigned int CHUNKS_PER_BLOCK = 2048u / 32u;  // LPE_SHARD_BLOCK / 32
    constexpr unsigned int QM = (unsigned int)T2_CAP - 1u;

    while (true) {
        unsigned int q = 0;
        if (CTA) {
            // the work unit is a block of four consecutive chunks, one per warp; every warp of the CTA stays in the loop
            // until the blocks are used up (the block-level phase needs all of them at its barriers)
            __syncthreads();   // the previous block's shared state is no longer in use
            if (threadIdx.x == 0) C.block = atomicAdd(&a.s->work_counter, 1u);
            __syncthreads();
            const unsigned int Q = C.block;
            if (Q * (unsigned int)T2_WARPS >= a.n_chunks_local) break;
            q = Q * (unsigned int)T2_WARPS + (threadIdx.x >> 5);
        } else {
            if (lane == 0) q = atomicAdd(&a.s->work_counter, 1u);
            q = __shfl_sync(0xFFFFFFFFu, q, 0);
            if (q >= a.n_chunks_local) break;
        }
        const unsigned int lblock = q / CHUNKS_PER_BLOCK, within = q % CHUNKS_PER_BLOCK;
        const unsigned int gblock = lblock * (unsigned int)c.shard_n + (unsigned int)c.shard_rank;
        const long long i = ((long long)gblock * CHUNKS_PER_BLOCK + within) * 32 + lane;
        // (a domain-decomposed rank knows its body count only on the device; the tail slots hold no bodies)
        bool valid = i < c.n && q < a.n_chunks_local;
        if (DD) {
            const unsigned int n_live = a.s->n_live;
            if (!CTA && q * 32u >= n_live) continue;
            valid = valid && (unsigned int)i < n_live;
        }

        unsigned int b = 0, self = LPE_NONE, cm = 0;
        double2 p = make_double2(0.0, 0.0);
        double bodyMass = 0.0;
        if (valid) {
            const Body sb = a.body[i];
            b = (unsigned int)i;
            cm = sb.comp;
            p = make_double2(sb.x, sb.y);
            if (STATS) bodyMass = sb.m;
            if (SELF) self = a.selfslot[i];
        }
        const bool target = valid && (cm & 1u) && (cm & 2u) && !(cm & 4u);   // barnes_hut.cpp:89
        const double pxs = p.x * c.invS, pys = p.y * c.invS;
        const float phx = (float)pxs, phy = (float)pys;
        const float nphx = -phx, nphy = -phy;
        const float nplx = -(float)(pxs - (double)phx), nply = -(float)(pys - (double)phy);
        LanePos2 LP;
        LP.nphx = pack2(nphx, nphx); LP.nphy = pack2(nphy, nphy);
        LP.nplx = pack2(nplx, nplx); LP.nply = pack2(nply, nply);
        LP.eps2 = pack2(eps2f, eps2f);
        unsigned int nacc = 0, nwarp = 0, cost = 0;   // cost: nodes the warp classified (load-balance weight)
        double fsum = 0.0;                            // STATS: DebugStats::updateForce, sum and max of G*M*m/distSq
        float fmaxq = 0.f;
        unsigned int kd[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // STATS: A-clean, A-dirty, O-dirty, M->all accept, M->all open, M->split, rounds, frontier nodes
        double AX = 0.0, AY = 0.0;
        bool overflow = c.test_overflow != 0;

        const unsigned int tmask = __ballot_sync(0xFFFFFFFFu, target);
        // ---- bounding box of the warp's targets (scaled units): fp32 edges rounded OUTWARD, so the box contains
        // every fp64 position and the group classification stays conservative ----
        float bx0 = target ? __double2float_rd(pxs) : INF, bx1 = target ? __double2float_ru(pxs) : -INF;
        float by0 = target ? __double2float_rd(pys) : INF, by1 = target ? __double2float_ru(pys) : -INF;
        if (CTA || tmask != 0u) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                bx0 = fminf(bx0, __shfl_xor_sync(0xFFFFFFFFu, bx0, o));
                bx1 = fmaxf(bx1, __shfl_xor_sync(0xFFFFFFFFu, bx1, o));
                by0 = fminf(by0, __shfl_xor_sync(0xFFFFFFFFu, by0, o));
                by1 = fmaxf(by1, __shfl_xor_sync(0xFFFFFFFFu, by1, o));
            }
        }
        unsigned int nSeeds = 0;
        if (CTA) {
            // ---------------- block-level phase: the far field of all 128 bodies, once ----------------
            const int warp = threadIdx.x >> 5;
            if (lane == 0) { C.box[warp][0] = bx0; C.box[warp][1] = bx1; C.box[warp][2] = by0; C.box[warp][3] = by1; }
            if (threadIdx.x == 0) {
                C.q[0] = 0u;   // the root
                C.head = 0u; C.tail = (n_nodes != 0u && !overflow) ? 1u : 0u;
                C.nA = 0u; C.nSeeds = 0u; C.ovf = 0u; C.nodes = 0u;
            }
            __syncthreads();
            float cx0 = C.box[0][0], cx1 = C.box[0][1], cy0 = C.box[0][2], cy1 = C.box[0][3];
#pragma unroll
            for (int ww = 1; ww < T2_WARPS; ++ww) {
                cx0 = fminf(cx0, C.box[ww][0]); cx1 = fmaxf(cx1, C.box[ww][1]);
                cy0 = fminf(cy0, C.box[ww][2]); cy1 = fmaxf(cy1, C.box[ww][3]);
            }
            const bool anyTarget = cx0 <= cx1;
            constexpr unsigned int CQM = (unsigned int)T3_QCAP - 1u;
            for (;;) {
                const unsigned int head = C.head, tail = C.tail, nA0 = C.nA, nS0 = C.nSeeds;
                if (head == tail || !anyTarget || C.ovf) break;
                // (a round never adds more entries than the accept list has room for)
                const unsigned int cnt = min(min((unsigned int)T2_THREADS, tail - head), (unsigned int)T3_ABUF - 2u - nA0);
                const bool has = threadIdx.x < cnt;
                unsigned int slot = 0;
                TravRec R;
                R.c = make_float4(0.f, 0.f, 0.f, 0.f); R.gm = 0.f; R.open_t = -1.f; R.node = 0; R.cblock = 0;
                bool toA = false, toS = false;
                unsigned int nch = 0, mA = 0, mS = 0, b0 = 0, b1 = 0, b2 = 0;
                float t = -1.f;
                const unsigned int le = lt | lanebit;
                // (measured: letting the warps without nodes skip the classification in the first, sparse rounds gains nothing)
                {
                    if (has) {
                        slot = C.q[(head + threadIdx.x) & CQM];
                        const uint4* src = reinterpret_cast<const uint4*>(a.rec + lpe_idx(slot, c.recSlots, 10, a.s));
                        const uint4 v0 = __ldg(src), v1 = __ldg(src + 1);
                        R.c = make_float4(__uint_as_float(v0.x), __uint_as_float(v0.y), __uint_as_float(v0.z), __uint_as_float(v0.w));
                        R.gm = __uint_as_float(v1.x); R.open_t = __uint_as_float(v1.y); R.node = v1.z; R.cblock = v1.w;
                    }
                    const float ax0 = (R.c.x - cx0) + R.c.z, ax1 = (R.c.x - cx1) + R.c.z;
                    const float ay0 = (R.c.y - cy0) + R.c.w, ay1 = (R.c.y - cy1) + R.c.w;
                    const float dxmin = fmaxf(fmaxf(-ax0, ax1), 0.f), dxmax = fmaxf(fabsf(ax0), fabsf(ax1));
                    const float dymin = fmaxf(fmaxf(-ay0, ay1), 0.f), dymax = fmaxf(fabsf(ay0), fabsf(ay1));
                    const float d2min = fmaf(dxmin, dxmin, fmaf(dymin, dymin, eps2f));
                    const float d2max = fmaf(dxmax, dxmax, fmaf(dymax, dymax, eps2f));
                    t = R.open_t;
                    const float tlo = t * (1.0f - OPEN_BAND), thi = t * (1.0f + OPEN_BAND);
                    const bool allAcc = (t < 0.f) || (d2min * (1.0f - T2_MARGIN) >= thi);
                    const bool allOpen = (t >= 0.f) && (d2max * (1.0f + T2_MARGIN) <= tlo);
                    toA = has && allAcc;
                    toS = has && !allAcc && !allOpen;      // mixed for the block: the warps decide
                    nch = (has && allOpen) ? (R.cblock & 3u) + 1u : 0u;
                    mA = __ballot_sync(0xFFFFFFFFu, toA); mS = __ballot_sync(0xFFFFFFFFu, toS);
                    b0 = __ballot_sync(0xFFFFFFFFu, (nch & 1u) != 0u);
                    b1 = __ballot_sync(0xFFFFFFFFu, (nch & 2u) != 0u);
                    b2 = __ballot_sync(0xFFFFFFFFu, (nch & 4u) != 0u);
                }
                if (lane == 0) C.wc[warp] = make_uint4(__popc(mA), __popc(mS), __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2), 0u);
                __syncthreads();
                unsigned int offA = nA0, offS = nS0, offC = 0, totA = 0, totS = 0, totC = 0;
#pragma unroll
                for (int ww = 0; ww < T2_WARPS; ++ww) {
                    const uint4 wv = C.wc[ww];
                    if (ww < warp) { offA += wv.x; offS += wv.y; offC += wv.z; }
                    totA += wv.x; totS += wv.y; totC += wv.z;
                }
                const bool fits = (tail - head - cnt) + totC <= (unsigned int)T3_QCAP && nS0 + totS <= (unsigned int)T3_SEEDS;
                if (fits) {
                    if (toA) {
                        const unsigned int pos = lpe_idx(offA + __popc(mA & lt), (unsigned int)T3_ABUF - 1u, 11, a.s);
                        APair& E = C.ap[pos >> 1];
                        const unsigned int h = pos & 1u;
                        E.xh[h] = R.c.x; E.yh[h] = R.c.y; E.xl[h] = R.c.z; E.yl[h] = R.c.w;
                        E.g[h] = R.gm;
                        C.aslot[pos] = slot | ((t == -2.0f) ? 0x80000000u : 0u);
                    }
                    if (toS) C.seeds[offS + __popc(mS & lt)] = slot;
                    if (nch) {
                        const unsigned int inc = __popc(b0 & le) + 2u * __popc(b1 & le) + 4u * __popc(b2 & le);
                        const unsigned int at = tail + offC + inc - nch;
                        const unsigned int cslot = (R.cblock >> 2) * 4u;
                        C.q[at & CQM] = cslot;
                        if (nch > 1u) C.q[(at + 1u) & CQM] = cslot + 1u;
                        if (nch > 2u) C.q[(at + 2u) & CQM] = cslot + 2u;
                        if (nch > 3u) C.q[(at + 3u) & CQM] = cslot + 3u;
                    }
                }
                const unsigned int nA1 = nA0 + totA;
                const bool lastRound = head + cnt == tail + totC;
                const bool flush = fits && (nA1 >= T3_FLUSH || lastRound) && nA1 != 0u;
                if (threadIdx.x == 0) {
                    if (!fits) C.ovf = 1u;   // block-level frontier or seed list too long: the depth-first kernel redoes the block
                    else {
                        C.head = head + cnt; C.tail = tail + totC; C.nSeeds = nS0 + totS; C.nodes += cnt;
                        C.nA = flush ? 0u : nA1;
                        if (flush && (nA1 & 1u)) {   // pad to an even count with an entry of zero mass
                            APair& E = C.ap[nA1 >> 1];
                            E.xh[1] = 4.f; E.yh[1] = 4.f; E.xl[1] = 0.f; E.yl[1] = 0.f; E.g[1] = 0.f;
                            C.aslot[nA1] = LPE_NONE;
                        }
                    }
                }
                __syncthreads();
                if (flush) {
                    if (tmask != 0u) {
                        // fp32 partial sums go to fp64 every 64 entries, as in the per-warp walk (measured: two
                        // straight-line loops instead of this nest are 3.5 % slower — register allocation)
                        for (unsigned int m0 = 0; m0 < nA1; m0 += 64u) {
                            f32x2_t AX2 = pack2(0.f, 0.f), AY2 = AX2;
                            const unsigned int m1 = min(nA1, m0 + 64u);
#pragma unroll 4
                            for (unsigned int m = m0; m < m1; m += 2)
                                t3_accept_pair<STATS, SELF>(C, m, self, LP, AX2, AY2, nacc, fsum, fmaxq);
                            AX += (double)(lo2(AX2) + hi2(AX2));
                            AY += (double)(lo2(AY2) + hi2(AY2));
                        }
                        if (STATS) { nwarp += nA1; if (warp == 0) kd[0] += nA1; }
                    }
                    __syncthreads();   // the list is consumed before the next round writes it
                }
            }
            nSeeds = C.nSeeds;
            if (STATS && warp == 0) { kd[6] += 1; kd[7] += C.nodes; }
            if (DD) cost = C.nodes / (unsigned int)T2_WARPS;
            if (C.ovf) overflow = true;
        }
        if (tmask != 0u && n_nodes != 0u && !overflow) {
            unsigned int head = 0, tail = 1, nA = 0;
            if (CTA) {
                // the per-warp walk starts from the nodes the block could not decide, each reached by every target
                tail = nSeeds;
                for (unsigned int k = lane; k < nSeeds; k += 32u) W.q[k] = make_uint2(C.seeds[k], tmask);
            } else {
                if (lane == 0) W.q[0] = make_uint2(0u, tmask);   // the root, reached by every target
            }
            __syncwarp();
            while (head != tail) {
                // ---------------- phase 1: one node per lane ----------------
                const unsigned int cnt = min(32u, tail - head);
                const bool has = (unsigned int)lane < cnt;
                unsigned int slot = 0, mask = tmask;
                TravRec R;
                R.c = make_float4(0.f, 0.f, 0.f, 0.f); R.gm = 0.f; R.open_t = -1.f; R.node = 0; R.cblock = 0;
                if (has) {
                    const uint2 e = W.q[(head + lane) & QM];
                    slot = e.x;
                    mask = e.y;
                    const uint4* src = reinterpret_cast<const uint4*>(a.rec + lpe_idx(slot, c.recSlots, 10, a.s));
                    const uint4 v0 = __ldg(src), v1 = __ldg(src + 1);
                    R.c = make_float4(__uint_as_float(v0.x), __uint_as_float(v0.y), __uint_as_float(v0.z), __uint_as_float(v0.w));
                    R.gm = __uint_as_float(v1.x); R.open_t = __uint_as_float(v1.y); R.node = v1.z; R.cblock = v1.w;
                }
                head += cnt;
                // distance bounds from the node centre to the targets' box
                const float ax0 = (R.c.x - bx0) + R.c.z, ax1 = (R.c.x - bx1) + R.c.z;
                const float ay0 = (R.c.y - by0) + R.c.w, ay1 = (R.c.y - by1) + R.c.w;
                const float dxmin = fmaxf(fmaxf(-ax0, ax1), 0.f), dxmax = fmaxf(fabsf(ax0), fabsf(ax1));
                const float dymin = fmaxf(fmaxf(-ay0, ay1), 0.f), dymax = fmaxf(fabsf(ay0), fabsf(ay1));
                const float d2min = fmaf(dxmin, dxmin, fmaf(dymin, dymin, eps2f));
                const float d2max = fmaf(dxmax, dxmax, fmaf(dymax, dymax, eps2f));
                const float t = R.open_t;
                const float tlo = t * (1.0f - OPEN_BAND), thi = t * (1.0f + OPEN_BAND);
                const bool allAcc = (t < 0.f) || (d2min * (1.0f - T2_MARGIN) >= thi);
                const bool allOpen = (t >= 0.f) && (d2max * (1.0f + T2_MARGIN) <= tlo);
                const bool dirty = mask != tmask;                    // some target accepted an ancestor
                const bool toA = has && allAcc;                      // clean or dirty: the entry carries the lane mask
                const bool toM = has && !allAcc && !allOpen;         // all-open nodes just pass their mask on
                const bool expand = has && !allAcc;

                const unsigned int maskA = __ballot_sync(0xFFFFFFFFu, toA);
                const unsigned int maskM = __ballot_sync(0xFFFFFFFFu, toM);
                const unsigned int posM = __popc(maskM & lt);
                const unsigned int cntM = __popc(maskM);
                if (toA) {
                    const unsigned int pos = lpe_idx(nA + __popc(maskA & lt), (unsigned int)T2_ABUF - 1u, 11, a.s);
                    APair& E = W.ap[pos >> 1];
                    const unsigned int h = pos & 1u;
                    E.xh[h] = R.c.x; E.yh[h] = R.c.y; E.xl[h] = R.c.z; E.yl[h] = R.c.w;
                    E.g[h] = R.gm;
                    E.mask[h] = mask;
                    if (SELF) W.aslot[pos] = slot | ((t == -2.0f) ? 0x80000000u : 0u);
                }
                nA += __popc(maskA);
                if (STATS) {
                    kd[0] += __popc(__ballot_sync(0xFFFFFFFFu, toA && !dirty));
                    kd[1] += __popc(__ballot_sync(0xFFFFFFFFu, toA && dirty));
                    kd[2] += __popc(__ballot_sync(0xFFFFFFFFu, expand && allOpen && dirty));
                    kd[6] += 1; kd[7] += cnt;
                }
                if (toM) {
                    MPair& E = W.mp[posM >> 1];
                    const unsigned int h = posM & 1u;
                    E.xh[h] = R.c.x; E.yh[h] = R.c.y; E.xl[h] = R.c.z; E.yl[h] = R.c.w;
                    E.g[h] = R.gm; E.lo[h] = tlo; E.hi[h] = thi;
                    E.mask[h] = mask;
                    W.mslot[posM] = slot | ((t == -2.0f) ? 0x80000000u : 0u);
                }
                if ((cntM & 1u) && lane == 0) {   // pad to an even count with an entry nobody reaches
                    MPair& E = W.mp[cntM >> 1];
                    E.xh[1] = 4.f; E.yh[1] = 4.f; E.xl[1] = 0.f; E.yl[1] = 0.f;
                    E.g[1] = 0.f; E.lo[1] = -1.f; E.hi[1] = -1.f;
                    E.mask[1] = 0u; W.mslot[cntM] = LPE_NONE;
                }
                __syncwarp();
                f32x2_t AX2 = pack2(0.f, 0.f), AY2 = AX2;   // this round's partial sums, two lanes of fp32 per axis

                // ---------------- phase 2a: the mixed nodes of this round, one body per lane ----------------
                // Run twice at most: the fast pass takes the fp32 decision everywhere and only REMEMBERS whether some
                // lane came inside a guard band; if one did (about 1 round in 500) the round's mixed nodes are redone
                // with the reference's fp64 test deciding those cases. No vote or branch per node in the fast pass.
                auto mixed = [&](auto exactTag) -> bool {
                constexpr bool EXACT = decltype(exactTag)::value;
                bool bandAny = false;
                AX2 = pack2(0.f, 0.f); AY2 = AX2;
                for (unsigned int m = 0; m < cntM; m += 2) {
                    const ulonglong2* mp = reinterpret_cast<const ulonglong2*>(&W.mp[m >> 1]);
                    const ulonglong2 vh = mp[0], vl = mp[1];
                    const float4 vg = *reinterpret_cast<const float4*>(mp + 2);
                    const uint4 vm = *reinterpret_cast<const uint4*>(mp + 3);
                    const f32x2_t xh = vh.x, yh = vh.y, xl = vl.x, yl = vl.y;
                    const float2 g = make_float2(vg.x, vg.y), tl = make_float2(vg.z, vg.w);
                    const float2 th = make_float2(__uint_as_float(vm.x), __uint_as_float(vm.y));
                    const uint2 mk = make_uint2(vm.z, vm.w);
                    const f32x2_t dx = add2(add2(xh, LP.nphx), add2(xl, LP.nplx));
                    const f32x2_t dy = add2(add2(yh, LP.nphy), add2(yl, LP.nply));
                    const f32x2_t d2p = fma2(dx, dx, fma2(dy, dy, LP.eps2));
                    const bool reached0 = (mk.x & lanebit) != 0u, reached1 = (mk.y & lanebit) != 0u;
                    // a lane that accepted an ancestor: never opens, contributes 0
                    const float d20 = reached0 ? lo2(d2p) : INF, d21 = reached1 ? hi2(d2p) : INF;
                    float lo0 = tl.x, lo1 = tl.y;
                    const bool band0 = d20 > lo0 && d20 < th.x, band1 = d21 > lo1 && d21 < th.y;
                    if (EXACT) {
                        if (band0)   // guard band: the reference's fp64 test decides
                            lo0 = exact_open_slot(&a, &c, W.mslot[m] & 0x7FFFFFFFu, pxs, pys) ? FMAXV : -1.f;
                        if (band1)
                            lo1 = exact_open_slot(&a, &c, W.mslot[m + 1] & 0x7FFFFFFFu, pxs, pys) ? FMAXV : -1.f;
                    } else {
                        bandAny = bandAny || band0 || band1;
                    }
                    const bool open0 = d20 <= lo0, open1 = d21 <= lo1;
                    const unsigned int om0 = __ballot_sync(0xFFFFFFFFu, open0);
                    const unsigned int om1 = __ballot_sync(0xFFFFFFFFu, open1);
                    if (lane == 0) *reinterpret_cast<uint2*>(&W.momask[m]) = make_uint2(om0, om1);
                    if (STATS) {
                        if (m < cntM) { if (om0 == 0u) kd[3]++; else if (om0 == mk.x) kd[4]++; else kd[5]++; }
                        if (m + 1 < cntM) { if (om1 == 0u) kd[3]++; else if (om1 == mk.y) kd[4]++; else kd[5]++; }
                    }
                    const f32x2_t rinv = pack2(rsqrt_approx(open0 ? INF : d20), rsqrt_approx(open1 ? INF : d21));
                    f32x2_t f = mul2(mul2(pack2(g.x, g.y), rinv), mul2(rinv, rinv));
                    if (SELF) {
                        const uint2 ms = *reinterpret_cast<const uint2*>(&W.mslot[m]);
                        const bool self0 = (ms.x & 0x7FFFFFFFu) == self, self1 = (ms.y & 0x7FFFFFFFu) == self;
                        f = pack2(self0 ? 0.f : lo2(f), self1 ? 0.f : hi2(f));
                        if (STATS) {
                            const bool c0 = reached0 && !open0 && !self0 && !(ms.x >> 31);
                            const bool c1 = reached1 && !open1 && !self1 && !(ms.y >> 31);
                            nacc += (c0 ? 1u : 0u) + (c1 ? 1u : 0u);
                            const float q0 = c0 ? g.x * lo2(rinv) * lo2(rinv) : 0.f, q1 = c1 ? g.y * hi2(rinv) * hi2(rinv) : 0.f;
                            fsum += (double)q0 + (double)q1;
                            fmaxq = fmaxf(fmaxq, fmaxf(q0, q1));
                        }
                    }
                    AX2 = fma2(dx, f, AX2);
                    AY2 = fma2(dy, f, AY2);
                }
                return bandAny;
                };
                bool redo = true;                       // counting runs take the exact pass only (counters tick once)
                if (!STATS) redo = mixed(std::false_type{});
                if (STATS || __any_sync(0xFFFFFFFFu, redo)) mixed(std::true_type{});
                if (STATS) nwarp += cntM;
                __syncwarp();

                // ---------------- children of opened nodes join the queue with the mask of the lanes that opened ----------------
                const unsigned int cmask = toM ? W.momask[posM] : mask;
                const unsigned int nch = (expand && cmask != 0u) ? (R.cblock & 3u) + 1u : 0u;
                // inclusive prefix of nch (0..4) over the lanes from three independent ballots, one per bit of nch:
                // same instruction count as a shuffle scan, a fifth of its dependent latency
                const unsigned int b0 = __ballot_sync(0xFFFFFFFFu, (nch & 1u) != 0u);
                const unsigned int b1 = __ballot_sync(0xFFFFFFFFu, (nch & 2u) != 0u);
                const unsigned int b2 = __ballot_sync(0xFFFFFFFFu, (nch & 4u) != 0u);
                const unsigned int le = lt | lanebit;
                const unsigned int inc = __popc(b0 & le) + 2u * __popc(b1 & le) + 4u * __popc(b2 & le);
                const unsigned int total = __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2);
                if ((tail - head) + total > (unsigned int)T2_CAP) { overflow = true; break; }
                if (nch) {
                    const unsigned int at = tail + inc - nch;
                    const unsigned int cslot = (R.cblock >> 2) * 4u;
                    W.q[at & QM] = make_uint2(cslot, cmask);
                    if (nch > 1u) W.q[(at + 1u) & QM] = make_uint2(cslot + 1u, cmask);
                    if (nch > 2u) W.q[(at + 2u) & QM] = make_uint2(cslot + 2u, cmask);
                    if (nch > 3u) W.q[(at + 3u) & QM] = make_uint2(cslot + 3u, cmask);
                }
                tail += total;

                // ---------------- phase 2b: the accept list, flushed when it is long enough ----------------
                if (nA >= 64u || head == tail) {
                    if ((nA & 1u) && lane == 0) {   // pad to an even count with an entry nobody reaches
                        APair& E = W.ap[nA >> 1];
                        E.xh[1] = 4.f; E.yh[1] = 4.f; E.xl[1] = 0.f; E.yl[1] = 0.f;   // outside the universe
                        E.g[1] = 0.f; E.mask[1] = 0u; W.aslot[nA] = LPE_NONE;
                    }
                    __syncwarp();
#pragma unroll 4
                    for (unsigned int m = 0; m < nA; m += 2)
                        t2_accept_pair<STATS, SELF>(W, m, lanebit, self, LP, AX2, AY2, nacc, fsum, fmaxq);
                    if (STATS) nwarp += nA;
                    nA = 0;
                }
                // fp32 partial sums go to fp64 every round
                AX += (double)(lo2(AX2) + hi2(AX2));
                AY += (double)(lo2(AY2) + hi2(AY2));
                __syncwarp();
            }
            if (DD) cost += tail;   // every node that entered the ring was classified once
        }

        if (overflow) {
            // frontier too wide for the shared-memory queue: the depth-first kernel redoes this chunk
            if (lane == 0 && q < a.n_chunks_local) ovf_list[atomicAdd(&a.s->ovf_count, 1u)] = q;
            continue;
        }

        if (DD && lane == 0 && q < a.n_chunks_local) a.chunk_cost[q] = cost;
        double2 v = make_double2(0.0, 0.0);
        if (valid && !STAGED) v = a.vel[b];   // (deferred kick: v is the velocity CHANGE, 0 + x is exact)
        const double accScale = c.G * massScale * c.invS * c.invS;   // a = G*sum M d/r^3; scaled units M/Ms, d/S
        if (target) {
            v.x = kick_step(v.x, AX * accScale, c.dtK);   // barnes_hut.cpp:284-286
            v.y = kick_step(v.y, AY * accScale, c.dtK);
        }
        if (STAGED) {   // host tick: k_finish_tick kicks and drifts once the velocities have arrived
            if (valid) a.stage_out[a.orig ? a.orig[b] : b] = make_double4(p.x, p.y, v.x, v.y);
        } else if (valid) {
            const bool mover = (cm & 2u) && !(cm & 4u) && !(cm & 8u);   // movement.cpp:20-29
            if (c.do_drift && mover) {
                p.x = drift_step(p.x, v.x, c.dtD);                                      // movement.cpp:32-33
                p.y = drift_step(p.y, v.y, c.dtD);
            }
            if (c.shard_n > 1) {
                const unsigned long long slotx = (unsigned long long)lblock * 2048ull + within * 32ull + lane;
                // The exchange is fused into the kick: the new state goes straight into every rank's receive buffer
                // (stores to NVLink peer memory overlap the rest of the traversal), no collective afterwards.
                const double4 out = make_double4(p.x, p.y, v.x, v.y);
                a.xchg_send[slotx] = out;
                // peers in rotating order, starting after this rank: at any moment the ranks address different peers
                for (int k = 1; k <= a.npeer; ++k) {
                    int r = c.shard_rank + k;
                    if (r >= a.npeer) r -= a.npeer;
                    double2* dst = reinterpret_cast<double2*>(a.peer[r] + slotx);
                    __stcs(dst, make_double2(out.x, out.y));       // streaming stores: nothing here is read again locally
                    __stcs(dst + 1, make_double2(out.z, out.w));
                }
            } else {
                if (target) a.vel[b] = v;
                if (c.do_drift && mover) *reinterpret_cast<double2*>(&a.body[b].x) = p;
            }
        }
        if (STATS && valid) {
            a.cntAcc[b] = target ? nacc : 0u;
            a.cntVis[b] = 0u;
        }
        if (STATS) {
            unsigned int tot = target ? nacc : 0u;
            // force = (G * Ms / S^2) * m_target * (gm / d2) in real units
            const double fscale = c.G * massScale * c.invS * c.invS * bodyMass;
            double fs = target ? fsum * fscale : 0.0, fm = target ? (double)fmaxq * fscale : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                tot += __shfl_xor_sync(0xFFFFFFFFu, tot, o);
                fs += __shfl_xor_sync(0xFFFFFFFFu, fs, o);
                fm = fmax(fm, __shfl_xor_sync(0xFFFFFFFFu, fm, o));
            }
            if (lane == 0) {
                atomicAdd(&a.s->force_sum, fs);
                atomicMax(&a.s->force_max_bits, (unsigned long long)__double_as_longlong(fm));
                atomicAdd(&a.s->interactions, (unsigned long long)tot);
                atomicAdd(&a.s->warp_visits, (unsigned long long)nwarp);
                for (int z = 0; z < 8; ++z) atomicAdd(&a.s->t2[z], (unsigned long long)kd[z]);
            }
        }
    }
}

}  // namespace lpe
