// bh_traverse.cuh — theta-criterion force traversal fused with the kick (and optional drift).
//
// Replaces BarnesHutSystem::calculateForce (reference barnes_hut.cpp:240-294) and, when do_drift is set,
// MovementSystem::update (movement.cpp:13-39).
//
// One warp walks the tree for 32 key-consecutive targets, depth first in the reference's child order
// (nw, ne, sw, se = Morton digit order), so nodes are met in pre-order. Each lane keeps its OWN accept/open
// decision, as the reference does per body: a lane that accepted a node at depth d ignores everything the walk meets
// below depth d until it is back at depth <= d; the warp descends while any lane still opens. Nodes are addressed by
// record slot only (no pre-order index), so the walk also runs over the record array of a domain-decomposed rank,
// whose top blocks and imported blocks have no place in the rank's own pre-order.
//
// Memory: the children of a cell sit side by side in one 128-byte CHILD BLOCK (4 x TravRec). Opening a cell is one
// coalesced 128-byte load by 8 lanes into the warp's frame stack in shared memory; every visit then reads its
// record with two broadcast LDS.128. All children of an opened cell are always visited, so nothing fetched is
// wasted (the pre-order array this replaces needed a fresh L2 round trip for 39 % of its visits).
//
// FAST precision: state stays fp64. Node centres and lane positions enter the inner loop as two-float (hi + lo)
// pairs, so d = (c_hi - p_hi) + (c_lo - p_lo) is the fp64 difference rounded once to fp32 (relative error 2^-24
// of |d|, independent of where in the universe the pair sits) without any FP64 or conversion instruction; d^2,
// rsqrt and the accumulation are fp32, flushed to fp64 whenever a child block is finished. The theta test is
// done in fp32 against two thresholds bracketing s^2/theta^2; between them the reference's own fp64 expression
// (barnes_hut.cpp:261-269) decides, so accept/open decisions are the reference's, not an approximation.
// STRICT precision: every interaction in fp64, in the reference's expression order.
#pragma once
#include "bh_common.cuh"

namespace lpe {

constexpr int TRAV_THREADS = 256;
constexpr int TRAV_WARPS = TRAV_THREADS / 32;
constexpr int TRAV_FRAMES = LPE_MAX_DEPTH + 2;   // a chain of branching cells is at most D+1 long

struct TravArgs {
    const TravRec* rec;        // child blocks
    const Agg* agg;            // [preorder] exact sums (+ level): fp64 centre / mass for the rare exact test and STRICT mode
    const unsigned int* selfslot;   // [sorted body] record slot of its own single-body leaf
    const unsigned int* chunk_list; // depth-first kernel only: when set, process these chunks (two-phase overflow)
    Body* body;                     // state in key order (positions are updated in place by the drift)
    double2* vel;
    double4* xchg_send;      // sharded mode: packed (x,y,vx,vy) of the own slice
    double4* peer[LPE_MAX_P2P];   // direct exchange: this rank's slice inside every rank's receive buffer (NVLink peer
    int npeer;                    //   memory); 0 = the caller runs a collective on xchg_send instead
    unsigned int* cntAcc;    // STATS only, creation order
    unsigned int* cntVis;
    Scal* s;
    unsigned int n_chunks_local;  // 32-body chunks this rank owns
    // domain-decomposed runs: record slots outside [localLo, localHi) belong to the top of the tree or were imported
    // from another rank; their exact centre / level come from xrec[slot] = {cx, cy, M, level} instead of agg / meta
    const double4* xrec;
    unsigned int localLo, localHi;
    unsigned int* chunk_cost;     // [chunk] list entries evaluated for the chunk (load-balance weight); may be null
    // host tick (FAST precision): the velocities are still on their way over PCIe while the tree is walked, so the kernel
    // kicks nothing: it stores {x, y, dvx, dvy} — the body's position and its velocity CHANGE — as one 32-byte record at
    // the body's creation index (orig, null = identity), and k_finish_tick does kick + drift in creation order with
    // streaming accesses once the velocities have arrived. Null = kick and drift in the kernel's own epilogue.
    double4* stage_out;
    const unsigned int* orig;
};

// The reference's test, barnes_hut.cpp:261-269, on exactly scaled operands (power-of-two scaling commutes
// with IEEE rounding): returns true when the node must be opened.
__device__ __noinline__ bool exact_open(const Agg* __restrict__ agg, unsigned int j,
                                        int quirk, double invS, double pxs, double pys, double eps2s, double Us,
                                        double theta2) {
    const Agg a = agg[j];
    const int level = agg_level(a);
    double M, cx, cy;
    node_centre(a, level, quirk, M, cx, cy);
    const double dxs = cx * invS - pxs, dys = cy * invS - pys;
    const double distSq = __dadd_rn(__dadd_rn(__dmul_rn(dxs, dxs), __dmul_rn(dys, dys)), eps2s);
    const double size = ldexp(Us, -level);
    const double sizeSq = __dmul_rn(size, size);
    return !(__ddiv_rn(sizeSq, distSq) < theta2);
}

// The same test for a record slot, wherever the node came from. Cold path (about one visit in 1e5): the kernel's
// parameter blocks are passed by address (they are __grid_constant__), so the call costs the hot loop no registers.
__device__ __noinline__ bool exact_open_slot(const TravArgs* __restrict__ a, const StepConst* __restrict__ c, unsigned int slot,
                                             double pxs, double pys) {
    const double Us = c->U * c->invS;
    if (slot >= a->localLo && slot < a->localHi)
        return exact_open(a->agg, a->rec[slot].node, c->quirk, c->invS, pxs, pys, c->eps2s, Us, c->theta2);
    const double4 x = a->xrec[slot];
    const double dxs = x.x * c->invS - pxs, dys = x.y * c->invS - pys;
    const double distSq = __dadd_rn(__dadd_rn(__dmul_rn(dxs, dxs), __dmul_rn(dys, dys)), c->eps2s);
    const double size = ldexp(Us, -(int)x.w);
    const double sizeSq = __dmul_rn(size, size);
    return !(__ddiv_rn(sizeSq, distSq) < c->theta2);
}

// 8 lanes copy one 128-byte child block into a frame of the warp's stack
__device__ __forceinline__ void load_block(const TravRec* __restrict__ rec, unsigned int block, TravRec* frame, int lane,
                                           const StepConst& c, const Scal* s) {
    __syncwarp();
    block = lpe_idx(block, c.recSlots >> 2, 12, s);
    if (lane < 8) {
        const uint4 v = reinterpret_cast<const uint4*>(rec + 4 * (size_t)block)[lane];
        reinterpret_cast<uint4*>(frame)[lane] = v;
    }
    __syncwarp();
}

template <int PREC, bool STATS>
__global__ void __launch_bounds__(TRAV_THREADS, 4) k_traverse(const __grid_constant__ StepConst c, const __grid_constant__ TravArgs a) {
    __shared__ TravRec sFrames[TRAV_WARPS][TRAV_FRAMES][4];
    __shared__ unsigned int sBlock[TRAV_WARPS][TRAV_FRAMES];   // child block held by each frame (record slot = 4 * block + k)
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    TravRec* const frames = &sFrames[warp][0][0];
    unsigned int* const fblock = &sBlock[warp][0];
    const unsigned int n_nodes = c.dd ? a.s->dd_nroots : a.s->n_term + a.s->n_internal;
    const long long n_bodies = c.dd ? (long long)a.s->n_live : (long long)c.n;
    const double massScale = 1.0 / mass_scale_inv(a.s->max_mass_bits);
    const float eps2f = c.eps2f;
    const double Us = c.U * c.invS;
    const float INF = __int_as_float(0x7f800000);
    constexpr unsigned int CHUNKS_PER_BLOCK = 2048u / 32u;  // LPE_SHARD_BLOCK / 32
    constexpr int ACTIVE = 1 << 20;   // accDepth of a lane that has accepted none of the current node's ancestors

    while (true) {
        unsigned int q = 0;
        if (a.chunk_list) {
            if (lane == 0) q = atomicAdd(&a.s->work_counter2, 1u);
            q = __shfl_sync(0xFFFFFFFFu, q, 0);
            if (q >= a.s->ovf_count) break;
            q = a.chunk_list[q];
        } else {
            if (lane == 0) q = atomicAdd(&a.s->work_counter, 1u);
            q = __shfl_sync(0xFFFFFFFFu, q, 0);
            if (q >= a.n_chunks_local) break;
        }
        // block-cyclic ownership of sorted positions (identity when shard_n == 1)
        const unsigned int lblock = q / CHUNKS_PER_BLOCK, within = q % CHUNKS_PER_BLOCK;
        const unsigned int gblock = lblock * (unsigned int)c.shard_n + (unsigned int)c.shard_rank;
        const long long i = ((long long)gblock * CHUNKS_PER_BLOCK + within) * 32 + lane;
        const bool valid = i < n_bodies;
        if (c.dd && (long long)q * 32 >= n_bodies) continue;   // tail slots of a domain-decomposed rank hold no bodies

        unsigned int b = 0, self = LPE_NONE, cm = 0;
        double2 p = make_double2(0.0, 0.0);
        double bodyMass = 1.0;
        if (valid) {
            const Body sb = a.body[i];
            b = (unsigned int)i;
            cm = sb.comp;
            p = make_double2(sb.x, sb.y);
            bodyMass = sb.m;
            if (c.need_self) self = a.selfslot[i];
        }
        // bodyView of update(): Position + Velocity + Mass, not Boundary (barnes_hut.cpp:89)
        const bool target = valid && (cm & 1u) && (cm & 2u) && !(cm & 4u);
        const double pxs = p.x * c.invS, pys = p.y * c.invS;
        // Per-lane decisions without a per-lane stack: a lane that ACCEPTED a node at depth d ignores every node met
        // at a greater depth until the walk is back at depth <= d (the node's next sibling or an ancestor's).
        int accDepth = target ? ACTIVE : -1;
        unsigned int nacc = 0, nvis = 0, nwarp = 0;
        double fsum = 0.0, fmaxd = 0.0;   // STATS: DebugStats::updateForce (barnes_hut.cpp:278), real units
        double2 v = make_double2(0.0, 0.0);
        if (valid && !(PREC == 0 && a.stage_out)) v = a.vel[b];   // (deferred kick: v stays the velocity CHANGE, 0 + x is exact)

        // ---- depth-first walk over child blocks; d, k are warp-uniform ----
        int d = 0, k = 0;
        unsigned long long kstack = 0;   // 2 bits per depth: slot being processed there
        unsigned long long cstack = 0;   // 2 bits per depth: (number of children in that frame) - 1
        const bool alive = n_nodes != 0;
        if (alive) {
            load_block(a.rec, 0u, frames, lane, c, a.s);
            if (lane == 0) fblock[0] = 0u;
            __syncwarp();
        }

        if constexpr (PREC == 0) {
            // two-float lane position, negated once: d = (c_hi - p_hi) + (c_lo - p_lo)
            const float phx = (float)pxs, phy = (float)pys;
            const float nphx = -phx, nphy = -phy;
            const float nplx = -(float)(pxs - (double)phx), nply = -(float)(pys - (double)phy);
            // With softening, a body's own leaf contributes exactly +0 (d = 0, finite f), so the reference's
            // "skip the leaf that holds the target" (barnes_hut.cpp:272) needs no test; without it d2 = 0 -> inf*0.
            const bool selfTest = STATS || !(eps2f > 0.f);
            const float bandLo = 1.0f - OPEN_BAND, bandHi = 1.0f + OPEN_BAND;
            double AX = 0.0, AY = 0.0;
            float ax = 0.f, ay = 0.f;
            while (alive) {
                if (k > (int)((cstack >> (2 * d)) & 3ull)) {
                    // child block finished: flush the fp32 partial sums and return to the parent's next slot
                    AX += (double)ax; AY += (double)ay;
                    ax = 0.f; ay = 0.f;
                    if (d == 0) break;
                    --d;
                    k = (int)((kstack >> (2 * d)) & 3ull) + 1;
                    continue;
                }
                const TravRec* R = frames + (d * 4 + k);
                const float4 C = R->c;
                const float4 Bq = *reinterpret_cast<const float4*>(&R->gm);   // gm, open_t, node, cblock
                const unsigned int cb = __float_as_uint(Bq.w);
                const unsigned int slot = 4u * fblock[d] + (unsigned int)k;
                const bool active = d <= accDepth;
                if (active) accDepth = ACTIVE;
                const float dx = (C.x + nphx) + (C.z + nplx);
                const float dy = (C.y + nphy) + (C.w + nply);
                float d2 = fmaf(dx, dx, fmaf(dy, dy, eps2f));
                // a lane that accepted an ancestor sees the node infinitely far away: never opens, contributes 0
                d2 = active ? d2 : INF;
                float lo = Bq.y * bandLo;
                if (d2 > lo && d2 < Bq.y * bandHi)   // rare: inside the guard band -> the reference's fp64 test decides
                    lo = exact_open_slot(&a, &c, slot, pxs, pys) ? INF : -1.f;
                const bool open = d2 <= lo;
                const bool anyopen = __any_sync(0xFFFFFFFFu, open);
                if (active && !open) accDepth = d;   // accepted: the subtree below is ignored
                float rinv;
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(open ? INF : d2));
                float f = (Bq.x * rinv) * (rinv * rinv);
                if (selfTest && slot == self) f = 0.f;
                ax = fmaf(dx, f, ax);
                ay = fmaf(dy, f, ay);
                if (STATS) {
                    nwarp++;
                    nvis += active ? 1u : 0u;
                    const bool counted = active && !open && slot != self && Bq.y != -2.0f;
                    nacc += counted ? 1u : 0u;
                    const double q = counted ? (double)(Bq.x * rinv * rinv) : 0.0;   // gm / d2, scaled units
                    fsum += q;
                    fmaxd = fmax(fmaxd, q);
                }
                if (anyopen) {
                    kstack = (kstack & ~(3ull << (2 * d))) | ((unsigned long long)k << (2 * d));
                    ++d;
                    cstack = (cstack & ~(3ull << (2 * d))) | ((unsigned long long)(cb & 3u) << (2 * d));
                    k = 0;
                    load_block(a.rec, cb >> 2, frames + lpe_idx((unsigned int)d, (unsigned int)TRAV_FRAMES, 13, a.s) * 4, lane, c, a.s);
                    if (lane == 0) fblock[d] = cb >> 2;
                    __syncwarp();
                } else {
                    ++k;
                }
            }
            // a = G * sum M d / r^3 ; scaled units: M/Ms, d/S  =>  factor G*Ms/S^2
            const double accScale = c.G * massScale * c.invS * c.invS;
            if (target) {
                v.x = kick_step(v.x, AX * accScale, c.dtK);   // barnes_hut.cpp:284-286
                v.y = kick_step(v.y, AY * accScale, c.dtK);
            }
            if (STATS) { fsum *= accScale * bodyMass; fmaxd *= accScale * bodyMass; }
        } else {
            // STRICT: the reference's arithmetic, operation for operation (barnes_hut.cpp:257-286), in real units.
            // The walk visits the children of a cell in key order; with Morton keys that is the reference's
            // nw,ne,sw,se recursion order, so the velocity sum has the same order too.
            const double m = bodyMass;
            const double eps2 = __dmul_rn(c.eps, c.eps);
            while (alive) {
                if (k > (int)((cstack >> (2 * d)) & 3ull)) {
                    if (d == 0) break;
                    --d;
                    k = (int)((kstack >> (2 * d)) & 3ull) + 1;
                    continue;
                }
                const TravRec* R = frames + (d * 4 + k);
                const float open_t = R->open_t;
                const unsigned int cblock = R->cblock;
                const unsigned int slot = 4u * fblock[d] + (unsigned int)k;
                // exact mass / centre / level of the node: from the rank's own aggregates, or (top of a decomposed tree,
                // imported cells) from the fp64 side record that travelled with the record
                int level;
                double M, cx, cy;
                if (slot >= a.localLo && slot < a.localHi) {
                    const unsigned int rn = R->node;
                    if (rn & LPE_LEAF_FLAG) {   // single-body leaf: no aggregate is stored, the node is the body
                        const Body lb = a.body[rn & ~LPE_LEAF_FLAG];
                        M = lb.m; cx = lb.x; cy = lb.y; level = -1;
                    } else {
                        const Agg ag = a.agg[rn];
                        level = agg_level(ag);
                        node_centre(ag, level, c.quirk, M, cx, cy);
                    }
                } else {
                    const double4 x = a.xrec[slot];
                    cx = x.x; cy = x.y; M = x.z; level = (int)x.w;
                }
                const double dx = (cx * c.invS - pxs) * c.S, dy = (cy * c.invS - pys) * c.S;   // exact: S is a power of two
                const double distSq = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), eps2);
                const bool active = d <= accDepth;
                if (active) accDepth = ACTIVE;
                bool open = false;
                if (level >= 0 && open_t != -2.0f) {
                    const double size = ldexp(c.U, -level);
                    open = !(__ddiv_rn(__dmul_rn(size, size), distSq) < c.theta2);
                }
                open = open && active;
                const bool anyopen = __any_sync(0xFFFFFFFFu, open);
                const bool acc = active && !open;
                if (acc) accDepth = d;
                if (acc && slot != self && open_t != -2.0f) {
                    const double dist = sqrt(distSq);
                    const double force = __ddiv_rn(__dmul_rn(__dmul_rn(c.G, M), m), distSq);
                    const double invDistMass = __ddiv_rn(force, __dmul_rn(m, dist));
                    v.x = __dadd_rn(v.x, __dmul_rn(__dmul_rn(dx, invDistMass), c.dtK));
                    v.y = __dadd_rn(v.y, __dmul_rn(__dmul_rn(dy, invDistMass), c.dtK));
                    if (STATS) { nacc++; fsum += force; fmaxd = fmax(fmaxd, force); }
                }
                if (STATS) {
                    nwarp++;
                    nvis += active ? 1u : 0u;
                }
                if (anyopen) {
                    kstack = (kstack & ~(3ull << (2 * d))) | ((unsigned long long)k << (2 * d));
                    ++d;
                    cstack = (cstack & ~(3ull << (2 * d))) | ((unsigned long long)(cblock & 3u) << (2 * d));
                    k = 0;
                    load_block(a.rec, cblock >> 2, frames + lpe_idx((unsigned int)d, (unsigned int)TRAV_FRAMES, 13, a.s) * 4, lane, c, a.s);
                    if (lane == 0) fblock[d] = cblock >> 2;
                    __syncwarp();
                } else {
                    ++k;
                }
            }
        }

        if (PREC == 0 && a.stage_out) {
            if (valid) a.stage_out[a.orig ? a.orig[b] : b] = make_double4(p.x, p.y, v.x, v.y);
        } else if (valid) {
            // STRICT reads single-body leaves straight from the state (a.body), so positions must not move while the
            // kernel runs: its drift is a separate elementwise pass (k_drift) after the traversal.
            const bool mover = (PREC == 0 || c.shard_n > 1) && (cm & 2u) && !(cm & 4u) && !(cm & 8u);   // movement.cpp:20-29
            if (c.do_drift && mover) {
                p.x = drift_step(p.x, v.x, c.dtD);                                      // movement.cpp:32-33
                p.y = drift_step(p.y, v.y, c.dtD);
            }
            if (c.shard_n > 1) {
                const unsigned long long slot = (unsigned long long)lblock * 2048ull + within * 32ull + lane;
                const double4 out = make_double4(p.x, p.y, v.x, v.y);
                a.xchg_send[slot] = out;
                for (int k = 1; k <= a.npeer; ++k) {   // rotating peer order, streaming stores (see bh_traverse2.cuh)
                    int r = c.shard_rank + k;
                    if (r >= a.npeer) r -= a.npeer;
                    double2* dst = reinterpret_cast<double2*>(a.peer[r] + slot);
                    __stcs(dst, make_double2(out.x, out.y));
                    __stcs(dst + 1, make_double2(out.z, out.w));
                }
            } else {
                if (target) a.vel[b] = v;
                if (c.do_drift && mover) *reinterpret_cast<double2*>(&a.body[b].x) = p;
            }
        }
        if (STATS && valid) {
            a.cntAcc[b] = nacc;
            a.cntVis[b] = nvis;
        }
        if (STATS) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                nacc += __shfl_xor_sync(0xFFFFFFFFu, nacc, o);
                nvis += __shfl_xor_sync(0xFFFFFFFFu, nvis, o);
                fsum += __shfl_xor_sync(0xFFFFFFFFu, fsum, o);
                fmaxd = fmax(fmaxd, __shfl_xor_sync(0xFFFFFFFFu, fmaxd, o));
            }
            if (lane == 0) {
                atomicAdd(&a.s->force_sum, fsum);
                atomicMax(&a.s->force_max_bits, (unsigned long long)__double_as_longlong(fmaxd));
                atomicAdd(&a.s->interactions, (unsigned long long)nacc);
                atomicAdd(&a.s->visits, (unsigned long long)nvis);
                atomicAdd(&a.s->warp_visits, (unsigned long long)nwarp);
            }
        }
        __syncwarp();   // the frame stack is reused by the next chunk
    }
}

}  // namespace lpe
