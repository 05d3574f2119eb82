// bh_traverse.cuh — theta-criterion force traversal fused with the kick (and optional drift).
//
// Replaces BarnesHutSystem::calculateForce (reference barnes_hut.cpp:240-294) and, when do_drift is set,
// MovementSystem::update (movement.cpp:13-39).
//
// One warp walks the pre-order node array for 32 Morton-consecutive targets. The walk is stackless: the
// node index only moves forward (j+1 = first child, skip[j] = next node outside the subtree). Each lane keeps
// its OWN accept/open decision, as the reference does per body: a lane that accepted a node ignores that
// node's descendants by remembering skipUntil = skip[j]; the warp descends while any lane still opens.
// All lanes read the same node -> one broadcast 2 x 16 B load per visited node, served from L1/L2.
//
// FAST precision: state and node centres stay fp64; the difference is formed in fp64 and rounded once to fp32
// (relative error 6e-8 of |d|, independent of where in the universe the pair sits); d^2, rsqrt and the
// accumulation are fp32 with periodic fp64 flushes. The theta test is done in fp32 with a guard band; inside
// the band the reference's own fp64 expression (barnes_hut.cpp:261-269) decides, so accept/open decisions are
// the reference's, not an approximation of them.
// STRICT precision: every interaction in fp64, in the reference's expression order.
#pragma once
#include "bh_common.cuh"

namespace lpe {

constexpr int TRAV_THREADS = 256;
constexpr float TRAV_BAND = 4e-6f;  // relative half-width of the fp32 guard band around s^2/theta^2

struct TravArgs {
    const double2* nodeA;
    const NodeB* nodeB;
    const double* nodeM;
    const double2* spos;
    const double* smass;
    const unsigned int* sidx;
    const unsigned int* selfnode;
    const unsigned char* comp;
    double2* pos;
    double2* vel;
    double4* xchg_send;      // sharded mode: packed (x,y,vx,vy) of the own slice
    unsigned int* cntAcc;    // STATS only, creation order
    unsigned int* cntVis;
    Scal* s;
    unsigned int n_chunks_local;  // 32-body chunks this rank owns
};

// The reference's test, barnes_hut.cpp:261-269, on exactly scaled operands (power-of-two scaling commutes
// with IEEE rounding): returns true when the node must be opened.
__device__ __noinline__ bool exact_open(double dxs, double dys, double eps2s, int level, double Us, double theta2) {
    const double distSq = __dadd_rn(__dadd_rn(__dmul_rn(dxs, dxs), __dmul_rn(dys, dys)), eps2s);
    const double size = ldexp(Us, -level);
    const double sizeSq = __dmul_rn(size, size);
    return !(__ddiv_rn(sizeSq, distSq) < theta2);
}

template <int PREC, bool STATS>
__global__ void __launch_bounds__(TRAV_THREADS) k_traverse(StepConst c, TravArgs a) {
    const int lane = threadIdx.x & 31;
    const unsigned int n_nodes = a.s->n_term + a.s->n_internal;
    const double massScale = 1.0 / mass_scale_inv(a.s->max_mass_bits);
    const float eps2f = (float)c.eps2s;
    const double Us = c.U * c.invS;
    constexpr unsigned int CHUNKS_PER_BLOCK = 2048u / 32u;  // LPE_SHARD_BLOCK / 32

    while (true) {
        unsigned int q = 0;
        if (lane == 0) q = atomicAdd(&a.s->work_counter, 1u);
        q = __shfl_sync(0xFFFFFFFFu, q, 0);
        if (q >= a.n_chunks_local) break;
        // block-cyclic ownership of sorted positions (identity when shard_n == 1)
        const unsigned int lblock = q / CHUNKS_PER_BLOCK, within = q % CHUNKS_PER_BLOCK;
        const unsigned int gblock = lblock * (unsigned int)c.shard_n + (unsigned int)c.shard_rank;
        const long long i = ((long long)gblock * CHUNKS_PER_BLOCK + within) * 32 + lane;
        const bool valid = i < c.n;

        unsigned int b = 0, self = LPE_NONE;
        unsigned char cm = 0;
        double2 p = make_double2(0.0, 0.0);
        if (valid) {
            b = a.sidx[i];
            cm = a.comp[b];
            p = a.spos[i];
            self = a.selfnode[i];
        }
        // bodyView of update(): Position + Velocity + Mass, not Boundary (barnes_hut.cpp:89)
        const bool target = valid && (cm & 1u) && (cm & 2u) && !(cm & 4u);
        const double pxs = p.x * c.invS, pys = p.y * c.invS;
        unsigned int skipUntil = target ? 0u : 0xFFFFFFFFu;
        unsigned int nacc = 0, nvis = 0;
        double2 v = make_double2(0.0, 0.0);
        if (valid) v = a.vel[b];

        if constexpr (PREC == 0) {
            float ax = 0.f, ay = 0.f;
            double AX = 0.0, AY = 0.0;
            unsigned int j = 0, it = 0;
            while (j < n_nodes) {
                const double2 A = a.nodeA[j];
                const NodeB B = a.nodeB[j];
                const double dxd = A.x - pxs, dyd = A.y - pys;
                const float dx = (float)dxd, dy = (float)dyd;
                const float d2 = fmaf(dx, dx, fmaf(dy, dy, eps2f));
                const bool active = j >= skipUntil;
                bool open = d2 <= B.open_d2;
                if (active && fabsf(d2 - B.open_d2) <= B.open_d2 * TRAV_BAND)
                    open = exact_open(dxd, dyd, c.eps2s, B.level, Us, c.theta2);
                open = open && active;
                const bool anyopen = __any_sync(0xFFFFFFFFu, open);
                const bool acc = active && !open;
                if (acc) skipUntil = B.skip;
                const float rinv = rsqrtf(d2);
                float f = B.gm * rinv * (rinv * rinv);
                const bool contrib = acc && (j != self);
                f = contrib ? f : 0.f;
                ax = fmaf(dx, f, ax);
                ay = fmaf(dy, f, ay);
                if (STATS) {
                    nvis += active ? 1u : 0u;
                    nacc += (contrib && B.level != -3) ? 1u : 0u;
                }
                j = anyopen ? j + 1u : max(B.skip, j + 1u);  // max(): a corrupt skip can never stall the walk
                if ((++it & 31u) == 0u) {
                    AX += (double)ax; AY += (double)ay;
                    ax = 0.f; ay = 0.f;
                }
            }
            AX += (double)ax; AY += (double)ay;
            // a = G * sum M d / r^3 ; scaled units: M/Ms, d/S  =>  factor G*Ms/S^2
            const double accScale = c.G * massScale * c.invS * c.invS;
            if (target) {
                v.x += (AX * accScale) * c.dtK;   // barnes_hut.cpp:284-286
                v.y += (AY * accScale) * c.dtK;
            }
        } else {
            // STRICT: the reference's arithmetic, operation for operation (barnes_hut.cpp:257-286), in real units.
            // Pre-order == the reference's nw,ne,sw,se recursion order, so the velocity sum has the same order too.
            const double m = valid ? a.smass[i] : 1.0;
            const double eps2 = __dmul_rn(c.eps, c.eps);
            unsigned int j = 0;
            while (j < n_nodes) {
                const double2 A = a.nodeA[j];
                const NodeB B = a.nodeB[j];
                const double M = a.nodeM[j];
                const double dx = (A.x - pxs) * c.S, dy = (A.y - pys) * c.S;   // exact: S is a power of two
                const double distSq = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), eps2);
                const bool active = j >= skipUntil;
                bool open = false;
                if (B.level >= 0) {
                    const double size = ldexp(c.U, -B.level);
                    open = !(__ddiv_rn(__dmul_rn(size, size), distSq) < c.theta2);
                }
                open = open && active;
                const bool anyopen = __any_sync(0xFFFFFFFFu, open);
                const bool acc = active && !open;
                if (acc) skipUntil = B.skip;
                if (acc && j != self && B.level != -3) {
                    const double dist = sqrt(distSq);
                    const double force = __ddiv_rn(__dmul_rn(__dmul_rn(c.G, M), m), distSq);
                    const double invDistMass = __ddiv_rn(force, __dmul_rn(m, dist));
                    v.x = __dadd_rn(v.x, __dmul_rn(__dmul_rn(dx, invDistMass), c.dtK));
                    v.y = __dadd_rn(v.y, __dmul_rn(__dmul_rn(dy, invDistMass), c.dtK));
                    if (STATS) nacc++;
                }
                if (STATS) nvis += active ? 1u : 0u;
                j = anyopen ? j + 1u : max(B.skip, j + 1u);
            }
        }

        if (valid) {
            const bool mover = (cm & 2u) && !(cm & 4u) && !(cm & 8u);   // movement.cpp:20-29
            if (c.do_drift && mover) {
                p.x += v.x * c.dtD;                                       // movement.cpp:32-33
                p.y += v.y * c.dtD;
            }
            if (c.shard_n > 1) {
                const unsigned long long slot = (unsigned long long)lblock * 2048ull + within * 32ull + lane;
                a.xchg_send[slot] = make_double4(p.x, p.y, v.x, v.y);
            } else {
                if (target) a.vel[b] = v;
                if (c.do_drift && mover) a.pos[b] = p;
            }
            if (STATS) {
                a.cntAcc[b] = nacc;
                a.cntVis[b] = nvis;
            }
        }
        if (STATS) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                nacc += __shfl_xor_sync(0xFFFFFFFFu, nacc, o);
                nvis += __shfl_xor_sync(0xFFFFFFFFu, nvis, o);
            }
            if (lane == 0) {
                atomicAdd(&a.s->interactions, (unsigned long long)nacc);
                atomicAdd(&a.s->visits, (unsigned long long)nvis);
            }
        }
    }
}

}  // namespace lpe
