// bh_common.cuh — shared device-side types for the B200 Barnes-Hut step.
//
// Data layout in HBM:
//   state:                   body Body[n] (32 B: x, y, m, rank, comp), vel double2[n], orig u32[n] (creation index of a
//                            slot) — creation order right after an upload, KEY ORDER after every step's gather
//   sort:                    keys u64[n], sidx u32[n]
//   terminals (t < n_term):  tkey u64, tfirst u32, delta i8, mask u32, wstart u32
//   nodes (pre-order index): agg Agg (64 B) for cells and aggregated terminals (it carries the cell's level and ordinal)
//   cells (ordinal):         child uint4;  records: rec TravRec[4 * (cells + 1)] in child blocks of 128 B
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define LPE_NONE 0xFFFFFFFFu
#define LPE_VAL_OUT 0x80000000u   // sort payload, 32-bit key mode: the body is not in the tree (sorts behind every cell)
#define LPE_MAX_DEPTH 30

// ---- checked build (liblpe_bh_checked.so, -DLPE_CHECKED) --------------------------------------------------------------
// compute-sanitizer is not available on this pool, so the index arithmetic of the kernels carries its own bounds checks:
// in the checked library every computed index into the big arrays is compared with the array's extent first; a
// violation raises a bit in Scal::check_fault (reported as an error by the next synchronising call) and the access is
// skipped. The GPU test-suite runs once against the checked library (tests/test_checked_build_gpu.py). In the
// production library the macro is empty.
#ifdef LPE_CHECKED
#define LPE_CHECK(cond, code, scal) \
    do { if (!(cond)) { atomicOr(&(scal)->check_fault, 1u << (code)); return; } } while (0)
#define LPE_CHECK_NR(cond, code, scal) \
    do { if (!(cond)) atomicOr(&(scal)->check_fault, 1u << (code)); } while (0)
#else
#define LPE_CHECK(cond, code, scal) ((void)0)
#define LPE_CHECK_NR(cond, code, scal) ((void)0)
#endif

namespace lpe {

// The state update of the tick, one rounding per operation like the reference's own code (g++ -O2 on x86-64 never fuses
// a*b+c): kick `vel += acc * dt` (barnes_hut.cpp:284-286) and drift `pos += vel * dt` (movement.cpp:32-33). Every
// kernel that kicks or drifts goes through these two, so the fused epilogue of the traversal kernels, the separate
// passes (k_drift, k_finish_tick) and the decomposed ranks produce the same bits.
__device__ __forceinline__ double kick_step(double v, double acc, double dt) { return __dadd_rn(v, __dmul_rn(acc, dt)); }
__device__ __forceinline__ double drift_step(double p, double v, double dt) { return __dadd_rn(p, __dmul_rn(v, dt)); }


constexpr int LPE_MAX_P2P = 8;   // ranks of one NVSwitch domain that can exchange by direct peer stores


// Scalars produced and consumed on the device inside one step (no host round trip).
struct Scal {
    // largest source mass as raw double bits (positive doubles order like integers). Set by the upload / pack kernels,
    // NOT cleared per step (the per-step memset starts after this word): masses only change through an upload.
    unsigned long long max_mass_bits;
    unsigned int n_in;          // sources inside [0,U)^2
    unsigned int n_term;        // distinct depth-D cells
    unsigned int n_internal;    // branching cells
    unsigned int work_counter;  // dynamic chunk dispenser of the traversal
    unsigned long long interactions;
    unsigned long long visits;
    unsigned long long warp_visits;    // node visits summed over warps (one per loop iteration)
    unsigned int ovf_count;     // chunks the two-phase traversal handed to the depth-first kernel (frontier overflow)
    unsigned int work_counter2; // dispenser of that second launch
    unsigned long long t2[8];   // two-phase diagnostics (STATS only): entries by kind
    // domain-decomposed runs (bh_dd.cuh)
    unsigned int n_live;        // bodies this rank owns in this step (slots [0, n_live) after the gather)
    unsigned int dd_nroots;     // roots of the top tree (0 = empty tree)
    unsigned int dd_myroots;    // roots this rank publishes
    unsigned int exp_count[8];  // child blocks exported to each rank
    unsigned int n_sort;        // slots holding this step's keys: last step's bodies + the migrants that arrived
    unsigned long long work_cost;   // list entries evaluated by this rank's traversal (load-balance weight)
    unsigned int dd_rounds;     // generations of the exporter's breadth-first walk (longest destination)
    unsigned int pad_dd2;
    // DebugStats::updateForce (reference include/core/debug.hpp:37-41, called per accepted node at barnes_hut.cpp:278):
    // max and sum of force = G*M*m/distSq over the accepted interactions (stats runs only; the count is `interactions`)
    unsigned long long force_max_bits;
    double force_sum;
    unsigned int check_fault;   // checked build: bit k = bounds check k failed in this step (see LPE_CHECK)
    unsigned int pad_chk;
};

// Traversal node record, 32 bytes = two broadcast 16-byte shared-memory loads per visited node.
//   c:      centre of mass in scaled units as a two-float (hi + lo) pair per axis: hi = fl32(c), lo = fl32(c - hi), so
//           (c_hi - p_hi) + (c_lo - p_lo) is the fp64 difference rounded once to fp32 (relative error 2^-24 of |d|,
//           wherever in the universe the pair sits) with no FP64 instruction and no conversion in the inner loop.
//   open_t: s^2/theta^2 in scaled units. d2 <= open_t*(1-band) opens, d2 >= open_t*(1+band) accepts, in between
//           the reference's own fp64 expression decides. -1 for leaves / terminals, -2 for the small-mass skip.
//   node:   where the exact (fp64) data of the node lives on the rank that built it: pre-order index of its aggregate,
//           or LPE_LEAF_FLAG | sorted position of the body for a single-body leaf. Read only by the rare fp64 re-test of a
//           borderline theta decision and by STRICT precision.
//   cblock: (index of the 128-byte block holding this cell's children) << 2 | (number of children - 1);
//           0 for leaves / terminals.
struct __align__(16) TravRec {
    float4 c;            // chx, chy, clx, cly
    float gm;            // node mass / mass scale (0 when the node is skipped by the small-mass rule)
    float open_t;
    unsigned int node;
    unsigned int cblock;
};
static_assert(sizeof(TravRec) == 32, "four records per 128-byte line");
constexpr float OPEN_BAND = 4e-6f;  // relative half-width of the fp32 guard band around s^2/theta^2

// One body = one 32-byte sector, so a gather by index costs one sector instead of one per component.
struct __align__(16) Body {
    double x, y, m;
    unsigned int rank;    // insertion rank (position in the reference's view iteration)
    unsigned int comp;    // LPE_HAS_MASS | LPE_HAS_VELOCITY | LPE_BOUNDARY | LPE_LIQUID
};
static_assert(sizeof(Body) == 32, "one sector per body");
#define LPE_LEAF_FLAG 0x80000000u   // child[] entry of a single-body leaf: LPE_LEAF_FLAG | sorted position of the body

// Per-node aggregate carried up the tree (exact sums; the quirk is applied only when a record is made). 64 bytes =
// two sectors; it carries the first occupant's own mass and position so that no second lookup is needed.
struct __align__(16) Agg {
    double m, sx;         // sum m, sum m*x over the bodies under the node
    double sy, mf;        // sum m*y; mass of the first occupant (minimum insertion rank)
    double xf, yf;        // position of the first occupant
    unsigned int frank;   // its insertion rank
    unsigned int fidx;    // its sorted position
    unsigned int ordinal; // cells: ordinal of the cell (its child block is blockBase + ordinal)
    unsigned int small;   // bit 0: every mass under the node is < small_mass_threshold; bits 1-2: children - 1 (cells);
                          // bits 8-13: level + 2 (a branching cell's level, 0 = aggregated terminal at the depth bound)
};
__host__ __device__ __forceinline__ int agg_level(const Agg& a) { return (int)((a.small >> 8) & 63u) - 2; }
static_assert(sizeof(Agg) == 64, "Agg is two 32-byte sectors");

struct StepConst {
    double U, invS, S;        // universe size; power-of-two length scale and its inverse
    double eps, eps2s;        // softening; (eps/S)^2
    double theta, theta2;
    double thr;               // small-mass threshold
    double G;
    double dtK, dtD;
    double h, invh;           // finest cell size U/2^D and its inverse
    int D;                    // key depth
    int quirk;
    int do_drift;
    int n;                    // bodies
    int shard_rank, shard_n;  // multi-GPU block-cyclic ownership of sorted positions
    int need_self;            // maintain selfnode / selfslot (eps == 0 or interaction counting)
    int test_overflow;        // tests only: pretend every two-phase frontier overflows
    int hilbert;              // sort key: 0 = Morton code, 1 = Hilbert index of the same depth-D cell
    int k32;                  // depth <= 16: keys travel as 32-bit words, "not in the tree" in the payload's top bit
    unsigned int recSlots;    // extent of the record array (slots), nodeCap of the per-node arrays, bodyCap of the per-body ones
    unsigned int nodeCap, bodyCap;
    int dd;                   // domain-decomposed rank: the local build leaves the root block and the top of the tree alone
    unsigned int blockBase;   // child block of the cell with ordinal q is blockBase + q (1 on a single GPU: block 0 = root)
    float eps2f;              // (float)eps2s, converted once on the host (the kernels would re-convert it in their loops)
};

// checked build: an index is compared with the array's extent before it is used; out of range raises bit `code` of
// Scal::check_fault and the access goes to element 0 instead. Production build: the index, untouched.
__device__ __forceinline__ unsigned int lpe_idx(unsigned int i, unsigned int extent, int code, const Scal* s) {
#ifdef LPE_CHECKED
    if (i >= extent) {
        atomicOr(&const_cast<Scal*>(s)->check_fault, 1u << code);
        return 0u;
    }
#else
    (void)extent; (void)code; (void)s;
#endif
    return i;
}

// Mass and centre of mass of a node as the traversal sees it, in real units.
//   leaf: exactly the body (barnes_hut.cpp:144-153); internal cell or aggregated terminal: with quirk, the first
//   occupant counted twice (barnes_hut.cpp:157-177).
__host__ __device__ __forceinline__ void node_centre(const Agg& a, int level, int quirk, double& M, double& cx, double& cy) {
    if (level == -1) {
        M = a.m; cx = a.xf; cy = a.yf;
        return;
    }
    double sx = a.sx, sy = a.sy;
    M = a.m;
    if (quirk) {
        M += a.mf;
        sx += a.mf * a.xf;
        sy += a.mf * a.yf;
    }
    cx = sx / M;
    cy = sy / M;
}

__device__ __forceinline__ unsigned long long spread_bits32(unsigned int v) {
    unsigned long long x = v;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}

// Sorted / unsorted key i, whichever container the step uses (StepConst::k32)
__device__ __forceinline__ unsigned long long load_key(const void* keys, int k32, long long i) {
    return k32 ? (unsigned long long)reinterpret_cast<const unsigned int*>(keys)[i]
               : reinterpret_cast<const unsigned long long*>(keys)[i];
}
// key + payload of body i as the sort wants them. key = cell index (in the tree), 1 << 2D (outside), (1 << 2D) | 1 (a
// dead slot of a domain-decomposed rank)
__device__ __forceinline__ void store_key(void* keys, unsigned int* vals, int k32, int D, long long i, unsigned long long key) {
    if (k32) {
        const bool out = (key >> (2 * D)) != 0ull;
        reinterpret_cast<unsigned int*>(keys)[i] = out ? (unsigned int)(key & 1ull) : (unsigned int)key;
        vals[i] = (unsigned int)i | (out ? LPE_VAL_OUT : 0u);
    } else {
        reinterpret_cast<unsigned long long*>(keys)[i] = key;
        vals[i] = (unsigned int)i;
    }
}

// Level of the lowest common ancestor cell of two distinct depth-D keys (= number of shared leading digits).
__device__ __forceinline__ int lca_level(unsigned long long a, unsigned long long b, int D) {
    const unsigned long long x = a ^ b;
    const int hb = 63 - __clzll((long long)x);
    return (2 * D - 1 - hb) >> 1;
}

// First terminal index whose level-L prefix equals that of terminal t (galloping + binary search to the left).
__device__ __forceinline__ int cell_first(const unsigned long long* __restrict__ tkey, int t, int shift) {
    const unsigned long long pre = tkey[t] >> shift;
    if (t == 0 || (tkey[t - 1] >> shift) != pre) return t;
    int hi = t - 1;  // known inside
    int lo = -1;     // known outside (or before the array)
    int step = 1;
    while (true) {
        const int probe = hi - step;
        if (probe < 0) break;
        if ((tkey[probe] >> shift) == pre) {
            hi = probe;
            step <<= 1;
        } else {
            lo = probe;
            break;
        }
    }
    while (hi - lo > 1) {
        const int mid = lo + ((hi - lo) >> 1);
        if ((tkey[mid] >> shift) == pre) hi = mid; else lo = mid;
    }
    return hi;
}

// Last terminal index (< n_term) whose level-L prefix equals that of terminal t.
__device__ __forceinline__ int cell_last(const unsigned long long* __restrict__ tkey, int t, int shift, int n_term) {
    const unsigned long long pre = tkey[t] >> shift;
    if (t == n_term - 1 || (tkey[t + 1] >> shift) != pre) return t;
    int lo = t + 1;   // known inside
    int hi = n_term;  // known outside (or past the array)
    int step = 1;
    while (true) {
        const long long probe = (long long)lo + step;
        if (probe >= n_term) break;
        if ((tkey[probe] >> shift) == pre) {
            lo = (int)probe;
            step <<= 1;
        } else {
            hi = (int)probe;
            break;
        }
    }
    while (hi - lo > 1) {
        const int mid = lo + ((hi - lo) >> 1);
        if ((tkey[mid] >> shift) == pre) lo = mid; else hi = mid;
    }
    return lo;
}

}  // namespace lpe
