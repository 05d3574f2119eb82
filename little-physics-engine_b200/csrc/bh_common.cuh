// bh_common.cuh — shared device-side types for the B200 Barnes-Hut step.
//
// Data layout in HBM (all SoA):
//   state (creation order):  pos double2[n], vel double2[n], mass f64[n], rank u32[n], comp u8[n]
//   sorted (Morton order):   keys u64[n], sidx u32[n], spos double2[n], smass f64[n], srank u32[n]
//   terminals (t < n_term):  tkey u64, tfirst u32, delta i8, mask u32, tnode u32
//   nodes (pre-order index): nodeA double2 (scaled COM), nodeB NodeB (16 B), nodeM f64, parent u32, child uint4, agg Agg
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define LPE_NONE 0xFFFFFFFFu
#define LPE_MAX_DEPTH 30

namespace lpe {

// Scalars produced and consumed on the device inside one step (no host round trip).
struct Scal {
    unsigned int n_in;          // sources inside [0,U)^2
    unsigned int n_term;        // distinct depth-D cells
    unsigned int n_internal;    // branching cells
    unsigned int work_counter;  // dynamic chunk dispenser of the traversal
    unsigned long long max_mass_bits;  // max source mass as raw double bits (positive doubles order like integers)
    unsigned long long interactions;
    unsigned long long visits;
    unsigned long long warp_visits;    // node visits summed over warps (one per loop iteration)
};

// Traversal node record, 32 bytes = two broadcast 16-byte loads per visited node.
//   NodeC: centre of mass in scaled units as a two-float (hi + lo) pair per axis. hi = fl32(c), lo = fl32(c - hi), so
//          (c_hi - p_hi) + (c_lo - p_lo) gives the fp64 difference rounded once to fp32 (relative error 2^-24 of |d|,
//          wherever in the universe the pair sits) with no FP64 instruction and no conversion in the inner loop.
//   NodeB: mass, the theta test as two thresholds around s^2/theta^2 (below open_lo: open; at or above open_hi:
//          accept; in between the reference's fp64 expression decides), and the skip pointer.
struct __align__(16) NodeB {
    float gm;            // node mass / mass scale (0 when the node is skipped by the small-mass rule)
    float open_lo;       // d2 <= open_lo  => every fp64 evaluation would open the node
    float open_hi;       // d2 >= open_hi  => every fp64 evaluation would accept it; -1 leaf/terminal, -2 small-mass skip
    unsigned int skip;   // pre-order index of the first node after this subtree
};

// Per-node aggregate carried up the tree (exact sums, quirk applied only when a record is finalised).
struct __align__(16) Agg {
    double m, sx;         // sum m, sum m*x over the bodies under the node
    double sy;            // sum m*y
    unsigned int frank;   // minimum insertion rank under the node
    unsigned int fidx;    // sorted position of that first occupant
    unsigned int count;   // bodies under the node
    unsigned int small;   // 1 if every mass under the node is < small_mass_threshold
    unsigned int pad[2];
};
static_assert(sizeof(Agg) == 48, "Agg is read back as three 16-byte words");

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
};

__device__ __forceinline__ unsigned long long spread_bits32(unsigned int v) {
    unsigned long long x = v;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
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
