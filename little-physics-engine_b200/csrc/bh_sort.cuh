// bh_sort.cuh — LSD radix sort (u64 key, u32 payload, 8-bit digits) and a device-wide exclusive scan.
//
// HBM-bound integer work: per pass the keys are read once for the per-tile digit histogram and keys+payload
// are read and written once by the scatter (algorithmic 8 + 12 + 12 = 32 B per element per pass).
// Tiles are 2048 elements (256 threads x 8); a warp owns a contiguous 256-element run of its tile, so loads
// are coalesced and the stable rank of an element is (warps before) + (earlier rounds of this warp) +
// (lower lanes with the same digit), found with __match_any_sync — no per-element atomics.
#pragma once
#include "bh_common.cuh"

namespace lpe {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 2048
constexpr int SORT_WARPS = SORT_THREADS / 32;

// table[d * numTiles + tile] = number of elements of `tile` whose digit is d; totals[d] += the same.
// `bins` is 256, or 512 for a 9-bit top digit (a 33-bit key then needs 4 passes instead of 5).
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_count(const unsigned long long* __restrict__ keys, int n, int shift, int bins, int numTiles,
             unsigned int* __restrict__ table, unsigned int* __restrict__ totals) {
    __shared__ unsigned int hist[512];
    const int tid = threadIdx.x;
    hist[tid] = 0;
    hist[tid + 256] = 0;
    __syncthreads();
    const unsigned int dmask = (unsigned int)bins - 1u;
    const long long base = (long long)blockIdx.x * SORT_TILE;
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const long long i = base + r * SORT_THREADS + tid;
        if (i < n) atomicAdd(&hist[(unsigned int)(keys[i] >> shift) & dmask], 1u);
    }
    __syncthreads();
    for (int d = tid; d < bins; d += SORT_THREADS) {
        const unsigned int c = hist[d];
        table[(size_t)d * numTiles + blockIdx.x] = c;
        if (c) atomicAdd(&totals[d], c);
    }
}

// One block per digit: exclusive scan of that digit's row of the table, offset by the count of all smaller digits.
__global__ void __launch_bounds__(256)
k_sort_scan(unsigned int* __restrict__ table, const unsigned int* __restrict__ totals, int numTiles) {
    __shared__ unsigned int sh[256];
    __shared__ unsigned int s_base;
    const int d = blockIdx.x;
    const int tid = threadIdx.x;
    // base = sum of totals of smaller digits (up to 512 digits)
    sh[tid] = ((tid < d) ? totals[tid] : 0u) + ((tid + 256 < d) ? totals[tid + 256] : 0u);
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) sh[tid] += sh[tid + o];
        __syncthreads();
    }
    if (tid == 0) s_base = sh[0];
    __syncthreads();
    const unsigned int base = s_base;
    unsigned int* row = table + (size_t)d * numTiles;
    const int per = (numTiles + 255) / 256;
    const int lo = tid * per;
    const int hi = min(lo + per, numTiles);
    unsigned int sum = 0;
    for (int i = lo; i < hi; ++i) sum += row[i];
    __syncthreads();
    sh[tid] = sum;
    __syncthreads();
    // inclusive Hillis-Steele over 256 partial sums
    for (int o = 1; o < 256; o <<= 1) {
        unsigned int v = (tid >= o) ? sh[tid - o] : 0u;
        __syncthreads();
        sh[tid] += v;
        __syncthreads();
    }
    unsigned int run = base + sh[tid] - sum;
    for (int i = lo; i < hi; ++i) {
        const unsigned int c = row[i];
        row[i] = run;
        run += c;
    }
}

template <int BINS>
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_scatter(const unsigned long long* __restrict__ keysIn, const unsigned int* __restrict__ valsIn,
               unsigned long long* __restrict__ keysOut, unsigned int* __restrict__ valsOut, int n, int shift,
               int numTiles, const unsigned int* __restrict__ table) {
    // The tile is first sorted by digit INSIDE shared memory, then written out: consecutive threads then store
    // consecutive addresses within each digit's run, so every 32-byte sector written is fully used (a direct
    // scatter from registers wrote 8-byte keys and 4-byte payloads to 32 different sectors per instruction).
    __shared__ unsigned int cnt[SORT_WARPS][BINS];     // per-warp digit counts -> tile-local offsets
    __shared__ unsigned int dbase[BINS];               // tile-local start of each digit's run
    __shared__ unsigned int gbase[BINS];               // global start of this tile's run of each digit
    __shared__ unsigned long long skey[SORT_TILE];
    __shared__ unsigned int sval[SORT_TILE];
    constexpr int bins = BINS;
    const int tid = threadIdx.x;
    const int w = tid >> 5;
    const int lane = tid & 31;
    const unsigned int dmask = (unsigned int)bins - 1u;
    for (int k = tid; k < SORT_WARPS * BINS; k += SORT_THREADS) (&cnt[0][0])[k] = 0;
    __syncthreads();

    const long long tbase = (long long)blockIdx.x * SORT_TILE;
    const long long wbase = tbase + (long long)w * (SORT_ITEMS * 32);
    unsigned long long key[SORT_ITEMS];
    unsigned int val[SORT_ITEMS];
    unsigned short rk[SORT_ITEMS];
    const unsigned int lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const long long i = wbase + r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? keysIn[i] : ~0ull;
        val[r] = ok ? valsIn[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const long long i = wbase + r * 32 + lane;
        const bool ok = i < n;
        const unsigned int d = ok ? ((unsigned int)(key[r] >> shift) & dmask) : 0xFFFFu;
        const unsigned int peers = __match_any_sync(0xFFFFFFFFu, d);
        const unsigned int before = __popc(peers & lt);
        const int leader = __ffs(peers) - 1;
        unsigned int old = 0;
        if (ok && lane == leader) {
            old = cnt[w][d];
            cnt[w][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xFFFFFFFFu, old, leader);
        rk[r] = (unsigned short)(old + before);
        __syncwarp();
    }
    __syncthreads();
    // per digit: total in this tile, and the exclusive offsets of the warps inside the digit's run
    for (int d = tid; d < bins; d += SORT_THREADS) {
        unsigned int run = 0;
#pragma unroll
        for (int ww = 0; ww < SORT_WARPS; ++ww) {
            const unsigned int c = cnt[ww][d];
            cnt[ww][d] = run;
            run += c;
        }
        dbase[d] = run;   // digit total for now
        gbase[d] = table[(size_t)d * numTiles + blockIdx.x];
    }
    __syncthreads();
    // exclusive scan of the digit totals (bins <= 512, 256 threads: two per thread, Hillis-Steele over pair sums)
    {
        __shared__ unsigned int pair[SORT_THREADS];
        constexpr int PER = BINS / SORT_THREADS;   // 1 or 2
        unsigned int v[PER];
        unsigned int sum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) { v[k] = dbase[tid * PER + k]; sum += v[k]; }
        pair[tid] = sum;
        __syncthreads();
        for (int o = 1; o < SORT_THREADS; o <<= 1) {
            const unsigned int t = (tid >= o) ? pair[tid - o] : 0u;
            __syncthreads();
            pair[tid] += t;
            __syncthreads();
        }
        unsigned int run = pair[tid] - sum;
#pragma unroll
        for (int k = 0; k < PER; ++k) { dbase[tid * PER + k] = run; run += v[k]; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const long long i = wbase + r * 32 + lane;
        if (i < n) {
            const unsigned int d = (unsigned int)(key[r] >> shift) & dmask;
            const unsigned int lp = dbase[d] + cnt[w][d] + rk[r];
            skey[lp] = key[r];
            sval[lp] = val[r];
        }
    }
    __syncthreads();
    const int tileCount = (int)min((long long)SORT_TILE, (long long)n - tbase);
    for (int j = tid; j < tileCount; j += SORT_THREADS) {
        const unsigned long long kx = skey[j];
        const unsigned int d = (unsigned int)(kx >> shift) & dmask;
        const unsigned int dst = gbase[d] + ((unsigned int)j - dbase[d]);
        keysOut[dst] = kx;
        valsOut[dst] = sval[j];
    }
}

// ------------------------------------------------------------------------------------------------
// Device-wide exclusive scan of a u32 sequence produced on the fly by `Load` (reduce / spine / apply).
// out[i] = sum_{j<i} load(j) for i in [0, n]; out[n] (= total) is also stored to *total when non-null.
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned int block_exclusive_scan_256(unsigned int v, unsigned int* sh, unsigned int* total) {
    // sh: 8 words (one per warp) + 1
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sh[w] = inc;
    __syncthreads();
    if (w == 0) {
        unsigned int s = (lane < 8) ? sh[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xFFFFFFFFu, s, o);
            if (lane >= o) s += t;
        }
        if (lane < 8) sh[lane] = s;  // inclusive warp sums
    }
    __syncthreads();
    const unsigned int wprefix = (w > 0) ? sh[w - 1] : 0u;
    if (total) *total = sh[7];
    const unsigned int ex = wprefix + inc - v;
    __syncthreads();
    return ex;
}

template <class Load>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(Load load, int n, unsigned int* __restrict__ tileSums) {
    __shared__ unsigned int sh[9];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    unsigned int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const long long i = base + k;
        if (i < n) s += load((int)i);
    }
    unsigned int tot;
    block_exclusive_scan_256(s, sh, &tot);
    if (threadIdx.x == 0) tileSums[blockIdx.x] = tot;
}

// single block: exclusive scan of tileSums[0..numTiles) in place; total to tileSums[numTiles] and *total
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_spine(unsigned int* __restrict__ tileSums, int numTiles,
                                                             unsigned int* __restrict__ total) {
    __shared__ unsigned int sh[9];
    unsigned int carry = 0;
    for (int base = 0; base < numTiles; base += SCAN_THREADS) {
        const int i = base + threadIdx.x;
        const unsigned int v = (i < numTiles) ? tileSums[i] : 0u;
        unsigned int tot;
        const unsigned int ex = block_exclusive_scan_256(v, sh, &tot);
        if (i < numTiles) tileSums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        tileSums[numTiles] = carry;
        if (total) *total = carry;
    }
}

template <class Load>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(Load load, int n, const unsigned int* __restrict__ tileSums,
                                                             unsigned int* __restrict__ out) {
    __shared__ unsigned int sh[9];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    unsigned int v[SCAN_ITEMS];
    unsigned int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const long long i = base + k;
        v[k] = (i < n) ? load((int)i) : 0u;
        s += v[k];
    }
    unsigned int run = tileSums[blockIdx.x] + block_exclusive_scan_256(s, sh, nullptr);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const long long i = base + k;
        if (i <= n) out[i] = run;  // note: i == n stores the grand total
        run += v[k];
    }
}

}  // namespace lpe
