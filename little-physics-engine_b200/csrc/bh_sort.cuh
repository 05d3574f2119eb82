// bh_sort.cuh — LSD radix sort (u32 or u64 key, u32 payload, 8-bit digits) and a device-wide exclusive scan.
//
// HBM-bound integer work by its bytes, latency-bound in practice. One histogram kernel reads the keys once and counts
// the digits of EVERY pass; each pass is then a single kernel that reads keys + payload once and writes them once
// (algorithmic 4 + passes * 16 B per element with 32-bit keys, 8 + passes * 24 B with 64-bit keys): a tile's position
// inside each digit's run comes from a decoupled look-back over the tiles before it (status word = epoch | flag | count,
// one word per tile and digit), not from a separate count + scan launch pair.
// Tiles are 4096 elements (512 threads x 8) for 32-bit keys, 2048 (256 x 8) for 64-bit keys; a warp owns a contiguous
// 256-element run of its tile, so loads are coalesced and the stable rank of an element is (warps before) + (earlier
// rounds of this warp) + (lower lanes with the same digit), found with __match_any_sync or, for small inputs, with one
// ballot per digit bit — no per-element atomics.
#pragma once
#include "bh_common.cuh"

namespace lpe {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 2048
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_MAX_PASSES = 8;
#ifndef SORT_BALLOT_MAX_N
#define SORT_BALLOT_MAX_N 2500000
#endif
constexpr int SORT_BALLOT_MAX = SORT_BALLOT_MAX_N;    // keys: up to here the passes rank by ballots (see k_sort_onesweep)
constexpr int SORT_HIST_STRIDE = 512;                 // words per pass in the histogram / base arrays

// ---- look-back status words: high half = epoch << 2 | state, low half = value -------------------------------
// A word written in an earlier step carries an older epoch and reads as "not there yet", so the status array is
// never cleared between steps.
constexpr unsigned int LB_AGGREGATE = 1u, LB_INCLUSIVE = 2u;
__device__ __forceinline__ unsigned long long lb_load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lb_store(unsigned long long* p, unsigned int epoch, unsigned int state, unsigned int value) {
    const unsigned long long v = ((unsigned long long)((epoch << 2) | state) << 32) | value;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Sum of the values published by the tiles before `tile` for one column (stride = words per tile). Tile ids are
// handed out by an atomic counter, so every earlier tile is already running and never waits on a later one.
__device__ __forceinline__ unsigned int lb_exclusive(const unsigned long long* column, size_t stride, unsigned int tile,
                                                     unsigned int epoch, unsigned int* __restrict__ fault) {
    // The walk looks at LB_WINDOW earlier tiles per round trip (independent loads in flight together) instead of
    // one: when a whole wave of tiles publishes its counts at the same moment, the chain of dependent L2 round
    // trips is what the tiles wait for, not the data.
    constexpr int LB_WINDOW = 8;
    unsigned int excl = 0;
    unsigned int spins = 0;
    unsigned int t = tile;   // tiles [t, tile) are already summed
    while (t > 0u) {
        unsigned long long w[LB_WINDOW];
#pragma unroll
        for (int k = 0; k < LB_WINDOW; ++k) {
            const unsigned int idx = t - 1u - (unsigned int)k;
            // before the first tile: an inclusive zero ends the walk
            w[k] = (idx < t) ? lb_load(column + (size_t)idx * stride)
                             : ((unsigned long long)((epoch << 2) | LB_INCLUSIVE) << 32);
        }
        bool done = false;
        unsigned int used = 0;
#pragma unroll
        for (int k = 0; k < LB_WINDOW; ++k) {
            const unsigned int hi = (unsigned int)(w[k] >> 32);
            const bool ready = (hi >> 2) == epoch;
            if (!done && used == (unsigned int)k && ready) {
                excl += (unsigned int)w[k];
                used = k + 1;
                if ((hi & 3u) == LB_INCLUSIVE) done = true;
            }
        }
        if (done) break;
        t -= used;   // (used <= t: the entries past tile 0 are inclusive and end the walk)
        if (used == 0u) __nanosleep(64);   // nothing new: leave the L2 to the tiles we are waiting for
        if (used == 0u && ++spins > (1u << 22)) {   // a bounded wait turns a protocol bug into an error, not a hang
            atomicExch(fault, 1u);
            return excl;
        }
    }
    return excl;
}

// exclusive scan of one value per thread across a 256-thread block (warp shuffles + one word per warp)
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

// Digit of a key in one pass. 64-bit keys carry everything (bit 2D = "not in the tree" lands in the top digit). 32-bit
// keys (depth <= 16: the cell index fills all 32 bits) keep that flag in the top bit of the payload instead; it joins
// the TOP pass as digit bit 8, so the top pass always has 512 bins and the containers are 4 + 4 bytes per element.
template <class KeyT>
__device__ __forceinline__ unsigned int sort_digit(KeyT key, unsigned int val, int shift, unsigned int dmask, bool top) {
    if constexpr (sizeof(KeyT) == 8) {
        (void)val; (void)top;
        return (unsigned int)(key >> shift) & dmask;
    } else {
        const unsigned int d = (key >> shift) & 255u;
        return top ? (d | ((val >> 31) << 8)) : d;
    }
}

// hist[p * 512 + d] += number of keys whose digit in pass p is d, for every pass at once (keys read once).
template <class KeyT>
__global__ void __launch_bounds__(256)
k_sort_hist(const KeyT* __restrict__ keys, const unsigned int* __restrict__ vals, int n, int passes, int lastBins,
            unsigned int* __restrict__ hist, const unsigned int* __restrict__ n_dev) {
    extern __shared__ unsigned int sh_hist[];   // passes * 512
    if (n_dev) n = (int)*n_dev;   // element count known only on the device (domain-decomposed ranks)
    const int words = passes * SORT_HIST_STRIDE;
    for (int k = threadIdx.x; k < words; k += blockDim.x) sh_hist[k] = 0;
    __syncthreads();
    const unsigned int lastMask = (unsigned int)lastBins - 1u;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const KeyT key = keys[i];
        const unsigned int val = (sizeof(KeyT) == 4) ? vals[i] : 0u;
        for (int p = 0; p < passes - 1; ++p)
            atomicAdd(&sh_hist[p * SORT_HIST_STRIDE + sort_digit<KeyT>(key, val, 8 * p, 255u, false)], 1u);
        atomicAdd(&sh_hist[(passes - 1) * SORT_HIST_STRIDE + sort_digit<KeyT>(key, val, 8 * (passes - 1), lastMask, true)], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < words; k += blockDim.x) {
        const unsigned int c = sh_hist[k];
        if (c) atomicAdd(&hist[k], c);
    }
}

// One block per pass: hist row -> exclusive prefix (global start of each digit's run), in place. Also the place where
// the step's look-back epoch advances: it lives in device memory (not in a kernel argument), so that a captured CUDA
// graph of the step can be replayed unchanged; every kernel that uses the epoch runs after this one.
__global__ void __launch_bounds__(512) k_sort_bases(unsigned int* __restrict__ hist, unsigned int* __restrict__ epoch) {
    __shared__ unsigned int sh[512];
    if (blockIdx.x == 0 && threadIdx.x == 0) *epoch += 1u;
    unsigned int* row = hist + blockIdx.x * SORT_HIST_STRIDE;
    const int tid = threadIdx.x;
    const unsigned int v = row[tid];
    sh[tid] = v;
    __syncthreads();
    for (int o = 1; o < 512; o <<= 1) {
        const unsigned int t = (tid >= o) ? sh[tid - o] : 0u;
        __syncthreads();
        sh[tid] += t;
        __syncthreads();
    }
    row[tid] = sh[tid] - v;
}

// exclusive scan of one value per thread across a block of WARPS warps (<= 32)
template <int WARPS>
__device__ __forceinline__ unsigned int block_exclusive_scan(unsigned int v, unsigned int* sh) {
    // sh: WARPS words
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
        unsigned int s = (lane < WARPS) ? sh[lane] : 0u;
#pragma unroll
        for (int o = 1; o < WARPS; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xFFFFFFFFu, s, o);
            if (lane >= o) s += t;
        }
        if (lane < WARPS) sh[lane] = s;  // inclusive warp sums
    }
    __syncthreads();
    const unsigned int wprefix = (w > 0) ? sh[w - 1] : 0u;
    return wprefix + inc - v;
}

// One pass = one kernel. THREADS x ITEMS keys per tile: 256 x 8 (2048) for 64-bit keys, 512 x 8 (4096) for 32-bit keys.
// The pass is latency-bound, not bandwidth-bound (ncu, 16 M keys): the ranking rounds are a chain of MATCH -> shared-memory
// read-modify-write -> SHFL per key of a thread, and the look-back walks one status column per digit. Hence: few keys per
// thread (short chains) in many warps (48 per SM at 512 x 8), and large tiles (with ~600 tiles in flight the finished
// prefix lags dozens of tiles behind, so every digit column of a tile walks that far: half the tiles, half the walks).
#ifdef SORT_TRACE
__device__ unsigned long long g_sort_trace[8];   // summed clock64 deltas of a tile's phases (thread 0), [7] = tiles
#define SORT_MARK(k) do { if (threadIdx.x == 0) { const long long _t = clock64(); atomicAdd(&g_sort_trace[k], (unsigned long long)(_t - _tprev)); _tprev = _t; } } while (0)
#else
#define SORT_MARK(k) do { } while (0)
#endif
template <int BINS, class KeyT, int THREADS, int ITEMS>
constexpr size_t sort_smem_bytes() {
    return (size_t)THREADS * ITEMS * (sizeof(KeyT) + 4) + (size_t)BINS * 8 + 128 + (size_t)(THREADS / 32) * BINS * 2;
}
// BALLOT: the lanes with the same digit come from one ballot per digit bit instead of MATCH.ANY — more instructions, less
// latency: faster while a pass is a single wave of tiles (-8 % at 1 M keys), slower at 16 M (+3 %); the host picks by size.
template <int BINS, class KeyT, int THREADS, int ITEMS, bool BALLOT>
__global__ void __launch_bounds__(THREADS, THREADS == 512 ? 3 : 1)
k_sort_onesweep(const KeyT* __restrict__ keysIn, const unsigned int* __restrict__ valsIn,
                KeyT* __restrict__ keysOut, unsigned int* __restrict__ valsOut, int n, int shift,
                const unsigned int* __restrict__ digitBase, unsigned long long* __restrict__ status,
                const unsigned int* __restrict__ epoch_ptr, unsigned int* __restrict__ tileCounter, unsigned int* __restrict__ fault,
                const unsigned int* __restrict__ n_dev) {
    if (n_dev) n = (int)*n_dev;   // tiles past the device-side element count take a ticket and leave
    const unsigned int epoch = *epoch_ptr;
    // The tile is first sorted by digit INSIDE shared memory, then written out: consecutive threads then store
    // consecutive addresses within each digit's run, so every 32-byte sector written is fully used (a direct
    // scatter from registers wrote 8-byte keys and 4-byte payloads to 32 different sectors per instruction).
    constexpr int WARPS = THREADS / 32;
    constexpr int TILE = THREADS * ITEMS;
    constexpr bool TOP = BINS == 512;                  // (32-bit keys: the 512-bin pass is the top pass, payload bit 31 joins the digit)
    // (dynamic shared memory: the 512-bin pass of a 4096-key tile needs 52 KB)
    extern __shared__ __align__(16) unsigned char sort_smem[];
    KeyT* skey = reinterpret_cast<KeyT*>(sort_smem);
    unsigned int* sval = reinterpret_cast<unsigned int*>(skey + TILE);
    unsigned int* dbase = sval + TILE;                 // tile-local start of each digit's run
    unsigned int* gbase = dbase + BINS;                // global start of this tile's run of each digit
    unsigned int* sh_scan = gbase + BINS;
    unsigned short (*cnt)[BINS] = reinterpret_cast<unsigned short (*)[BINS]>(sh_scan + 32);   // per-warp digit counts -> tile-local offsets (a tile has <= 4096 keys)
    const int tid = threadIdx.x;
    const int w = tid >> 5;
    const int lane = tid & 31;
    constexpr unsigned int dmask = (unsigned int)BINS - 1u;
    __shared__ unsigned int s_tile;
    if (tid == 0) s_tile = atomicAdd(tileCounter, 1u);   // tiles in start order: look-back never waits on a later block
    for (int k = tid; k < WARPS * BINS / 2; k += THREADS) reinterpret_cast<unsigned int*>(&cnt[0][0])[k] = 0;
    __syncthreads();
    const unsigned int tile = s_tile;
    if ((long long)tile * TILE >= n) return;
#ifdef SORT_TRACE
    long long _tprev = clock64();
    if (threadIdx.x == 0) atomicAdd(&g_sort_trace[7], 1ull);
#endif

    const long long tbase = (long long)tile * TILE;
    const long long wbase = tbase + (long long)w * (ITEMS * 32);
    KeyT key[ITEMS];
    unsigned int val[ITEMS];
    unsigned short rk[ITEMS];
    const unsigned int lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const long long i = wbase + r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? keysIn[i] : (KeyT)~(KeyT)0;
        val[r] = ok ? valsIn[i] : 0u;
    }
    SORT_MARK(0);   // (loads issued)
    // the MATCHes of all rounds are independent of each other: issue them together, then run the counter chain
    unsigned int peers[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const bool ok = wbase + r * 32 + lane < n;
        const unsigned int d = ok ? sort_digit<KeyT>(key[r], val[r], shift, dmask, TOP) : 0xFFFFu;
        if constexpr (BALLOT) {
            // the lanes with the same digit, from one ballot per digit bit (+ one for "in range")
            constexpr int BITS = BINS == 512 ? 9 : 8;
            unsigned int pm = 0xFFFFFFFFu;
#pragma unroll
            for (int b = 0; b < BITS; ++b) {
                const bool bit = (d >> b) & 1u;
                const unsigned int bal = __ballot_sync(0xFFFFFFFFu, bit);
                pm &= bit ? bal : ~bal;
            }
            const unsigned int balOk = __ballot_sync(0xFFFFFFFFu, ok);
            peers[r] = pm & (ok ? balOk : ~balOk);
        } else {
            peers[r] = __match_any_sync(0xFFFFFFFFu, d);
        }
    }
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const bool ok = wbase + r * 32 + lane < n;
        const unsigned int d = sort_digit<KeyT>(key[r], val[r], shift, dmask, TOP);
        const unsigned int before = __popc(peers[r] & lt);
        const int leader = __ffs(peers[r]) - 1;
        unsigned int old = 0;
        if (ok && lane == leader) {
            old = cnt[w][d];
            cnt[w][d] = (unsigned short)(old + __popc(peers[r]));
        }
        old = __shfl_sync(0xFFFFFFFFu, old, leader);
        rk[r] = (unsigned short)(old + before);
        __syncwarp();
    }
    __syncthreads();
    SORT_MARK(1);   // (keys arrived, ranked)
    // per digit: total in this tile, and the exclusive offsets of the warps inside the digit's run. The tile's counts
    // are published at once (aggregate), the wait for the tiles before it comes after the local reordering.
    constexpr int PER = BINS > THREADS ? BINS / THREADS : 1;   // digits per thread: tid * PER + k (threads past BINS idle)
    const bool owner = tid * PER < BINS;
    unsigned int total[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        total[k] = 0;
        if (owner) {
            const int d = tid * PER + k;
            unsigned int run = 0;
#pragma unroll
            for (int ww = 0; ww < WARPS; ++ww) {
                const unsigned int c = cnt[ww][d];
                cnt[ww][d] = (unsigned short)run;
                run += c;
            }
            total[k] = run;
            lb_store(status + (size_t)tile * BINS + d, epoch, tile == 0u ? LB_INCLUSIVE : LB_AGGREGATE, run);
        }
    }
    // exclusive scan of the digit totals -> start of each digit's run inside the tile
    {
        unsigned int sum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) sum += total[k];
        unsigned int run = block_exclusive_scan<WARPS>(sum, sh_scan);
        if (owner) {
#pragma unroll
            for (int k = 0; k < PER; ++k) { dbase[tid * PER + k] = run; run += total[k]; }
        }
    }
    __syncthreads();
    SORT_MARK(2);   // (counts published, scanned)
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const long long i = wbase + r * 32 + lane;
        if (i < n) {
            const unsigned int d = sort_digit<KeyT>(key[r], val[r], shift, dmask, TOP);
            const unsigned int lp = dbase[d] + cnt[w][d] + rk[r];
            skey[lp] = key[r];
            sval[lp] = val[r];
        }
    }
    SORT_MARK(3);   // (scattered into shared memory)
    // now add up the tiles before this one and publish the inclusive value for the tiles after it
    if (owner) {
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int d = tid * PER + k;
            unsigned int excl = 0;
            if (tile != 0u) {
                excl = lb_exclusive(status + d, BINS, tile, epoch, fault);
                lb_store(status + (size_t)tile * BINS + d, epoch, LB_INCLUSIVE, excl + total[k]);
            }
            gbase[d] = digitBase[d] + excl;
        }
    }
    __syncthreads();
    SORT_MARK(4);   // (look-back done)
    const int tileCount = (int)min((long long)TILE, (long long)n - tbase);
    for (int j = tid; j < tileCount; j += THREADS) {
        const KeyT kx = skey[j];
        const unsigned int vx = sval[j];
        const unsigned int d = sort_digit<KeyT>(kx, vx, shift, dmask, TOP);
        const unsigned int dst = gbase[d] + ((unsigned int)j - dbase[d]);
#ifdef LPE_CHECKED
        if (dst >= (unsigned int)n) { atomicOr(fault, 2u); continue; }   // (bit 1 of the sort's fault word)
#endif
        keysOut[dst] = kx;
        valsOut[dst] = vx;
    }
    SORT_MARK(5);   // (stores issued)
}

// ------------------------------------------------------------------------------------------------
// Device-wide exclusive scan of a u32 sequence produced on the fly by `Load`.
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// ------------------------------------------------------------------------------------------------
// Single pass (decoupled look-back, one status word per tile): every element is loaded once and handed to
// `sink(i, exclusive_prefix, value)` for i in [0, n] (i == n carries the grand total), so the consumer of the scan
// is fused into it. One kernel instead of reduce + spine + apply + the consumer's own launch.
// ------------------------------------------------------------------------------------------------
// warp-wide look-back over one column: lane k inspects tile - 1 - k
__device__ __forceinline__ unsigned int lb_exclusive_warp(const unsigned long long* status, unsigned int tile,
                                                          unsigned int epoch, unsigned int* __restrict__ fault) {
    const unsigned int lane = threadIdx.x & 31u;
    unsigned int excl = 0, spins = 0;
    unsigned int t = tile;   // tiles [t, tile) are already summed
    while (t > 0u) {
        const bool inrange = lane < t;
        // before the first tile: an inclusive zero ends the walk
        const unsigned long long w = inrange ? lb_load(status + (t - 1u - lane))
                                             : ((unsigned long long)((epoch << 2) | LB_INCLUSIVE) << 32);
        const unsigned int hi = (unsigned int)(w >> 32);
        const bool ready = (hi >> 2) == epoch;
        const unsigned int notReady = __ballot_sync(0xFFFFFFFFu, !ready);
        const unsigned int usable = notReady ? ((1u << (__ffs(notReady) - 1)) - 1u) : 0xFFFFFFFFu;   // lanes before the first gap
        const unsigned int incl = __ballot_sync(0xFFFFFFFFu, ready && (hi & 3u) == LB_INCLUSIVE) & usable;
        const unsigned int take = incl ? ((2u << (__ffs(incl) - 1)) - 1u) : usable;                  // up to the first inclusive one
        excl += __reduce_add_sync(0xFFFFFFFFu, ((take >> lane) & 1u) ? (unsigned int)w : 0u);
        if (incl) break;
        t -= __popc(usable);
        if (usable == 0u) {
            __nanosleep(64);
            if (++spins > (1u << 22)) {   // bounded wait: a protocol bug becomes an error flag, not a hang
                if (lane == 0) atomicExch(fault, 1u);
                break;
            }
        }
    }
    return excl;
}

template <class Load, class Sink>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_chained(Load load, Sink sink, int n, unsigned long long* __restrict__ status, const unsigned int* __restrict__ epoch_ptr,
               unsigned int* __restrict__ ticket, unsigned int* __restrict__ fault) {
    const unsigned int epoch = *epoch_ptr;
    __shared__ unsigned int sh[9];
    __shared__ unsigned int s_tile, s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);   // tiles in start order
    __syncthreads();
    const unsigned int tile = s_tile;
    const long long base = (long long)tile * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    unsigned int v[SCAN_ITEMS];
    unsigned int sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const long long i = base + k;
        v[k] = (i < n) ? load((int)i) : 0u;
        sum += v[k];
    }
    unsigned int tot;
    const unsigned int ex = block_exclusive_scan_256(sum, sh, &tot);
    if (threadIdx.x == 0) lb_store(status + tile, epoch, tile == 0u ? LB_INCLUSIVE : LB_AGGREGATE, tot);
    if (threadIdx.x < 32) {
        unsigned int excl = 0;
        if (tile != 0u) {
            excl = lb_exclusive_warp(status, tile, epoch, fault);
            if (threadIdx.x == 0) lb_store(status + tile, epoch, LB_INCLUSIVE, excl + tot);
        }
        if (threadIdx.x == 0) s_prefix = excl;
    }
    __syncthreads();
    unsigned int run = s_prefix + ex;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const long long i = base + k;
        if (i <= n) sink((int)i, run, v[k]);
        run += v[k];
    }
}

}  // namespace lpe

namespace lpe {
__global__ void k_epoch_next(unsigned int* __restrict__ epoch) { *epoch += 1u; }   // (scans outside a step: the decomposed upload)
__global__ void k_epoch_set(unsigned int* __restrict__ epoch, unsigned int v) { *epoch = v; }
}
