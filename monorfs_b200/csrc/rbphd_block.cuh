// rbphd_block.cuh -- CTA-wide primitives used by the per-particle kernels:
// reductions, exclusive scans, a bitonic (key,val) sort and a small uniform 3-D cell grid.
// All functions must be called by every thread of the CTA (they contain __syncthreads()).
#pragma once
#include "rbphd_math.cuh"

namespace rbphd {

// Launch shape of the per-particle kernels.  The defaults are the production build; build.py can override
// them (-DRBPHD_BLOCK=... etc.) for occupancy experiments.
#ifndef RBPHD_BLOCK
#define RBPHD_BLOCK 1024
#endif
#ifndef RBPHD_CTAS_PER_SM
#define RBPHD_CTAS_PER_SM 1
#endif
#ifndef RBPHD_GRID_CELLS
#define RBPHD_GRID_CELLS 4864   // with RBPHD_SORT_CAP 6656: 163 KB of shared memory at 500 measurements (see rbphd_kernels.cu)
#endif
#ifndef RBPHD_SORT_BUCKETS
#define RBPHD_SORT_BUCKETS 4096
#endif
constexpr int kBlock = RBPHD_BLOCK;  // threads per CTA of the per-particle kernels
constexpr int kCtasPerSm = RBPHD_CTAS_PER_SM;
constexpr int kWarps = kBlock / 32;
constexpr int kGridMaxDim = 64;      // cells per axis of a cell grid
constexpr int kGridMaxCells = RBPHD_GRID_CELLS;  // cells of a cell grid; the offsets (kGridMaxCells + 1 ints) live in shared memory

struct BlockShared {                 // small fixed scratch in shared memory
    int    warp_i[kWarps + 1];
    double warp_d[2 * kWarps];
    int    carry;
    int    counter[8];
    double bb[6];
};

__device__ __forceinline__ int warp_incl_scan(int v)
{
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (unsigned)o) v += t;
    }
    return v;
}

// exclusive scan of one int per thread; returns prefix, total in *total
__device__ __forceinline__ int block_excl_scan(BlockShared& sh, int v, int* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = warp_incl_scan(v);
    __syncthreads();
    if (lane == 31) sh.warp_i[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < kWarps) ? sh.warp_i[lane] : 0;
        int wi = warp_incl_scan(w);
        if (lane < kWarps) sh.warp_i[lane] = wi - w;
        if (lane == kWarps - 1) sh.warp_i[kWarps] = wi;
    }
    __syncthreads();
    int res = incl - v + sh.warp_i[warp];
    *total = sh.warp_i[kWarps];
    return res;
}

// in-place exclusive scan of an int array of length n (any memory space); returns the total.
// Every thread owns a run of consecutive entries (all loads in flight at once, ONE block-wide scan of the run
// totals); arrays longer than kScanRun * kBlock fall back to kBlock-sized chunks with a carry.
constexpr int kScanRun = 8;
__device__ inline int block_scan_array(BlockShared& sh, int* a, int n)
{
    if (n <= kScanRun * kBlock) {
        const int per = (n + kBlock - 1) / kBlock;
        const int i0 = threadIdx.x * per;
        int v[kScanRun];
        int run = 0;
#pragma unroll
        for (int j = 0; j < kScanRun; j++) {
            v[j] = (j < per && i0 + j < n) ? a[i0 + j] : 0;
            run += v[j];
        }
        int tot;
        int ex = block_excl_scan(sh, run, &tot);
#pragma unroll
        for (int j = 0; j < kScanRun; j++) {
            if (j < per && i0 + j < n) a[i0 + j] = ex;
            ex += v[j];
        }
        __syncthreads();
        return tot;
    }
    int carry = 0;
    for (int base = 0; base < n; base += kBlock) {
        int i = base + threadIdx.x;
        int v = (i < n) ? a[i] : 0;
        int tot;
        int ex = block_excl_scan(sh, v, &tot);
        if (i < n) a[i] = carry + ex;
        carry += tot;
    }
    __syncthreads();
    return carry;
}

// in-place exclusive scans of two int arrays of length n at once (one pass, the two sums packed into
// the halves of a 64-bit add; each total must stay below 2^31); totals in *tot_a / *tot_b
__device__ inline void block_scan_array2(BlockShared& sh, int* a, int* b, int n, int* tot_a, int* tot_b)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long carry = 0;
    unsigned long long* wsum = reinterpret_cast<unsigned long long*>(sh.warp_d);   // kWarps + 1 entries
    if (n <= kScanRun * kBlock) {
        // runs of consecutive entries per thread: one block-wide scan (see block_scan_array)
        const int per = (n + kBlock - 1) / kBlock;
        const int i0 = threadIdx.x * per;
        unsigned long long v[kScanRun];
        unsigned long long run = 0;
#pragma unroll
        for (int j = 0; j < kScanRun; j++) {
            v[j] = (j < per && i0 + j < n) ? ((unsigned long long)(unsigned)a[i0 + j] | ((unsigned long long)(unsigned)b[i0 + j] << 32)) : 0ull;
            run += v[j];
        }
        unsigned long long incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        __syncthreads();
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const unsigned long long w = (lane < kWarps) ? wsum[lane] : 0ull;
            unsigned long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            if (lane < kWarps) wsum[lane] = wi - w;
            if (lane == kWarps - 1) wsum[kWarps] = wi;
        }
        __syncthreads();
        unsigned long long ex = wsum[warp] + incl - run;
#pragma unroll
        for (int j = 0; j < kScanRun; j++) {
            if (j < per && i0 + j < n) { a[i0 + j] = (int)(unsigned)(ex & 0xffffffffull); b[i0 + j] = (int)(unsigned)(ex >> 32); }
            ex += v[j];
        }
        carry = wsum[kWarps];
        __syncthreads();
        *tot_a = (int)(unsigned)(carry & 0xffffffffull);
        *tot_b = (int)(unsigned)(carry >> 32);
        return;
    }
    for (int base = 0; base < n; base += kBlock) {
        const int i = base + threadIdx.x;
        unsigned long long v = 0;
        if (i < n) v = (unsigned long long)(unsigned)a[i] | ((unsigned long long)(unsigned)b[i] << 32);
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        __syncthreads();
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const unsigned long long w = (lane < kWarps) ? wsum[lane] : 0ull;
            unsigned long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            if (lane < kWarps) wsum[lane] = wi - w;
            if (lane == kWarps - 1) wsum[kWarps] = wi;
        }
        __syncthreads();
        const unsigned long long ex = carry + wsum[warp] + incl - v;
        if (i < n) { a[i] = (int)(unsigned)(ex & 0xffffffffull); b[i] = (int)(unsigned)(ex >> 32); }
        carry += wsum[kWarps];
    }
    __syncthreads();
    *tot_a = (int)(unsigned)(carry & 0xffffffffull);
    *tot_b = (int)(unsigned)(carry >> 32);
}

__device__ __forceinline__ double block_sum(BlockShared& sh, double v)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sh.warp_d[warp] = v;
    __syncthreads();
    double s = 0;
    for (int w = 0; w < kWarps; w++) s += sh.warp_d[w];
    return s;
}

__device__ __forceinline__ int block_sum_int(BlockShared& sh, int v)
{
    int tot;
    block_excl_scan(sh, v, &tot);
    return tot;
}

__device__ __forceinline__ int next_pow2(int n)
{
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

// Bitonic sort of (key,val) pairs ascending by (key, then val); n must be a power of two
// (pad with key = ~0ull, val = ~0u).  Works on shared or global arrays.
__device__ inline void block_bitonic_sort(unsigned long long* key, unsigned int* val, int n)
{
    __syncthreads();
    if (n <= kBlock) {
        // one element per thread, held in registers: the stages with partner distance < 32 exchange by warp
        // shuffles (no barrier), only the few wider ones go through the arrays (n = 512: 10 of 45 stages)
        const int t = threadIdx.x;
        unsigned long long kk = (t < n) ? key[t] : ~0ull;
        unsigned int vv = (t < n) ? val[t] : ~0u;
        const int nn = n < 32 ? 32 : n;     // lanes past n in the last warp carry padding (largest key)
        for (int k = 2; k <= nn; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                unsigned long long ko;
                unsigned int vo;
                if (j >= 32) {
                    __syncthreads();            // everyone has read its partner of the previous wide stage
                    if (t < n) { key[t] = kk; val[t] = vv; }
                    __syncthreads();
                    ko = (t < n) ? key[t ^ j] : ~0ull;
                    vo = (t < n) ? val[t ^ j] : ~0u;
                }
                else {
                    ko = __shfl_xor_sync(0xffffffffu, kk, j);
                    vo = __shfl_xor_sync(0xffffffffu, vv, j);
                }
                const bool asc = ((t & k) == 0);
                const bool lower = ((t & j) == 0);
                const bool mine_gt = (kk > ko) || (kk == ko && vv > vo);
                // the lower index of the pair keeps the smaller element in an ascending run, the larger otherwise
                const bool take_other = (lower == asc) ? mine_gt : !mine_gt && !(kk == ko && vv == vo);
                if (take_other) { kk = ko; vv = vo; }
            }
        }
        __syncthreads();
        if (t < n) { key[t] = kk; val[t] = vv; }
        __syncthreads();
        return;
    }
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1, lj = 31 - __clz(k >> 1); j > 0; j >>= 1, lj--) {
            for (int t = threadIdx.x; t < (n >> 1); t += kBlock) {
                int i = ((t >> lj) << (lj + 1)) + (t & (j - 1));
                int p = i + j;
                bool asc = ((i & k) == 0);
                unsigned long long ki = key[i], kp = key[p];
                unsigned int vi = val[i], vp = val[p];
                bool gt = (ki > kp) || (ki == kp && vi > vp);
                if (gt == asc) {
                    key[i] = kp; key[p] = ki;
                    val[i] = vp; val[p] = vi;
                }
            }
            __syncthreads();
        }
    }
}

// lanes of the warp whose 8-bit digit equals this lane's (invalid lanes match nobody): eight ballots, which
// issue at full rate, instead of MATCH.ANY
__device__ __forceinline__ unsigned warp_match_digit(unsigned d, bool valid)
{
    unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const bool bit = (d >> b) & 1u;
        const unsigned bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
    }
    return valid ? peers : 0u;
}

// Stable LSD radix sort (8-bit digits) of n (key,val) pairs, ascending by key; equal keys keep their input
// order.  (k0,v0) holds the input, (k1,v1) is scratch of the same length (any memory space); the return value
// says which of the two holds the result.  Every warp owns a contiguous chunk of the input, counts its digits
// into its own 256-bin histogram (whist: kWarps*256 ints of shared memory; dbase: 256 ints) and ranks equal
// digits inside a 32-element row with match_any, so no atomics are needed and the order is deterministic.
// Digits on which all keys agree are skipped (weights span a few binades: the leading byte usually does).
__device__ inline int block_radix_sort(BlockShared& sh, unsigned long long* k0, unsigned int* v0,
                                       unsigned long long* k1, unsigned int* v1, int n, int* whist, int* dbase)
{

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long o = 0, a = ~0ull;
    for (int e = threadIdx.x; e < n; e += kBlock) { const unsigned long long k = k0[e]; o |= k; a &= k; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { o |= __shfl_xor_sync(0xffffffffu, o, d); a &= __shfl_xor_sync(0xffffffffu, a, d); }
    __syncthreads();
    if (lane == 0) { sh.warp_d[warp] = __longlong_as_double((long long)o); sh.warp_d[kWarps + warp] = __longlong_as_double((long long)a); }
    __syncthreads();
    o = 0; a = ~0ull;
    for (int w = 0; w < kWarps; w++) {
        o |= (unsigned long long)__double_as_longlong(sh.warp_d[w]);
        a &= (unsigned long long)__double_as_longlong(sh.warp_d[kWarps + w]);
    }
    const unsigned long long diff = (n > 1) ? (o ^ a) : 0ull;
    const int C = (((n + kWarps - 1) / kWarps) + 31) & ~31;
    const int beg = min(n, warp * C), end = min(n, beg + C);
    int* wh = whist + warp * 256;
    const unsigned lt = (1u << lane) - 1u;
    int cur = 0;
    for (int shift = 0; shift < 64; shift += 8) {
        if (((diff >> shift) & 255ull) == 0) continue;
        const unsigned long long* ks = cur ? k1 : k0;
        const unsigned int* vs = cur ? v1 : v0;
        unsigned long long* kd = cur ? k0 : k1;
        unsigned int* vd = cur ? v0 : v1;
        __syncthreads();
        for (int b = threadIdx.x; b < kWarps * 256; b += kBlock) whist[b] = 0;
        __syncthreads();
        // rows of 32 keys, kRows rows loaded ahead of their use (the source may be global memory)
        constexpr int kRows = 8;
        for (int base = beg; base < end; base += 32 * kRows) {
            unsigned long long kk[kRows];
#pragma unroll
            for (int r = 0; r < kRows; r++) { const int e = base + 32 * r + lane; kk[r] = (e < end) ? ks[e] : 0ull; }
#pragma unroll
            for (int r = 0; r < kRows; r++) {
                const int e = base + 32 * r + lane;
                const bool valid = e < end;
                const unsigned d = (unsigned)((kk[r] >> shift) & 255ull);
                const unsigned peers = warp_match_digit(d, valid);
                if (valid && lane == __ffs(peers) - 1) wh[d] += __popc(peers);
                __syncwarp();
            }
        }
        __syncthreads();
            int tot = 0;   // digit-major offsets: all warps' counts of digit t, in warp order
        if (threadIdx.x < 256)
#pragma unroll
            for (int w = 0; w < kWarps; w++) { const int c = whist[w * 256 + threadIdx.x]; whist[w * 256 + threadIdx.x] = tot; tot += c; }
        int total;
        const int basep = block_excl_scan(sh, tot, &total);
        if (threadIdx.x < 256) dbase[threadIdx.x] = basep;
        __syncthreads();
            constexpr int kRowsS = 4;
        for (int base = beg; base < end; base += 32 * kRowsS) {
            unsigned long long kk[kRowsS];
            unsigned int vv[kRowsS];
#pragma unroll
            for (int r = 0; r < kRowsS; r++) {
                const int e = base + 32 * r + lane;
                kk[r] = 0; vv[r] = 0;
                if (e < end) { kk[r] = ks[e]; vv[r] = vs[e]; }
            }
#pragma unroll
            for (int r = 0; r < kRowsS; r++) {
                const int e = base + 32 * r + lane;
                const bool valid = e < end;
                const unsigned d = (unsigned)((kk[r] >> shift) & 255ull);
                const unsigned peers = warp_match_digit(d, valid);
                const int off = valid ? wh[d] : 0;
                __syncwarp();
                if (valid && lane == __ffs(peers) - 1) wh[d] = off + __popc(peers);
                __syncwarp();
                if (valid) { const int pos = dbase[d] + off + __popc(peers & lt); kd[pos] = kk[r]; vd[pos] = vv[r]; }
            }
        }
        cur ^= 1;
        __syncthreads();
        }
    __syncthreads();
    return cur;
}

// Bucket sort of n (key,val) pairs, ascending by (key, then val); vals must be distinct.  One pass spreads the
// keys over kSortBuckets equal slices of [min key, max key] (histogram and cursors: shared-memory atomics),
// then every bucket is finished on its own: up to kBucketThreadMax entries by one thread (insertion sort), up
// to kBucketWarpMax by one warp (rank sort held in registers), longer ones (a group of equal weights, e.g. the
// frame's births) by the whole CTA (rank sort through tmpk/tmpv).  Far cheaper than a full radix or bitonic
// sort when the keys are spread out, which weights are.  Returns false -- leaving (kin,vin) untouched -- when
// one bucket is so long that its O(len^2) rank sort would cost more than the radix sort the caller then uses.
// (kout,vout): n entries, ideally shared memory; hist: kSortBuckets + 1 ints of shared memory;
// biglist: kBigBuckets + 2 ints of shared memory (buckets longer than kBucketThreadMax); wscratch: when
// (kout,vout) are in global memory, kWarps * 1.5 * kBucketWarpMax 64-bit words of shared memory.
constexpr int kSortBuckets = RBPHD_SORT_BUCKETS;
constexpr int kBucketThreadMax = 6;
constexpr int kBucketWarpMax = 128;
constexpr int kBigBuckets = kSortBuckets + 64;   // every bucket can be listed (medium ones in front, long ones at the back)
constexpr int kLongBuckets = 64;    // share of the list kept for buckets longer than kBucketWarpMax
constexpr int kInPlaceRows = 8;     // in-place sorts (kin == kout) hold n <= kInPlaceRows * kBlock entries

__device__ inline bool block_bucket_sort(BlockShared& sh, const unsigned long long* kin, const unsigned int* vin,
                                         unsigned long long* kout, unsigned int* vout, int n, int* hist, int* biglist,
                                         unsigned long long* tmpk, unsigned int* tmpv,
                                         unsigned long long* wscratch = nullptr)
{

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (kin == kout && n > kInPlaceRows * kBlock) return false;
    unsigned long long lo = ~0ull, hi = 0ull;
    for (int e = threadIdx.x; e < n; e += kBlock) { const unsigned long long k = kin[e]; lo = min(lo, k); hi = max(hi, k); }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    __syncthreads();
    if (lane == 0) { sh.warp_d[warp] = __longlong_as_double((long long)lo); sh.warp_d[kWarps + warp] = __longlong_as_double((long long)hi); }
    for (int b = threadIdx.x; b <= kSortBuckets; b += kBlock) hist[b] = 0;
    if (threadIdx.x == 0) { biglist[kBigBuckets] = 0; biglist[kBigBuckets + 1] = 0; }
    __syncthreads();
    for (int w = 0; w < kWarps; w++) {
        lo = min(lo, (unsigned long long)__double_as_longlong(sh.warp_d[w]));
        hi = max(hi, (unsigned long long)__double_as_longlong(sh.warp_d[kWarps + w]));
    }
    // Bucket boundaries follow the key distribution: a coarse histogram over 256 equal slices of the key range
    // first, then every slice gets fine buckets in proportion to its count (weights are roughly uniform in
    // value, so equal slices of the bit pattern would put most keys of a large map into one binade's buckets)
    const unsigned long long range = hi - lo;
    const int shift0 = max(0, (64 - __clzll((long long)range)) - 8);   // (key - lo) >> shift0 < 256
    const int sh2 = max(0, shift0 - 20);                               // offset inside a slice, top 20 bits
    int* coarse = biglist;          // 256 counts           (the list itself is not in use yet)
    int* nbk = biglist + 256;       // 256 fine buckets per slice
    int* cbase = biglist + 512;     // 256 first fine bucket of a slice
    if (threadIdx.x < 256) coarse[threadIdx.x] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < n; e += kBlock) atomicAdd(&coarse[(int)((kin[e] - lo) >> shift0)], 1);
    __syncthreads();
    {
        int mine = 0;
        if (threadIdx.x < 256) {
            mine = (int)(((long long)coarse[threadIdx.x] * (kSortBuckets - 512) + n / 2) / max(n, 1));
            mine = max(mine, 1);
            nbk[threadIdx.x] = mine;
        }
        int tot;
        const int ex = block_excl_scan(sh, mine, &tot);   // tot <= kSortBuckets - 512 + 128 + 256
        if (threadIdx.x < 256) cbase[threadIdx.x] = ex;
        __syncthreads();
    }
    auto bucket_of = [&](unsigned long long k) {
        const unsigned long long x = k - lo;
        const int cb = (int)(x >> shift0);
        const unsigned long long off = x - ((unsigned long long)cb << shift0);
        return cbase[cb] + (int)(((off >> sh2) * (unsigned long long)nbk[cb]) >> (shift0 - sh2));
    };
    for (int e = threadIdx.x; e < n; e += kBlock) atomicAdd(&hist[bucket_of(kin[e])], 1);
    __syncthreads();
    int longest = 0, nlong_mine = 0;
    for (int b = threadIdx.x; b < kSortBuckets; b += kBlock) {
        longest = max(longest, hist[b]);
        nlong_mine += (hist[b] > kBucketWarpMax) ? 1 : 0;
    }
    if (__syncthreads_or((long long)longest * longest > 600ll * n && longest > kBucketWarpMax)) return false;
    if (__syncthreads_count(nlong_mine > 0) > kLongBuckets / 4) {   // (an upper bound is enough: <= 4 buckets per thread)
        if (block_sum_int(sh, nlong_mine) > kLongBuckets) return false;
    }
    block_scan_array(sh, hist, kSortBuckets + 1);   // hist[b] = start of bucket b; used as the scatter cursor
    if (kin == kout) {
        // in place (the input already sits in the shared-memory buffer): every thread takes its entries into
        // registers, a barrier, then the scatter
        unsigned long long kr[kInPlaceRows];
        unsigned int vr[kInPlaceRows];
#pragma unroll
        for (int r = 0; r < kInPlaceRows; r++) {
            const int e = threadIdx.x + r * kBlock;
            kr[r] = 0; vr[r] = 0;
            if (e < n) { kr[r] = kin[e]; vr[r] = vin[e]; }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kInPlaceRows; r++) {
            const int e = threadIdx.x + r * kBlock;
            if (e < n) {
                const int pos = atomicAdd(&hist[bucket_of(kr[r])], 1);
                kout[pos] = kr[r];
                vout[pos] = vr[r];
            }
        }
    }
    else {
        for (int e = threadIdx.x; e < n; e += kBlock) {
            const unsigned long long k = kin[e];
            const int pos = atomicAdd(&hist[bucket_of(k)], 1);
            kout[pos] = k;
            vout[pos] = vin[e];
        }
    }
    __syncthreads();
    // after the scatter hist[b] = end of bucket b = start of bucket b + 1
    for (int b = threadIdx.x; b < kSortBuckets; b += kBlock) {
        const int beg = (b == 0) ? 0 : hist[b - 1], end = hist[b];
        const int len = end - beg;
        if (len <= 1) continue;
        if (len > kBucketThreadMax) {
            // medium buckets are listed from the front, long ones from the back (fixed share) of the same list
            if (len <= kBucketWarpMax) {
                const int slot = atomicAdd(&biglist[kBigBuckets], 1);
                if (slot < kBigBuckets - kLongBuckets) { biglist[slot] = b; continue; }
            }
            else {
                const int slot = atomicAdd(&biglist[kBigBuckets + 1], 1);
                if (slot < kLongBuckets) { biglist[kBigBuckets - 1 - slot] = b; continue; }
            }
        }
        for (int a = beg + 1; a < end; a++) {   // (also the overflow of the list: slow, still correct)
            const unsigned long long k = kout[a];
            const unsigned int v = vout[a];
            int q = a - 1;
            while (q >= beg && (kout[q] > k || (kout[q] == k && vout[q] > v))) { kout[q + 1] = kout[q]; vout[q + 1] = vout[q]; q--; }
            kout[q + 1] = k;
            vout[q + 1] = v;
        }
    }
    __syncthreads();
    const int nlong = min(biglist[kBigBuckets + 1], kLongBuckets);
    const int nmed = min(biglist[kBigBuckets], kBigBuckets - kLongBuckets);
    for (int t = warp; t < nmed; t += kWarps) {
        const int b = biglist[t];
        const int beg = (b == 0) ? 0 : hist[b - 1], end = hist[b];
        constexpr int R = kBucketWarpMax / 32;
        unsigned long long k[R];
        unsigned int v[R];
        int rank[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int a = beg + lane + 32 * r;
            k[r] = ~0ull; v[r] = ~0u; rank[r] = 0;
            if (a < end) { k[r] = kout[a]; v[r] = vout[a]; }
        }
        if (wscratch) {   // (kout,vout) live in global memory: rank against a shared-memory copy of the bucket
            unsigned long long* wk = wscratch + (size_t)warp * (kBucketWarpMax + kBucketWarpMax / 2);
            unsigned int* wv = reinterpret_cast<unsigned int*>(wk + kBucketWarpMax);
#pragma unroll
            for (int r = 0; r < R; r++)
                if (beg + lane + 32 * r < end) { wk[lane + 32 * r] = k[r]; wv[lane + 32 * r] = v[r]; }
            __syncwarp();
            const int len = end - beg;
            for (int q = 0; q < len; q++) {
                const unsigned long long kq = wk[q];
                const unsigned int vq = wv[q];
#pragma unroll
                for (int r = 0; r < R; r++) rank[r] += (kq < k[r] || (kq == k[r] && vq < v[r])) ? 1 : 0;
            }
        }
        else {
            for (int q = beg; q < end; q++) {
                const unsigned long long kq = kout[q];
                const unsigned int vq = vout[q];
#pragma unroll
                for (int r = 0; r < R; r++) rank[r] += (kq < k[r] || (kq == k[r] && vq < v[r])) ? 1 : 0;
            }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < R; r++)
            if (beg + lane + 32 * r < end) { kout[beg + rank[r]] = k[r]; vout[beg + rank[r]] = v[r]; }
        __syncwarp();
    }
    if (nlong > 0) __syncthreads();   // the long-bucket staging reuses the warps' scratch
    for (int t = 0; t < nlong; t++) {   // the whole CTA on one long bucket
        const int b = biglist[kBigBuckets - 1 - t];
        const int beg = (b == 0) ? 0 : hist[b - 1], end = hist[b];
        const int len = end - beg;
        // rank against a shared-memory copy when (kout,vout) are in global memory (the scratch then spans the
        // idle shared sort buffer: kWarps * 1.5 * kBucketWarpMax words >= 1.5 * len is checked)
        const bool staged = wscratch && (size_t)len * 3 <= (size_t)kWarps * 3 * kBucketWarpMax;
        const unsigned long long* rk = kout + beg;
        const unsigned int* rv = vout + beg;
        if (staged) {
            unsigned long long* wk = wscratch;
            unsigned int* wv = reinterpret_cast<unsigned int*>(wscratch + len);
            for (int a = threadIdx.x; a < len; a += kBlock) { wk[a] = kout[beg + a]; wv[a] = vout[beg + a]; }
            __syncthreads();
            rk = wk; rv = wv;
        }
        for (int a = threadIdx.x; a < len; a += kBlock) {
            const unsigned long long k = rk[a];
            const unsigned int v = rv[a];
            int rank = 0;
            for (int q = 0; q < len; q++) {
                const unsigned long long kq = rk[q];
                rank += (kq < k || (kq == k && rv[q] < v)) ? 1 : 0;
            }
            tmpk[beg + rank] = k;
            tmpv[beg + rank] = v;
        }
        __syncthreads();
        for (int a = beg + threadIdx.x; a < end; a += kBlock) { kout[a] = tmpk[a]; vout[a] = tmpv[a]; }
        __syncthreads();
    }
    __syncthreads();
    return true;
}

// Keep the `want` smallest (key, then val) of n candidates when n exceeds what a later sort can hold:
// radix-select the want-th key (8 bits per pass, histogram in shared memory) and copy every candidate
// with key <= it to (okey, oval).  Returns the number copied (>= want; more only on ties), or -1 if it
// would exceed ocap (nothing useful copied).  hist: 256 ints of shared memory, scratch2: 2 ints.
__device__ inline int block_select_smallest(const unsigned long long* key, const unsigned int* val, int n,
                                            int want, unsigned long long* okey, unsigned int* oval, int ocap,
                                            int* hist, int* scratch2, unsigned long long* selkey)
{
    unsigned long long prefix = 0;
    int remaining = want;
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        for (int b = threadIdx.x; b < 256; b += kBlock) hist[b] = 0;
        __syncthreads();
        for (int base = 0; base < n; base += kBlock) {
            const int e = base + threadIdx.x;
            unsigned digit = 0xffffffffu;   // inactive lanes share one pseudo digit
            if (e < n) {
                const unsigned long long kk = key[e];
                if (pass == 0 || (kk >> (shift + 8)) == (prefix >> (shift + 8))) digit = (unsigned)((kk >> shift) & 255);
            }
            // keys share their leading bytes (weights span a few binades): aggregate equal digits per warp
            const unsigned peers = __match_any_sync(0xffffffffu, digit);
            if (digit != 0xffffffffu && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1))
                atomicAdd(&hist[(int)digit], __popc(peers));
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int cum = 0, d = 0;
            for (; d < 256; d++) { if (cum + hist[d] >= remaining) break; cum += hist[d]; }
            if (d > 255) d = 255;
            scratch2[0] = remaining - cum;
            *selkey = prefix | ((unsigned long long)d << shift);
        }
        __syncthreads();
        remaining = scratch2[0];
        prefix = *selkey;
        __syncthreads();
    }
    if (threadIdx.x == 0) scratch2[1] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < n; e += kBlock) {
        if (key[e] <= prefix) {
            int idx = atomicAdd(&scratch2[1], 1);
            if (idx < ocap) { okey[idx] = key[e]; oval[idx] = val[e]; }
        }
    }
    __syncthreads();
    int cnt = scratch2[1];
    __syncthreads();
    return (cnt <= ocap) ? cnt : -1;
}

// first index in sorted key[0..n) with key >= k
__device__ __forceinline__ int lower_bound_u64(const unsigned long long* key, int n, unsigned long long k)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (key[mid] < k) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------------
// Uniform 3-D cell grid over n points given as three arrays.  Cell offsets live in shared memory
// (start[ncell+1]); the point indices of each cell, in no particular order, in `items` (any memory).
// ---------------------------------------------------------------------------------------------
// order-preserving 64-bit encoding of a double (for integer atomicMin / atomicMax)
__device__ __forceinline__ unsigned long long ordered_key(double x)
{
    const long long b = __double_as_longlong(x);
    return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
}
__device__ __forceinline__ double ordered_value(unsigned long long k)
{
    const long long b = (k & 0x8000000000000000ull) ? (long long)(k ^ 0x8000000000000000ull) : (long long)~k;
    return __longlong_as_double(b);
}

struct CellGrid {
    double org[3];
    double inv[3];     // 1 / cell size per axis
    double hi[3];      // bounding box maximum
    int    dim[3];
    int    ncell;
    int    n;
};

__device__ __forceinline__ int grid_coord(const CellGrid& g, int a, double x)
{
    // round down and clamp in the integer domain (the conversion saturates; NaN gives 0)
    return max(0, min(g.dim[a] - 1, __double2int_rd((x - g.org[a]) * g.inv[a])));
}
__device__ __forceinline__ int grid_cell(const CellGrid& g, double x, double y, double z)
{
    return (grid_coord(g, 2, z) * g.dim[1] + grid_coord(g, 1, y)) * g.dim[0] + grid_coord(g, 0, x);
}

// Build.  mincell = smallest useful cell edge (a typical query radius).  start must hold
// maxcells + 1 ints of shared memory (maxcells <= kCellsBound); items n ints.
template <int kCellsBound = kGridMaxCells>
__device__ inline void grid_build(BlockShared& sh, CellGrid& g, int* start, int* items, const double* px,
                                  const double* py, const double* pz, int n, double mincell0,
                                  double mincell1, double mincell2, int maxcells = kGridMaxCells)
{
    const double mincell3[3] = {mincell0, mincell1, mincell2};
    // bounding box
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = threadIdx.x; i < n; i += kBlock) {
        double x = px[i], y = py[i], z = pz[i];
        lo[0] = fmin(lo[0], x); hi[0] = fmax(hi[0], x);
        lo[1] = fmin(lo[1], y); hi[1] = fmax(hi[1], y);
        lo[2] = fmin(lo[2], z); hi[2] = fmax(hi[2], z);
    }
    // block-wide bounding box: warp shuffles, then one shared-memory atomic per warp and bound on an order-preserving
    // 64-bit encoding of the doubles (two barriers instead of two per axis)
    const int lane = threadIdx.x & 31;
    unsigned long long* bbk = reinterpret_cast<unsigned long long*>(sh.bb);
    if (threadIdx.x < 6) bbk[threadIdx.x] = (threadIdx.x < 3) ? ordered_key(INFINITY) : ordered_key(-INFINITY);
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fmin(lo[a], __shfl_down_sync(0xffffffffu, lo[a], o));
            hi[a] = fmax(hi[a], __shfl_down_sync(0xffffffffu, hi[a], o));
        }
        if (lane == 0) { atomicMin(&bbk[a], ordered_key(lo[a])); atomicMax(&bbk[3 + a], ordered_key(hi[a])); }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // cell edge per axis = mincell * f with the smallest common f >= 1 that keeps the grid within
        // kGridMaxCells cells (and kGridMaxDim per axis)
        double ext[3], mc[3];
        for (int a = 0; a < 3; a++) {
            double l = ordered_value(bbk[a]), h = ordered_value(bbk[3 + a]);
            if (!(l <= h)) { l = 0; h = 0; }
            g.org[a] = l; g.hi[a] = h;
            ext[a] = g.hi[a] - g.org[a];
            mc[a] = (mincell3[a] > 0 && mincell3[a] < INFINITY) ? mincell3[a] : ext[a] / 16.0;
            if (!(mc[a] > 0)) mc[a] = 1.0;
        }
        // start from the closed-form factor (volume / maxcells, longest axis / kGridMaxDim); the loop only
        // corrects for axes thinner than one cell.  This section runs on one thread: FP64 divisions are the
        // cost, so the per-axis ratios are formed once and scaled by 1 / f.
        double da[3];
        double f = 1.0;
        {
            double vol = 1.0;
            for (int a = 0; a < 3; a++) {
                da[a] = ext[a] / mc[a];
                f = fmax(f, da[a] * (1.0 / (double)kGridMaxDim));
                vol *= fmax(da[a], 1.0);
            }
            f = fmax(f, cbrt(vol / (double)maxcells));
            if (!(f >= 1.0) || isinf(f)) f = 1.0;
        }
        int dsel[3] = {1, 1, 1};
        for (int it = 0; it < 200; it++) {
            const double inv_f = 1.0 / f;
            long cells = 1;
            bool ok = true;
            for (int a = 0; a < 3; a++) {
                double dd = floor(da[a] * inv_f);   // cell edge >= mincell * f
                if (!(dd >= 1.0)) dd = 1.0;
                if (dd > (double)kGridMaxDim) { ok = false; dd = (double)kGridMaxDim; }
                dsel[a] = (int)dd;
                cells *= (long)dd;
            }
            if (ok && cells <= maxcells) break;
            f *= 1.2;
        }
        for (int a = 0; a < 3; a++) {
            const int d = dsel[a];
            // (x - org) * inv < d for every x <= hi
            g.dim[a] = d;
            g.inv[a] = (ext[a] > 0) ? (double)d / (ext[a] * (1.0 + 1e-12) + 1e-300) : 1.0;
        }
        if ((long)g.dim[0] * g.dim[1] * g.dim[2] > maxcells) { g.dim[0] = g.dim[1] = g.dim[2] = (maxcells >= 4096) ? 16 : 4; 
            for (int a = 0; a < 3; a++) { double cs = ext[a] / g.dim[a]; if (!(cs > 0)) cs = 1.0; g.inv[a] = 1.0 / (cs * (1.0 + 1e-12) + 1e-300); } }
    }
    __syncthreads();
    if (threadIdx.x == 0) { g.ncell = g.dim[0] * g.dim[1] * g.dim[2]; g.n = n; }
    __syncthreads();
    const int ncell = g.ncell;
    for (int c = threadIdx.x; c <= ncell; c += kBlock) start[c] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kBlock) atomicAdd(&start[grid_cell(g, px[i], py[i], pz[i])], 1);
    __syncthreads();
    // inclusive scan of the ncell + 1 counts (start[c] = END of cell c): every thread owns a run of consecutive
    // cells, so one block-wide scan of the run totals suffices
    const int per = (ncell + 1 + kBlock - 1) / kBlock;
    const int c0 = threadIdx.x * per;
    {
        int run = 0;
        for (int a = 0; a < per; a++) { const int cc = c0 + a; if (cc <= ncell) run += start[cc]; }
        int tot;
        int prefix = block_excl_scan(sh, run, &tot);
        for (int a = 0; a < per; a++) {
            const int cc = c0 + a;
            if (cc <= ncell) { prefix += start[cc]; start[cc] = prefix; }
        }
    }
    __syncthreads();
    // scatter downwards from the ends: afterwards start[c] is the BEGIN of cell c (and start[ncell] = n, no point
    // maps to that slot)
    for (int i = threadIdx.x; i < n; i += kBlock) {
        int c = grid_cell(g, px[i], py[i], pz[i]);
        int slot = atomicSub(&start[c], 1) - 1;
        items[slot] = i;
    }
    __syncthreads();
    // (The order of the indices inside a cell is whatever the atomics produced.  Every consumer either
    // sorts what it derives from the walk -- gated pairs, merge edges, likelihood edges -- or only
    // accumulates, so no per-cell sort is needed.)
}

// cell range covered by the ball (x,y,z; r); returns false if it misses the bounding box
__device__ __forceinline__ bool grid_range(const CellGrid& g, double x, double y, double z, double r, int* lo,
                                           int* hi)
{
    const double q[3] = {x, y, z};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if (!(q[a] + r >= g.org[a]) || !(q[a] - r <= g.hi[a])) return false;
        lo[a] = grid_coord(g, a, q[a] - r);
        hi[a] = grid_coord(g, a, q[a] + r);
    }
    return true;
}

}  // namespace rbphd
