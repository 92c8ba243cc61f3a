// Radix partition of 32-bit keys: the reusable primitive behind sorted / deferred
// accumulation.
//
// What it replaces: cuburn/code/sort.py:385-520 (`Sorter.sort`: one pass that groups keys by
// the RBITS = 8 bits above `lo_bit`, optionally dropping keys equal to 0xffffffff, built from
// prefix_scan / prefix_sum_condense / prefix_sum_inner / prefix_sum_distribute /
// radix_sort kernels over groups of 8192 keys, sort.py:33-382) and the experiments in
// helpers/sortbench.cu.  The reference's pass is "mildly unstable" (sort.py:27-29), which is
// why its multi-pass sort is marked broken (sort.py:437-441, 462-466).  This one is stable
// -- keys with equal digits keep their order -- so least-significant-digit passes compose
// into a full sort (Sorter.multisort).
//
// Three kernels per pass over groups of GROUP = 8192 keys (one CTA of 256 threads each):
//   k_sort_count    per-group digit histogram (shared-memory ATOMS) -> counts[digit][group]
//   k_scan_*        exclusive prefix sum over counts in digit-major order (three small kernels)
//   k_sort_scatter  the group's keys once more: every warp owns 1024 consecutive keys and
//                   ranks them in order (match.any per 32 keys + a running per-warp digit
//                   count), warps are prefix-summed per digit, and every key goes to
//                   offset[digit][group] + keys of that digit in earlier warps + its rank.
// Scratch: (2^bits * groups + 2 * ceil(that / 1024) + 8) 32-bit words -- n / 8 bytes for
// 8-bit digits, the same as the reference's `dpfxs` (sort.py:424).
#include "cb_common.h"

#define SORT_GROUP 8192
#define SORT_THREADS 256
#define SORT_WARPS (SORT_THREADS / 32)
#define SORT_WARP_KEYS (SORT_GROUP / SORT_WARPS)
#define SORT_MAX_BITS 8
#define SCAN_BLOCK 1024

__device__ __forceinline__ unsigned int sort_digit(unsigned int key, int lo_bit, int bits) {
    return (key >> lo_bit) & ((1u << bits) - 1u);
}

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_count(unsigned int *counts, const unsigned int *keys, unsigned int n, int lo_bit, int bits,
             int ignore_max, unsigned int ngroups) {
    __shared__ unsigned int hist[1 << SORT_MAX_BITS];
    const unsigned int ndig = 1u << bits;
    for (unsigned int d = threadIdx.x; d < ndig; d += SORT_THREADS) hist[d] = 0u;
    __syncthreads();
    const unsigned int base = blockIdx.x * SORT_GROUP;
#pragma unroll 4
    for (unsigned int i = threadIdx.x; i < SORT_GROUP; i += SORT_THREADS) {
        const unsigned int j = base + i;
        if (j < n) {
            const unsigned int k = keys[j];
            if (!(ignore_max && k == 0xffffffffu)) atomicAdd(&hist[sort_digit(k, lo_bit, bits)], 1u);
        }
    }
    __syncthreads();
    for (unsigned int d = threadIdx.x; d < ndig; d += SORT_THREADS)
        counts[d * ngroups + blockIdx.x] = hist[d];
}

// ---- exclusive scan of `len` words, in place: block sums, scan of the sums, add ---------
__device__ __forceinline__ unsigned int block_exclusive_scan(unsigned int v, unsigned int *total) {
    __shared__ unsigned int warp_sums[SCAN_BLOCK / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned int w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        warp_sums[lane] = w;
    }
    __syncthreads();
    const unsigned int before = warp ? warp_sums[warp - 1] : 0u;
    if (total) *total = warp_sums[SCAN_BLOCK / 32 - 1];
    __syncthreads();
    return before + incl - v;
}

__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan_blocks(unsigned int *data, unsigned int *block_sums, unsigned int len) {
    const unsigned int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const unsigned int v = i < len ? data[i] : 0u;
    __shared__ unsigned int total;
    const unsigned int ex = block_exclusive_scan(v, threadIdx.x == 0 ? &total : nullptr);
    if (i < len) data[i] = ex;
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// one CTA walks the block sums (at most a few thousand) chunk by chunk
__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan_sums(unsigned int *block_sums, unsigned int nblocks, unsigned int *grand_total) {
    __shared__ unsigned int total;
    unsigned int carry = 0u;
    for (unsigned int base = 0; base < nblocks; base += SCAN_BLOCK) {
        const unsigned int i = base + threadIdx.x;
        const unsigned int v = i < nblocks ? block_sums[i] : 0u;
        const unsigned int ex = block_exclusive_scan(v, threadIdx.x == 0 ? &total : nullptr);
        if (i < nblocks) block_sums[i] = carry + ex;
        __syncthreads();
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan_add(unsigned int *data, const unsigned int *block_sums, unsigned int len) {
    const unsigned int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    if (i < len) data[i] += block_sums[blockIdx.x];
}

// ---- stable scatter ----------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_scatter(unsigned int *dst, const unsigned int *keys, const unsigned int *offsets,
               unsigned int n, int lo_bit, int bits, int ignore_max, unsigned int ngroups) {
    // running digit counts per warp; afterwards: keys of that digit in earlier warps
    __shared__ unsigned int wcount[SORT_WARPS][1 << SORT_MAX_BITS];
    __shared__ unsigned int goff[1 << SORT_MAX_BITS];
    const unsigned int ndig = 1u << bits;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned int d = threadIdx.x; d < ndig; d += SORT_THREADS) {
        goff[d] = offsets[d * ngroups + blockIdx.x];
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) wcount[w][d] = 0u;
    }
    __syncthreads();

    const unsigned int wbase = blockIdx.x * SORT_GROUP + warp * SORT_WARP_KEYS;
    unsigned int key[SORT_WARP_KEYS / 32];
    unsigned short rank[SORT_WARP_KEYS / 32];
#pragma unroll
    for (int r = 0; r < SORT_WARP_KEYS / 32; r++) {
        const unsigned int j = wbase + r * 32 + lane;
        const bool live = j < n;
        key[r] = live ? keys[j] : 0xffffffffu;
        const bool use = live && !(ignore_max && key[r] == 0xffffffffu);
        // lanes that do not take part get a digit nobody else has
        const unsigned int d = use ? sort_digit(key[r], lo_bit, bits) : (ndig + lane);
        const unsigned int same = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(same) - 1;
        unsigned int first = 0u;
        if (use && lane == leader) {
            first = wcount[warp][d];
            wcount[warp][d] = first + __popc(same);
        }
        first = __shfl_sync(0xffffffffu, first, leader);
        rank[r] = (unsigned short)(first + __popc(same & ((1u << lane) - 1u)));
        if (!use) rank[r] = 0xffffu;
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over the warps, on top of the group's global offset
    for (unsigned int d = threadIdx.x; d < ndig; d += SORT_THREADS) {
        unsigned int run = goff[d];
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            const unsigned int c = wcount[w][d];
            wcount[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_WARP_KEYS / 32; r++)
        if (rank[r] != 0xffffu)
            dst[wcount[warp][sort_digit(key[r], lo_bit, bits)] + rank[r]] = key[r];
}

extern "C" {

int cb_sort_scratch_words(uint64_t max_keys, int bits, uint64_t *words) {
    CB_REQUIRE(words && bits >= 1 && bits <= SORT_MAX_BITS, "1 <= bits <= 8");
    const uint64_t groups = (max_keys + SORT_GROUP - 1) / SORT_GROUP;
    const uint64_t len = (groups ? groups : 1) << bits;
    const uint64_t nblocks = (len + SCAN_BLOCK - 1) / SCAN_BLOCK;
    *words = len + nblocks + 8;
    return CB_OK;
}

int cb_sort_pass(cb_dptr dst, cb_dptr src, uint64_t n, int lo_bit, int bits, int ignore_max,
                 cb_dptr scratch, cb_stream s) {
    CB_REQUIRE(bits >= 1 && bits <= SORT_MAX_BITS && lo_bit >= 0 && lo_bit + bits <= 32,
               "digit must lie inside the 32-bit key, 1 <= bits <= 8");
    CB_REQUIRE(n < (1ull << 32), "at most 2^32 - 1 keys");
    CB_REQUIRE(dst && src && scratch && dst != src, "dst, src, scratch must be distinct buffers");
    const unsigned int groups = (unsigned int)((n + SORT_GROUP - 1) / SORT_GROUP);
    const unsigned int len = (groups ? groups : 1u) << bits;
    const unsigned int nblocks = (len + SCAN_BLOCK - 1) / SCAN_BLOCK;
    unsigned int *counts = cb_ptr<unsigned int>(scratch);
    unsigned int *sums = counts + len;
    unsigned int *total = sums + nblocks;           // total[0] = keys kept
    if (n == 0) {
        CB_CUDA(cudaMemsetAsync(total, 0, 4, cb_cs(s)));
        return CB_OK;
    }
    k_sort_count<<<groups, SORT_THREADS, 0, cb_cs(s)>>>(
        counts, cb_ptr<const unsigned int>(src), (unsigned int)n, lo_bit, bits, ignore_max, groups);
    CB_LAUNCH_CHECK();
    k_scan_blocks<<<nblocks, SCAN_BLOCK, 0, cb_cs(s)>>>(counts, sums, len);
    CB_LAUNCH_CHECK();
    k_scan_sums<<<1, SCAN_BLOCK, 0, cb_cs(s)>>>(sums, nblocks, total);
    CB_LAUNCH_CHECK();
    k_scan_add<<<nblocks, SCAN_BLOCK, 0, cb_cs(s)>>>(counts, sums, len);
    CB_LAUNCH_CHECK();
    k_sort_scatter<<<groups, SORT_THREADS, 0, cb_cs(s)>>>(
        cb_ptr<unsigned int>(dst), cb_ptr<const unsigned int>(src), counts, (unsigned int)n,
        lo_bit, bits, ignore_max, groups);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

}  // extern "C"
