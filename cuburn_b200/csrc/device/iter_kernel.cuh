// The chaos-game kernel skeleton (sm_100a, compiled per genome by NVRTC).
//
// Included at the end of a generated translation unit that defines
//   NSLOTS                number of packed parameter floats per temporal sample
//   void camera_coefs(float &xx, ..., float &yo)   the camera affine's coefficients
//   HAS_FINAL             0/1
//   PARAMS_CONST          1: one parameter block for the whole launch, read from
//                            __constant__ memory (stills: every temporal sample is
//                            identical); parameters become constant-bank operands
//                         0: the block of the unit's temporal sample is staged in
//                            shared memory (motion blur)
//   POINTS                points per thread (1 or 2): a thread carries POINTS trajectories
//                         through every round, all through the xform its warp chose, so
//                         that the choice, the parameter fetches and the loop overhead
//                         are paid once per POINTS samples
//   void chaos_step(float sel, point_set &pt, mwc_st &rng, bool (&vis)[POINTS])
//                         the weighted xform choice + application to all points of the
//                         thread; reads P[slot]; vis[p]: whether the new point is visible
//                         (xform opacity, genome/specs.py:17); pt.last[p] is the index of
//                         the xform the point went through before (only read when XAOS)
//                         and is updated
//   void final_step(point_set &pt, mwc_st &rng)   (if HAS_FINAL)
//   XAOS                  0/1: xform choice depends on the previous xform of the
//                         trajectory (iter.py:32-54,233-257); as in the reference the
//                         choice is then per thread and points are not exchanged
//   HOT_BINS              0/1: samples for the bins listed in iter_args::hot_tags are
//                         accumulated in shared memory (see below)
// and which #includes "iter_params.cuh" *before* those functions so that P exists.
//
// What it computes is the reference `iter` kernel (cuburn/code/iter.py:157-418):
// per-warp xform choice, the variations, an inter-warp point exchange, the
// optional final xform on a copy, camera affine, round-to-nearest-even binning
// with the unsigned bounds test against (astride, aheight), the dithered
// palette index, and accumulation.  How it does it is different:
//   * persistent CTAs walk "units" of 256 threads x UNIT_ROUNDS rounds; the unit
//     index selects the temporal sample (params block) and palette row, so one
//     launch renders a whole frame and every RNG stream / trajectory belongs to
//     one thread for the frame (no ring buffer, no block-slot atomics)
//   * accumulation is a single 16-byte red.global.add.v4.f32 per sample straight
//     into the float4 histogram, which is L2-resident on B200 up to 1080p+: integer
//     palette levels and a count, kept exact by a sweep that moves full bins aside
//     (spill_sweep below) instead of the reference's per-sample overflow check
//     (iter.py:359-407); hotspot thinning is replaced by HOT_BINS
//   * the point exchange is double-buffered so a round costs one barrier, and
//     the unit's palette row is staged in shared memory so the per-sample colour
//     fetch is an LDS.128 instead of a scattered global load
#pragma once

struct iter_dims { int width, height, awidth, aheight, astride; };

struct iter_args {
    float4 *hist;
    mwc_st *seeds;
    float4 *points;
    const float *params;
    const float4 *palette;
    iter_dims dim;
    int param_stride;
    int nts;
    int pal_rows;
    int fuse_rounds;
    int swizzle_bins;       // bins [0, swizzle_bins) use the slice-balancing layout
    unsigned long long first_sample;
    unsigned long long nsamples;
    unsigned long long total_samples;
    unsigned long long *cells;                 // ACC_PACKED: u64 [aheight][astride]
    const unsigned long long *palette_packed;  // u64 [pal_rows][256] (ACC_PACKED, PAL_COMPACT)
    const int *hot_tags;                       // HOT_BINS: int [HOT_SLOTS + 1]: bin or -1, then the hash multiplier
    int first_round;        // phase of the exchange permutation to start from
    float4 *spill;          // float4 path: where swept bins are moved to (layout of hist), or 0
    int spill_bins;         // bins of hist every unit examines
    float spill_count;      // a bin holding at least this many samples is moved
    unsigned int *tickets;  // [0]: next unit to hand out (dynamic), [1]: sweep windows begun
    int dynamic;            // 0: CTA b runs units b, b + grid, ...; 1: units are claimed
};

#ifndef ACC_PACKED
#define ACC_PACKED 0
#endif
#ifndef XAOS
#define XAOS 0
#endif
#ifndef HOT_BINS
#define HOT_BINS 0
#endif
// PAL_COMPACT: the unit's palette row is staged as 8-bit levels (Y << 16 | U << 8 | V, one
// 32-bit word per entry) instead of float4: the per-sample colour fetch becomes an LDS.32
// (1 KB table, ~3 shared-memory wavefronts per warp instead of ~9 for the 16-byte
// gather) followed by three byte-to-float conversions.  The float4 table holds exactly
// level * (1 / 255) (cb_interp_palette), so the contribution is bit-identical.  Pays
// where the kernel is bound by L1TEX wavefronts (motion blur: parameters in shared
// memory), costs issue slots where it is not (profiles/r02_iter_variants.md).
// Since the kernel became L1TEX-bound at 86 % (dynamic units, two points per thread) it also
// pays for the still variant with two points: G6F 21.47 -> 21.03 ms; with one point it still
// costs (21.59 -> 21.9), profiles/r02_iter_variants_dyn.txt.
#ifndef PAL_COMPACT
#define PAL_COMPACT (HOT_BINS || !PARAMS_CONST || POINTS == 2)
#endif
#if ACC_PACKED
#undef PAL_COMPACT
#define PAL_COMPACT 0
#endif
// Slots of the per-CTA hot-bin table: a direct-mapped multiplicative hash of the bin
// index; cb_hot_scan picks, per frame, the multiplier (out of eight) that places most
// hot bins and stores it behind the tags.
#define HOT_SLOTS 1024
#define HOT_HASH_SHIFT 22
// 20 KB of cells and tags per CTA: six CTAs per SM instead of eight.  Flames with hot
// bins are bound by atomics, not by issue slots (profiles/r02_hot_bins.md).
#define HOT_MIN_CTAS 6
// Units between two flushes of the table into the histogram: keeps the 32-bit level
// sums of a cell (<= 255 x samples) and the float conversion of its count exact.
#define HOT_FLUSH_UNITS 32

#define ITER_WARPS (ITER_THREADS / 32)
#define UNIT_SAMPLES (ITER_THREADS * UNIT_ROUNDS)
#if XAOS && POINTS != 1
#error "xaos chooses per point: POINTS must be 1"
#endif
// iter_args::points holds POINTS planes of this many trajectories (= RNG streams)
#define POINT_STRIDE 262144

// RED_POLICY 1: the per-sample reductions carry an L2 "evict last" priority.  A scattered
// reduction stream into a grid that is a large fraction of L2 runs up to 46 % faster with it
// (tools/red_policy_microbench.py: uniform addresses, 100 MiB grid 1.24e11 -> 1.82e11/s,
// 120 MiB 0.99e11 -> 1.45e11/s; nothing changes for grids up to half of L2): without the
// hint the cache keeps writing dirty sectors back that the next reduction dirties again.
#ifndef RED_POLICY
#define RED_POLICY 0
#endif
__device__ __forceinline__ unsigned long long red_policy() {
    unsigned long long pol = 0ull;
#if RED_POLICY
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
#endif
    return pol;
}

__device__ __forceinline__ void red_add_f32x4(float4 *addr, float4 v) {
#if RED_POLICY
    asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(red_policy()) : "memory");
#else
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
#endif
}

// ---- exact sums in a float4 histogram -------------------------------------------------
// The float4 path adds *integer-valued* floats: 8-bit palette levels and a count of 1 (the
// reference's addends, interp.py:428-429, iter.py:334-351; the division by 255 of
// iter.py:395-406 happens once per bin in cb_hist_finish).  Such a sum is exact while it
// stays below 2^24, i.e. for 65793 samples of the brightest level, and a float add has no
// way of telling that it is about to round.  So the grid is *swept*: every unit looks at
// spill_bins bins (a window that advances with the unit index and wraps around the grid)
// and moves the ones holding >= spill_count samples -- one 128-bit exchange with zero and
// one reduction into the `spill` grid, which receives a bin's sum in chunks of < 2^24 and
// therefore rounds each chunk by at most 2^-24 of the total.  Windows are numbered by a
// ticket counter in the order in which units *begin* (CTAs run at different speeds, the
// unit index is not a clock), and the host sizes them so that the grid is swept once per
// ~2^26 samples of the launch: a bin would have to collect more than 65793 - spill_count
// samples between two sweeps to lose a bit (1/1088 of all samples; bins much hotter than
// that are what HOT_BINS is for).  The reference spills its 10-bit counters at 512 the same
// way (iter.py:359-407), driven by the value its atom.add returns; a sweep costs a
// 4-byte copy per bin and sweep instead of an atomic round trip per sample.
__device__ __forceinline__ float4 exch_zero_f32x4(float4 *addr) {
    unsigned long long lo, hi;
    asm volatile("{\n\t.reg .b128 z, o;\n\tmov.b128 z, {%3, %3};\n\t"
                 "atom.global.exch.b128 o, [%2], z;\n\tmov.b128 {%0, %1}, o;\n\t}"
                 : "=l"(lo), "=l"(hi) : "l"(addr), "l"(0ull) : "memory");
    return make_float4(__uint_as_float((unsigned int)lo), __uint_as_float((unsigned int)(lo >> 32)),
                       __uint_as_float((unsigned int)hi), __uint_as_float((unsigned int)(hi >> 32)));
}

// Split in two so that nobody waits for the counts: at the start of a unit every thread
// asks for the counts of its bins of the window with 4-byte asynchronous copies into
// shared memory (LDGSTS: no register, no scoreboard), at the end of the unit -- 128 rounds
// later -- it looks at them.  A count read a unit ago only errs on the low side.
#define SWEEP_MAX_WINDOW (4 * ITER_THREADS)

__device__ __forceinline__ int sweep_bin(int base, int i, int nbins) {
    const int b = base + i;
    return b >= nbins ? b - nbins : b;
}

// (Streaming / last-use / evict-first hints on these loads were measured at 4K, where a sweep
// drags the whole grid through L2: no difference, profiles/r02_schedule_and_sweep.md.)
__device__ __forceinline__ void sweep_begin(float *counts, const float4 *hist, int nbins, int base,
                                            int window, int tid) {
#pragma unroll
    for (int k = 0; k < SWEEP_MAX_WINDOW / ITER_THREADS; k++) {
        const int i = tid + k * ITER_THREADS;
        if (i < window)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;"
                         :: "r"((unsigned int)__cvta_generic_to_shared(counts + i)),
                            "l"(&hist[sweep_bin(base, i, nbins)].w) : "memory");
    }
}

__device__ __forceinline__ void sweep_end(const float *counts, float4 *hist, float4 *spill,
                                          int nbins, int base, int window, float count, int tid) {
    asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
    for (int k = 0; k < SWEEP_MAX_WINDOW / ITER_THREADS; k++) {
        const int i = tid + k * ITER_THREADS;
        if (i < window && counts[i] >= count) {
            const int b = sweep_bin(base, i, nbins);
            const float4 v = exch_zero_f32x4(hist + b);
            red_add_f32x4(spill + b, v);
        }
    }
}

// ---- packed accumulation (grids far larger than L2) ---------------------------------
// The reference's cell format (cuburn/code/iter.py:334-407, interp.py:428-429): one
// u64 per bin, count:10 | sum Y:18 | sum U:18 | sum V:18 of 8-bit palette levels, good
// for 1023 samples.  Half the bytes per bin of the float4 histogram, so twice as many
// bins stay L2-resident.  The reference lets 3 % of the warps check the returned count
// and drain the cell once it has passed 512; between that check and the drain a hot
// bin keeps filling (one L2 round trip, hundreds of adds at B200 rates), and a warp
// that has converged onto one bin adds 32 at a time.  Here, in every round ONE lane of
// each warp (rotating with the warp's random word) drains unconditionally: it swaps
// the cell for zero and moves what it held, plus its own sample, into the float4
// histogram.  Every bin is therefore drained after ~32 adds on average whatever its
// rate; running past 1023 would take 1023 consecutive undrained adds, (31/32)^1023.
// cb_flush_packed adds what is left in the cells at the end (flush_atom,
// iter.py:420-479).
__device__ __forceinline__ void accumulate_packed(unsigned long long *cell, float4 *hist,
                                                  unsigned long long v, bool drain) {
    if (!drain) {
#if RED_POLICY
        asm volatile("red.global.add.L2::cache_hint.u64 [%0], %1, %2;"
                     :: "l"(cell), "l"(v), "l"(red_policy()) : "memory");
#else
        asm volatile("red.global.add.u64 [%0], %1;" :: "l"(cell), "l"(v) : "memory");
#endif
        return;
    }
    unsigned long long old = atomicExch(cell, 0ull);
    float cnt = (float)((unsigned int)(old >> 54) + 1u);
    float y = (float)((unsigned int)((old >> 36) & 0x3ffffull) + (unsigned int)((v >> 36) & 0xffull));
    float u = (float)((unsigned int)((old >> 18) & 0x3ffffull) + (unsigned int)((v >> 18) & 0xffull));
    float w = (float)((unsigned int)(old & 0x3ffffull) + (unsigned int)(v & 0xffull));
    const float k = 1.0f / 255.0f;
    red_add_f32x4(hist, make_float4(y * k, u * k, w * k, cnt));
}

__device__ __forceinline__ bool point_is_bad(float x, float y) {
    return !isfinite(fabsf(x) + fabsf(y));
}

__device__ __forceinline__ void reseed_point(float &x, float &y, float &c, mwc_st &rng) {
    x = mwc_next_11(rng);
    y = mwc_next_11(rng);
    c = mwc_next_01(rng);
}

// Camera affine + round-to-nearest-even binning (iter.py:302-317, trunca in
// code/util.py:194-200).  The unsigned compare also rejects negative
// coordinates; the x bound is astride, not awidth, as in the reference.
__device__ __forceinline__ int sample_bin(float x, float y, int astride, int aheight) {
    float xx, xy, xo, yx, yy, yo;
    camera_coefs(xx, xy, xo, yx, yy, yo);
    float cx = __fmaf_rn(xx, x, __fmaf_rn(xy, y, xo));
    float cy = __fmaf_rn(yx, x, __fmaf_rn(yy, y, yo));
    unsigned int ix = (unsigned int)__float2int_rn(cx);
    unsigned int iy = (unsigned int)__float2int_rn(cy);
    if (ix >= (unsigned int)astride || iy >= (unsigned int)aheight) return -1;
    return (int)(iy * (unsigned int)astride + ix);
}

// Slice-balancing histogram layout.  An L2 slice retires ~1e9 float4 reductions/s
// and addresses map to slices in 256-byte chunks (16 horizontally adjacent bins), so
// with the natural layout a bright region of the flame overloads a few slices
// (measured: +-15 % kernel time depending on the buffer's address).  Multiplying the
// low 16 bits of the bin index by an odd constant scatters neighbouring bins over
// the whole 1 MiB block; cb_hist_unswizzle restores the linear layout afterwards.
#define HIST_SWZ_MUL 40503u
__device__ __forceinline__ int swizzle_bin(int bin, int swizzle_bins) {
    unsigned int b = (unsigned int)bin;
    unsigned int s = (b & 0xffff0000u) | ((b * HIST_SWZ_MUL) & 0xffffu);
    return bin < swizzle_bins ? (int)s : bin;
}

// Palette column: rni(color * 255 + dither), clamped to the table (iter.py:346-351).
__device__ __forceinline__ unsigned int color_index(float color, float dither) {
    return min(__float2uint_rn(__fmaf_rn(color, 255.0f, dither)), 255u);
}

// Exchange buffers: (x, y) pairs and colours in separate arrays so that both the
// permuted write and the linear read are bank-conflict free (64-bit accesses are
// served per half-warp; the lane permutation below is a bijection mod 16).
struct xchg_buf {
    float2 xy[POINTS * ITER_THREADS];
    float c[POINTS * ITER_THREADS];
};

#ifndef XCHG_MODE
#define XCHG_MODE 1
#endif
__device__ __forceinline__ int exchange_slot(int tid, int warp, int lane, int round) {
#if XCHG_MODE == 0
    int dw = (warp + lane + (lane >> 3) * (round & 3) + (round >> 2)) & (ITER_WARPS - 1);
    int dl = ((2 * ((round >> 1) & 3) + 1) * lane + round) & 31;
    return dw * 32 + dl;
#else
    // slot = (M_r * tid + A_r) mod 256 with a round-dependent odd multiplier: a
    // bijection on the CTA's slots whose low 5 (resp. 4) bits are a bijection over
    // the lanes of a warp (half-warp), i.e. conflict-free for 4- and 8-byte stores.
    int m = ((round * 0x5b) & 0xfe) | 0x21;
    return (m * tid + round * 37) & (ITER_THREADS - 1);
#endif
}

// ---- hot bins ----------------------------------------------------------------------
// One histogram address absorbs ~6.5e8 reductions/s (profiles/r01_red_microbench.json),
// so a flame that sends more than ~0.4 % of its samples to one bin is bound by that
// bin, not by the GPU.  The reference thins such bins (1 sample in 2 / 8 / 32 with a
// 2 / 8 / 32-fold weight, iter.py:319-329 with the flags of iter.py:442-526); here the
// bins a short pilot pass found hot (cb_hot_scan) get a private cell in the CTA's
// shared memory -- sm_100a retires 7e11 shared 2 x ATOMS.ADD/s against 1.9e11 L2
// reductions (profiles/r02_smem_atomics.md) -- holding integer level sums, and each
// CTA folds its cells into the histogram every HOT_FLUSH_UNITS units.  Exact where the
// float4 path rounds: a hot bin's colour sums are integers until the flush.
#if HOT_BINS
#if ACC_PACKED
#error "HOT_BINS is an option of the float4 accumulation"
#endif
struct hot_table {
    int tag[HOT_SLOTS];                     // bin index owning the slot, -1 = none
    unsigned int cell[HOT_SLOTS][4];        // count, sum Y, sum U, sum V of 8-bit levels
};

__device__ __forceinline__ unsigned int hot_slot(int bin, unsigned int mul) {
    return ((unsigned int)bin * mul) >> HOT_HASH_SHIFT;
}

// Called by the whole CTA between two barriers.
__device__ __forceinline__ void hot_flush(hot_table *ht, const iter_args &a, int tid) {
    for (int s = tid; s < HOT_SLOTS; s += ITER_THREADS) {
        const unsigned int n = ht->cell[s][0];
        if (n) {
            red_add_f32x4(a.hist + swizzle_bin(ht->tag[s], a.swizzle_bins),
                          make_float4((float)ht->cell[s][1], (float)ht->cell[s][2],
                                      (float)ht->cell[s][3], (float)n));
            ht->cell[s][0] = 0u; ht->cell[s][1] = 0u; ht->cell[s][2] = 0u; ht->cell[s][3] = 0u;
        }
    }
}
#endif

// ---- one round = push, record, pull ------------------------------------------------
// A round transforms the point, hands it to another thread of the CTA (push ... barrier
// ... pull) and records one sample.  The sample recorded is the point the thread just
// produced, not the one it receives (every transformed point is still recorded exactly
// once): its coordinates are dead once published, so the final xform and the camera run
// without the trajectory in registers.  Where the reduction is issued relative to the
// barrier is a per-genome choice (RED_BEFORE_PULL, set by the code generator): a barrier
// right behind a global reduction waits for its L2 round trip, which paces the
// reductions of heavy genomes usefully (-3 % G6F, -2..9 % G24H: the warps of a CTA run
// different xforms and reach the barrier spread out) and starves light ones whose warps
// run in lockstep (+11..16 % G3), see profiles/r01_iter_variants.md.  A split-phase
// mbarrier (arrive, record, wait) was measured too: its polling wait costs more issue
// slots than the barrier stall it removes (25.0 vs 23.4 ms, G6F).
// First half of a round: transform this thread's point and publish it.  Returns the
// warp's random word (its low bits pick the lane that drains packed cells); `visible`
// says whether the new point may be recorded (xform opacity).
__device__ __forceinline__ unsigned int chaos_push(xchg_buf *xb, int tid, int warp, int lane, int round,
                                                   point_set &pt, mwc_st &rng,
                                                   bool (&visible)[POINTS]) {
#pragma unroll
    for (int p = 0; p < POINTS; p++)
        if (point_is_bad(pt.x[p], pt.y[p])) reseed_point(pt.x[p], pt.y[p], pt.c[p], rng);

#if XAOS
    // the choice depends on the trajectory's previous xform: one random per thread,
    // and the point stays with its thread (iter.py:236-257)
    chaos_step(mwc_next_01(rng), pt, rng, visible);
    return (unsigned int)round;
#else
    // one random word per warp per round (iter.py:197-201,261)
    unsigned int word = 0;
    if (lane == 0) word = mwc_next(rng);
    word = __shfl_sync(0xffffffffu, word, 0);
    float sel = __uint2float_rn(word) * 2.3283064365386962890625e-10f;
    chaos_step(sel, pt, rng, visible);

    xchg_buf *b = xb + (round & 1);
#pragma unroll
    for (int p = 0; p < POINTS; p++) {
        // the thread's p-th points circulate among the p-th points of the CTA, each
        // population under its own sequence of permutations
        int slot = p * ITER_THREADS + exchange_slot(tid, warp, lane, POINTS * round + p);
        b->xy[slot] = make_float2(pt.x[p], pt.y[p]);
        b->c[slot] = pt.c[p];
    }
    return word;
#endif
}

// Second half: take the points other threads published this round.
__device__ __forceinline__ void chaos_pull(xchg_buf *xb, int tid, int round, point_set &pt) {
#if !XAOS
    __syncthreads();
    xchg_buf *b = xb + (round & 1);
#pragma unroll
    for (int p = 0; p < POINTS; p++) {
        float2 q = b->xy[p * ITER_THREADS + tid];
        pt.x[p] = q.x;
        pt.y[p] = q.y;
        pt.c[p] = b->c[p * ITER_THREADS + tid];
    }
#endif
}

#if ACC_PACKED
typedef unsigned long long pal_entry;
#else
typedef float4 pal_entry;
#endif
#ifndef RED_BEFORE_PULL
#define RED_BEFORE_PULL 0
#endif

struct iter_smem {
    xchg_buf xb[2];
#if PAL_COMPACT
    unsigned int palc[256];             // the unit's palette row as 8-bit levels
#else
    pal_entry pal[256];
#endif
#if HOT_BINS
    hot_table hot;
    unsigned int hot_mul;
#endif
    unsigned int claim[2];              // unit and sweep window of the CTA's current unit
#if !ACC_PACKED
    float sweep_counts[SWEEP_MAX_WINDOW];   // counts of the window's bins (spill sweep)
#endif
};

__device__ __forceinline__ unsigned int compact_levels(unsigned long long packed) {
    return ((unsigned int)(packed >> 36) & 0xffu) << 16 |
           ((unsigned int)(packed >> 18) & 0xffu) << 8 | ((unsigned int)packed & 0xffu);
}

__device__ __forceinline__ void record_sample(const iter_args &a, iter_smem &sm, int bin,
                                              unsigned int cidx, unsigned int word, int lane) {
#if ACC_PACKED
    accumulate_packed(a.cells + bin, a.hist + bin, sm.pal[cidx], (word & 31u) == (unsigned int)lane);
#else
#if PAL_COMPACT
    const unsigned int lv = sm.palc[cidx];
#if HOT_BINS
    const unsigned int hs = hot_slot(bin, sm.hot_mul);
    if (sm.hot.tag[hs] == bin) {
        atomicAdd(&sm.hot.cell[hs][0], 1u);
        atomicAdd(&sm.hot.cell[hs][1], lv >> 16);
        atomicAdd(&sm.hot.cell[hs][2], (lv >> 8) & 0xffu);
        atomicAdd(&sm.hot.cell[hs][3], lv & 0xffu);
        return;
    }
#endif
    red_add_f32x4(a.hist + swizzle_bin(bin, a.swizzle_bins),
                  make_float4(__uint2float_rn(lv >> 16), __uint2float_rn((lv >> 8) & 0xffu),
                              __uint2float_rn(lv & 0xffu), 1.0f));
#else
    red_add_f32x4(a.hist + swizzle_bin(bin, a.swizzle_bins), sm.pal[cidx]);
#endif
#endif
}

#if HOT_BINS
#define ITER_CTAS_PER_SM HOT_MIN_CTAS
#else
#define ITER_CTAS_PER_SM ITER_MIN_CTAS
#endif
extern "C" __global__ void __launch_bounds__(ITER_THREADS, ITER_CTAS_PER_SM)
cb_iter(const __grid_constant__ iter_args a) {
    __shared__ iter_smem sm;
    xchg_buf *xb = sm.xb;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int gtid = blockIdx.x * ITER_THREADS + tid;

    mwc_st rng = a.seeds[gtid];
    point_set pt;
#ifdef CTA_TIMELINE
    if (tid == 0 && a.hot_tags) {
        unsigned long long t; unsigned int smid;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        unsigned long long *tt = (unsigned long long *)(a.hot_tags + 1024);
        tt[blockIdx.x * 3] = t; tt[blockIdx.x * 3 + 2] = smid;
    }
#endif

    const unsigned long long unit0 = a.first_sample / UNIT_SAMPLES;
    const unsigned long long nunits = (a.nsamples + UNIT_SAMPLES - 1) / UNIT_SAMPLES;
    const unsigned long long frame_units =
        (a.total_samples + UNIT_SAMPLES - 1) / UNIT_SAMPLES;

    int round_ctr = a.first_round;
    int cur_row = -1;
    bool fresh = a.fuse_rounds > 0;
#pragma unroll
    for (int p = 0; p < POINTS; p++) {
        pt.last[p] = 0;               // iter.py:209
        if (fresh) {
            reseed_point(pt.x[p], pt.y[p], pt.c[p], rng);
        } else {
            float4 q = a.points[p * POINT_STRIDE + gtid];
            pt.x[p] = q.x; pt.y[p] = q.y; pt.c[p] = q.z;
        }
    }
#if HOT_BINS
    for (int s = tid; s < HOT_SLOTS; s += ITER_THREADS) {
        sm.hot.tag[s] = a.hot_tags[s];
        sm.hot.cell[s][0] = 0u; sm.hot.cell[s][1] = 0u; sm.hot.cell[s][2] = 0u; sm.hot.cell[s][3] = 0u;
    }
    if (tid == 0) sm.hot_mul = (unsigned int)a.hot_tags[HOT_SLOTS];
    int units_done = 0;
#endif

    // Which CTA runs which unit.  The CTAs of a persistent grid do not run at one speed:
    // the warp schedulers favour the warps that were launched first, and on an issue-bound
    // kernel the first CTA of an SM runs twice as fast as the last (measured: with units
    // dealt out statically the first CTAs are done after 48 % of the launch and the tail
    // runs on a quarter of the warps, tools/cta_timeline.py).  dynamic = 1: a CTA claims
    // its next unit from a counter when it is ready for it, so that all finish within a
    // unit of each other; which stream draws which unit then depends on timing.
    // dynamic = 0: unit lu belongs to CTA lu mod grid -- the sample set is a pure function
    // of the seeds (parity tests, reproducible frames).
    unsigned long long lu = blockIdx.x;
    for (;; lu += gridDim.x) {
        if (tid == 0) {
            if (a.dynamic) sm.claim[0] = atomicAdd(&a.tickets[0], 1u);
#if !ACC_PACKED
            if (a.spill) sm.claim[1] = atomicAdd(&a.tickets[1], 1u);
#endif
        }
        __syncthreads();            // everyone is done with the previous unit's tables
        if (a.dynamic) lu = sm.claim[0];
        if (!(lu < nunits || fresh)) break;
        // temporal sample of this unit: contiguous runs of units per sample
        unsigned long long u = unit0 + (lu < nunits ? lu : 0);
        int ts = (int)((u * (unsigned long long)a.nts) / frame_units);
        int row = ts * a.pal_rows / a.nts;
#if !PARAMS_CONST
        for (int i = tid; i < NSLOTS; i += ITER_THREADS)
            s_params[i] = a.params[(size_t)ts * a.param_stride + i];
#endif
        if (row != cur_row) {
#if ACC_PACKED
            sm.pal[tid] = a.palette_packed[row * 256 + tid];
#elif PAL_COMPACT
            sm.palc[tid] = compact_levels(a.palette_packed[row * 256 + tid]);
#else
            {
                // level / 255 (cb_interp_palette) back to the level itself: exact
                const float4 q = a.palette[row * 256 + tid];
                sm.pal[tid] = make_float4(rintf(q.x * 255.0f), rintf(q.y * 255.0f),
                                          rintf(q.z * 255.0f), q.w);
            }
#endif
            cur_row = row;
        }
#if HOT_BINS
        if (units_done && units_done % HOT_FLUSH_UNITS == 0) hot_flush(&sm.hot, a, tid);
        units_done++;
#endif
#if !ACC_PACKED
        const int nbins = a.dim.aheight * a.dim.astride;
        const bool sweeping = a.spill && lu < nunits;
        int sweep_base = 0;
        if (sweeping) {
            sweep_base = (int)(((unsigned long long)sm.claim[1] * (unsigned long long)a.spill_bins)
                               % (unsigned long long)nbins);
            sweep_begin(sm.sweep_counts, a.hist, nbins, sweep_base, a.spill_bins, tid);
        }
#endif
        __syncthreads();

        if (fresh) {
            // settle new trajectories without recording them (iter.py:211-216)
            bool vis[POINTS];
            for (int r = 0; r < a.fuse_rounds; r++, round_ctr++) {
                chaos_push(xb, tid, warp, lane, round_ctr, pt, rng, vis);
                chaos_pull(xb, tid, round_ctr, pt);
            }
            fresh = false;
            if (lu >= nunits) break;
        }

        // samples of this unit that belong to the request: round r of the unit produces
        // samples (r * POINTS + p) * ITER_THREADS + tid
        unsigned long long done = lu * UNIT_SAMPLES;
        unsigned long long left = a.nsamples - done;
        int live = left >= UNIT_SAMPLES ? UNIT_SAMPLES : (int)left;
        int rounds = (live + POINTS * ITER_THREADS - 1) / (POINTS * ITER_THREADS);

        const float color_dither = 0.49f * mwc_next_11(rng);      // iter.py:185

        for (int r = 0; r < rounds; r++, round_ctr++) {
            bool visible[POINTS];
            unsigned int word = chaos_push(xb, tid, warp, lane, round_ctr, pt, rng, visible);
            // The samples of the points this thread just produced.  Their coordinates are
            // dead once published (the thread continues with the points it receives), so
            // the final xform and the camera run with them off the register file.
            int bin[POINTS];
            unsigned int cidx[POINTS];
            point_set f = pt;
#if HAS_FINAL
            final_step(f, rng);
#endif
#pragma unroll
            for (int p = 0; p < POINTS; p++) {
                bin[p] = -1;
                cidx[p] = 0;
                if (visible[p] && (r * POINTS + p) * ITER_THREADS + tid < live) {
                    bin[p] = sample_bin(f.x[p], f.y[p], a.dim.astride, a.dim.aheight);
                    cidx[p] = color_index(f.c[p], color_dither);
                }
            }
#if RED_BEFORE_PULL
#pragma unroll
            for (int p = 0; p < POINTS; p++)
                if (bin[p] >= 0) record_sample(a, sm, bin[p], cidx[p], word, lane);
            chaos_pull(xb, tid, round_ctr, pt);
#else
            chaos_pull(xb, tid, round_ctr, pt);
#pragma unroll
            for (int p = 0; p < POINTS; p++)
                if (bin[p] >= 0) record_sample(a, sm, bin[p], cidx[p], word, lane);
#endif
        }
#if !ACC_PACKED
        if (sweeping)
            sweep_end(sm.sweep_counts, a.hist, a.spill, nbins, sweep_base, a.spill_bins,
                      a.spill_count, tid);
#endif
    }
#if HOT_BINS
    __syncthreads();
    hot_flush(&sm.hot, a, tid);
#endif

#pragma unroll
    for (int p = 0; p < POINTS; p++)
        a.points[p * POINT_STRIDE + gtid] = make_float4(pt.x[p], pt.y[p], pt.c[p], 0.0f);
    a.seeds[gtid] = rng;
#ifdef CTA_TIMELINE
    if (tid == 0 && a.hot_tags) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        ((unsigned long long *)(a.hot_tags + 1024))[blockIdx.x * 3 + 1] = t;
    }
#endif
}

#if !PARAMS_CONST
// ---- probes used by the parity tests (shared-memory parameter variant only) ------
__device__ __forceinline__ void probe_load_params(const float *params) {
    for (int i = threadIdx.x; i < NSLOTS; i += blockDim.x) s_params[i] = params[i];
    __syncthreads();
}

// Apply the genome's weighted choice with a given selector to explicit points.
extern "C" __global__ void cb_probe_xform(const float *params, float *xs, float *ys,
                                          float *cs, mwc_st *seeds, int n, float sel,
                                          int use_final, int last_xf, int *visible,
                                          int *last_out) {
    probe_load_params(params);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    mwc_st rng = seeds[i];
    point_set pt;               // the probes are built with POINTS == 1
    pt.x[0] = xs[i]; pt.y[0] = ys[i]; pt.c[0] = cs[i]; pt.last[0] = last_xf;
#if HAS_FINAL
    if (use_final) final_step(pt, rng);
    else
#endif
    {
        bool vis[POINTS];
        chaos_step(sel, pt, rng, vis);
        if (visible) visible[i] = vis[0] ? 1 : 0;
        if (last_out) last_out[i] = pt.last[0];
    }
    xs[i] = pt.x[0]; ys[i] = pt.y[0]; cs[i] = pt.c[0];
    seeds[i] = rng;
}

// Camera + binning + palette column for explicit points.
extern "C" __global__ void cb_probe_bins(const float *params, const float *xs,
                                         const float *ys, const float *cs,
                                         const float *dithers, int n, int astride,
                                         int aheight, int *bins, int *cidx) {
    probe_load_params(params);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bins[i] = sample_bin(xs[i], ys[i], astride, aheight);
    cidx[i] = (int)color_index(cs[i], dithers[i]);
}

#if ACC_PACKED
// Packed accumulation of an explicit sample list: thread i adds palette entry cidx[i]
// to bin bins[i], draining the cell when drain[i] is set.
extern "C" __global__ void cb_probe_accumulate(unsigned long long *cells, float4 *hist,
                                               const int *bins, const int *cidx,
                                               const int *drain, const unsigned long long *pal,
                                               int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    accumulate_packed(cells + bins[i], hist + bins[i], pal[cidx[i]], drain[i] != 0);
}
#endif
#endif
