// The chaos-game kernel skeleton (sm_100a, compiled per genome by NVRTC).
//
// Included at the end of a generated translation unit that defines
//   NSLOTS                number of packed parameter floats per temporal sample
//   CAM_XX .. CAM_YO      slot indices of the camera affine
//   HAS_FINAL             0/1
//   void chaos_step(const float *P, float sel, float &x, float &y, float &c, mwc_st &rng)
//                         the weighted xform choice + application
//   void final_step(const float *P, float &x, float &y, float &c, mwc_st &rng)   (if HAS_FINAL)
//
// What it computes is the reference `iter` kernel (cuburn/code/iter.py:157-418):
// per-warp xform choice, the variations, an inter-warp point exchange, the
// optional final xform on a copy, camera affine, round-to-nearest-even binning
// with the unsigned bounds test against (astride, aheight), the dithered
// palette index, and accumulation.  How it does it is different:
//   * persistent CTAs walk "units" of 256 threads x 256 rounds; the unit index
//     selects the temporal sample (params row) and palette row, so one launch
//     renders a whole frame and every RNG stream / trajectory belongs to one
//     thread for the frame (no ring buffer, no block-slot atomics)
//   * accumulation is a single 16-byte red.global.add.v4.f32 per sample straight
//     into the float4 histogram, which is L2-resident on B200 up to 1080p+;
//     the packed-u64 cells, the overflow spill and the flush kernel of the
//     reference (iter.py:332-407, 420-544) disappear, as does hotspot thinning
//   * the point exchange is double-buffered so a round costs one barrier
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
    unsigned long long first_sample;
    unsigned long long nsamples;
    unsigned long long total_samples;
};

#define ITER_THREADS 256
#define ITER_WARPS 8
#define UNIT_ROUNDS 256
#define UNIT_SAMPLES (ITER_THREADS * UNIT_ROUNDS)

#ifndef ITER_MIN_CTAS
#define ITER_MIN_CTAS 4
#endif

__device__ __forceinline__ void red_add_f32x4(float4 *addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
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
__device__ __forceinline__ int sample_bin(const float *P, float x, float y,
                                          int astride, int aheight) {
    float cx = __fmaf_rn(P[CAM_XX], x, __fmaf_rn(P[CAM_XY], y, P[CAM_XO]));
    float cy = __fmaf_rn(P[CAM_YX], x, __fmaf_rn(P[CAM_YY], y, P[CAM_YO]));
    unsigned int ix = (unsigned int)__float2int_rn(cx);
    unsigned int iy = (unsigned int)__float2int_rn(cy);
    if (ix >= (unsigned int)astride || iy >= (unsigned int)aheight) return -1;
    return (int)(iy * (unsigned int)astride + ix);
}

// Palette column: rni(color * 255 + dither), clamped to the table (iter.py:346-351).
__device__ __forceinline__ unsigned int color_index(float color, float dither) {
    return min(__float2uint_rn(__fmaf_rn(color, 255.0f, dither)), 255u);
}

// Exchange buffers: structure-of-arrays so that both the permuted write and the
// linear read are bank-conflict free.
struct xchg_buf {
    float x[ITER_THREADS];
    float y[ITER_THREADS];
    float c[ITER_THREADS];
};

__device__ __forceinline__ int exchange_slot(int warp, int lane, int round) {
    int dw = (warp + lane + (lane >> 3) * (round & 3) + (round >> 2)) & (ITER_WARPS - 1);
    int dl = ((2 * ((round >> 1) & 3) + 1) * lane + round) & 31;
    return dw * 32 + dl;
}

// Apply one round of the chaos game to this thread's point and swap points
// across the CTA.  `round` only steers the permutation.
__device__ __forceinline__ void chaos_round(const float *P, xchg_buf *xb, int tid,
                                            int warp, int lane, int round,
                                            float &x, float &y, float &c, mwc_st &rng) {
    if (point_is_bad(x, y)) reseed_point(x, y, c, rng);

    // one xform choice per warp per round (iter.py:197-201,261)
    float sel = 0.0f;
    if (lane == 0) sel = mwc_next_01(rng);
    sel = __shfl_sync(0xffffffffu, sel, 0);
    chaos_step(P, sel, x, y, c, rng);

    xchg_buf *b = xb + (round & 1);
    int slot = exchange_slot(warp, lane, round);
    b->x[slot] = x;
    b->y[slot] = y;
    b->c[slot] = c;
    __syncthreads();
    x = b->x[tid];
    y = b->y[tid];
    c = b->c[tid];
}

extern "C" __global__ void __launch_bounds__(ITER_THREADS, ITER_MIN_CTAS)
cb_iter(const __grid_constant__ iter_args a) {
    __shared__ float P[NSLOTS > 0 ? NSLOTS : 1];
    __shared__ xchg_buf xb[2];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int gtid = blockIdx.x * ITER_THREADS + tid;

    mwc_st rng = a.seeds[gtid];
    float x, y, c;

    const unsigned long long unit0 = a.first_sample / UNIT_SAMPLES;
    const unsigned long long nunits = (a.nsamples + UNIT_SAMPLES - 1) / UNIT_SAMPLES;
    const unsigned long long frame_units =
        (a.total_samples + UNIT_SAMPLES - 1) / UNIT_SAMPLES;

    int round_ctr = 0;
    bool fresh = a.fuse_rounds > 0;
    if (fresh) {
        reseed_point(x, y, c, rng);
    } else {
        float4 p = a.points[gtid];
        x = p.x; y = p.y; c = p.z;
    }

    for (unsigned long long lu = blockIdx.x; lu < nunits || fresh; lu += gridDim.x) {
        unsigned long long u = unit0 + (lu < nunits ? lu : 0);
        int ts = (frame_units >= (unsigned long long)a.nts)
                     ? (int)(u % (unsigned long long)a.nts)
                     : (int)((u * (unsigned long long)a.nts) / frame_units);
        __syncthreads();        // previous unit's readers are done with P
        for (int i = tid; i < NSLOTS; i += ITER_THREADS)
            P[i] = a.params[(size_t)ts * a.param_stride + i];
        __syncthreads();

        if (fresh) {
            // settle new trajectories without recording them (iter.py:211-216)
            for (int r = 0; r < a.fuse_rounds; r++, round_ctr++)
                chaos_round(P, xb, tid, warp, lane, round_ctr, x, y, c, rng);
            fresh = false;
            if (lu >= nunits) break;
        }

        // samples of this unit that belong to the request
        unsigned long long done = lu * UNIT_SAMPLES;
        unsigned long long left = a.nsamples - done;
        int live = left >= UNIT_SAMPLES ? UNIT_SAMPLES : (int)left;
        int rounds = (live + ITER_THREADS - 1) / ITER_THREADS;

        const float4 *pal = a.palette + (ts * a.pal_rows / a.nts) * 256;
        const float color_dither = 0.49f * mwc_next_11(rng);      // iter.py:185

        for (int r = 0; r < rounds; r++, round_ctr++) {
            chaos_round(P, xb, tid, warp, lane, round_ctr, x, y, c, rng);
            if (r * ITER_THREADS + tid >= live) continue;

            float fx = x, fy = y, fc = c;
#if HAS_FINAL
            final_step(P, fx, fy, fc, rng);
#endif
            int bin = sample_bin(P, fx, fy, a.dim.astride, a.dim.aheight);
            if (bin < 0) continue;
            float4 col = __ldg(pal + color_index(fc, color_dither));
            red_add_f32x4(a.hist + bin, col);
        }
    }

    a.points[gtid] = make_float4(x, y, c, 0.0f);
    a.seeds[gtid] = rng;
}

// ---- probes used by the parity tests -------------------------------------------
// Apply the genome's weighted choice with a given selector to explicit points.
extern "C" __global__ void cb_probe_xform(const float *params, float *xs, float *ys,
                                          float *cs, mwc_st *seeds, int n, float sel,
                                          int use_final) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    mwc_st rng = seeds[i];
    float x = xs[i], y = ys[i], c = cs[i];
#if HAS_FINAL
    if (use_final) final_step(params, x, y, c, rng);
    else
#endif
        chaos_step(params, sel, x, y, c, rng);
    xs[i] = x; ys[i] = y; cs[i] = c;
    seeds[i] = rng;
}

// Camera + binning + palette column for explicit points.
extern "C" __global__ void cb_probe_bins(const float *params, const float *xs,
                                         const float *ys, const float *cs,
                                         const float *dithers, int n, int astride,
                                         int aheight, int *bins, int *cidx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bins[i] = sample_bin(params, xs[i], ys[i], astride, aheight);
    cidx[i] = (int)color_index(cs[i], dithers[i]);
}
