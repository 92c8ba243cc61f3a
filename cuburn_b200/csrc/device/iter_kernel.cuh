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
            float cx = __fmaf_rn(P[CAM_XX], fx, __fmaf_rn(P[CAM_XY], fy, P[CAM_XO]));
            float cy = __fmaf_rn(P[CAM_YX], fx, __fmaf_rn(P[CAM_YY], fy, P[CAM_YO]));
            // round to nearest even; the unsigned compare also rejects negatives
            unsigned int ix = (unsigned int)__float2int_rn(cx);
            unsigned int iy = (unsigned int)__float2int_rn(cy);
            if (ix >= (unsigned int)a.dim.astride || iy >= (unsigned int)a.dim.aheight)
                continue;
            unsigned int ci = min(__float2uint_rn(__fmaf_rn(fc, 255.0f, color_dither)), 255u);
            float4 col = __ldg(pal + ci);
            red_add_f32x4(a.hist + (size_t)iy * a.dim.astride + ix, col);
        }
    }

    a.points[gtid] = make_float4(x, y, c, 0.0f);
    a.seeds[gtid] = rng;
}
