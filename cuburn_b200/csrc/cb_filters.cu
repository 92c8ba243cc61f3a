// The filter chain on the float4 accumulation buffer: YUV->RGB, directional
// Gaussian blurs, the directional bilateral density-estimation filter,
// log-scale, and the colorclip / haloclip / smearclip / plainclip / logencode
// tone-mapping kernels.
//
// Semantics follow the reference kernels (cuburn/code/filters.py:4-413,
// cuburn/code/color.py:12-42); launch recipes live in cuburn_b200/filters.py.
// The reference samples through 2-D texture references with unnormalised
// coordinates; CUDA 12 has no texture references, and on B200 plain vector
// loads through L1/L2 are the natural replacement: neighbours are addressed
// with clamp-to-edge indices.  This file is compiled with
// --use_fast_math like the reference's modules (code/util.py:96).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "cb_common.h"

#define K_SQRT2 1.41421353816986f

// 16 filter directions in image addressing ((0,0) upper left, +y down):
// 0, 90, +-45, +-22.5, 67.5/112.5, +-30, 60/120, +-15, 75/105 degrees
// (code/filters.py:8-17).
__constant__ float2 c_dirs[16] = {
    {1.0f, 0.0f},        {0.0f, 1.0f},
    {1.0f, 1.0f},        {-1.0f, 1.0f},
    {1.0f, 0.5f},        {-0.5f, 1.0f},
    {1.0f, -0.5f},       {0.5f, 1.0f},
    {1.0f, 0.666667f},   {-0.666667f, 1.0f},
    {1.0f, -0.666667f},  {0.666667f, 1.0f},
    {1.0f, 0.333333f},   {-0.333333f, 1.0f},
    {1.0f, -0.333333f},  {0.333333f, 1.0f},
};

struct coefs7 { float c[7]; };

// Offset of a tap `radius` steps along direction `pattern`: each component is
// rounded to nearest-even *before* the pixel position is added, so the tap
// pattern is identical at every pixel (tex_shear, code/filters.py:22-35).
__device__ __forceinline__ int2 shear_offset(int pattern, float radius) {
    float2 d = c_dirs[pattern];
    return make_int2(__float2int_rn(d.x * radius), __float2int_rn(d.y * radius));
}

__device__ __forceinline__ int clamp_idx(int x, int y, int w, int h) {
    x = min(max(x, 0), w - 1);
    y = min(max(y, 0), h - 1);
    return y * w + x;
}

#define PIX_XY()                                                  \
    int xi = blockIdx.x * 32 + threadIdx.x;                       \
    int yi = blockIdx.y * 8 + threadIdx.y;                        \
    int gi = yi * dim.astride + xi

// Inverse of the slice-balancing accumulation layout (device/iter_kernel.cuh).
struct op_unswizzle_index {
    int swizzle_bins;
    __device__ __forceinline__ int index(int i) const {
        unsigned int u = (unsigned int)i;
        unsigned int j = (u & 0xffff0000u) | ((u * 40503u) & 0xffffu);
        return i < swizzle_bins ? (int)j : i;
    }
};

// ---- pointwise kernels -------------------------------------------------------------
// Each filter is a small functor; the three kernel templates below apply it to
// PW_PER x 256 consecutive bins per CTA, with every thread issuing all of its
// 16-byte loads before it uses any of them.  That keeps >= 64 KB in flight per SM,
// which HBM3e needs to approach its copy bandwidth.  nbins is a multiple of 512
// (astride % 32 == 0, aheight % 16 == 0); the last CTA is bounds-tested.
#define PW_PER 4

template <typename Op>
__global__ void __launch_bounds__(256)
k_map4(float4 *dst, const float4 *src, int n, Op op) {
    const int base = blockIdx.x * (256 * PW_PER) + threadIdx.x;
    float4 v[PW_PER];
#pragma unroll
    for (int k = 0; k < PW_PER; k++) {
        int i = base + k * 256;
        if (i < n) v[k] = src[op.index(i)];
    }
#pragma unroll
    for (int k = 0; k < PW_PER; k++) {
        int i = base + k * 256;
        if (i < n) dst[i] = op(v[k]);
    }
}

// End of the float4 accumulation (device/iter_kernel.cuh): the histogram holds integer
// level sums in the accumulation layout, full bins were moved to a second grid in the same
// layout; the filters want (sum Y / 255, sum U / 255, sum V / 255, count) in linear layout
// (iter.py:395-406).
__global__ void __launch_bounds__(256)
k_hist_finish(float4 *dst, const float4 *hist, const float4 *spill, int n, int swizzle_bins,
              float k) {
    op_unswizzle_index ix;
    ix.swizzle_bins = swizzle_bins;
    const int base = blockIdx.x * (256 * PW_PER) + threadIdx.x;
    float4 v[PW_PER], w[PW_PER];
#pragma unroll
    for (int q = 0; q < PW_PER; q++) {
        int i = base + q * 256;
        w[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (i < n) {
            const int j = ix.index(i);
            v[q] = hist[j];
            if (spill) w[q] = spill[j];
        }
    }
#pragma unroll
    for (int q = 0; q < PW_PER; q++) {
        int i = base + q * 256;
        if (i < n)
            dst[i] = make_float4((v[q].x + w[q].x) * k, (v[q].y + w[q].y) * k,
                                 (v[q].z + w[q].z) * k, v[q].w + w[q].w);
    }
}

// second input T2 (float or float4) read at the same index
template <typename Op, typename T2>
__global__ void __launch_bounds__(256)
k_map4x2(float4 *dst, const float4 *src, const T2 *src2, int n, Op op) {
    const int base = blockIdx.x * (256 * PW_PER) + threadIdx.x;
    float4 v[PW_PER];
    T2 w[PW_PER];
#pragma unroll
    for (int k = 0; k < PW_PER; k++) {
        int i = base + k * 256;
        if (i < n) { v[k] = src[i]; w[k] = src2[i]; }
    }
#pragma unroll
    for (int k = 0; k < PW_PER; k++) {
        int i = base + k * 256;
        if (i < n) dst[i] = op(v[k], w[k]);
    }
}

// float4 in, one float out
template <typename Op>
__global__ void __launch_bounds__(256)
k_map4to1(float *dst, const float4 *src, int n, Op op) {
    const int base = blockIdx.x * (256 * PW_PER) + threadIdx.x;
    float4 v[PW_PER];
#pragma unroll
    for (int k = 0; k < PW_PER; k++) {
        int i = base + k * 256;
        if (i < n) v[k] = src[i];
    }
#pragma unroll
    for (int k = 0; k < PW_PER; k++) {
        int i = base + k * 256;
        if (i < n) dst[i] = op(v[k]);
    }
}

struct op_base { __device__ __forceinline__ int index(int i) const { return i; } };

// flush_atom (code/iter.py:420-479): add what is left in the packed u64 cells
// (count:10 | Y:18 | U:18 | V:18 of 8-bit levels) into the float4 histogram.
struct op_flush_packed : op_base {
    __device__ __forceinline__ float4 operator()(float4 h, unsigned long long v) const {
        const float k = 1.0f / 255.0f;
        h.x += (float)(unsigned int)((v >> 36) & 0x3ffffull) * k;
        h.y += (float)(unsigned int)((v >> 18) & 0x3ffffull) * k;
        h.z += (float)(unsigned int)(v & 0x3ffffull) * k;
        h.w += (float)(unsigned int)(v >> 54);
        return h;
    }
};

__global__ void __launch_bounds__(256)
k_palette_pack(unsigned long long *out, const float4 *pal, int n) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float4 p = pal[i];      // exact 8-bit levels / 255 (cb_interp_palette)
    unsigned long long y = __float2uint_rn(p.x * 255.0f), u = __float2uint_rn(p.y * 255.0f),
                       v = __float2uint_rn(p.z * 255.0f);
    out[i] = (1ull << 54) | (y << 36) | (u << 18) | v;
}

__device__ __forceinline__ float4 scaled(float4 p, float s) {
    return make_float4(p.x * s, p.y * s, p.z * s, p.w * s);
}

struct op_unswizzle : op_unswizzle_index {
    __device__ __forceinline__ float4 operator()(float4 p) const { return p; }
};

// yuvo2rgb (code/color.py:33-40): remove the +0.5 per-sample chroma bias, then JPEG
// full-range YUV -> RGB, clamped at 0; the density channel is kept.
struct op_yuv_to_rgb : op_base {
    __device__ __forceinline__ float4 operator()(float4 p) const {
        float u = p.y - 0.5f * p.w, v = p.z - 0.5f * p.w;
        float r = p.x + 1.402f * v;
        float g = p.x - 0.34414f * u - 0.71414f * v;
        float b = p.x + 1.772f * u;
        return make_float4(fmaxf(0.0f, r), fmaxf(0.0f, g), fmaxf(0.0f, b), p.w);
    }
};

struct op_logscale : op_base {
    float k1, k2;
    __device__ __forceinline__ float4 operator()(float4 p) const {
        return scaled(p, fmaxf(0.0f, k1 * logf(1.0f + p.w * k2) / p.w));
    }
};

struct op_logencode : op_base {
    float degamma;
    __device__ __forceinline__ float enc(float x) const {
        return log2f(powf(x, degamma)) / 12.0f + 1.0f;
    }
    __device__ __forceinline__ float4 operator()(float4 p) const {
        return make_float4(enc(p.x), enc(p.y), enc(p.z), enc(p.w));
    }
};

struct op_apply_gamma : op_base {
    float gamma;
    __device__ __forceinline__ float operator()(float4 p) const { return powf(p.x, gamma); }
};

struct op_haloclip : op_base {
    float gamma_m_1;
    __device__ __forceinline__ float4 operator()(float4 p, float area) const {
        if (p.w <= 0.0f) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        return scaled(p, powf(p.w, gamma_m_1) / fmaxf(1.0f, area));
    }
};

struct op_gamma_full_hi : op_base {
    __device__ __forceinline__ float4 operator()(float4 p) const {
        float ls = 0.0f;
        if (p.w > 0.0f) ls = fmaxf(0.0f, p.w - 1.0f) / p.w;
        return scaled(p, ls);
    }
};

__device__ __forceinline__ float gamma_toe(float w, float gamma_m_1, float linrange,
                                           float lingam) {
    float ls = powf(w, gamma_m_1);
    if (w < linrange) {
        float frac = w / linrange;
        ls = (1.0f - frac) * lingam + frac * ls;
    }
    return ls;
}

struct op_smearclip : op_base {
    float gamma_m_1, linrange, lingam;
    __device__ __forceinline__ float4 operator()(float4 p, float4 a) const {
        p.x += a.x; p.y += a.y; p.z += a.z; p.w += a.w;
        if (p.w <= 0.0f) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        return scaled(p, gamma_toe(p.w, gamma_m_1, linrange, lingam));
    }
};

struct op_plainclip : op_base {
    float gamma_m_1, linrange, lingam, brightness;
    __device__ __forceinline__ float4 operator()(float4 p) const {
        if (p.w <= 0.0f) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        return scaled(p, gamma_toe(p.w, gamma_m_1, linrange, lingam) * brightness);
    }
};

// flam3-style gamma / vibrancy / highlight-power clip (code/filters.py:354-412)
struct op_colorclip : op_base {
    float vibrance, highpow, gamma, linrange, lingam;
    __device__ __forceinline__ float4 operator()(float4 p) const {
        if (p.w <= 0.0f) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        float4 o = p;
        float alpha = powf(p.w, gamma);
        if (p.w < linrange) {
            float frac = p.w / linrange;
            alpha = (1.0f - frac) * p.w * lingam + frac * alpha;
        }
        float ls = vibrance * alpha / p.w;
        alpha = fminf(1.0f, fmaxf(0.0f, alpha));

        float maxc = fmaxf(p.x, fmaxf(p.y, p.z));
        float maxa = maxc * ls;
        float newls = 1.0f / maxc;
        if (maxa > 1.0f && highpow >= 0.0f) {
            // desaturate towards white in proportion to the overshoot
            float lsratio = powf(newls / ls, highpow);
            p.x = maxc - (maxc - p.x * newls) * lsratio;
            p.y = maxc - (maxc - p.y * newls) * lsratio;
            p.z = maxc - (maxc - p.z * newls) * lsratio;
        } else {
            float adjhlp = -highpow;
            if (adjhlp > 1.0f || maxa <= 1.0f) adjhlp = 1.0f;
            if (maxc > 0.0f) {
                float adj = (1.0f - adjhlp) * newls + adjhlp * ls;
                p.x *= adj; p.y *= adj; p.z *= adj;
            }
        }
        float rest = 1.0f - vibrance;
        p.x += rest * powf(o.x, gamma);
        p.y += rest * powf(o.y, gamma);
        p.z += rest * powf(o.z, gamma);
        return make_float4(fminf(1.0f, p.x), fminf(1.0f, p.y), fminf(1.0f, p.z), alpha);
    }
};

// ---- hot-bin scan (iterate support) -------------------------------------------------
// After a short pilot pass of the chaos game, find the bins that hold at least
// `threshold` samples and enter them into the direct-mapped table the HOT_BINS variant
// of the iterate kernel keeps in shared memory (device/iter_kernel.cuh).  Eight hash
// multipliers are tried at once; where two hot bins share a slot the hotter one wins
// (atomicMax on density << 32 | bin) and the other stays on the global-reduction path;
// the multiplier that places most bins is the one the frame uses.  `count` receives the
// number of bins at or above `trigger` (>= threshold): the host switches the variant on
// only when some bin is hot enough to be bound by single-address atomic throughput.
// The reference's counterpart is the hotspot flag computation at the end of flush_atom
// (code/iter.py:481-526).
#define HOT_SLOTS 1024
#define HOT_HASH_SHIFT 22
#define HOT_TRIES 8
#define HIST_SWZ_INV 30599u          // 40503 * 30599 = 1 (mod 65536)
__constant__ unsigned int c_hot_muls[HOT_TRIES] = {
    2654435761u, 2246822519u, 3266489917u, 668265263u,
    374761393u, 1540483477u, 2891336453u, 4182319919u};

__global__ void __launch_bounds__(256)
k_hot_scan(unsigned long long *best, int *ntrigger, const float4 *hist, const float4 *spill,
           int nbins, int swizzle_bins, float threshold, float trigger) {
    const int i = blockIdx.x * 256 + threadIdx.x;      // storage index
    if (i >= nbins) return;
    const float w = hist[i].w + (spill ? spill[i].w : 0.0f);
    if (w < threshold) return;
    unsigned int u = (unsigned int)i;
    if (i < swizzle_bins) u = (u & 0xffff0000u) | ((u * HIST_SWZ_INV) & 0xffffu);
    const unsigned long long key = ((unsigned long long)__float_as_uint(w) << 32) | u;
#pragma unroll
    for (int t = 0; t < HOT_TRIES; t++)
        atomicMax(best + t * HOT_SLOTS + ((u * c_hot_muls[t]) >> HOT_HASH_SHIFT), key);
    if (w >= trigger) atomicAdd(ntrigger, 1);
}

// count[0] = bins >= trigger (accumulated by k_hot_scan in count[1]), count[1] reset
__global__ void __launch_bounds__(HOT_SLOTS)
k_hot_finish(int *tags, int *count, unsigned long long *best) {
    __shared__ int placed[HOT_TRIES];
    __shared__ int pick;
    const int s = threadIdx.x;
    for (int t = 0; t < HOT_TRIES; t++) {
        const int n = __syncthreads_count(best[t * HOT_SLOTS + s] != 0ull);
        if (s == 0) placed[t] = n;
    }
    if (s == 0) {
        int p = 0;
        for (int t = 1; t < HOT_TRIES; t++) if (placed[t] > placed[p]) p = t;
        pick = p;
        tags[HOT_SLOTS] = (int)c_hot_muls[p];
        count[0] = count[1];
        count[1] = 0;
        count[2] = placed[p];
    }
    __syncthreads();
    const unsigned long long b = best[pick * HOT_SLOTS + s];
    tags[s] = b ? (int)(unsigned int)b : -1;
    for (int t = 0; t < HOT_TRIES; t++) best[t * HOT_SLOTS + s] = 0ull;     // next frame
}

// ---- 7-tap directional blurs ----------------------------------------------------
__global__ void __launch_bounds__(256)
k_den_blur(float *dst, const float4 *src, int pattern, int upsample, coefs7 k,
           cb_dims dim) {
    PIX_XY();
    float den = 0.0f;
#pragma unroll
    for (int i = 0; i < 7; i++) {
        int2 o = shear_offset(pattern, (float)((i - 3) * (1 << upsample)));
        den += src[clamp_idx(xi + o.x, yi + o.y, dim.astride, dim.aheight)].w * k.c[i];
    }
    dst[gi] = den;
}

__global__ void __launch_bounds__(256)
k_den_blur_1c(float *dst, const float *src, int pattern, int upsample, coefs7 k,
              cb_dims dim) {
    PIX_XY();
    float den = 0.0f;
#pragma unroll
    for (int i = 0; i < 7; i++) {
        int2 o = shear_offset(pattern, (float)((i - 3) * (1 << upsample)));
        den += src[clamp_idx(xi + o.x, yi + o.y, dim.astride, dim.aheight)] * k.c[i];
    }
    dst[gi] = den;
}

__global__ void __launch_bounds__(256)
k_full_blur(float4 *dst, const float4 *src, int pattern, int upsample, coefs7 k,
            cb_dims dim) {
    PIX_XY();
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
    for (int i = 0; i < 7; i++) {
        int2 o = shear_offset(pattern, (float)((i - 3) * (1 << upsample)));
        float4 p = src[clamp_idx(xi + o.x, yi + o.y, dim.astride, dim.aheight)];
        acc.x += p.x * k.c[i];
        acc.y += p.y * k.c[i];
        acc.z += p.z * k.c[i];
        acc.w += p.w * k.c[i];
    }
    dst[gi] = acc;
}

// ---- directional bilateral filter (code/filters.py:166-264) -------------------
// Weighted mean of the 2*radius+1 taps along one direction.  Weight = spatial term x
// colour-distance term x density-distance term x (for r != 0) a Gompertz gradient term
// that pulls energy uphill.
// ---- restructured evaluation (same result, ~3x fewer SFU ops) --------------------
// The reference kernel spends ~8 MUFU operations per tap (powf of the tap density,
// three exponentials, two reciprocals).  Everything that depends on one pixel only
// is hoisted into two small prologue kernels, and the per-tap weight is evaluated
// as ONE exp2 of a sum of log2-domain terms:
//   factor = exp2( log2 spa[|r|] + log2e*cscale*cdiff + dscale*|cpow - pow_r|
//                  - [r != 0] exp2(gspeed * grad) )
// Prologue 1: aux[i].x = 7-tap blur of density (as den_blur), aux[i].y = w^dpow.
// Prologue 2: side[i] = (1 / (two-octave blur + 1e-6), w^dpow)  (den_blur_1c, up = 1)
// Both prologues are short and latency-bound, so a thread handles PREP_PER pixels
// (rows 8 apart in a 32 x 32 tile) and issues all of their loads before it uses any.
#define PREP_PER 4
__global__ void __launch_bounds__(256)
k_bilat_prep1(float2 *aux, const float4 *src, int pattern, coefs7 k, float dpow,
              cb_dims dim) {
    const int xi = blockIdx.x * 32 + threadIdx.x;
    const int y0 = blockIdx.y * (8 * PREP_PER) + threadIdx.y;
    float w[PREP_PER][7], cw[PREP_PER];
#pragma unroll
    for (int p = 0; p < PREP_PER; p++) {
        const int yi = min(y0 + 8 * p, dim.aheight - 1);
#pragma unroll
        for (int i = 0; i < 7; i++) {
            int2 o = shear_offset(pattern, (float)(i - 3));
            w[p][i] = src[clamp_idx(xi + o.x, yi + o.y, dim.astride, dim.aheight)].w;
        }
        cw[p] = src[yi * dim.astride + xi].w;
    }
#pragma unroll
    for (int p = 0; p < PREP_PER; p++) {
        const int yi = y0 + 8 * p;
        if (yi >= dim.aheight) break;
        float den = 0.0f;
#pragma unroll
        for (int i = 0; i < 7; i++) den += w[p][i] * k.c[i];
        aux[yi * dim.astride + xi] = make_float2(den, powf(cw[p], dpow));
    }
}

__global__ void __launch_bounds__(256)
k_bilat_prep2(float2 *side, const float2 *aux, int pattern, coefs7 k, cb_dims dim) {
    const int xi = blockIdx.x * 32 + threadIdx.x;
    const int y0 = blockIdx.y * (8 * PREP_PER) + threadIdx.y;
    float d[PREP_PER][7], cp[PREP_PER];
#pragma unroll
    for (int p = 0; p < PREP_PER; p++) {
        const int yi = min(y0 + 8 * p, dim.aheight - 1);
#pragma unroll
        for (int i = 0; i < 7; i++) {
            int2 o = shear_offset(pattern, (float)((i - 3) * 2));
            d[p][i] = aux[clamp_idx(xi + o.x, yi + o.y, dim.astride, dim.aheight)].x;
        }
        cp[p] = aux[yi * dim.astride + xi].y;
    }
#pragma unroll
    for (int p = 0; p < PREP_PER; p++) {
        const int yi = y0 + 8 * p;
        if (yi >= dim.aheight) break;
        float den = 0.0f;
#pragma unroll
        for (int i = 0; i < 7; i++) den += d[p][i] * k.c[i];
        side[yi * dim.astride + xi] = make_float2(1.0f / (den + 1.0e-6f), cp[p]);
    }
}

// The same per-pixel record from a two-octave blur plane the caller computed itself
// (cb_den_blur + cb_den_blur_1c, the reference's launch sequence filters.py:80-92).
__global__ void __launch_bounds__(256)
k_bilat_side(float2 *side, const float4 *src, const float *blur, float dpow, int n) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) side[i] = make_float2(1.0f / (blur[i] + 1.0e-6f), powf(src[i].w, dpow));
}

// RADIUS > 0: compile-time radius (the loop unrolls); RADIUS == 0: run-time radius.
// INTERIOR: every tap of every pixel of the block is inside the grid, so taps are
// addressed with precomputed linear offsets and no clamping.
template <int RADIUS, bool INTERIOR>
__device__ __forceinline__ void bilateral_fast_body(
        float4 *dst, const float4 *src, const float2 *side, int pattern, int radius_rt,
        float cscale2, float dscale, float gspeed, const float *lspa, const int2 *offs,
        const int *loffs, cb_dims dim) {
    PIX_XY();
    const int radius = RADIUS > 0 ? RADIUS : radius_rt;
    const int W = dim.astride, H = dim.aheight;

    float4 cen = src[gi];
    float cpow = side[gi].y;
    float cdrcp = 1.0f / (cen.w + 1.0e-6f);
    cen.x *= cdrcp; cen.y *= cdrcp; cen.z *= cdrcp;
    const bool cen_live = cen.w > 0.0f;

    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float wsum = 0.0f;

    auto tap = [&](int k) -> int {
        if (INTERIOR) return gi + loffs[k];
        int2 o = offs[k];
        return clamp_idx(xi + o.x, yi + o.y, W, H);
    };
    float4 pix = src[tap(0)];
    int ni = tap(1);
    float4 next = src[ni];
    float2 nside = side[ni];

#pragma unroll
    for (int r = -radius; r <= radius; r++) {
        float prev = pix.w;
        pix = next;
        float2 ps = nside;
        ni = tap(r + radius + 2);
        next = src[ni];
        nside = side[ni];

        float cdiff = 0.5f;
        if (pix.w > 0.0f && cen_live) {
            float pdrcp = 1.0f / pix.w;
            float yd = pix.x * pdrcp - cen.x;
            float ud = pix.y * pdrcp - cen.y;
            float vd = pix.z * pdrcp - cen.z;
            cdiff = yd * yd + ud * ud + vd * vd;
        }
        float e = lspa[r < 0 ? -r : r] + cscale2 * cdiff + dscale * fabsf(cpow - ps.y);
        if (r != 0) {
            float grad = (next.w - prev) * ps.x;
            if (r < 0) grad = -grad;
            e -= exp2f(gspeed * grad);
        }
        float f = exp2f(e);
        wsum += f;
        acc.x += f * pix.x;
        acc.y += f * pix.y;
        acc.z += f * pix.z;
        acc.w += f * pix.w;
    }
    float rcp = 1.0f / (wsum + 1e-10f);
    dst[gi] = make_float4(acc.x * rcp, acc.y * rcp, acc.z * rcp, acc.w * rcp);
}

#ifndef BILAT_MIN_CTAS
#define BILAT_MIN_CTAS 4
#endif
template <int RADIUS>
__global__ void __launch_bounds__(256, BILAT_MIN_CTAS)
k_bilateral_fast(float4 *dst, const float4 *src, const float2 *side, int pattern,
                 int radius_rt, float sstd, float cstd, float dstd, float gspeed,
                 cb_dims dim) {
    __shared__ float lspa[32];
    __shared__ int2 offs[36];
    __shared__ int loffs[36];
    const float log2e = 1.44269502162933f;
    const int radius = RADIUS > 0 ? RADIUS : radius_rt;
    if (threadIdx.y == 0) {
        float df = (float)threadIdx.x;
        lspa[threadIdx.x] = log2e * df * df / (-K_SQRT2 * sstd);
    }
    // tap k is the pixel (k - radius - 1) steps along the direction, k = 0 .. 2*radius+2
    for (int k = threadIdx.y * 32 + threadIdx.x; k <= 2 * radius + 2; k += 256) {
        int2 o = shear_offset(pattern, (float)(k - radius - 1));
        offs[k] = o;
        loffs[k] = o.y * dim.astride + o.x;
    }
    __syncthreads();
    const float cscale2 = log2e / (-K_SQRT2 * 3.0f * cstd);
    const float dscale = -0.5f / dstd;
    // reach of the farthest tap (|dir| <= 1 per axis, radius + 1 steps)
    const int reach = radius + 1;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    const bool interior = x0 - reach >= 0 && x0 + 31 + reach < dim.astride &&
                          y0 - reach >= 0 && y0 + 7 + reach < dim.aheight;
    if (interior)
        bilateral_fast_body<RADIUS, true>(dst, src, side, pattern, radius_rt, cscale2, dscale,
                                          gspeed, lspa, offs, loffs, dim);
    else
        bilateral_fast_body<RADIUS, false>(dst, src, side, pattern, radius_rt, cscale2, dscale,
                                           gspeed, lspa, offs, loffs, dim);
}

// ---- sliding-window bilateral pass for the y-major directions ---------------------
// k_bilateral_fast is bound by L1 wavefronts: every pixel loads 2 x 31 records that its
// neighbours along the direction load as well.  For directions whose taps advance one
// row per step, a thread here owns BW_P pixels spaced BW_S steps apart along the
// direction (lanes run along x, so every load stays coalesced) and walks ONE window of
// 33 + BW_S * (BW_P - 1) records that serves all of them: 9 (S = 1) or 11.25 (S = 4)
// record loads per pixel instead of 33, and everything that depends on the record only
// (its reciprocal density, the two Gompertz exponentials of its gradient) is computed
// once per record instead of once per (pixel, tap) pair.
//   S = 1: directions (0,1), (1,1), (-1,1) -- tap s of the window is s * dir exactly.
//   S = 4: directions (-.5,1), (.5,1) -- tex_shear rounds half-steps to even, which is
//          invariant under shifts by 4 steps (= 2 whole columns), so pixels 4 steps
//          apart see the same window: shear(4j + r) = shear(4j) + shear(r).
// A block covers 32 rows x 32 columns, sheared: the pixel j of a thread sits at column
// x + off(S j).x (mod astride).  Blocks whose window leaves the grid, and rows beyond
// aheight, take the per-pixel clamped path with identical arithmetic.
#define BW_P 4
#define BW_MAXREC (33 + 4 * (BW_P - 1))
struct bilat_tab {
    float lspa[16];             // log2 of the spatial kernel at |r|
    int loff[BW_MAXREC];        // linear offset of window record k (s = k - 16)
    int loff2[BW_MAXREC];       // the same in the pitch of a staged `side` tile (tile kernel)
    short2 off[BW_MAXREC];      // (dx, dy) of window record k
    int xlo, xhi, ylo, yhi;     // extent of the window offsets
};

struct bilat_consts { float cscale2, dscale, gspeed; };

// Packed fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2): two lanes of arithmetic per issue
// slot.  The window kernel is issue-bound, so the two pixels of a pair share every
// instruction of the colour distance and of the accumulation.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// What a record contributes, broadcast into both lanes.
struct bilat_rec {
    f32x2 x, y, z, w, pdrcp;    // the pixel (raw sums) and 1 / density
    float pow, gplus, gminus;   // density^dpow, Gompertz terms for taps after / before
    bool live;
};
// Two pixels' running state.
struct bilat_duo {
    f32x2 ncx, ncy, ncz;        // minus the centres' normalised colours
    float cpow0, cpow1;
    bool live0, live1;
    f32x2 ax, ay, az, aw, ws;
};

// log2 weight of one (record, pixel) pair but for the colour term; R = tap index of the
// record relative to the pixel.
template <int R>
__device__ __forceinline__ float bilat_weight(float cdiff, bool both_live, float cpow,
                                              const bilat_rec &rec, const bilat_tab &tab,
                                              const bilat_consts &kc) {
    if (R < -15 || R > 15) return 0.0f;          // not a tap of this pixel
    if (!both_live) cdiff = 0.5f;
    float e = kc.dscale * fabsf(cpow - rec.pow) + tab.lspa[R < 0 ? -R : R];
    e += kc.cscale2 * cdiff;
    if (R > 0) e -= rec.gplus;
    if (R < 0) e -= rec.gminus;
    return exp2f(e);
}

template <int S, int K, int D>
struct bilat_duos {
    // pixel pairs D .. BW_P/2-1 against window record K (s = K - 16)
    static __device__ __forceinline__ void run(const bilat_rec &rec, bilat_duo *duo,
                                               const bilat_tab &tab, const bilat_consts &kc) {
        constexpr int R0 = K - 16 - S * (2 * D), R1 = K - 16 - S * (2 * D + 1);
        constexpr bool in0 = R0 >= -15 && R0 <= 15, in1 = R1 >= -15 && R1 <= 15;
        if (in0 || in1) {
            bilat_duo &d = duo[D];
            f32x2 yd = fma2(rec.x, rec.pdrcp, d.ncx);
            f32x2 ud = fma2(rec.y, rec.pdrcp, d.ncy);
            f32x2 vd = fma2(rec.z, rec.pdrcp, d.ncz);
            f32x2 cd = fma2(vd, vd, fma2(ud, ud, mul2(yd, yd)));
            float c0, c1;
            unpk2(cd, c0, c1);
            float f0 = bilat_weight<R0>(c0, rec.live && d.live0, d.cpow0, rec, tab, kc);
            float f1 = bilat_weight<R1>(c1, rec.live && d.live1, d.cpow1, rec, tab, kc);
            f32x2 f = pk2(f0, f1);
            d.ws = add2(d.ws, f);
            d.ax = fma2(f, rec.x, d.ax);
            d.ay = fma2(f, rec.y, d.ay);
            d.az = fma2(f, rec.z, d.az);
            d.aw = fma2(f, rec.w, d.aw);
        }
        bilat_duos<S, K, D + 1>::run(rec, duo, tab, kc);
    }
};
template <int S, int K>
struct bilat_duos<S, K, BW_P / 2> {
    static __device__ __forceinline__ void run(const bilat_rec &, bilat_duo *,
                                               const bilat_tab &, const bilat_consts &) {}
};

// Where window record k of a thread lives: interior blocks add a linear offset, blocks
// at the edge of the grid clamp each coordinate like the reference's texture fetch.
// Clamping commutes with sharing: tap r of pixel j and record S j + r + 16 are the
// same position before the clamp, hence after it.
// MODE 0: clamped (blocks at the edge of the grid), 1: interior of the global planes,
// 2: a tile staged in shared memory (src and side tiles have their own pitches, the
// result goes to the global plane).
template <int MODE>
struct bilat_addr {
    int c0, x, y0, W, H;        // source position (MODE 2: inside the tile)
    int c1;                     // MODE 2: base index into the side tile
    int xd, yd, Wd;             // MODE 2: where the thread's first pixel lives in dst
    __device__ __forceinline__ int operator()(const bilat_tab &tab, int k) const {
        if (MODE != 0) return c0 + tab.loff[k];
        return clamp_idx(x + tab.off[k].x, y0 + tab.off[k].y, W, H);
    }
    __device__ __forceinline__ int side_idx(const bilat_tab &tab, int k) const {
        if (MODE == 2) return c1 + tab.loff2[k];
        return (*this)(tab, k);
    }
    __device__ __forceinline__ int store_idx(short2 o) const {
        // MODE 2: results go to an output tile indexed (lane, column in the block); the
        // row shift of the sheared block shape is applied when the tile is written out
        if (MODE == 2) return yd * Wd + xd + o.x;
        return (y0 + o.y) * W + x + o.x;
    }
};

// Loads run BW_AHEAD records in front of the arithmetic (a ring of registers), so that
// with two or three CTAs per SM the L2 latency hides behind the pairs of earlier records.
#ifndef BW_AHEAD
#define BW_AHEAD 2
#endif
struct bilat_raw { float4 pix; float2 side; };

template <int K, int KEND, int INTERIOR>
__device__ __forceinline__ bilat_raw bilat_load(const float4 *src, const float2 *side,
                                                const bilat_addr<INTERIOR> &at,
                                                const bilat_tab &tab) {
    bilat_raw r;
    r.pix = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    r.side = make_float2(0.0f, 0.0f);
    if (K < KEND) {                 // a tap: the whole record
        r.pix = src[at(tab, K)];
        r.side = side[at.side_idx(tab, K)];
    } else if (K == KEND) {         // past the last tap: only its density is read
        r.pix.w = src[at(tab, K)].w;
    }
    return r;
}

template <int S, int K, int KEND, int INTERIOR>
struct bilat_window {
    // ring[i] holds record K + i, i = 0 .. BW_AHEAD
    static __device__ __forceinline__ void run(
            const float4 *src, const float2 *side, const bilat_addr<INTERIOR> &at,
            float wprev, bilat_raw *ring, bilat_duo *duo, const bilat_tab &tab,
            const bilat_consts &kc) {
        const bilat_raw fresh = bilat_load<K + BW_AHEAD + 1, KEND, INTERIOR>(src, side, at, tab);
        const float4 cur = ring[0].pix;
        const float2 cs = ring[0].side;
        bilat_rec rec;
        rec.live = cur.w > 0.0f;
        const float pdrcp = 1.0f / cur.w;
        // the next record's density feeds this record's gradient
        const float ga = kc.gspeed * ((ring[1].pix.w - wprev) * cs.x);
        rec.gplus = exp2f(ga);
        rec.gminus = exp2f(-ga);
        rec.pow = cs.y;
        rec.x = pk2(cur.x, cur.x);
        rec.y = pk2(cur.y, cur.y);
        rec.z = pk2(cur.z, cur.z);
        rec.w = pk2(cur.w, cur.w);
        rec.pdrcp = pk2(pdrcp, pdrcp);
        bilat_duos<S, K, 0>::run(rec, duo, tab, kc);
#pragma unroll
        for (int i = 0; i < BW_AHEAD; i++) ring[i] = ring[i + 1];
        ring[BW_AHEAD] = fresh;
        bilat_window<S, K + 1, KEND, INTERIOR>::run(src, side, at, cur.w, ring, duo, tab, kc);
    }
};
template <int S, int KEND, int INTERIOR>
struct bilat_window<S, KEND, KEND, INTERIOR> {
    static __device__ __forceinline__ void run(const float4 *, const float2 *,
                                               const bilat_addr<INTERIOR> &, float, bilat_raw *,
                                               bilat_duo *, const bilat_tab &,
                                               const bilat_consts &) {}
};

// Per-pixel path with clamped taps (blocks at the edge of the grid); same arithmetic
// as one pixel of the window, taps r = -15 .. 15 are window records r + 16.
__device__ __noinline__ void bilat_pixel_clamped(
        float4 *dst, const float4 *src, const float2 *side, int xi, int yi,
        const bilat_tab &tab, const bilat_consts &kc, int W, int H) {
    const int gi = yi * W + xi;
    const float4 cen = src[gi];
    const float cpow = side[gi].y;
    const float cdrcp = 1.0f / (cen.w + 1.0e-6f);
    const float cnx = cen.x * cdrcp, cny = cen.y * cdrcp, cnz = cen.z * cdrcp;
    const bool cen_live = cen.w > 0.0f;
    float ax = 0.0f, ay = 0.0f, az = 0.0f, aw = 0.0f, wsum = 0.0f;
    auto at = [&](int k) { return clamp_idx(xi + tab.off[k].x, yi + tab.off[k].y, W, H); };
    float wprev = src[at(0)].w;
    int ci = at(1);
    float4 pix = src[ci];
    float2 ps = side[ci];
#pragma unroll 1
    for (int r = -15; r <= 15; r++) {
        const int ni = at(r + 17);
        const float4 nxt = src[ni];
        const float2 ns = side[ni];
        float yd = pix.x * (1.0f / pix.w) - cnx;
        float ud = pix.y * (1.0f / pix.w) - cny;
        float vd = pix.z * (1.0f / pix.w) - cnz;
        float cdiff = yd * yd + ud * ud + vd * vd;
        if (!(pix.w > 0.0f && cen_live)) cdiff = 0.5f;
        float e = kc.dscale * fabsf(cpow - ps.y) + tab.lspa[r < 0 ? -r : r];
        e += kc.cscale2 * cdiff;
        const float ga = kc.gspeed * ((nxt.w - wprev) * ps.x);
        if (r > 0) e -= exp2f(ga);
        if (r < 0) e -= exp2f(-ga);
        float f = exp2f(e);
        wsum += f;
        ax += f * pix.x;
        ay += f * pix.y;
        az += f * pix.z;
        aw += f * pix.w;
        wprev = pix.w;
        pix = nxt;
        ps = ns;
    }
    const float rcp = 1.0f / (wsum + 1e-10f);
    dst[gi] = make_float4(ax * rcp, ay * rcp, az * rcp, aw * rcp);
}

// The window of one thread.  Pixels whose sheared position leaves the grid in x belong
// to the other side of the row (the block shapes tile the row cyclically): the few
// threads that own such pixels redo them with the per-pixel path.
template <int S, int INTERIOR>
__device__ __forceinline__ void bilat_window_thread(
        float4 *dst, const float4 *src, const float2 *side, const bilat_addr<INTERIOR> &at,
        const bilat_tab &tab, const bilat_consts &kc) {
    const int x = at.x, y0 = at.y0, W = at.W, H = at.H;
    bilat_duo duo[BW_P / 2];
#pragma unroll
    for (int d = 0; d < BW_P / 2; d++) {
        float nc[2][3];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int k = S * (2 * d + h) + 16;
            const float4 cen = src[at(tab, k)];
            const float cdrcp = 1.0f / (cen.w + 1.0e-6f);
            nc[h][0] = -(cen.x * cdrcp); nc[h][1] = -(cen.y * cdrcp); nc[h][2] = -(cen.z * cdrcp);
            (h ? duo[d].cpow1 : duo[d].cpow0) = side[at.side_idx(tab, k)].y;
            (h ? duo[d].live1 : duo[d].live0) = cen.w > 0.0f;
        }
        duo[d].ncx = pk2(nc[0][0], nc[1][0]);
        duo[d].ncy = pk2(nc[0][1], nc[1][1]);
        duo[d].ncz = pk2(nc[0][2], nc[1][2]);
        duo[d].ax = duo[d].ay = duo[d].az = duo[d].aw = duo[d].ws = pk2(0.0f, 0.0f);
    }
    constexpr int KEND = 32 + S * (BW_P - 1);      // records 1 .. KEND-1 are taps
    const float wprev = src[at(tab, 0)].w;
    bilat_raw ring[BW_AHEAD + 1];
    ring[0] = bilat_load<1, KEND, INTERIOR>(src, side, at, tab);
    ring[1] = bilat_load<2, KEND, INTERIOR>(src, side, at, tab);
#if BW_AHEAD >= 2
    ring[2] = bilat_load<3, KEND, INTERIOR>(src, side, at, tab);
#endif
#if BW_AHEAD >= 3
    ring[3] = bilat_load<4, KEND, INTERIOR>(src, side, at, tab);
#endif
    bilat_window<S, 1, KEND, INTERIOR>::run(src, side, at, wprev, ring, duo, tab, kc);
#pragma unroll
    for (int d = 0; d < BW_P / 2; d++) {
        float ax[2], ay[2], az[2], aw[2], ws[2];
        unpk2(duo[d].ax, ax[0], ax[1]);
        unpk2(duo[d].ay, ay[0], ay[1]);
        unpk2(duo[d].az, az[0], az[1]);
        unpk2(duo[d].aw, aw[0], aw[1]);
        unpk2(duo[d].ws, ws[0], ws[1]);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const short2 o = tab.off[S * (2 * d + h) + 16];
            const int xj = x + o.x, yj = y0 + o.y;
            if (INTERIOR == 0 && (yj >= H || xj < 0 || xj >= W)) continue;
            const float rcp = 1.0f / (ws[h] + 1e-10f);
            dst[at.store_idx(o)] = make_float4(ax[h] * rcp, ay[h] * rcp, az[h] * rcp, aw[h] * rcp);
        }
    }
    if (INTERIOR == 0) {
#pragma unroll 1
        for (int j = 0; j < BW_P; j++) {
            const short2 o = tab.off[S * j + 16];
            const int xj = x + o.x, yj = y0 + o.y;
            if (yj < H && (xj < 0 || xj >= W))
                bilat_pixel_clamped(dst, src, side, xj < 0 ? xj + W : xj - W, yj, tab, kc, W, H);
        }
    }
}

#ifndef BW_MIN_CTAS
#define BW_MIN_CTAS 2          // 128 registers: the window state fits without spills
#endif
template <int S>
__global__ void __launch_bounds__(256, BW_MIN_CTAS)
k_bilateral_window(float4 *dst, const float4 *src, const float2 *side,
                   const __grid_constant__ bilat_tab tab, bilat_consts kc, cb_dims dim) {
    const int W = dim.astride, H = dim.aheight;
    const int x0 = blockIdx.x * 32, yb = blockIdx.y * 32;
    const int x = x0 + threadIdx.x;
    // first row of this thread: S = 1 -> rows y0 .. y0+3; S = 4 -> rows y0, y0+4, ...
    const int y0 = S == 1 ? yb + threadIdx.y * BW_P
                          : yb + (threadIdx.y >> 2) * 16 + (threadIdx.y & 3);
    const bool interior = x0 + tab.xlo >= 0 && x0 + 31 + tab.xhi < W &&
                          yb + tab.ylo >= 0 && yb + 31 + tab.yhi < H;
    if (interior) {
        bilat_addr<1> at;
        at.c0 = y0 * W + x; at.x = x; at.y0 = y0; at.W = W; at.H = H;
        bilat_window_thread<S, 1>(dst, src, side, at, tab, kc);
    } else if (y0 < H) {
        bilat_addr<0> at;
        at.c0 = y0 * W + x; at.x = x; at.y0 = y0; at.W = W; at.H = H;
        bilat_window_thread<S, 0>(dst, src, side, at, tab, kc);
    }
}

// ---- x-major directions: the same window over a tile staged in shared memory ----------
// For the directions whose taps advance one column per step -- (1,0): S = 1; (1,+-.5):
// S = 4 -- the records a thread walks lie along a row, so lanes along x would read the
// same addresses and lanes along y would read global memory with a stride of one row.
// Here a block stages the rows it needs (64 columns of float4 and float2 records, 32 to
// 54 rows) in shared memory with one bulk asynchronous copy per row and plane
// (cp.async.bulk, the TMA engine; completion on an mbarrier), and lanes run along y:
// row pitches of 65 float4 / 66 float2 keep the column-wise reads free of bank conflicts.
// Results are staged in a 32 x 32 output tile (pitch 33) and written out with lanes along
// x, so the stores are coalesced as well.
// A thread owns BW_P pixels S steps apart along the direction as in the y-major kernel;
// for S = 4 a block therefore covers a sheared set of pixels (column c of the block is
// shifted by +-2 * ((c % 16) / 4) rows) that tiles the plane vertically with period
// 32 * gridDim.y.  Blocks whose tile leaves the grid run the per-pixel clamped path.
#define BT_COLS 64
#define BT_PITCH4 65            // float4 elements per tile row (1040 B: 16 B bank shift per row)
#define BT_PITCH2 66            // float2 elements per tile row (528 B, a multiple of 16)
#define BT_MAXROWS 54
#define BT_OPITCH 33            // float4 elements per row of the output tile
#define BT_OUT_BYTES (32 * BT_OPITCH * 16)

__device__ __forceinline__ unsigned int smem_u32(const void *p) {
    return (unsigned int)__cvta_generic_to_shared(p);
}

template <int S>
__global__ void __launch_bounds__(256, BW_MIN_CTAS)
k_bilateral_tile(float4 *dst, const float4 *src, const float2 *side,
                 const __grid_constant__ bilat_tab tab, bilat_consts kc, cb_dims dim) {
    extern __shared__ __align__(128) unsigned char bt_smem[];
    const int W = dim.astride, H = dim.aheight;
    const int x0 = blockIdx.x * 32, yb = blockIdx.y * 32;
    const int lane = threadIdx.x, warp = threadIdx.y, tid = warp * 32 + lane;
    // first column of this thread's pixels inside the block
    const int cs = S == 1 ? warp * BW_P : (warp >> 2) * 16 + (warp & 3);
    const int nrows = 32 + tab.yhi - tab.ylo;
    const bool interior = x0 - 16 >= 0 && x0 + 48 <= W && yb + tab.ylo >= 0 &&
                          yb + 31 + tab.yhi < H;
    if (interior) {
        float4 *t4 = reinterpret_cast<float4 *>(bt_smem);
        float2 *t2 = reinterpret_cast<float2 *>(bt_smem + nrows * BT_PITCH4 * 16);
        unsigned long long *bar = reinterpret_cast<unsigned long long *>(
            bt_smem + nrows * (BT_PITCH4 * 16 + BT_PITCH2 * 8));
        float4 *tout = reinterpret_cast<float4 *>(
            bt_smem + nrows * (BT_PITCH4 * 16 + BT_PITCH2 * 8) + 16);
        const unsigned int bar_a = smem_u32(bar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_a));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                         :: "r"(bar_a), "r"(nrows * (BT_COLS * 16 + BT_COLS * 8)) : "memory");
        if (tid < nrows) {
            const size_t g = (size_t)(yb + tab.ylo + tid) * W + (x0 - 16);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
                         "[%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(t4 + tid * BT_PITCH4)), "l"(src + g),
                            "r"(BT_COLS * 16), "r"(bar_a) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
                         "[%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(t2 + tid * BT_PITCH2)), "l"(side + g),
                            "r"(BT_COLS * 8), "r"(bar_a) : "memory");
        }
        unsigned int done;
        do {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
                         " selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar_a) : "memory");
        } while (!done);
        bilat_addr<2> at;
        const int row = lane - tab.ylo, col = cs + 16;
        at.c0 = row * BT_PITCH4 + col;
        at.c1 = row * BT_PITCH2 + col;
        at.x = col; at.y0 = row; at.W = BT_PITCH4; at.H = nrows;
        at.xd = cs; at.yd = lane; at.Wd = BT_OPITCH;
        bilat_window_thread<S, 2>(tout, t4, t2, at, tab, kc);
        __syncthreads();
        // column c of the block belongs to pixel j of its thread: shifted by off(S j).y rows
        const int j = S == 1 ? (lane & 3) : ((lane & 15) >> 2);
        const int shift = tab.off[S * j + 16].y;
#pragma unroll
        for (int r = warp; r < 32; r += 8)
            dst[(size_t)(yb + r + shift) * W + x0 + lane] = tout[r * BT_OPITCH + lane];
    } else {
        const int wrap = 32 * gridDim.y;
#pragma unroll 1
        for (int j = 0; j < BW_P; j++) {
            const short2 o = tab.off[S * j + 16];
            int yj = yb + lane + o.y;
            yj = yj < 0 ? yj + wrap : (yj >= wrap ? yj - wrap : yj);
            if (yj < H)
                bilat_pixel_clamped(dst, src, side, x0 + cs + o.x, yj, tab, kc, W, H);
        }
    }
}

// Host copy of the first eight entries of c_dirs, for the window tables.
static const float h_dirs[8][2] = {
    {1.0f, 0.0f},  {0.0f, 1.0f},  {1.0f, 1.0f},  {-1.0f, 1.0f},
    {1.0f, 0.5f},  {-0.5f, 1.0f}, {1.0f, -0.5f}, {0.5f, 1.0f},
};

// Window step of a direction: 1 or 4 for the y-major directions handled by
// k_bilateral_window, 0 for the others.
static int bilat_window_step(int pattern) {
    if (pattern == 1 || pattern == 2 || pattern == 3) return 1;
    if (pattern == 5 || pattern == 7) return 4;
    return 0;
}

// Direction step of the x-major directions handled by k_bilateral_tile, 0 for the others.
static int bilat_tile_step(int pattern) {
    if (pattern == 0) return 1;
    if (pattern == 4 || pattern == 6) return 4;
    return 0;
}

// pitch / pitch2: elements per row of the plane (or staged tile) the linear offsets index
static bilat_tab make_bilat_tab(int pattern, int step, float sstd, int pitch, int pitch2 = 0) {
    bilat_tab t;
    memset(&t, 0, sizeof(t));
    const float log2e = 1.44269502162933f;
    for (int r = 0; r < 16; r++)
        t.lspa[r] = log2e * (float)(r * r) / (-K_SQRT2 * sstd);
    const int nrec = 33 + step * (BW_P - 1);
    for (int k = 0; k < nrec; k++) {
        // shear_offset on the host: float product, round to nearest even
        const float s = (float)(k - 16);
        const int dx = (int)nearbyintf(h_dirs[pattern][0] * s);
        const int dy = (int)nearbyintf(h_dirs[pattern][1] * s);
        t.off[k] = make_short2((short)dx, (short)dy);
        t.loff[k] = dy * pitch + dx;
        t.loff2[k] = dy * pitch2 + dx;
        t.xlo = dx < t.xlo ? dx : t.xlo;
        t.xhi = dx > t.xhi ? dx : t.xhi;
        t.ylo = dy < t.ylo ? dy : t.ylo;
        t.yhi = dy > t.yhi ? dy : t.yhi;
    }
    return t;
}

// ---- C ABI -------------------------------------------------------------------
static inline int nbins(const cb_dims *d) { return d->aheight * d->astride; }
static inline dim3 grid2(const cb_dims *d) { return dim3(d->astride / 32, d->aheight / 8); }

#define CHECK_DIM(d)                                                           \
    CB_REQUIRE((d) && (d)->astride > 0 && (d)->astride % 32 == 0 &&            \
               (d)->aheight > 0 && (d)->aheight % 16 == 0,                     \
               "dims must come from cb_calc_dim")

static inline int pw_grid(const cb_dims *d) {
    return (nbins(d) + 256 * PW_PER - 1) / (256 * PW_PER);
}

#define MAP4(dst, src, op)                                                        \
    do {                                                                          \
        CHECK_DIM(dim);                                                           \
        k_map4<<<pw_grid(dim), 256, 0, cb_cs(s)>>>(cb_ptr<float4>(dst),           \
                                                   cb_ptr<const float4>(src),     \
                                                   nbins(dim), op);               \
        CB_LAUNCH_CHECK();                                                        \
        return CB_OK;                                                             \
    } while (0)

extern "C" {

int cb_palette_pack(cb_dptr palette_packed, cb_dptr palette4, int nrows, cb_stream s) {
    CB_REQUIRE(nrows > 0, "no palette rows");
    k_palette_pack<<<nrows, 256, 0, cb_cs(s)>>>(cb_ptr<unsigned long long>(palette_packed),
                                                cb_ptr<const float4>(palette4), nrows * 256);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_flush_packed(cb_dptr hist4, cb_dptr cells, const cb_dims *dim, cb_stream s) {
    CHECK_DIM(dim);
    k_map4x2<<<pw_grid(dim), 256, 0, cb_cs(s)>>>(
        cb_ptr<float4>(hist4), cb_ptr<const float4>(hist4),
        cb_ptr<const unsigned long long>(cells), nbins(dim), op_flush_packed());
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_hot_scan(cb_dptr tags, cb_dptr count, cb_dptr scratch, cb_dptr hist4, cb_dptr spill4,
                int swizzle_bins, float threshold, float trigger, const cb_dims *dim,
                cb_stream s) {
    CHECK_DIM(dim);
    CB_REQUIRE(swizzle_bins >= 0 && swizzle_bins % 65536 == 0 && swizzle_bins <= nbins(dim),
               "swizzle_bins must be a multiple of 65536 inside the grid");
    CB_REQUIRE(threshold >= 1.0f && trigger >= threshold, "need 1 <= threshold <= trigger");
    k_hot_scan<<<(nbins(dim) + 255) / 256, 256, 0, cb_cs(s)>>>(
        cb_ptr<unsigned long long>(scratch), cb_ptr<int>(count) + 1, cb_ptr<const float4>(hist4),
        cb_ptr<const float4>(spill4), nbins(dim), swizzle_bins, threshold, trigger);
    CB_LAUNCH_CHECK();
    k_hot_finish<<<1, HOT_SLOTS, 0, cb_cs(s)>>>(cb_ptr<int>(tags), cb_ptr<int>(count),
                                                cb_ptr<unsigned long long>(scratch));
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_hist_unswizzle(cb_dptr dst4, cb_dptr src4, int swizzle_bins, const cb_dims *dim,
                      cb_stream s) {
    CB_REQUIRE(swizzle_bins >= 0 && swizzle_bins % 65536 == 0, "swizzle_bins must be a multiple of 65536");
    CB_REQUIRE(dim && swizzle_bins <= dim->aheight * dim->astride, "swizzle_bins exceeds the grid");
    op_unswizzle op;
    op.swizzle_bins = swizzle_bins;
    MAP4(dst4, src4, op);
}

int cb_hist_finish(cb_dptr dst4, cb_dptr hist4, cb_dptr spill4, int swizzle_bins,
                   float level_scale, const cb_dims *dim, cb_stream s) {
    CHECK_DIM(dim);
    CB_REQUIRE(swizzle_bins >= 0 && swizzle_bins % 65536 == 0 && swizzle_bins <= nbins(dim),
               "swizzle_bins must be a multiple of 65536 inside the grid");
    CB_REQUIRE(swizzle_bins == 0 || dst4 != hist4, "in place only for the linear layout");
    k_hist_finish<<<pw_grid(dim), 256, 0, cb_cs(s)>>>(
        cb_ptr<float4>(dst4), cb_ptr<const float4>(hist4), cb_ptr<const float4>(spill4),
        nbins(dim), swizzle_bins, level_scale);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_yuv_to_rgb(cb_dptr dst, cb_dptr src, const cb_dims *dim, cb_stream s) {
    MAP4(dst, src, op_yuv_to_rgb());
}

int cb_logscale(cb_dptr dst4, cb_dptr src4, float k1, float k2, const cb_dims *dim,
                cb_stream s) {
    op_logscale op;
    op.k1 = k1; op.k2 = k2;
    MAP4(dst4, src4, op);
}

int cb_logencode(cb_dptr dst4, cb_dptr src4, float degamma, const cb_dims *dim,
                 cb_stream s) {
    op_logencode op;
    op.degamma = degamma;
    MAP4(dst4, src4, op);
}

int cb_apply_gamma(cb_dptr dst1, cb_dptr src4, float gamma, const cb_dims *dim,
                   cb_stream s) {
    CHECK_DIM(dim);
    op_apply_gamma op;
    op.gamma = gamma;
    k_map4to1<<<pw_grid(dim), 256, 0, cb_cs(s)>>>(cb_ptr<float>(dst1),
                                                  cb_ptr<const float4>(src4), nbins(dim), op);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_haloclip(cb_dptr pix4, cb_dptr den1, float gamma_m_1, const cb_dims *dim,
                cb_stream s) {
    CHECK_DIM(dim);
    op_haloclip op;
    op.gamma_m_1 = gamma_m_1;
    k_map4x2<<<pw_grid(dim), 256, 0, cb_cs(s)>>>(cb_ptr<float4>(pix4), cb_ptr<const float4>(pix4),
                                                 cb_ptr<const float>(den1), nbins(dim), op);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_apply_gamma_full_hi(cb_dptr dst4, cb_dptr src4, float gamma_m_1,
                           const cb_dims *dim, cb_stream s) {
    (void)gamma_m_1;        // unused by the reference kernel too (code/filters.py:294-303)
    MAP4(dst4, src4, op_gamma_full_hi());
}

int cb_smearclip(cb_dptr pix4, cb_dptr smear4, float gamma_m_1, float linrange,
                 float lingam, const cb_dims *dim, cb_stream s) {
    CHECK_DIM(dim);
    op_smearclip op;
    op.gamma_m_1 = gamma_m_1; op.linrange = linrange; op.lingam = lingam;
    k_map4x2<<<pw_grid(dim), 256, 0, cb_cs(s)>>>(cb_ptr<float4>(pix4), cb_ptr<const float4>(pix4),
                                                 cb_ptr<const float4>(smear4), nbins(dim), op);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_plainclip(cb_dptr pix4, float gamma_m_1, float linrange, float lingam,
                 float brightness, const cb_dims *dim, cb_stream s) {
    op_plainclip op;
    op.gamma_m_1 = gamma_m_1; op.linrange = linrange; op.lingam = lingam;
    op.brightness = brightness;
    MAP4(pix4, pix4, op);
}

int cb_colorclip(cb_dptr pix4, float vibrance, float highpow, float gamma,
                 float linrange, float lingam, const cb_dims *dim, cb_stream s) {
    op_colorclip op;
    op.vibrance = vibrance; op.highpow = highpow; op.gamma = gamma;
    op.linrange = linrange; op.lingam = lingam;
    MAP4(pix4, pix4, op);
}

static coefs7 load_coefs(const float c[7]) {
    coefs7 k;
    for (int i = 0; i < 7; i++) k.c[i] = c[i];
    return k;
}

int cb_den_blur(cb_dptr dst1, cb_dptr src4, int pattern, int upsample,
                const float coefs[7], const cb_dims *dim, cb_stream s) {
    CHECK_DIM(dim);
    CB_REQUIRE(pattern >= 0 && pattern < 16 && coefs, "bad blur arguments");
    k_den_blur<<<grid2(dim), dim3(32, 8), 0, cb_cs(s)>>>(
        cb_ptr<float>(dst1), cb_ptr<const float4>(src4), pattern, upsample,
        load_coefs(coefs), *dim);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_den_blur_1c(cb_dptr dst1, cb_dptr src1, int pattern, int upsample,
                   const float coefs[7], const cb_dims *dim, cb_stream s) {
    CHECK_DIM(dim);
    CB_REQUIRE(pattern >= 0 && pattern < 16 && coefs, "bad blur arguments");
    k_den_blur_1c<<<grid2(dim), dim3(32, 8), 0, cb_cs(s)>>>(
        cb_ptr<float>(dst1), cb_ptr<const float>(src1), pattern, upsample,
        load_coefs(coefs), *dim);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_full_blur(cb_dptr dst4, cb_dptr src4, int pattern, int upsample,
                 const float coefs[7], const cb_dims *dim, cb_stream s) {
    CHECK_DIM(dim);
    CB_REQUIRE(pattern >= 0 && pattern < 16 && coefs, "bad blur arguments");
    k_full_blur<<<grid2(dim), dim3(32, 8), 0, cb_cs(s)>>>(
        cb_ptr<float4>(dst4), cb_ptr<const float4>(src4), pattern, upsample,
        load_coefs(coefs), *dim);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// The 31-tap pass over (src, side): sliding-window kernel for the y-major directions,
// one pixel per thread otherwise.
static int bilateral_main(float4 *dst, const float4 *src, const float2 *side, int pattern,
                          int radius, float sstd, float cstd, float dstd, float gspeed,
                          const cb_dims *dim, cb_stream s) {
    const int step = radius == 15 ? bilat_window_step(pattern) : 0;
    const int tstep = radius == 15 ? bilat_tile_step(pattern) : 0;
    bilat_consts kc;
    kc.cscale2 = 1.44269502162933f / (-K_SQRT2 * 3.0f * cstd);
    kc.dscale = -0.5f / dstd;
    kc.gspeed = gspeed;
    const dim3 grid(dim->astride / 32, (dim->aheight + 31) / 32);
    // Measured (profiles/r02_filter_kernels.md): the staged tile wins on 4K-wide frames, whose
    // planes no longer fit L2 (375 vs 436 us for (1,0)); at 1080p the one-pixel-per-thread kernel,
    // whose loads overlap its arithmetic, is faster (127 vs 153 us).  CB_BILAT_TILE=0/1
    // forces the choice (tests run both).
    const char *force = getenv("CB_BILAT_TILE");
    // (the width decides, not the plane: a row band of a frame -- multi-GPU stills -- must
    // run the same kernel as the whole frame to stay bit-identical to it)
    const bool tile = force ? force[0] == '1' : dim->astride >= 3072;
    if (tstep && tile && dim->astride >= 96) {
        const bilat_tab tab = make_bilat_tab(pattern, tstep, sstd, BT_PITCH4, BT_PITCH2);
        const int nrows = 32 + tab.yhi - tab.ylo;
        const int smem = nrows * (BT_PITCH4 * 16 + BT_PITCH2 * 8) + 16 + BT_OUT_BYTES;
        static bool configured = false;
        if (!configured) {
            CB_CUDA(cudaFuncSetAttribute(k_bilateral_tile<1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         BT_MAXROWS * (BT_PITCH4 * 16 + BT_PITCH2 * 8) + 16 +
                                             BT_OUT_BYTES));
            CB_CUDA(cudaFuncSetAttribute(k_bilateral_tile<4>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         BT_MAXROWS * (BT_PITCH4 * 16 + BT_PITCH2 * 8) + 16 +
                                             BT_OUT_BYTES));
            configured = true;
        }
        if (tstep == 1)
            k_bilateral_tile<1><<<grid, dim3(32, 8), smem, cb_cs(s)>>>(dst, src, side, tab, kc, *dim);
        else
            k_bilateral_tile<4><<<grid, dim3(32, 8), smem, cb_cs(s)>>>(dst, src, side, tab, kc, *dim);
    } else if (step) {
        const bilat_tab tab = make_bilat_tab(pattern, step, sstd, dim->astride);
        if (step == 1)
            k_bilateral_window<1><<<grid, dim3(32, 8), 0, cb_cs(s)>>>(dst, src, side, tab, kc, *dim);
        else
            k_bilateral_window<4><<<grid, dim3(32, 8), 0, cb_cs(s)>>>(dst, src, side, tab, kc, *dim);
    } else if (radius == 15)  // the reference's fixed radius (cuburn/filters.py:59): unrolled
        k_bilateral_fast<15><<<grid2(dim), dim3(32, 8), 0, cb_cs(s)>>>(
            dst, src, side, pattern, radius, sstd, cstd, dstd, gspeed, *dim);
    else
        k_bilateral_fast<0><<<grid2(dim), dim3(32, 8), 0, cb_cs(s)>>>(
            dst, src, side, pattern, radius, sstd, cstd, dstd, gspeed, *dim);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// The reference's `bilateral` launch (cuburn/filters.py:86-94): the caller has already
// blurred the density twice into blur1.  Scratch for the per-pixel records comes from
// the stream-ordered allocator.
int cb_bilateral(cb_dptr dst4, cb_dptr src4, cb_dptr blur1, int pattern, int radius,
                 float sstd, float cstd, float dstd, float dpow, float gspeed,
                 const cb_dims *dim, cb_stream s) {
    CHECK_DIM(dim);
    CB_REQUIRE(pattern >= 0 && pattern < 16, "bad direction");
    CB_REQUIRE(radius >= 0 && radius <= 16, "radius must be at most 16");
    float2 *side = nullptr;
    CB_CUDA(cudaMallocAsync((void **)&side, sizeof(float2) * (size_t)nbins(dim), cb_cs(s)));
    k_bilat_side<<<(nbins(dim) + 255) / 256, 256, 0, cb_cs(s)>>>(
        side, cb_ptr<const float4>(src4), cb_ptr<const float>(blur1), dpow, nbins(dim));
    g_cb_launches++;
    int rc = bilateral_main(cb_ptr<float4>(dst4), cb_ptr<const float4>(src4), side, pattern,
                            radius, sstd, cstd, dstd, gspeed, dim, s);
    cudaFreeAsync(side, cb_cs(s));
    return rc;
}

// One direction of the bilateral filter = den_blur + den_blur_1c + bilateral of the
// reference recipe (cuburn/filters.py:80-94) with the per-pixel terms hoisted;
// scratch4 is any float4-sized scratch buffer (holds two float2 planes).
int cb_bilateral_direction(cb_dptr dst4, cb_dptr src4, cb_dptr scratch4, int pattern,
                           int radius, const float coefs[7], float sstd, float cstd,
                           float dstd, float dpow, float gspeed, const cb_dims *dim,
                           cb_stream s) {
    CHECK_DIM(dim);
    CB_REQUIRE(pattern >= 0 && pattern < 16 && coefs, "bad direction / coefs");
    CB_REQUIRE(radius >= 0 && radius <= 16, "radius must be at most 16");
    float2 *aux = cb_ptr<float2>(scratch4);
    float2 *side = aux + (size_t)nbins(dim);
    coefs7 k = load_coefs(coefs);
    const dim3 pgrid(dim->astride / 32, (dim->aheight + 8 * PREP_PER - 1) / (8 * PREP_PER));
    // (a fused single-launch prologue through a shared-memory tile was built and measured:
    // bit-identical, but 31.7 us against 13.8 + 9.9 us -- the halo makes it evaluate the
    // first blur 1.9 x as often, and it is issue-bound; profiles/r02_filter_kernels.md)
    k_bilat_prep1<<<pgrid, dim3(32, 8), 0, cb_cs(s)>>>(
        aux, cb_ptr<const float4>(src4), pattern, k, dpow, *dim);
    CB_LAUNCH_CHECK();
    k_bilat_prep2<<<pgrid, dim3(32, 8), 0, cb_cs(s)>>>(side, aux, pattern, k, *dim);
    CB_LAUNCH_CHECK();
    return bilateral_main(cb_ptr<float4>(dst4), cb_ptr<const float4>(src4), side, pattern,
                          radius, sstd, cstd, dstd, gspeed, dim, s);
}

}  // extern "C"
