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

// Inverse of the slice-balancing accumulation layout (device/iter_kernel.cuh).
struct op_unswizzle {
    int swizzle_bins;
    __device__ __forceinline__ int index(int i) const {
        unsigned int u = (unsigned int)i;
        unsigned int j = (u & 0xffff0000u) | ((u * 40503u) & 0xffffu);
        return i < swizzle_bins ? (int)j : i;
    }
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
// Weighted mean of the 2*radius+1 taps along one direction.  Weight = spatial
// term x colour-distance term x density-distance term x (for r != 0) a Gompertz
// gradient term that pulls energy uphill.
__global__ void __launch_bounds__(256)
k_bilateral(float4 *dst, const float4 *src, const float *blur, int pattern,
            int radius, float sstd, float cstd, float dstd, float dpow,
            float gspeed, cb_dims dim) {
    PIX_XY();
    __shared__ float spa[32];
    if (threadIdx.y == 0) {
        float df = (float)threadIdx.x;
        spa[threadIdx.x] = expf(df * df / (-K_SQRT2 * sstd));
    }
    const int W = dim.astride, H = dim.aheight;
    const float cscale = 1.0f / (-K_SQRT2 * 3.0f * cstd);
    const float dscale = -0.5f / dstd;

    float4 cen = src[gi];
    float cdrcp = 1.0f / (cen.w + 1.0e-6f);
    cen.x *= cdrcp; cen.y *= cdrcp; cen.z *= cdrcp;
    float cpowden = powf(cen.w, dpow);

    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float wsum = 0.0f;
    __syncthreads();

    int2 o = shear_offset(pattern, (float)(-radius - 1));
    float4 pix = src[clamp_idx(xi + o.x, yi + o.y, W, H)];
    o = shear_offset(pattern, (float)(-radius));
    float4 next = src[clamp_idx(xi + o.x, yi + o.y, W, H)];

    for (int r = -radius; r <= radius; r++) {
        float prev = pix.w;
        pix = next;
        o = shear_offset(pattern, (float)(r + 1));
        next = src[clamp_idx(xi + o.x, yi + o.y, W, H)];

        float cdiff = 0.5f;
        if (pix.w > 0.0f && cen.w > 0.0f) {
            float pdrcp = 1.0f / pix.w;
            float yd = pix.x * pdrcp - cen.x;
            float ud = pix.y * pdrcp - cen.y;
            float vd = pix.z * pdrcp - cen.z;
            cdiff = yd * yd + ud * ud + vd * vd;
        }
        float powden = powf(pix.w, dpow);
        float dfact = exp2f(dscale * fabsf(cpowden - powden));

        o = shear_offset(pattern, (float)r);
        float avg = blur[clamp_idx(xi + o.x, yi + o.y, W, H)];
        float grad = (next.w - prev) / (avg + 1.0e-6f);
        if (r < 0) grad = -grad;
        float gfact = exp2f(-exp2f(gspeed * grad));

        float f = spa[abs(r)] * expf(cscale * cdiff) * dfact;
        if (r != 0) f *= gfact;
        wsum += f;
        acc.x += f * pix.x;
        acc.y += f * pix.y;
        acc.z += f * pix.z;
        acc.w += f * pix.w;
    }
    float rcp = 1.0f / (wsum + 1e-10f);
    dst[gi] = make_float4(acc.x * rcp, acc.y * rcp, acc.z * rcp, acc.w * rcp);
}

// ---- restructured bilateral direction (same result, ~3x fewer SFU ops) -----------
// The reference kernel spends ~8 MUFU operations per tap (powf of the tap density,
// three exponentials, two reciprocals).  Everything that depends on one pixel only
// is hoisted into two small prologue kernels, and the per-tap weight is evaluated
// as ONE exp2 of a sum of log2-domain terms:
//   factor = exp2( log2 spa[|r|] + log2e*cscale*cdiff + dscale*|cpow - pow_r|
//                  - [r != 0] exp2(gspeed * grad) )
// Prologue 1: aux[i].x = 7-tap blur of density (as den_blur), aux[i].y = w^dpow.
// Prologue 2: side[i] = (1 / (two-octave blur + 1e-6), w^dpow)  (den_blur_1c, up = 1)
__global__ void __launch_bounds__(256)
k_bilat_prep1(float2 *aux, const float4 *src, int pattern, coefs7 k, float dpow,
              cb_dims dim) {
    PIX_XY();
    float den = 0.0f;
#pragma unroll
    for (int i = 0; i < 7; i++) {
        int2 o = shear_offset(pattern, (float)(i - 3));
        den += src[clamp_idx(xi + o.x, yi + o.y, dim.astride, dim.aheight)].w * k.c[i];
    }
    aux[gi] = make_float2(den, powf(src[gi].w, dpow));
}

__global__ void __launch_bounds__(256)
k_bilat_prep2(float2 *side, const float2 *aux, int pattern, coefs7 k, cb_dims dim) {
    PIX_XY();
    float den = 0.0f;
#pragma unroll
    for (int i = 0; i < 7; i++) {
        int2 o = shear_offset(pattern, (float)((i - 3) * 2));
        den += aux[clamp_idx(xi + o.x, yi + o.y, dim.astride, dim.aheight)].x * k.c[i];
    }
    side[gi] = make_float2(1.0f / (den + 1.0e-6f), aux[gi].y);
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

int cb_hist_unswizzle(cb_dptr dst4, cb_dptr src4, int swizzle_bins, const cb_dims *dim,
                      cb_stream s) {
    CB_REQUIRE(swizzle_bins >= 0 && swizzle_bins % 65536 == 0, "swizzle_bins must be a multiple of 65536");
    CB_REQUIRE(dim && swizzle_bins <= dim->aheight * dim->astride, "swizzle_bins exceeds the grid");
    op_unswizzle op;
    op.swizzle_bins = swizzle_bins;
    MAP4(dst4, src4, op);
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

int cb_bilateral(cb_dptr dst4, cb_dptr src4, cb_dptr blur1, int pattern, int radius,
                 float sstd, float cstd, float dstd, float dpow, float gspeed,
                 const cb_dims *dim, cb_stream s) {
    CHECK_DIM(dim);
    CB_REQUIRE(pattern >= 0 && pattern < 16, "bad direction");
    CB_REQUIRE(radius >= 0 && radius < 32, "radius must be below 32");
    k_bilateral<<<grid2(dim), dim3(32, 8), 0, cb_cs(s)>>>(
        cb_ptr<float4>(dst4), cb_ptr<const float4>(src4), cb_ptr<const float>(blur1),
        pattern, radius, sstd, cstd, dstd, dpow, gspeed, *dim);
    CB_LAUNCH_CHECK();
    return CB_OK;
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
    k_bilat_prep1<<<grid2(dim), dim3(32, 8), 0, cb_cs(s)>>>(
        aux, cb_ptr<const float4>(src4), pattern, k, dpow, *dim);
    CB_LAUNCH_CHECK();
    k_bilat_prep2<<<grid2(dim), dim3(32, 8), 0, cb_cs(s)>>>(side, aux, pattern, k, *dim);
    CB_LAUNCH_CHECK();
    if (radius == 15)       // the reference's fixed radius (cuburn/filters.py:59): unrolled
        k_bilateral_fast<15><<<grid2(dim), dim3(32, 8), 0, cb_cs(s)>>>(
            cb_ptr<float4>(dst4), cb_ptr<const float4>(src4), side, pattern, radius, sstd,
            cstd, dstd, gspeed, *dim);
    else
        k_bilateral_fast<0><<<grid2(dim), dim3(32, 8), 0, cb_cs(s)>>>(
            cb_ptr<float4>(dst4), cb_ptr<const float4>(src4), side, pattern, radius, sstd,
            cstd, dstd, gspeed, *dim);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

}  // extern "C"
