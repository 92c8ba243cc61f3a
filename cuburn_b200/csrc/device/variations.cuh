// The flam3 variation library for the chaos-game kernel (sm_100a, NVRTC).
//
// Every function adds  w * V(tx, ty)  into (ox, oy).  Formulas, argument order
// of atan2 and the order of RNG draws are those of the reference variation
// table (cuburn/code/variations.py:22-988, catalogue in SURVEY.md appendix A);
// the code is written for this kernel: plain inline functions with explicit
// parameters (the per-genome generator passes packed-parameter slots), shared
// polar helpers, and single-precision math that maps onto the SFU pipe when the
// module is built with --use_fast_math (sin/cos/ex2/lg2/rcp/rsq -> MUFU.*).
#pragma once

// VFN_NOINLINE (experiment, off): one copy of every variation body per module, called,
// instead of one per use.  A heavy genome (G24H: 72 variation uses in 24 xforms) compiles to
// 0.5 MB of straight-line code that every warp walks through a different part of, and ncu
// shows its warps waiting for instructions as often as for anything else (stall
// no_instruction 7.7 per issue at 8K).  Calls shrink the module to 0.34 MB but pass the
// accumulators and the RNG through the stack: G24H 11.37 -> 11.26 ms (nothing), G6F 21.6 ->
// 49.4 ms.  Measured and rejected (profiles/r02_iter_variants_dyn.txt).
#ifdef VFN_NOINLINE
#define VFN __device__ __noinline__ void
#else
#define VFN __device__ __forceinline__ void
#endif

#define CB_PI      3.14159274101257f
#define CB_PI_2    1.57079637050629f
#define CB_1_PI    0.31830987334251f
#define CB_2_PI    0.63661974668503f
#define CB_2PI     6.28318548202515f
#define CB_LOG2E   1.44269502162933f

// ---- compact elementary functions ------------------------------------------------------
// --use_fast_math turns sin/cos/exp/log/pow/div/sqrt into 1-2 SFU instructions, but
// atan2f (~42 instructions), sinhf/coshf (~23 each) and fmodf (~60) stay library
// routines and dominate the heavier variations.  These replacements keep the error
// at the level of the SFU intrinsics used everywhere else (atan2: 3.2e-7 rad max;
// sinh/cosh: a few ulp; fmod: exact quotient truncation in float) at a third of the
// issue slots.  Define CB_FAST_LIBM 0 to fall back to the CUDA math library.
#ifndef CB_FAST_LIBM
#define CB_FAST_LIBM 1
#endif

#if CB_FAST_LIBM
__device__ __forceinline__ float v_atan2(float y, float x) {
    float ax = fabsf(x), ay = fabsf(y);
    float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float a = __fdividef(mn, fmaxf(mx, 1.0e-37f));          // in [0, 1]
    float s = a * a;
    // minimax fit of atan(a)/a in s = a^2 on [0, 1]
    float p = 0.006811772007495165f;
    p = fmaf(p, s, -0.033604152500629425f);
    p = fmaf(p, s, 0.07962358742952347f);
    p = fmaf(p, s, -0.13233336806297302f);
    p = fmaf(p, s, 0.19807814061641693f);
    p = fmaf(p, s, -0.3331736922264099f);
    p = fmaf(p, s, 0.9999961256980896f);
    float r = p * a;
    if (ay > ax) r = 1.57079637050629f - r;
    if (x < 0.0f) r = 3.14159274101257f - r;
    return copysignf(r, y);
}

__device__ __forceinline__ void v_sinhcosh(float x, float &sh, float &ch) {
    float e = __expf(x);
    float ie = __fdividef(1.0f, e);
    ch = 0.5f * (e + ie);
    sh = 0.5f * (e - ie);
    if (fabsf(x) < 0.25f) {                                  // avoid cancellation near 0
        float x2 = x * x;
        sh = x * fmaf(x2, fmaf(x2, fmaf(x2, 1.0f / 5040.0f, 1.0f / 120.0f), 1.0f / 6.0f), 1.0f);
    }
}

__device__ __forceinline__ float v_fmod(float x, float y) {
    return x - y * truncf(__fdividef(x, y));
}
#else
__device__ __forceinline__ float v_atan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ void v_sinhcosh(float x, float &sh, float &ch) { sh = sinhf(x); ch = coshf(x); }
__device__ __forceinline__ float v_fmod(float x, float y) { return fmodf(x, y); }
#endif
__device__ __forceinline__ float v_sinh(float x) { float s, c; v_sinhcosh(x, s, c); return s; }
__device__ __forceinline__ float v_cosh(float x) { float s, c; v_sinhcosh(x, s, c); return c; }

__device__ __forceinline__ float v_r2(float x, float y) { return x * x + y * y; }
__device__ __forceinline__ float v_r(float x, float y) { return sqrtf(x * x + y * y); }
// flam3's "atan2(x, y)" convention (angle measured from the +y axis)
__device__ __forceinline__ float v_atan_xy(float x, float y) { return v_atan2(x, y); }
// the mathematical convention
__device__ __forceinline__ float v_atan_yx(float x, float y) { return v_atan2(y, x); }

// ---- 0..12 -----------------------------------------------------------------
VFN var_linear(float tx, float ty, float w, float &ox, float &oy) {
    ox += w * tx;
    oy += w * ty;
}

VFN var_sinusoidal(float tx, float ty, float w, float &ox, float &oy) {
    ox += w * sinf(tx);
    oy += w * sinf(ty);
}

VFN var_spherical(float tx, float ty, float w, float &ox, float &oy) {
    float k = w / v_r2(tx, ty);
    ox += k * tx;
    oy += k * ty;
}

VFN var_swirl(float tx, float ty, float w, float &ox, float &oy) {
    float rr = v_r2(tx, ty);
    float s = sinf(rr), c = cosf(rr);
    ox += w * (s * tx - c * ty);
    oy += w * (c * tx + s * ty);
}

VFN var_horseshoe(float tx, float ty, float w, float &ox, float &oy) {
    float k = w / v_r(tx, ty);
    ox += k * (tx - ty) * (tx + ty);
    oy += 2.0f * tx * ty * k;
}

VFN var_polar(float tx, float ty, float w, float &ox, float &oy) {
    ox += w * v_atan_xy(tx, ty) * CB_1_PI;
    oy += w * (v_r(tx, ty) - 1.0f);
}

VFN var_handkerchief(float tx, float ty, float w, float &ox, float &oy) {
    float a = v_atan_xy(tx, ty), r = v_r(tx, ty);
    ox += w * r * sinf(a + r);
    oy += w * r * cosf(a - r);
}

VFN var_heart(float tx, float ty, float w, float &ox, float &oy) {
    float r = v_r(tx, ty);
    float a = r * v_atan_xy(tx, ty);
    float k = w * r;
    ox += k * sinf(a);
    oy -= k * cosf(a);
}

VFN var_disc(float tx, float ty, float w, float &ox, float &oy) {
    float a = w * v_atan_xy(tx, ty) * CB_1_PI;
    float r = CB_PI * v_r(tx, ty);
    ox += sinf(r) * a;
    oy += cosf(r) * a;
}

VFN var_spiral(float tx, float ty, float w, float &ox, float &oy) {
    float a = v_atan_xy(tx, ty), r = v_r(tx, ty);
    float k = w / r;
    ox += k * (cosf(a) + sinf(r));
    oy += k * (sinf(a) - cosf(r));
}

VFN var_hyperbolic(float tx, float ty, float w, float &ox, float &oy) {
    float a = v_atan_xy(tx, ty), r = v_r(tx, ty);
    ox += w * sinf(a) / r;
    oy += w * cosf(a) * r;
}

VFN var_diamond(float tx, float ty, float w, float &ox, float &oy) {
    float a = v_atan_xy(tx, ty), r = v_r(tx, ty);
    ox += w * sinf(a) * cosf(r);
    oy += w * cosf(a) * sinf(r);
}

VFN var_ex(float tx, float ty, float w, float &ox, float &oy) {
    float a = v_atan_xy(tx, ty), r = v_r(tx, ty);
    float p = sinf(a + r), q = cosf(a - r);
    float p3 = p * p * p * r, q3 = q * q * q * r;
    ox += w * (p3 + q3);
    oy += w * (p3 - q3);
}

// ---- 13..22 ----------------------------------------------------------------
VFN var_julia(float tx, float ty, float w, float &ox, float &oy, mwc_st &rng) {
    float a = 0.5f * v_atan_xy(tx, ty);
    if (mwc_next(rng) & 1u) a += CB_PI;
    float r = w * sqrtf(v_r(tx, ty));
    ox += r * cosf(a);
    oy += r * sinf(a);
}

VFN var_bent(float tx, float ty, float w, float &ox, float &oy) {
    float sx = tx < 0.0f ? 2.0f : 1.0f;
    float sy = ty < 0.0f ? 0.5f : 1.0f;
    ox += w * sx * tx;
    oy += w * sy * ty;
}

// c10, c11: pre-affine xy, yy coefficients; dx2, dy2: 1/(offset^2 + 1e-20)
VFN var_waves(float tx, float ty, float w, float &ox, float &oy,
              float c10, float c11, float dx2, float dy2) {
    ox += w * (tx + c10 * sinf(ty * dx2));
    oy += w * (ty + c11 * sinf(tx * dy2));
}

VFN var_fisheye(float tx, float ty, float w, float &ox, float &oy) {
    float k = 2.0f * w / (v_r(tx, ty) + 1.0f);
    ox += k * ty;
    oy += k * tx;
}

// xo, yo: pre-affine offsets
VFN var_popcorn(float tx, float ty, float w, float &ox, float &oy,
                float xo, float yo) {
    float dx = tanf(3.0f * ty), dy = tanf(3.0f * tx);
    ox += w * (tx + xo * sinf(dx));
    oy += w * (ty + yo * sinf(dy));
}

VFN var_exponential(float tx, float ty, float w, float &ox, float &oy) {
    float d = w * expf(tx - 1.0f);
    if (isfinite(d)) {
        float a = CB_PI * ty;
        ox += d * cosf(a);
        oy += d * sinf(a);
    }
}

VFN var_power(float tx, float ty, float w, float &ox, float &oy) {
    float a = v_atan_xy(tx, ty);
    float sa = sinf(a);
    float r = w * powf(v_r(tx, ty), sa);
    ox += r * cosf(a);
    oy += r * sa;
}

VFN var_cosine(float tx, float ty, float w, float &ox, float &oy) {
    float a = CB_PI * tx;
    ox += w * cosf(a) * v_cosh(ty);
    oy -= w * sinf(a) * v_sinh(ty);
}

VFN var_rings(float tx, float ty, float w, float &ox, float &oy, float xo) {
    float d = xo * xo;
    float r = v_r(tx, ty), a = v_atan_xy(tx, ty);
    r = w * (v_fmod(r + d, 2.0f * d) - d + r * (1.0f - d));
    ox += r * cosf(a);
    oy += r * sinf(a);
}

VFN var_fan(float tx, float ty, float w, float &ox, float &oy,
            float xo, float yo) {
    float d = xo * xo * CB_PI;
    float h = 0.5f * d;
    float a = v_atan_xy(tx, ty);
    a += (v_fmod(a + yo, d) > h) ? -h : h;
    float r = w * v_r(tx, ty);
    ox += r * cosf(a);
    oy += r * sinf(a);
}

// ---- 23..39 ----------------------------------------------------------------
VFN var_blob(float tx, float ty, float w, float &ox, float &oy,
             float low, float high, float waves) {
    float r = v_r(tx, ty), a = v_atan_xy(tx, ty);
    float half = 0.5f * (high - low);
    r *= w * (low + half * (1.0f + sinf(waves * a)));
    ox += sinf(a) * r;
    oy += cosf(a) * r;
}

VFN var_pdj(float tx, float ty, float w, float &ox, float &oy,
            float a, float b, float c, float d) {
    float nx1 = cosf(b * tx), nx2 = sinf(c * tx);
    float ny1 = sinf(a * ty), ny2 = cosf(d * ty);
    ox += w * (ny1 - nx1);
    oy += w * (nx2 - ny2);
}

VFN var_fan2(float tx, float ty, float w, float &ox, float &oy,
             float fx, float fy) {
    float d = fx * fx * CB_PI;
    float h = 0.5f * d;
    float a = v_atan_xy(tx, ty);
    float r = w * v_r(tx, ty);
    float t = a + fy - d * truncf((a + fy) / d);
    a += (t > h) ? -h : h;
    ox += r * sinf(a);
    oy += r * cosf(a);
}

VFN var_rings2(float tx, float ty, float w, float &ox, float &oy, float val) {
    float d = val * val;
    float r = v_r(tx, ty), a = v_atan_xy(tx, ty);
    r += -2.0f * d * (float)(int)((r + d) / (2.0f * d)) + r * (1.0f - d);
    ox += w * sinf(a) * r;
    oy += w * cosf(a) * r;
}

VFN var_eyefish(float tx, float ty, float w, float &ox, float &oy) {
    float k = 2.0f * w / (v_r(tx, ty) + 1.0f);
    ox += k * tx;
    oy += k * ty;
}

VFN var_bubble(float tx, float ty, float w, float &ox, float &oy) {
    float k = w / (0.25f * v_r2(tx, ty) + 1.0f);
    ox += k * tx;
    oy += k * ty;
}

VFN var_cylinder(float tx, float ty, float w, float &ox, float &oy) {
    ox += w * sinf(tx);
    oy += w * ty;
}

// mdist = max(1e-9, dist); psin = sin(angle*pi/2); pcos = mdist*cos(angle*pi/2)
VFN var_perspective(float tx, float ty, float w, float &ox, float &oy,
                    float mdist, float psin, float pcos) {
    float t = 1.0f / (mdist - ty * psin);
    ox += w * mdist * tx * t;
    oy += w * pcos * ty * t;
}

VFN var_noise(float tx, float ty, float w, float &ox, float &oy, mwc_st &rng) {
    float a = mwc_next_01(rng) * 2.0f * CB_PI;
    float r = w * mwc_next_01(rng);
    ox += tx * r * cosf(a);
    oy += ty * r * sinf(a);
}

// cn = dist / (2 power)
VFN var_julian(float tx, float ty, float w, float &ox, float &oy, mwc_st &rng,
               float power, float cn) {
    float k = truncf(mwc_next_01(rng) * fabsf(power));
    float a = (v_atan_yx(tx, ty) + 2.0f * CB_PI * k) / power;
    float r = w * powf(v_r2(tx, ty), cn);
    ox += r * cosf(a);
    oy += r * sinf(a);
}

VFN var_juliascope(float tx, float ty, float w, float &ox, float &oy,
                   mwc_st &rng, float power, float cn) {
    float ang = v_atan_yx(tx, ty);
    float k = truncf(mwc_next_01(rng) * fabsf(power));
    if (mwc_next(rng) & 1u) ang = -ang;
    float a = (2.0f * CB_PI * k + ang) / power;
    float r = w * powf(v_r2(tx, ty), cn);
    ox += r * cosf(a);
    oy += r * sinf(a);
}

VFN var_blur(float tx, float ty, float w, float &ox, float &oy, mwc_st &rng) {
    float a = mwc_next_01(rng) * 2.0f * CB_PI;
    float r = w * mwc_next_01(rng);
    ox += r * cosf(a);
    oy += r * sinf(a);
}

// Box-Muller radius with the 0.57736 stdev correction the reference applies
__device__ __forceinline__ float v_gauss_radius(float w, mwc_st &rng) {
    return w * 0.57736f * sqrtf(-2.0f * log2f(mwc_next_01(rng)) / CB_LOG2E);
}

VFN var_gaussian_blur(float tx, float ty, float w, float &ox, float &oy,
                      mwc_st &rng) {
    float a = mwc_next_01(rng) * 2.0f * CB_PI;
    float r = v_gauss_radius(w, rng);
    ox += r * cosf(a);
    oy += r * sinf(a);
}

VFN var_radial_blur(float tx, float ty, float w, float &ox, float &oy,
                    mwc_st &rng, float angle) {
    float ba = angle * CB_PI * 0.5f;
    float spin = sinf(ba), zoom = cosf(ba);
    float r = v_gauss_radius(w, rng);
    float ra = v_r(tx, ty);
    float a = v_atan_yx(tx, ty) + spin * r;
    float rz = zoom * r - 1.0f;
    ox += ra * cosf(a) + rz * tx;
    oy += ra * sinf(a) + rz * ty;
}

VFN var_pie(float tx, float ty, float w, float &ox, float &oy, mwc_st &rng,
            float slices, float rotation, float thickness) {
    float sl = truncf(mwc_next_01(rng) * slices + 0.5f);
    float a = rotation
            + 2.0f * CB_PI * (sl + mwc_next_01(rng) * thickness) / slices;
    float r = w * mwc_next_01(rng);
    ox += r * cosf(a);
    oy += r * sinf(a);
}

VFN var_ngon(float tx, float ty, float w, float &ox, float &oy,
             float sides, float power, float circle, float corners) {
    float hp = power * 0.5f;
    float b = 2.0f * CB_PI / sides;
    float rf = powf(v_r2(tx, ty), hp);
    float theta = v_atan_yx(tx, ty);
    float phi = theta - b * floorf(theta / b);
    if (phi > b / 2.0f) phi -= b;
    float amp = (corners * (1.0f / cosf(phi) - 1.0f) + circle) / rf;
    ox += w * tx * amp;
    oy += w * ty * amp;
}

VFN var_curl(float tx, float ty, float w, float &ox, float &oy,
             float c1, float c2) {
    float re = 1.0f + c1 * tx + c2 * (tx * tx - ty * ty);
    float im = c1 * ty + 2.0f * c2 * tx * ty;
    float k = w / (re * re + im * im);
    ox += k * (tx * re + ty * im);
    oy += k * (ty * re - tx * im);
}

// ---- 40..58 ----------------------------------------------------------------
VFN var_rectangles(float tx, float ty, float w, float &ox, float &oy,
                   float rx, float ry) {
    ox += w * ((rx == 0.0f) ? tx : rx * (2.0f * floorf(tx / rx) + 1.0f) - tx);
    oy += w * ((ry == 0.0f) ? ty : ry * (2.0f * floorf(ty / ry) + 1.0f) - ty);
}

VFN var_arch(float tx, float ty, float w, float &ox, float &oy, mwc_st &rng) {
    float a = mwc_next_01(rng) * w * CB_PI;
    float s = sinf(a);
    ox += w * s;
    oy += w * s * s / cosf(a);
}

VFN var_tangent(float tx, float ty, float w, float &ox, float &oy) {
    ox += w * sinf(tx) / cosf(ty);
    oy += w * tanf(ty);
}

VFN var_square(float tx, float ty, float w, float &ox, float &oy, mwc_st &rng) {
    ox += w * (mwc_next_01(rng) - 0.5f);
    oy += w * (mwc_next_01(rng) - 0.5f);
}

VFN var_rays(float tx, float ty, float w, float &ox, float &oy, mwc_st &rng) {
    float a = w * mwc_next_01(rng) * CB_PI;
    float k = w / v_r2(tx, ty);
    float t = w * tanf(a) * k;
    ox += t * cosf(tx);
    oy += t * sinf(ty);
}

VFN var_blade(float tx, float ty, float w, float &ox, float &oy, mwc_st &rng) {
    float r = mwc_next_01(rng) * w * v_r(tx, ty);
    float c = cosf(r), s = sinf(r);
    ox += w * tx * (c + s);
    oy += w * tx * (c - s);
}

VFN var_secant2(float tx, float ty, float w, float &ox, float &oy) {
    float cr = cosf(w * v_r(tx, ty));
    float icr = 1.0f / cr;
    icr += (cr < 0.0f) ? 1.0f : -1.0f;
    ox += w * tx;
    oy += w * icr;
}

VFN var_cross(float tx, float ty, float w, float &ox, float &oy) {
    float s = tx * tx - ty * ty;
    float k = w * sqrtf(1.0f / (s * s));
    ox += k * tx;
    oy += k * ty;
}

VFN var_disc2(float tx, float ty, float w, float &ox, float &oy,
              float rot, float twist) {
    float rotpi = rot * CB_PI;
    float st = sinf(twist);
    float ct = cosf(twist) - 1.0f;
    if (twist > 2.0f * CB_PI) {
        float k = 1.0f + twist - 2.0f * CB_PI;
        st *= k;
        ct *= k;
    }
    if (twist < -2.0f * CB_PI) {
        float k = 1.0f + twist + 2.0f * CB_PI;
        st *= k;
        ct *= k;
    }
    float t = rotpi * (tx + ty);
    float r = w * v_atan_xy(tx, ty) / CB_PI;
    ox += r * (sinf(t) + ct);
    oy += r * (cosf(t) + st);
}

VFN var_super_shape(float tx, float ty, float w, float &ox, float &oy,
                    mwc_st &rng, float rnd, float m, float n1, float n2,
                    float n3, float holes) {
    float theta = 0.25f * (m * v_atan_yx(tx, ty) + CB_PI);
    float t1 = powf(fabsf(cosf(theta)), n2);
    float t2 = powf(fabsf(sinf(theta)), n3);
    float d = v_r(tx, ty);
    float r = w * ((rnd * mwc_next_01(rng) + (1.0f - rnd) * d) - holes)
                * powf(t1 + t2, -1.0f / n1) / d;
    ox += r * tx;
    oy += r * ty;
}

VFN var_flower(float tx, float ty, float w, float &ox, float &oy, mwc_st &rng,
               float holes, float petals) {
    float r = w * (mwc_next_01(rng) - holes)
                * cosf(petals * v_atan_yx(tx, ty)) / v_r(tx, ty);
    ox += r * tx;
    oy += r * ty;
}

VFN var_conic(float tx, float ty, float w, float &ox, float &oy, mwc_st &rng,
              float holes, float eccen) {
    float d = v_r(tx, ty);
    float ct = tx / d;
    float r = w * (mwc_next_01(rng) - holes) * eccen / (1.0f + eccen * ct) / d;
    ox += r * tx;
    oy += r * ty;
}

VFN var_parabola(float tx, float ty, float w, float &ox, float &oy,
                 mwc_st &rng, float height, float width) {
    float r = v_r(tx, ty);
    float sr = sinf(r), cr = cosf(r);
    ox += height * w * sr * sr * mwc_next_01(rng);
    oy += width * w * cr * mwc_next_01(rng);
}

VFN var_bent2(float tx, float ty, float w, float &ox, float &oy,
              float bx, float by) {
    float sx = tx < 0.0f ? bx : 1.0f;
    float sy = ty < 0.0f ? by : 1.0f;
    ox += w * sx * tx;
    oy += w * sy * ty;
}

VFN var_bipolar(float tx, float ty, float w, float &ox, float &oy,
                float shift) {
    float rr = v_r2(tx, ty);
    float t = rr + 1.0f;
    float x2 = tx * 2.0f;
    float ps = -CB_PI_2 * shift;
    float y = 0.5f * v_atan2(2.0f * ty, rr - 1.0f) + ps;
    if (y > CB_PI_2)
        y = -CB_PI_2 + v_fmod(y + CB_PI_2, CB_PI);
    else if (y < -CB_PI_2)
        y = CB_PI_2 - v_fmod(CB_PI_2 - y, CB_PI);
    ox += w * 0.25f * CB_2_PI * logf((t + x2) / (t - x2));
    oy += w * CB_2_PI * y;
}

VFN var_boarders(float tx, float ty, float w, float &ox, float &oy,
                 mwc_st &rng) {
    float rx = rintf(tx), ry = rintf(ty);
    float fx = tx - rx, fy = ty - ry;
    float hx = fx * 0.5f + rx, hy = fy * 0.5f + ry;
    if (mwc_next_01(rng) > 0.75f) {
        ox += w * hx;
        oy += w * hy;
    } else if (fabsf(fx) >= fabsf(fy)) {
        float sgn = (fx >= 0.0f) ? 0.25f : -0.25f;
        ox += w * (hx + sgn);
        oy += w * (hy + sgn * fy / fx);
    } else {
        float sgn = (fy >= 0.0f) ? 0.25f : -0.25f;
        oy += w * (hy + sgn);
        ox += w * (hx + fx / fy * sgn);
    }
}

VFN var_butterfly(float tx, float ty, float w, float &ox, float &oy) {
    float wx = w * 1.3029400317411197908970256609023f;   // 4/sqrt(3 pi)
    float y2 = ty * 2.0f;
    float r = wx * sqrtf(fabsf(ty * tx) / (tx * tx + y2 * y2));
    ox += r * tx;
    oy += r * y2;
}

VFN var_cell(float tx, float ty, float w, float &ox, float &oy, float size) {
    float inv = 1.0f / size;
    float cx = floorf(tx * inv), cy = floorf(ty * inv);
    float dx = tx - cx * size, dy = ty - cy * size;
    // interleave negative and positive cells
    cx = (cx >= 0.0f) ? 2.0f * cx : -(2.0f * cx + 1.0f);
    cy = (cy >= 0.0f) ? 2.0f * cy : -(2.0f * cy + 1.0f);
    ox += w * (dx + cx * size);
    oy -= w * (dy + cy * size);
}

// ---- 59..77 ----------------------------------------------------------------
VFN var_cpow(float tx, float ty, float w, float &ox, float &oy, mwc_st &rng,
             float cr, float ci, float cpower) {
    float a = v_atan_yx(tx, ty);
    float lnr = 0.5f * logf(v_r2(tx, ty));
    float ip = 1.0f / cpower;
    float va = 2.0f * CB_PI * ip;
    float vc = cr * ip, vd = ci * ip;
    float ang = vc * a + vd * lnr + va * floorf(ip * mwc_next_01(rng));
    float m = w * expf(vc * lnr - vd * a);
    ox += m * cosf(ang);
    oy += m * sinf(ang);
}

// x2 = 1/max(1e-20, xlength^2), y2 likewise
VFN var_curve(float tx, float ty, float w, float &ox, float &oy,
              float xamp, float yamp, float x2, float y2) {
    ox += w * (tx + xamp * expf(-ty * ty * x2));
    oy += w * (ty + yamp * expf(-tx * tx * y2));
}

VFN var_edisc(float tx, float ty, float w, float &ox, float &oy) {
    float t = v_r2(tx, ty) + 1.0f;
    float t2 = 2.0f * tx;
    float xmax = (sqrtf(t + t2) + sqrtf(t - t2)) * 0.5f;
    float a1 = logf(xmax + sqrtf(xmax - 1.0f));
    float a2 = -acosf(tx / xmax);
    float nw = w / 11.57034632f;
    float sn = sinf(a1), cs = cosf(a1);
    if (ty > 0.0f) sn = -sn;
    ox += nw * v_cosh(a2) * cs;
    oy += nw * v_sinh(a2) * sn;
}

VFN var_elliptic(float tx, float ty, float w, float &ox, float &oy) {
    float t = v_r2(tx, ty) + 1.0f;
    float x2 = 2.0f * tx;
    float xmax = 0.5f * (sqrtf(t + x2) + sqrtf(t - x2));
    float a = tx / xmax;
    float b = 1.0f - a * a;
    float ssx = xmax - 1.0f;
    float nw = w / CB_PI_2;
    b = (b < 0.0f) ? 0.0f : sqrtf(b);
    ssx = (ssx < 0.0f) ? 0.0f : sqrtf(ssx);
    ox += nw * v_atan2(a, b);
    float l = nw * logf(xmax + ssx);
    if (ty > 0.0f) oy += l; else oy -= l;
}

VFN var_escher(float tx, float ty, float w, float &ox, float &oy, float beta) {
    float a = v_atan_yx(tx, ty);
    float lnr = 0.5f * logf(v_r2(tx, ty));
    float vc = 0.5f * (1.0f + cosf(beta));
    float vd = 0.5f * sinf(beta);
    float m = w * expf(vc * lnr - vd * a);
    float n = vc * a + vd * lnr;
    ox += m * cosf(n);
    oy += m * sinf(n);
}

VFN var_foci(float tx, float ty, float w, float &ox, float &oy) {
    float ex = expf(tx) * 0.5f;
    float enx = 0.25f / ex;
    float k = w / (ex + enx - cosf(ty));
    ox += k * (ex - enx);
    oy += k * sinf(ty);
}

VFN var_lazysusan(float tx, float ty, float w, float &ox, float &oy,
                  float lx, float ly, float twist, float space, float spin) {
    float x = tx - lx, y = ty + ly;
    float r = v_r(x, y);
    if (r < w) {
        float a = v_atan2(y, x) + spin + twist * (w - r);
        ox += w * (r * cosf(a) + lx);
        oy += w * (r * sinf(a) - ly);
    } else {
        float k = 1.0f + space / r;
        ox += w * (k * x + lx);
        oy += w * (k * y - ly);
    }
}

VFN var_loonie(float tx, float ty, float w, float &ox, float &oy) {
    float rr = v_r2(tx, ty), ww = w * w;
    float k = (rr < ww) ? w * sqrtf(ww / rr - 1.0f) : w;
    ox += k * tx;
    oy += k * ty;
}

// Perturbs the input point for the variations that follow it (name order).
VFN var_pre_blur(float &tx, float &ty, float w, float &ox, float &oy,
                 mwc_st &rng) {
    float g = w * (mwc_next_01(rng) + mwc_next_01(rng)
                 + mwc_next_01(rng) + mwc_next_01(rng) - 2.0f);
    float a = mwc_next_01(rng) * 2.0f * CB_PI;
    tx += g * cosf(a);
    ty += g * sinf(a);
}

__device__ __forceinline__ float v_modulus1(float t, float m) {
    float span = 2.0f * m;
    if (t > m) return -m + v_fmod(t + m, span);
    if (t < -m) return m - v_fmod(m - t, span);
    return t;
}

VFN var_modulus(float tx, float ty, float w, float &ox, float &oy,
                float mx, float my) {
    ox += w * v_modulus1(tx, mx);
    oy += w * v_modulus1(ty, my);
}

VFN var_oscope(float tx, float ty, float w, float &ox, float &oy,
               float separation, float frequency, float amplitude,
               float damping) {
    float tpf = 2.0f * CB_PI * frequency;
    float t = amplitude * expf(-fabsf(tx) * damping) * cosf(tpf * tx)
            + separation;
    ox += w * tx;
    if (fabsf(ty) <= t) oy -= w * ty; else oy += w * ty;
}

VFN var_polar2(float tx, float ty, float w, float &ox, float &oy) {
    float k = w / CB_PI;
    ox += k * v_atan_xy(tx, ty);
    oy += 0.5f * k * logf(v_r2(tx, ty));
}

VFN var_popcorn2(float tx, float ty, float w, float &ox, float &oy,
                 float px, float py, float pc) {
    ox += w * (tx + px * sinf(tanf(ty * pc)));
    oy += w * (ty + py * sinf(tanf(tx * pc)));
}

VFN var_scry(float tx, float ty, float w, float &ox, float &oy) {
    float t = v_r2(tx, ty);
    float k = 1.0f / (sqrtf(t) * (t + 1.0f / w));
    ox += tx * k;
    oy += ty * k;
}

__device__ __forceinline__ float v_separation1(float t, float s, float inside) {
    float h = sqrtf(t * t + s * s);
    return (t > 0.0f) ? (h - t * inside) : -(h + t * inside);
}

VFN var_separation(float tx, float ty, float w, float &ox, float &oy,
                   float sx, float xin, float sy, float yin) {
    ox += w * v_separation1(tx, sx, xin);
    oy += w * v_separation1(ty, sy, yin);
}

VFN var_split(float tx, float ty, float w, float &ox, float &oy,
              float xsize, float ysize) {
    if (cosf(tx * xsize * CB_PI) >= 0.0f) oy += w * ty; else oy -= w * ty;
    if (cosf(ty * ysize * CB_PI) >= 0.0f) ox += w * tx; else ox -= w * tx;
}

VFN var_splits(float tx, float ty, float w, float &ox, float &oy,
               float sx, float sy) {
    ox += w * (tx + copysignf(sx, tx));
    oy += w * (ty + copysignf(sy, ty));
}

VFN var_stripes(float tx, float ty, float w, float &ox, float &oy,
                float space, float warp) {
    float rx = floorf(tx + 0.5f);
    float fx = tx - rx;
    ox += w * (fx * (1.0f - space) + rx);
    oy += w * (ty + fx * fx * warp);
}

VFN var_wedge(float tx, float ty, float w, float &ox, float &oy,
              float angle, float hole, float count, float swirl) {
    float r = v_r(tx, ty);
    float a = v_atan_yx(tx, ty) + swirl * r;
    float c = floorf((count * a + CB_PI) * CB_1_PI * 0.5f);
    float comp = 1.0f - angle * count * CB_1_PI * 0.5f;
    a = a * comp + c * angle;
    r = w * (r + hole);
    ox += r * cosf(a);
    oy += r * sinf(a);
}

// ---- 80..98 ----------------------------------------------------------------
VFN var_whorl(float tx, float ty, float w, float &ox, float &oy,
              float inside, float outside) {
    float r = v_r(tx, ty);
    float a = v_atan_yx(tx, ty);
    a += ((r < w) ? inside : outside) / (w - r);
    ox += w * r * cosf(a);
    oy += w * r * sinf(a);
}

VFN var_waves2(float tx, float ty, float w, float &ox, float &oy,
               float scalex, float scaley, float freqx, float freqy) {
    ox += w * (tx + scalex * sinf(ty * freqx));
    oy += w * (ty + scaley * sinf(tx * freqy));
}

VFN var_exp(float tx, float ty, float w, float &ox, float &oy) {
    float e = expf(tx);
    ox += w * e * cosf(ty);
    oy += w * e * sinf(ty);
}

VFN var_log(float tx, float ty, float w, float &ox, float &oy) {
    ox += w * 0.5f * logf(v_r2(tx, ty));
    oy += w * v_atan_yx(tx, ty);
}

VFN var_sin(float tx, float ty, float w, float &ox, float &oy) {
    ox += w * sinf(tx) * v_cosh(ty);
    oy += w * cosf(tx) * v_sinh(ty);
}

VFN var_cos(float tx, float ty, float w, float &ox, float &oy) {
    ox += w * cosf(tx) * v_cosh(ty);
    oy -= w * sinf(tx) * v_sinh(ty);
}

VFN var_tan(float tx, float ty, float w, float &ox, float &oy) {
    float k = 1.0f / (cosf(2.0f * tx) + v_cosh(2.0f * ty));
    ox += w * k * sinf(2.0f * tx);
    oy += w * k * v_sinh(2.0f * ty);
}

VFN var_sec(float tx, float ty, float w, float &ox, float &oy) {
    float k = 2.0f / (cosf(2.0f * tx) + v_cosh(2.0f * ty));
    ox += w * k * cosf(tx) * v_cosh(ty);
    oy += w * k * sinf(tx) * v_sinh(ty);
}

VFN var_csc(float tx, float ty, float w, float &ox, float &oy) {
    float k = 2.0f / (v_cosh(2.0f * ty) - cosf(2.0f * tx));
    ox += w * k * sinf(tx) * v_cosh(ty);
    oy -= w * k * cosf(tx) * v_sinh(ty);
}

VFN var_cot(float tx, float ty, float w, float &ox, float &oy) {
    float k = 1.0f / (v_cosh(2.0f * ty) - cosf(2.0f * tx));
    ox += w * k * sinf(2.0f * tx);
    oy += w * k * -1.0f * v_sinh(2.0f * ty);
}

VFN var_sinh(float tx, float ty, float w, float &ox, float &oy) {
    ox += w * v_sinh(tx) * cosf(ty);
    oy += w * v_cosh(tx) * sinf(ty);
}

VFN var_cosh(float tx, float ty, float w, float &ox, float &oy) {
    ox += w * v_cosh(tx) * cosf(ty);
    oy += w * v_sinh(tx) * sinf(ty);
}

VFN var_tanh(float tx, float ty, float w, float &ox, float &oy) {
    float k = 1.0f / (cosf(2.0f * ty) + v_cosh(2.0f * tx));
    ox += w * k * v_sinh(2.0f * tx);
    oy += w * k * sinf(2.0f * ty);
}

VFN var_sech(float tx, float ty, float w, float &ox, float &oy) {
    float k = 2.0f / (cosf(2.0f * ty) + v_cosh(2.0f * tx));
    ox += w * k * cosf(ty) * v_cosh(tx);
    oy -= w * k * sinf(ty) * v_sinh(tx);
}

VFN var_csch(float tx, float ty, float w, float &ox, float &oy) {
    float k = 2.0f / (v_cosh(2.0f * tx) - cosf(2.0f * ty));
    ox += w * k * v_sinh(tx) * cosf(ty);
    oy -= w * k * v_cosh(tx) * sinf(ty);
}

VFN var_coth(float tx, float ty, float w, float &ox, float &oy) {
    float k = 1.0f / (v_cosh(2.0f * tx) - cosf(2.0f * ty));
    ox += w * k * v_sinh(2.0f * tx);
    oy += w * k * sinf(2.0f * ty);
}

VFN var_flux(float tx, float ty, float w, float &ox, float &oy, float spread) {
    float xp = tx + w, xm = tx - w;
    float r = w * (2.0f + spread)
            * sqrtf(sqrtf(ty * ty + xp * xp) / sqrtf(ty * ty + xm * xm));
    float a = (v_atan2(ty, xm) - v_atan2(ty, xp)) * 0.5f;
    ox += r * cosf(a);
    oy += r * sinf(a);
}

VFN var_mobius(float tx, float ty, float w, float &ox, float &oy,
               float re_a, float im_a, float re_b, float im_b,
               float re_c, float im_c, float re_d, float im_d) {
    float re_u = re_a * tx - im_a * ty + re_b;
    float im_u = re_a * ty + im_a * tx + im_b;
    float re_v = re_c * tx - im_c * ty + re_d;
    float im_v = re_c * ty + im_c * tx + im_d;
    float k = w / (re_v * re_v + im_v * im_v);
    ox += k * (re_u * re_v + im_u * im_v);
    oy += k * (im_u * re_v - re_u * im_v);
}
