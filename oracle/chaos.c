/*
 * ORACLE -- test infrastructure only.  Nothing under cuburn_b200/ may import,
 * link or call this file; it exists so tests, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py have an independent CPU answer to compare with.
 *
 * CPU restatement of cuburn's chaos game (reference: cuburn/code/iter.py:157-418
 * for the iteration, cuburn/code/variations.py:22-988 for the 95 variations,
 * cuburn/code/mwc.py:56-77 for the RNG).  The reference itself cannot run here
 * (Python 2 + PyCUDA + texture references), so parity is pinned on
 * the reference's MWC known-answer model and on line-by-line fidelity to the
 * cited sources; variation formulas below are written from those sources.
 *
 * Structure: plain scalar C, one trajectory per RNG stream, one xform choice
 * per iteration (the per-warp choice and the point shuffle of the GPU kernels
 * are de-correlation devices with no effect on the distribution).
 *
 *   gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC -o liboracle_chaos.so chaos.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PI_F      3.14159274101257f
#define PI_2_F    1.57079637050629f
#define INV_PI_F  0.31830987334251f
#define TWO_PI_INV_F 0.63661974668503f   /* 2/pi */
#define LOG2E_F   1.44269502162933f

/* ---- MWC (code/mwc.py:56-77) ------------------------------------------------ */
typedef struct { uint32_t mul, state, carry; } mwc_t;

static inline uint32_t mwc_u32(mwc_t *s) {
    uint64_t t = (uint64_t)s->mul * s->state + s->carry;
    s->state = (uint32_t)t;
    s->carry = (uint32_t)(t >> 32);
    return s->state;
}
static inline float mwc_01(mwc_t *s) { return (float)mwc_u32(s) * (1.0f / 4294967296.0f); }
static inline float mwc_11(mwc_t *s) { return (float)(int32_t)mwc_u32(s) * (1.0f / 2147483648.0f); }

/* test_mwc (code/mwc.py:81-87): sum of `rounds` outputs per stream */
void oracle_mwc_sums(uint32_t *seeds, int nstreams, int rounds, uint64_t *sums) {
    for (int i = 0; i < nstreams; i++) {
        mwc_t s = {seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2]};
        uint64_t acc = 0;
        for (int k = 0; k < rounds; k++) acc += mwc_u32(&s);
        sums[i] = acc;
        seeds[3 * i + 1] = s.state;
        seeds[3 * i + 2] = s.carry;
    }
}

/* ---- xform record layout (floats), see oracle/flame_ref.py ------------------ */
#define XF_PRE      0    /* xx xy xo yx yy yo */
#define XF_POST     6
#define XF_HASPOST  12
#define XF_COLOR    13
#define XF_CSPEED   14
#define XF_NVARS    15
#define XF_VARS     16
#define VAR_STRIDE  12   /* id, weight, 10 args */
#define MAX_VARS    16
#define XF_OPACITY  (XF_VARS + MAX_VARS * VAR_STRIDE)   /* has_opacity, opacity */
#define XF_FLOATS   (XF_OPACITY + 2)
#define MAX_XF      64
/* frame record: camera[6], density[MAX_XF], then (nxf + has_final) xforms */
#define FR_CAM      0
#define FR_DEN      6
#define FR_XF       (6 + MAX_XF)

int oracle_xf_floats(void) { return XF_FLOATS; }
int oracle_frame_floats(int nxf_total) { return FR_XF + nxf_total * XF_FLOATS; }

/*
 * One variation: adds w * V(tx, ty) to (ox, oy).  `id` is the flam3 number;
 * a[] holds the arguments in the order the reference body reads them:
 *   waves: pre.xy, pre.yy, dx2, dy2        popcorn: pre.xo, pre.yo
 *   rings: pre.xo                          fan: pre.xo, pre.yo
 *   perspective: mdist, sin, cos           julian/juliascope: power, cn
 *   curve: xamp, yamp, x2, y2              others: schema order
 * tx, ty are pointers because pre_blur (67) perturbs the input point.
 */
static void variation(int id, const float *a, float w, float *ptx, float *pty,
                      float *ox, float *oy, mwc_t *rng) {
    float tx = *ptx, ty = *pty;
    switch (id) {
    case 0: *ox += tx * w; *oy += ty * w; break;
    case 1: *ox += w * sinf(tx); *oy += w * sinf(ty); break;
    case 2: { float r2 = w / (tx * tx + ty * ty); *ox += tx * r2; *oy += ty * r2; break; }
    case 3: {
        float r2 = tx * tx + ty * ty, c1 = sinf(r2), c2 = cosf(r2);
        *ox += w * (c1 * tx - c2 * ty); *oy += w * (c2 * tx + c1 * ty); break;
    }
    case 4: {
        float r = w / sqrtf(tx * tx + ty * ty);
        *ox += r * (tx - ty) * (tx + ty); *oy += 2.0f * tx * ty * r; break;
    }
    case 5:
        *ox += w * atan2f(tx, ty) * INV_PI_F;
        *oy += w * (sqrtf(tx * tx + ty * ty) - 1.0f); break;
    case 6: {
        float an = atan2f(tx, ty), r = sqrtf(tx * tx + ty * ty);
        *ox += w * r * sinf(an + r); *oy += w * r * cosf(an - r); break;
    }
    case 7: {
        float sq = sqrtf(tx * tx + ty * ty), an = sq * atan2f(tx, ty), r = w * sq;
        *ox += r * sinf(an); *oy -= r * cosf(an); break;
    }
    case 8: {
        float an = w * atan2f(tx, ty) * INV_PI_F, r = PI_F * sqrtf(tx * tx + ty * ty);
        *ox += sinf(r) * an; *oy += cosf(r) * an; break;
    }
    case 9: {
        float an = atan2f(tx, ty), r = sqrtf(tx * tx + ty * ty), r1 = w / r;
        *ox += r1 * (cosf(an) + sinf(r)); *oy += r1 * (sinf(an) - cosf(r)); break;
    }
    case 10: {
        float an = atan2f(tx, ty), r = sqrtf(tx * tx + ty * ty);
        *ox += w * sinf(an) / r; *oy += w * cosf(an) * r; break;
    }
    case 11: {
        float an = atan2f(tx, ty), r = sqrtf(tx * tx + ty * ty);
        *ox += w * sinf(an) * cosf(r); *oy += w * cosf(an) * sinf(r); break;
    }
    case 12: {
        float an = atan2f(tx, ty), r = sqrtf(tx * tx + ty * ty);
        float n0 = sinf(an + r), n1 = cosf(an - r);
        float m0 = n0 * n0 * n0 * r, m1 = n1 * n1 * n1 * r;
        *ox += w * (m0 + m1); *oy += w * (m0 - m1); break;
    }
    case 13: {
        float an = 0.5f * atan2f(tx, ty);
        if (mwc_u32(rng) & 1) an += PI_F;
        float r = w * sqrtf(sqrtf(tx * tx + ty * ty));
        *ox += r * cosf(an); *oy += r * sinf(an); break;
    }
    case 14: {
        float nx = tx < 0.0f ? 2.0f : 1.0f, ny = ty < 0.0f ? 0.5f : 1.0f;
        *ox += w * nx * tx; *oy += w * ny * ty; break;
    }
    case 15:
        *ox += w * (tx + a[0] * sinf(ty * a[2]));
        *oy += w * (ty + a[1] * sinf(tx * a[3])); break;
    case 16: {
        float r = sqrtf(tx * tx + ty * ty);
        r = 2.0f * w / (r + 1.0f);
        *ox += r * ty; *oy += r * tx; break;
    }
    case 17: {
        float dx = tanf(3.0f * ty), dy = tanf(3.0f * tx);
        *ox += w * (tx + a[0] * sinf(dx)); *oy += w * (ty + a[1] * sinf(dy)); break;
    }
    case 18: {
        float dx = w * expf(tx - 1.0f);
        if (isfinite(dx)) { float dy = PI_F * ty; *ox += dx * cosf(dy); *oy += dx * sinf(dy); }
        break;
    }
    case 19: {
        float an = atan2f(tx, ty), sa = sinf(an);
        float r = w * powf(sqrtf(tx * tx + ty * ty), sa);
        *ox += r * cosf(an); *oy += r * sa; break;
    }
    case 20: {
        float an = PI_F * tx;
        *ox += w * cosf(an) * coshf(ty); *oy -= w * sinf(an) * sinhf(ty); break;
    }
    case 21: {
        float dx = a[0] * a[0];
        float r = sqrtf(tx * tx + ty * ty), an = atan2f(tx, ty);
        r = w * (fmodf(r + dx, 2.0f * dx) - dx + r * (1.0f - dx));
        *ox += r * cosf(an); *oy += r * sinf(an); break;
    }
    case 22: {
        float dx = a[0] * a[0] * PI_F, dx2 = 0.5f * dx, dy = a[1];
        float an = atan2f(tx, ty);
        an += (fmodf(an + dy, dx) > dx2) ? -dx2 : dx2;
        float r = w * sqrtf(tx * tx + ty * ty);
        *ox += r * cosf(an); *oy += r * sinf(an); break;
    }
    case 23: {  /* blob: low, high, waves */
        float r = sqrtf(tx * tx + ty * ty), an = atan2f(tx, ty);
        float bdiff = 0.5f * (a[1] - a[0]);
        r *= w * (a[0] + bdiff * (1.0f + sinf(a[2] * an)));
        *ox += sinf(an) * r; *oy += cosf(an) * r; break;
    }
    case 24: {  /* pdj: a b c d */
        float nx1 = cosf(a[1] * tx), nx2 = sinf(a[2] * tx);
        float ny1 = sinf(a[0] * ty), ny2 = cosf(a[3] * ty);
        *ox += w * (ny1 - nx1); *oy += w * (nx2 - ny2); break;
    }
    case 25: {  /* fan2: x y */
        float dy = a[1], dx = a[0] * a[0] * PI_F, dx2 = 0.5f * dx;
        float an = atan2f(tx, ty), r = w * sqrtf(tx * tx + ty * ty);
        float t = an + dy - dx * truncf((an + dy) / dx);
        if (t > dx2) an -= dx2; else an += dx2;
        *ox += r * sinf(an); *oy += r * cosf(an); break;
    }
    case 26: {  /* rings2: val */
        float dx = a[0] * a[0];
        float r = sqrtf(tx * tx + ty * ty), an = atan2f(tx, ty);
        r += -2.0f * dx * (int)((r + dx) / (2.0f * dx)) + r * (1.0f - dx);
        *ox += w * sinf(an) * r; *oy += w * cosf(an) * r; break;
    }
    case 27: {
        float r = 2.0f * w / (sqrtf(tx * tx + ty * ty) + 1.0f);
        *ox += r * tx; *oy += r * ty; break;
    }
    case 28: {
        float r = w / (0.25f * (tx * tx + ty * ty) + 1.0f);
        *ox += r * tx; *oy += r * ty; break;
    }
    case 29: *ox += w * sinf(tx); *oy += w * ty; break;
    case 30: {  /* perspective: mdist sin cos */
        float t = 1.0f / (a[0] - ty * a[1]);
        *ox += w * a[0] * tx * t; *oy += w * a[2] * ty * t; break;
    }
    case 31: {
        float tmpr = mwc_01(rng) * 2.0f * PI_F, r = w * mwc_01(rng);
        *ox += tx * r * cosf(tmpr); *oy += ty * r * sinf(tmpr); break;
    }
    case 32: {  /* julian: power cn */
        float power = a[0];
        float t_rnd = truncf(mwc_01(rng) * fabsf(power));
        float an = atan2f(ty, tx);
        float tmpr = (an + 2.0f * PI_F * t_rnd) / power;
        float r = w * powf(tx * tx + ty * ty, a[1]);
        *ox += r * cosf(tmpr); *oy += r * sinf(tmpr); break;
    }
    case 33: {  /* juliascope: power cn */
        float ang = atan2f(ty, tx), power = a[0];
        float t_rnd = truncf(mwc_01(rng) * fabsf(power));
        if (mwc_u32(rng) & 1) ang = -ang;
        float tmpr = (2.0f * PI_F * t_rnd + ang) / power;
        float r = w * powf(tx * tx + ty * ty, a[1]);
        *ox += r * cosf(tmpr); *oy += r * sinf(tmpr); break;
    }
    case 34: {
        float tmpr = mwc_01(rng) * 2.0f * PI_F, r = w * mwc_01(rng);
        *ox += r * cosf(tmpr); *oy += r * sinf(tmpr); break;
    }
    case 35: {
        float ang = mwc_01(rng) * 2.0f * PI_F;
        float r = w * 0.57736f * sqrtf(-2.0f * log2f(mwc_01(rng)) / LOG2E_F);
        *ox += r * cosf(ang); *oy += r * sinf(ang); break;
    }
    case 36: {  /* radial_blur: angle */
        float ba = a[0] * PI_F * 0.5f, spinvar = sinf(ba), zoomvar = cosf(ba);
        float r = w * 0.57736f * sqrtf(-2.0f * log2f(mwc_01(rng)) / LOG2E_F);
        float ra = sqrtf(tx * tx + ty * ty);
        float tmpa = atan2f(ty, tx) + spinvar * r, rz = zoomvar * r - 1.0f;
        *ox += ra * cosf(tmpa) + rz * tx; *oy += ra * sinf(tmpa) + rz * ty; break;
    }
    case 37: {  /* pie: slices rotation thickness */
        float slices = a[0];
        float sl = truncf(mwc_01(rng) * slices + 0.5f);
        float an = a[1] + 2.0f * PI_F * (sl + mwc_01(rng) * a[2]) / slices;
        float r = w * mwc_01(rng);
        *ox += r * cosf(an); *oy += r * sinf(an); break;
    }
    case 38: {  /* ngon: sides power circle corners */
        float power = a[1] * 0.5f, b = 2.0f * PI_F / a[0];
        float corners = a[3], circle = a[2];
        float r_factor = powf(tx * tx + ty * ty, power);
        float theta = atan2f(ty, tx);
        float phi = theta - b * floorf(theta / b);
        if (phi > b / 2.0f) phi -= b;
        float amp = (corners * (1.0f / cosf(phi) - 1.0f) + circle) / r_factor;
        *ox += w * tx * amp; *oy += w * ty * amp; break;
    }
    case 39: {  /* curl: c1 c2 */
        float re = 1.0f + a[0] * tx + a[1] * (tx * tx - ty * ty);
        float im = a[0] * ty + 2.0f * a[1] * tx * ty;
        float r = w / (re * re + im * im);
        *ox += r * (tx * re + ty * im); *oy += r * (ty * re - tx * im); break;
    }
    case 40: {  /* rectangles: x y */
        float rx = a[0], ry = a[1];
        *ox += w * ((rx == 0.0f) ? tx : rx * (2.0f * floorf(tx / rx) + 1.0f) - tx);
        *oy += w * ((ry == 0.0f) ? ty : ry * (2.0f * floorf(ty / ry) + 1.0f) - ty); break;
    }
    case 41: {
        float ang = mwc_01(rng) * w * PI_F;
        *ox += w * sinf(ang); *oy += w * sinf(ang) * sinf(ang) / cosf(ang); break;
    }
    case 42: *ox += w * sinf(tx) / cosf(ty); *oy += w * tanf(ty); break;
    case 43: *ox += w * (mwc_01(rng) - 0.5f); *oy += w * (mwc_01(rng) - 0.5f); break;
    case 44: {
        float ang = w * mwc_01(rng) * PI_F;
        float r = w / (tx * tx + ty * ty);
        float tanr = w * tanf(ang) * r;
        *ox += tanr * cosf(tx); *oy += tanr * sinf(ty); break;
    }
    case 45: {
        float r = mwc_01(rng) * w * sqrtf(tx * tx + ty * ty);
        *ox += w * tx * (cosf(r) + sinf(r)); *oy += w * tx * (cosf(r) - sinf(r)); break;
    }
    case 46: {
        float r = w * sqrtf(tx * tx + ty * ty), cr = cosf(r), icr = 1.0f / cr;
        icr += (cr < 0 ? 1 : -1);
        *ox += w * tx; *oy += w * icr; break;
    }
    case 48: {
        float s = tx * tx - ty * ty, r = w * sqrtf(1.0f / (s * s));
        *ox += r * tx; *oy += r * ty; break;
    }
    case 49: {  /* disc2: rot twist */
        float twist = a[1], rotpi = a[0] * PI_F;
        float sintwist = sinf(twist), costwist = cosf(twist) - 1.0f;
        if (twist > 2.0f * PI_F) { float k = 1.0f + twist - 2.0f * PI_F; sintwist *= k; costwist *= k; }
        if (twist < -2.0f * PI_F) { float k = 1.0f + twist + 2.0f * PI_F; sintwist *= k; costwist *= k; }
        float t = rotpi * (tx + ty), r = w * atan2f(tx, ty) / PI_F;
        *ox += r * (sinf(t) + costwist); *oy += r * (cosf(t) + sintwist); break;
    }
    case 50: {  /* super_shape: rnd m n1 n2 n3 holes */
        float ang = atan2f(ty, tx);
        float theta = 0.25f * (a[1] * ang + PI_F);
        float t1 = powf(fabsf(cosf(theta)), a[3]);
        float t2 = powf(fabsf(sinf(theta)), a[4]);
        float myrnd = a[0], d = sqrtf(tx * tx + ty * ty);
        float r = w * ((myrnd * mwc_01(rng) + (1.0f - myrnd) * d) - a[5])
                    * powf(t1 + t2, -1.0f / a[2]) / d;
        *ox += r * tx; *oy += r * ty; break;
    }
    case 51: {  /* flower: holes petals */
        float r = w * (mwc_01(rng) - a[0]) * cosf(a[1] * atan2f(ty, tx))
                    / sqrtf(tx * tx + ty * ty);
        *ox += r * tx; *oy += r * ty; break;
    }
    case 52: {  /* conic: holes eccentricity */
        float d = sqrtf(tx * tx + ty * ty), ct = tx / d;
        float r = w * (mwc_01(rng) - a[0]) * a[1] / (1.0f + a[1] * ct) / d;
        *ox += r * tx; *oy += r * ty; break;
    }
    case 53: {  /* parabola: height width */
        float r = sqrtf(tx * tx + ty * ty), sr = sinf(r), cr = cosf(r);
        *ox += a[0] * w * sr * sr * mwc_01(rng);
        *oy += a[1] * w * cr * mwc_01(rng); break;
    }
    case 54: {  /* bent2: x y */
        float nx = tx < 0.0f ? a[0] : 1.0f, ny = ty < 0.0f ? a[1] : 1.0f;
        *ox += w * nx * tx; *oy += w * ny * ty; break;
    }
    case 55: {  /* bipolar: shift */
        float x2y2 = tx * tx + ty * ty, t = x2y2 + 1.0f, x2 = tx * 2.0f;
        float ps = -PI_2_F * a[0];
        float y = 0.5f * atan2f(2.0f * ty, x2y2 - 1.0f) + ps;
        if (y > PI_2_F) y = -PI_2_F + fmodf(y + PI_2_F, PI_F);
        else if (y < -PI_2_F) y = PI_2_F - fmodf(PI_2_F - y, PI_F);
        *ox += w * 0.25f * TWO_PI_INV_F * logf((t + x2) / (t - x2));
        *oy += w * TWO_PI_INV_F * y; break;
    }
    case 56: {
        float roundX = rintf(tx), roundY = rintf(ty);
        float offsetX = tx - roundX, offsetY = ty - roundY;
        if (mwc_01(rng) > 0.75f) {
            *ox += w * (offsetX * 0.5f + roundX); *oy += w * (offsetY * 0.5f + roundY);
        } else if (fabsf(offsetX) >= fabsf(offsetY)) {
            if (offsetX >= 0.0f) {
                *ox += w * (offsetX * 0.5f + roundX + 0.25f);
                *oy += w * (offsetY * 0.5f + roundY + 0.25f * offsetY / offsetX);
            } else {
                *ox += w * (offsetX * 0.5f + roundX - 0.25f);
                *oy += w * (offsetY * 0.5f + roundY - 0.25f * offsetY / offsetX);
            }
        } else {
            if (offsetY >= 0.0f) {
                *oy += w * (offsetY * 0.5f + roundY + 0.25f);
                *ox += w * (offsetX * 0.5f + roundX + offsetX / offsetY * 0.25f);
            } else {
                *oy += w * (offsetY * 0.5f + roundY - 0.25f);
                *ox += w * (offsetX * 0.5f + roundX - offsetX / offsetY * 0.25f);
            }
        }
        break;
    }
    case 57: {
        float wx = w * 1.3029400317411197908970256609023f, y2 = ty * 2.0f;
        float r = wx * sqrtf(fabsf(ty * tx) / (tx * tx + y2 * y2));
        *ox += r * tx; *oy += r * y2; break;
    }
    case 58: {  /* cell: size */
        float cs = a[0], ics = 1.0f / cs;
        float x = floorf(tx * ics), y = floorf(ty * ics);
        float dx = tx - x * cs, dy = ty - y * cs;
        if (y >= 0.0f) {
            if (x >= 0.0f) { y *= 2.0f; x *= 2.0f; }
            else { y *= 2.0f; x = -(2.0f * x + 1.0f); }
        } else {
            if (x >= 0.0f) { y = -(2.0f * y + 1.0f); x *= 2.0f; }
            else { y = -(2.0f * y + 1.0f); x = -(2.0f * x + 1.0f); }
        }
        *ox += w * (dx + x * cs); *oy -= w * (dy + y * cs); break;
    }
    case 59: {  /* cpow: r i power */
        float an = atan2f(ty, tx), lnr = 0.5f * logf(tx * tx + ty * ty);
        float power = 1.0f / a[2], va = 2.0f * PI_F * power;
        float vc = a[0] * power, vd = a[1] * power;
        float ang = vc * an + vd * lnr + va * floorf(power * mwc_01(rng));
        float m = w * expf(vc * lnr - vd * an);
        *ox += m * cosf(ang); *oy += m * sinf(ang); break;
    }
    case 60:    /* curve: xamp yamp x2 y2 */
        *ox += w * (tx + a[0] * expf(-ty * ty * a[2]));
        *oy += w * (ty + a[1] * expf(-tx * tx * a[3])); break;
    case 61: {
        float tmp = tx * tx + ty * ty + 1.0f, tmp2 = 2.0f * tx;
        float r1 = sqrtf(tmp + tmp2), r2 = sqrtf(tmp - tmp2), xmax = (r1 + r2) * 0.5f;
        float a1 = logf(xmax + sqrtf(xmax - 1.0f)), a2 = -acosf(tx / xmax);
        float neww = w / 11.57034632f;
        float snv = sinf(a1), csv = cosf(a1);
        if (ty > 0.0f) snv = -snv;
        *ox += neww * coshf(a2) * csv; *oy += neww * sinhf(a2) * snv; break;
    }
    case 62: {
        float tmp = tx * tx + ty * ty + 1.0f, x2 = 2.0f * tx;
        float xmax = 0.5f * (sqrtf(tmp + x2) + sqrtf(tmp - x2));
        float aa = tx / xmax, b = 1.0f - aa * aa, ssx = xmax - 1.0f;
        float neww = w / PI_2_F;
        b = b < 0.0f ? 0.0f : sqrtf(b);
        ssx = ssx < 0.0f ? 0.0f : sqrtf(ssx);
        *ox += neww * atan2f(aa, b);
        if (ty > 0.0f) *oy += neww * logf(xmax + ssx);
        else *oy -= neww * logf(xmax + ssx);
        break;
    }
    case 63: {  /* escher: beta */
        float an = atan2f(ty, tx), lnr = 0.5f * logf(tx * tx + ty * ty);
        float seb = sinf(a[0]), ceb = cosf(a[0]);
        float vc = 0.5f * (1.0f + ceb), vd = 0.5f * seb;
        float m = w * expf(vc * lnr - vd * an), n = vc * an + vd * lnr;
        *ox += m * cosf(n); *oy += m * sinf(n); break;
    }
    case 64: {
        float expx = expf(tx) * 0.5f, expnx = 0.25f / expx;
        float sn = sinf(ty), cn = cosf(ty);
        float tmp = w / (expx + expnx - cn);
        *ox += tmp * (expx - expnx); *oy += tmp * sn; break;
    }
    case 65: {  /* lazysusan: x y twist space spin */
        float lx = a[0], ly = a[1];
        float x = tx - lx, y = ty + ly, r = sqrtf(x * x + y * y);
        if (r < w) {
            float an = atan2f(y, x) + a[4] + a[2] * (w - r);
            *ox += w * (r * cosf(an) + lx); *oy += w * (r * sinf(an) - ly);
        } else {
            r = 1.0f + a[3] / r;
            *ox += w * (r * x + lx); *oy += w * (r * y - ly);
        }
        break;
    }
    case 66: {
        float r2 = tx * tx + ty * ty, w2 = w * w;
        if (r2 < w2) { float r = w * sqrtf(w2 / r2 - 1.0f); *ox += r * tx; *oy += r * ty; }
        else { *ox += w * tx; *oy += w * ty; }
        break;
    }
    case 67: {
        float rndG = w * (mwc_01(rng) + mwc_01(rng) + mwc_01(rng) + mwc_01(rng) - 2.0f);
        float rndA = mwc_01(rng) * 2.0f * PI_F;
        *ptx = tx + rndG * cosf(rndA); *pty = ty + rndG * sinf(rndA); break;
    }
    case 68: {  /* modulus: x y */
        float mx = a[0], my = a[1], xr = 2.0f * mx, yr = 2.0f * my;
        if (tx > mx) *ox += w * (-mx + fmodf(tx + mx, xr));
        else if (tx < -mx) *ox += w * (mx - fmodf(mx - tx, xr));
        else *ox += w * tx;
        if (ty > my) *oy += w * (-my + fmodf(ty + my, yr));
        else if (ty < -my) *oy += w * (my - fmodf(my - ty, yr));
        else *oy += w * ty;
        break;
    }
    case 69: {  /* oscope: separation frequency amplitude damping */
        float tpf = 2.0f * PI_F * a[1];
        float t = a[2] * expf(-fabsf(tx) * a[3]) * cosf(tpf * tx) + a[0];
        *ox += w * tx;
        if (fabsf(ty) <= t) *oy -= w * ty; else *oy += w * ty;
        break;
    }
    case 70: {
        float p2v = w / PI_F;
        *ox += p2v * atan2f(tx, ty); *oy += 0.5f * p2v * logf(tx * tx + ty * ty); break;
    }
    case 71:    /* popcorn2: x y c */
        *ox += w * (tx + a[0] * sinf(tanf(ty * a[2])));
        *oy += w * (ty + a[1] * sinf(tanf(tx * a[2]))); break;
    case 72: {
        float t = tx * tx + ty * ty, r = 1.0f / (sqrtf(t) * (t + 1.0f / w));
        *ox += tx * r; *oy += ty * r; break;
    }
    case 73: {  /* separation: x xinside y yinside */
        float sx2 = a[0] * a[0], sy2 = a[2] * a[2];
        if (tx > 0.0f) *ox += w * (sqrtf(tx * tx + sx2) - tx * a[1]);
        else *ox -= w * (sqrtf(tx * tx + sx2) + tx * a[1]);
        if (ty > 0.0f) *oy += w * (sqrtf(ty * ty + sy2) - ty * a[3]);
        else *oy -= w * (sqrtf(ty * ty + sy2) + ty * a[3]);
        break;
    }
    case 74:    /* split: xsize ysize */
        if (cosf(tx * a[0] * PI_F) >= 0.0f) *oy += w * ty; else *oy -= w * ty;
        if (cosf(ty * a[1] * PI_F) >= 0.0f) *ox += w * tx; else *ox -= w * tx;
        break;
    case 75:    /* splits: x y */
        *ox += w * (tx + copysignf(a[0], tx)); *oy += w * (ty + copysignf(a[1], ty)); break;
    case 76: {  /* stripes: space warp */
        float roundx = floorf(tx + 0.5f), offsetx = tx - roundx;
        *ox += w * (offsetx * (1.0f - a[0]) + roundx);
        *oy += w * (ty + offsetx * offsetx * a[1]); break;
    }
    case 77: {  /* wedge: angle hole count swirl */
        float r = sqrtf(tx * tx + ty * ty);
        float an = atan2f(ty, tx) + a[3] * r;
        float wc = a[2], wa = a[0];
        float c = floorf((wc * an + PI_F) * INV_PI_F * 0.5f);
        float comp_fac = 1 - wa * wc * INV_PI_F * 0.5f;
        an = an * comp_fac + c * wa;
        r = w * (r + a[1]);
        *ox += r * cosf(an); *oy += r * sinf(an); break;
    }
    case 80: {  /* whorl: inside outside */
        float r = sqrtf(tx * tx + ty * ty), an = atan2f(ty, tx);
        if (r < w) an += a[0] / (w - r); else an += a[1] / (w - r);
        *ox += w * r * cosf(an); *oy += w * r * sinf(an); break;
    }
    case 81:    /* waves2: scalex scaley freqx freqy */
        *ox += w * (tx + a[0] * sinf(ty * a[2]));
        *oy += w * (ty + a[1] * sinf(tx * a[3])); break;
    case 82: { float e = expf(tx); *ox += w * e * cosf(ty); *oy += w * e * sinf(ty); break; }
    case 83: *ox += w * 0.5f * logf(tx * tx + ty * ty); *oy += w * atan2f(ty, tx); break;
    case 84: *ox += w * sinf(tx) * coshf(ty); *oy += w * cosf(tx) * sinhf(ty); break;
    case 85: *ox += w * cosf(tx) * coshf(ty); *oy -= w * sinf(tx) * sinhf(ty); break;
    case 86: {
        float d = 1.0f / (cosf(2.0f * tx) + coshf(2.0f * ty));
        *ox += w * d * sinf(2.0f * tx); *oy += w * d * sinhf(2.0f * ty); break;
    }
    case 87: {
        float d = 2.0f / (cosf(2.0f * tx) + coshf(2.0f * ty));
        *ox += w * d * cosf(tx) * coshf(ty); *oy += w * d * sinf(tx) * sinhf(ty); break;
    }
    case 88: {
        float d = 2.0f / (coshf(2.0f * ty) - cosf(2.0f * tx));
        *ox += w * d * sinf(tx) * coshf(ty); *oy -= w * d * cosf(tx) * sinhf(ty); break;
    }
    case 89: {
        float d = 1.0f / (coshf(2.0f * ty) - cosf(2.0f * tx));
        *ox += w * d * sinf(2.0f * tx); *oy += w * d * -1.0f * sinhf(2.0f * ty); break;
    }
    case 90: *ox += w * sinhf(tx) * cosf(ty); *oy += w * coshf(tx) * sinf(ty); break;
    case 91: *ox += w * coshf(tx) * cosf(ty); *oy += w * sinhf(tx) * sinf(ty); break;
    case 92: {
        float d = 1.0f / (cosf(2.0f * ty) + coshf(2.0f * tx));
        *ox += w * d * sinhf(2.0f * tx); *oy += w * d * sinf(2.0f * ty); break;
    }
    case 93: {
        float d = 2.0f / (cosf(2.0f * ty) + coshf(2.0f * tx));
        *ox += w * d * cosf(ty) * coshf(tx); *oy -= w * d * sinf(ty) * sinhf(tx); break;
    }
    case 94: {
        float d = 2.0f / (coshf(2.0f * tx) - cosf(2.0f * ty));
        *ox += w * d * sinhf(tx) * cosf(ty); *oy -= w * d * coshf(tx) * sinf(ty); break;
    }
    case 95: {
        float d = 1.0f / (coshf(2.0f * tx) - cosf(2.0f * ty));
        *ox += w * d * sinhf(2.0f * tx); *oy += w * d * sinf(2.0f * ty); break;
    }
    case 97: {  /* flux: spread */
        float xpw = tx + w, xmw = tx - w;
        float avgr = w * (2.0f + a[0])
                   * sqrtf(sqrtf(ty * ty + xpw * xpw) / sqrtf(ty * ty + xmw * xmw));
        float avga = (atan2f(ty, xmw) - atan2f(ty, xpw)) * 0.5f;
        *ox += avgr * cosf(avga); *oy += avgr * sinf(avga); break;
    }
    case 98: {  /* mobius: re_a im_a re_b im_b re_c im_c re_d im_d */
        float re_u = a[0] * tx - a[1] * ty + a[2], im_u = a[0] * ty + a[1] * tx + a[3];
        float re_v = a[4] * tx - a[5] * ty + a[6], im_v = a[4] * ty + a[5] * tx + a[7];
        float rad_v = w / (re_v * re_v + im_v * im_v);
        *ox += rad_v * (re_u * re_v + im_u * im_v);
        *oy += rad_v * (im_u * re_v - re_u * im_v); break;
    }
    default: break;
    }
}

/* One variation on arrays of points, for checking this file against the
 * reference's own variation bodies compiled for the CPU (oracle/build_ref.py). */
void oracle_variation(int id, const float *args, float w, float *txs, float *tys,
                      float *oxs, float *oys, uint32_t *seeds, int n) {
    for (int i = 0; i < n; i++) {
        mwc_t s = {seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2]};
        variation(id, args, w, &txs[i], &tys[i], &oxs[i], &oys[i], &s);
        seeds[3 * i + 1] = s.state;
        seeds[3 * i + 2] = s.carry;
    }
}

/* apply_xf (code/iter.py:121-149) */
static void apply_xform(const float *xf, float *x, float *y, float *color, mwc_t *rng) {
    const float *p = xf + XF_PRE;
    float tx = p[0] * *x + p[1] * *y + p[2];
    float ty = p[3] * *x + p[4] * *y + p[5];
    float ox = 0.0f, oy = 0.0f;
    int nv = (int)xf[XF_NVARS];
    for (int v = 0; v < nv; v++) {
        const float *rec = xf + XF_VARS + v * VAR_STRIDE;
        variation((int)rec[0], rec + 2, rec[1], &tx, &ty, &ox, &oy, rng);
    }
    if (xf[XF_HASPOST] != 0.0f) {
        const float *q = xf + XF_POST;
        tx = ox; ty = oy;
        ox = q[0] * tx + q[1] * ty + q[2];
        oy = q[3] * tx + q[4] * ty + q[5];
    }
    *x = ox; *y = oy;
    float csp = xf[XF_CSPEED];
    *color = *color * (1.0f - csp) + xf[XF_COLOR] * csp;
}

/* Apply one named xform to an array of points with given RNG streams: the
 * harness behind the per-variation parity tests. */
void oracle_apply_xform(const float *xf, float *xs, float *ys, float *cs,
                        uint32_t *seeds, int n) {
    for (int i = 0; i < n; i++) {
        mwc_t s = {seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2]};
        apply_xform(xf, &xs[i], &ys[i], &cs[i], &s);
        seeds[3 * i + 1] = s.state;
        seeds[3 * i + 2] = s.carry;
    }
}

/* cvt.rni.s32.f32: round to nearest even, saturate, NaN -> 0 (code/util.py:194-200) */
static inline int32_t rni_s32(float f) {
    if (isnan(f)) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)lrintf(f);
}

/* Camera transform + binning + palette index for given points
 * (code/iter.py:302-351).  bin = -1 when rejected by the bounds test. */
void oracle_point_to_bin(const float *cam, const float *xs, const float *ys,
                         const float *cs, const float *dithers, int n,
                         int astride, int aheight, int32_t *bins, int32_t *cidx) {
    for (int i = 0; i < n; i++) {
        float cx = fmaf(cam[0], xs[i], fmaf(cam[1], ys[i], cam[2]));
        float cy = fmaf(cam[3], xs[i], fmaf(cam[4], ys[i], cam[5]));
        uint32_t ix = (uint32_t)rni_s32(cx), iy = (uint32_t)rni_s32(cy);
        if (ix >= (uint32_t)astride || iy >= (uint32_t)aheight) bins[i] = -1;
        else bins[i] = (int32_t)(iy * (uint32_t)astride + ix);
        float cf = fmaf(cs[i], 255.0f, dithers[i]);
        int32_t ci;
        if (isnan(cf) || cf <= -0.5f) ci = 0;
        else if (cf >= 255.0f) ci = 255;
        else { ci = (int32_t)lrintf(cf); if (ci > 255) ci = 255; if (ci < 0) ci = 0; }
        cidx[i] = ci;
    }
}

static inline int bad_point(float x, float y) { return !isfinite(fabsf(x) + fabsf(y)); }

/*
 * The chaos game.  `ntraj` trajectories split the nsamples recorded iterations
 * into contiguous ranges; trajectory j owns RNG stream j, starts from a random
 * point, runs `fuse` unrecorded iterations, then its range.  Sample k uses
 * temporal sample (k * nts) / nsamples and palette row ts * pal_rows / nts.
 * hist is float4 [aheight][astride], accumulated with atomic float adds.
 *
 * xaos (NULL or float [nts][nxf][nxf-1]): cumulative densities of the next xform
 * given the previous one (precalc_chaos, code/iter.py:32-54; the chain of
 * iter.py:236-257); a trajectory starts with previous xform 0 (iter.py:209).
 * An xform whose record carries an opacity draws its point with that probability
 * (genome/specs.py:17); the point stays on the trajectory either way.
 */
void oracle_iterate_ex(const float *frames, int frame_stride, int nts, int nxf,
                       int has_final, const float *palette, int pal_rows,
                       float *hist, int astride, int aheight, uint32_t *seeds,
                       int ntraj, uint64_t nsamples, int fuse, int nthreads,
                       const float *xaos) {
    if (nthreads > 0) {
#ifdef _OPENMP
        extern void omp_set_num_threads(int);
        omp_set_num_threads(nthreads);
#endif
    }
    uint64_t per = (nsamples + (uint64_t)ntraj - 1) / (uint64_t)ntraj;
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = 0; j < ntraj; j++) {
        uint64_t k0 = (uint64_t)j * per;
        uint64_t k1 = k0 + per < nsamples ? k0 + per : nsamples;
        if (k0 >= k1) continue;
        mwc_t rng = {seeds[3 * j], seeds[3 * j + 1], seeds[3 * j + 2]};
        float color_dither = 0.49f * mwc_11(&rng);
        float x = mwc_11(&rng), y = mwc_11(&rng), c = mwc_01(&rng);
        int last = 0;
        for (int64_t k = -(int64_t)fuse; k < (int64_t)(k1 - k0); k++) {
            uint64_t ks = k0 + (k < 0 ? 0 : (uint64_t)k);
            int ts = (int)((ks * (uint64_t)nts) / nsamples);
            const float *fr = frames + (size_t)ts * frame_stride;
            if (bad_point(x, y)) { x = mwc_11(&rng); y = mwc_11(&rng); c = mwc_01(&rng); }
            float sel = mwc_01(&rng);
            const float *den = xaos ? xaos + ((size_t)ts * nxf + last) * (nxf - 1) : fr + FR_DEN;
            int pick = nxf - 1;
            for (int i = 0; i < nxf - 1; i++)
                if (sel <= den[i]) { pick = i; break; }
            const float *xf = fr + FR_XF + pick * XF_FLOATS;
            apply_xform(xf, &x, &y, &c, &rng);
            last = pick;
            int visible = 1;
            if (xf[XF_OPACITY] != 0.0f) {
                float op = xf[XF_OPACITY + 1];
                visible = op >= 1.0f || mwc_01(&rng) < op;
            }
            if (k < 0 || !visible) continue;

            float fx = x, fy = y, fc = c;
            if (has_final) apply_xform(fr + FR_XF + nxf * XF_FLOATS, &fx, &fy, &fc, &rng);
            const float *cam = fr + FR_CAM;
            float cx = fmaf(cam[0], fx, fmaf(cam[1], fy, cam[2]));
            float cy = fmaf(cam[3], fx, fmaf(cam[4], fy, cam[5]));
            uint32_t ix = (uint32_t)rni_s32(cx), iy = (uint32_t)rni_s32(cy);
            if (ix >= (uint32_t)astride || iy >= (uint32_t)aheight) continue;
            float cf = fmaf(fc, 255.0f, color_dither);
            int ci = (isnan(cf) || cf <= -0.5f) ? 0 : (cf >= 255.0f ? 255 : (int)lrintf(cf));
            if (ci > 255) ci = 255;
            const float *pe = palette + ((size_t)(ts * pal_rows / nts) * 256 + ci) * 4;
            float *h = hist + ((size_t)iy * astride + ix) * 4;
            for (int q = 0; q < 4; q++) {
#pragma omp atomic
                h[q] += pe[q];
            }
        }
        seeds[3 * j + 1] = rng.state;
        seeds[3 * j + 2] = rng.carry;
    }
}

void oracle_iterate(const float *frames, int frame_stride, int nts, int nxf,
                    int has_final, const float *palette, int pal_rows,
                    float *hist, int astride, int aheight, uint32_t *seeds,
                    int ntraj, uint64_t nsamples, int fuse, int nthreads) {
    oracle_iterate_ex(frames, frame_stride, nts, nxf, has_final, palette, pal_rows, hist,
                      astride, aheight, seeds, ntraj, nsamples, fuse, nthreads, NULL);
}
