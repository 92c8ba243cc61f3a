// Device-side interpolation of the packed genome: Catmull-Rom rows, the precalc
// program, and the dithered palette table.
//
// Replaces interp_iter_params and interp_palette_flat
// (cuburn/code/interp.py:235-271, 372-433) and the precalc hunks templated into
// them (cuburn/code/iter.py:12-30,56-95; cuburn/code/variations.py precalcs).
// Instead of generating one interpolation kernel per genome, the genome's
// structure is data: a table of knot rows and a small program of precalc ops
// (cuburn_b200/code/packer.py), interpreted here.  All arithmetic is explicit
// single-rounded IEEE (device/det_math.cuh) so oracle/flame_ref.py can match the
// packed parameters bit for bit.
#include "cb_common.h"
#include "device/det_math.cuh"
#include "device/mwc.cuh"

#define KNOTS 32

// Rightmost index i in [0, 32) with hay[i] < needle, by 5 halving steps
// (bitwise_binsearch, code/util.py:220-229).
__device__ __forceinline__ int knot_search(const float *hay, float needle) {
    int lo = 0;
#pragma unroll
    for (int step = KNOTS / 2; step > 0; step >>= 1)
        if (needle > hay[lo + step]) lo += step;
    return lo;
}

#define MAG_ELBOW 0.0625f
#define MAG_OFFSET 5.0f

__device__ __forceinline__ float mag_fwd(float x) {
    if (x > MAG_ELBOW) return FA(det_log2f(x), MAG_OFFSET);
    if (x < -MAG_ELBOW) return -FA(det_log2f(-x), MAG_OFFSET);
    return FD(x, MAG_ELBOW);
}

__device__ __forceinline__ float mag_inv(float v) {
    if (v >= 1.0f) return det_exp2f(FS(v, MAG_OFFSET));
    if (v <= -1.0f) return -det_exp2f(FS(-v, MAG_OFFSET));
    return FM(v, MAG_ELBOW);
}

__device__ __forceinline__ float mag_slope(float x, float m) {
    if (x >= MAG_ELBOW) return FD(m, x);
    if (x <= -MAG_ELBOW) return FD(m, -x);
    return FD(m, MAG_ELBOW);
}

// catmull_rom_base (interp.py:318-355)
__device__ float spline_eval(const float *times, const float *knots, float t, bool mag) {
    int idx = max(knot_search(times, t), 1);
    int i3 = min(idx + 2, KNOTS - 1);
    float t1 = times[idx];
    float t2 = FS(times[idx + 1], t1);
    float rt2 = FD(1.0f, t2);
    float t0 = FM(FS(times[idx - 1], t1), rt2);
    float t3 = FM(FS(times[i3], t1), rt2);
    float u = FM(FS(t, t1), rt2);

    float k0 = knots[idx - 1], k1 = knots[idx], k2 = knots[idx + 1], k3 = knots[i3];
    float m1 = FD(FS(k2, k0), FS(1.0f, t0));
    float m2 = FD(FS(k3, k1), t3);
    if (mag) {
        m1 = mag_slope(k1, m1);
        m2 = mag_slope(k2, m2);
        k1 = mag_fwd(k1);
        k2 = mag_fwd(k2);
    }
    float uu = FM(u, u), uuu = FM(uu, u);
    float b1 = FA(FS(uuu, FM(2.0f, uu)), u);                    // u^3 - 2u^2 + u
    float b2 = FA(FS(FM(2.0f, uuu), FM(3.0f, uu)), 1.0f);       // 2u^3 - 3u^2 + 1
    float b3 = FS(uuu, uu);                                     // u^3 - u^2
    float b4 = FA(FM(-2.0f, uuu), FM(3.0f, uu));                // -2u^3 + 3u^2
    float r = FA(FA(FA(FM(m1, b1), FM(k1, b2)), FM(m2, b3)), FM(k2, b4));
    if (mag) r = mag_inv(r);
    return r;
}

__device__ __forceinline__ float sample_time(float tstart, float tstep, int i) {
    return __fmaf_rn((float)i, tstep, tstart);
}

__global__ void __launch_bounds__(256)
k_interp_rows(float *vals, const float *times, const float *knots,
              const int *row_mag, int nrows, float tstart, float tstep, int nts) {
    int total = nrows * nts;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += gridDim.x * blockDim.x) {
        int ts = i / nrows, row = i - ts * nrows;
        float t = sample_time(tstart, tstep, ts);
        vals[i] = spline_eval(times + row * KNOTS, knots + row * KNOTS, t,
                              row_mag[row] != 0);
    }
}

// Program word layout: {op, out_slot, in[0..7], aux0, aux1}
#define PROG_W 12
enum { OP_DIRECT = 0, OP_DIRECT_MAG, OP_AFFINE, OP_CAMERA, OP_DENSITY,
       OP_WAVES, OP_PERSPECTIVE, OP_JULIAN_CN, OP_CURVE, OP_XAOS };

#define DEG2RAD(a) FD(FM((a), 3.14159274101257f), 180.0f)

#define OPS_PER_THREAD 8
__global__ void __launch_bounds__(128)
k_interp_params(float *params, int stride, const float *vals, int nrows,
                const int *prog, int nops, cb_dims dim, int nts) {
    // one thread per (temporal sample, group of OPS_PER_THREAD ops): ops are independent
    // (each reads spline rows and writes its own slots)
    int ts = blockIdx.x * blockDim.x + threadIdx.x;
    if (ts >= nts) return;
    const float *v = vals + (size_t)ts * nrows;
    float *out = params + (size_t)ts * stride;
    const int o0 = blockIdx.y * OPS_PER_THREAD;
    const int o1 = min(o0 + OPS_PER_THREAD, nops);
    for (int o = o0; o < o1; o++) {
        const int *w = prog + o * PROG_W;
        int op = w[0], dst = w[1];
        const int *in = w + 2;
        switch (op) {
        case OP_DIRECT:
        case OP_DIRECT_MAG:
            out[dst] = v[in[0]];
            break;
        case OP_AFFINE: {
            // inputs: angle, spread, magnitude.x, magnitude.y, offset.x, offset.y
            // (precalc_xf_affine, code/iter.py:81-95); outputs xx xy xo yx yy yo
            float pri = DEG2RAD(v[in[0]]), spr = DEG2RAD(v[in[1]]);
            float magx = v[in[2]], magy = v[in[3]];
            float sm, cm, sp, cp;
            det_sincosf(FS(pri, spr), &sm, &cm);
            det_sincosf(FA(pri, spr), &sp, &cp);
            out[dst + 0] = FM(magx, cm);
            out[dst + 1] = FM(-magy, cp);
            out[dst + 2] = v[in[4]];
            out[dst + 3] = FM(-magx, sm);
            out[dst + 4] = FM(magy, sp);
            out[dst + 5] = -v[in[5]];
            break;
        }
        case OP_CAMERA: {
            // inputs: rotation, center.x, center.y, scale (precalc_camera,
            // code/iter.py:56-79); outputs xx xy xo yx yy yo
            float rs, rc;
            det_sincosf(DEG2RAD(v[in[0]]), &rs, &rc);
            float cenx = v[in[1]], ceny = v[in[2]];
            float scale = FM(v[in[3]], (float)dim.width);
            out[dst + 0] = FM(scale, rc);
            out[dst + 1] = FM(scale, -rs);
            out[dst + 2] = FA(FM(scale, FS(FM(rs, ceny), FM(rc, cenx))),
                              FM(0.5f, (float)dim.awidth));
            out[dst + 3] = FM(scale, rs);
            out[dst + 4] = FM(scale, rc);
            out[dst + 5] = FA(FM(scale, -FA(FM(rs, cenx), FM(rc, ceny))),
                              FM(0.5f, (float)dim.aheight));
            break;
        }
        case OP_DENSITY: {
            // in[0] = first weight row, aux0 = xform count; writes count-1
            // cumulative normalised weights (precalc_densities, iter.py:12-30)
            int n = w[10];
            float sum = 0.0f;
            for (int k = 0; k < n; k++) sum = FA(sum, v[in[0] + k]);
            float rsum = FD(1.0f, sum);
            sum = 0.0f;
            for (int k = 0; k < n - 1; k++) {
                sum = FA(sum, FM(v[in[0] + k], rsum));
                out[dst + k] = sum;
            }
            break;
        }
        case OP_XAOS: {
            // in[0] = first weight row, in[1] = first row of the previous xform's chaos
            // multipliers, aux0 = xform count; writes count-1 cumulative normalised
            // weight * chaos products (precalc_chaos, iter.py:32-54)
            int n = w[10];
            float sum = 0.0f;
            for (int k = 0; k < n; k++) sum = FA(sum, FM(v[in[0] + k], v[in[1] + k]));
            float rsum = FD(1.0f, sum);
            sum = 0.0f;
            for (int k = 0; k < n - 1; k++) {
                sum = FA(sum, FM(FM(v[in[0] + k], v[in[1] + k]), rsum));
                out[dst + k] = sum;
            }
            break;
        }
        case OP_WAVES: {
            float dx = v[in[0]], dy = v[in[1]];
            out[dst + 0] = FD(1.0f, FA(FM(dx, dx), 1.0e-20f));
            out[dst + 1] = FD(1.0f, FA(FM(dy, dy), 1.0e-20f));
            break;
        }
        case OP_PERSPECTIVE: {
            float ang = FM(v[in[0]], 1.57079637050629f);
            float pdist = fmaxf(1e-9f, v[in[1]]);
            float sn, cs;
            det_sincosf(ang, &sn, &cs);
            out[dst + 0] = pdist;
            out[dst + 1] = sn;
            out[dst + 2] = FM(pdist, cs);
            break;
        }
        case OP_JULIAN_CN:
            out[dst] = FD(v[in[0]], FM(2.0f, v[in[1]]));
            break;
        case OP_CURVE: {
            float xl = v[in[0]], yl = v[in[1]];
            out[dst + 0] = FD(1.0f, fmaxf(1e-20f, FM(xl, xl)));
            out[dst + 1] = FD(1.0f, fmaxf(1e-20f, FM(yl, yl)));
            break;
        }
        default:
            break;
        }
    }
}

// One block per palette row, one thread per colour index.
__global__ void __launch_bounds__(256)
k_interp_palette(float4 *out, mwc_st *seeds, const float *ptimes,
                 const float4 *pals, float tstart, float tstep) {
    int row = blockIdx.x, c = threadIdx.x;
    int sid = row * 256 + c;
    mwc_st rng = seeds[sid];

    float t = sample_time(tstart, tstep, row);
    int idx = max(knot_search(ptimes, t) + 1, 1);
    float tr = ptimes[idx];
    float lf = FD(FS(tr, t), FS(tr, ptimes[idx - 1]));
    float rf = FS(1.0f, lf);
    float4 left = pals[(idx - 1) * 256 + c];
    float4 right = left;
    if (tr > 1.0f) {
        lf = 1.0f;
        rf = 0.0f;
    } else {
        right = pals[idx * 256 + c];
    }
    // JPEG full-range RGB->YUV (code/color.py:18-23), blended in YUV
    float ly = FA(FA(FM(0.299f, left.x), FM(0.587f, left.y)), FM(0.114f, left.z));
    float lu = FA(FS(FM(-0.168736f, left.x), FM(0.331264f, left.y)), FM(0.5f, left.z));
    float lv = FS(FS(FM(0.5f, left.x), FM(0.418688f, left.y)), FM(0.081312f, left.z));
    float ry = FA(FA(FM(0.299f, right.x), FM(0.587f, right.y)), FM(0.114f, right.z));
    float ru = FA(FS(FM(-0.168736f, right.x), FM(0.331264f, right.y)), FM(0.5f, right.z));
    float rv = FS(FS(FM(0.5f, right.x), FM(0.418688f, right.y)), FM(0.081312f, right.z));
    float y = FA(FM(ly, lf), FM(ry, rf));
    float u = FA(FA(FM(lu, lf), FM(ru, rf)), 0.5f);
    float v = FA(FA(FM(lv, lf), FM(rv, rf)), 0.5f);

    float qy = FA(FM(y, 255.0f), FM(0.49f, mwc_next_11(rng)));
    float qu = FA(FM(u, 255.0f), FM(0.49f, mwc_next_11(rng)));
    float qv = FA(FM(v, 255.0f), FM(0.49f, mwc_next_11(rng)));
    // truncate toward zero, saturate to [0, 255]
    float iy = fminf(255.0f, fmaxf(0.0f, truncf(qy)));
    float iu = fminf(255.0f, fmaxf(0.0f, truncf(qu)));
    float iv = fminf(255.0f, fmaxf(0.0f, truncf(qv)));
    const float k = 1.0f / 255.0f;
    out[sid] = make_float4(FM(iy, k), FM(iu, k), FM(iv, k), 1.0f);
    seeds[sid] = rng;
}

extern "C" {

int cb_interp_rows(cb_dptr vals, cb_dptr times, cb_dptr knots, cb_dptr row_mag,
                   int nrows, float tstart, float tstep, int nts, cb_stream s) {
    CB_REQUIRE(nrows > 0 && nts > 0, "empty interpolation request");
    int total = nrows * nts;
    int grid = (total + 255) / 256;
    int cap = cb_sm_count() * 8;
    if (grid > cap) grid = cap;
    k_interp_rows<<<grid, 256, 0, cb_cs(s)>>>(
        cb_ptr<float>(vals), cb_ptr<const float>(times), cb_ptr<const float>(knots),
        cb_ptr<const int>(row_mag), nrows, tstart, tstep, nts);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_interp_params(cb_dptr params, int param_stride, cb_dptr vals, int nrows,
                     cb_dptr program, int nops, const cb_dims *dim, int nts,
                     cb_stream s) {
    CB_REQUIRE(dim && nts > 0 && nops >= 0, "bad interp_params request");
    if (nops == 0) return CB_OK;
    const dim3 grid((nts + 127) / 128, (nops + OPS_PER_THREAD - 1) / OPS_PER_THREAD);
    k_interp_params<<<grid, 128, 0, cb_cs(s)>>>(
        cb_ptr<float>(params), param_stride, cb_ptr<const float>(vals), nrows,
        cb_ptr<const int>(program), nops, *dim, nts);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_interp_palette(cb_dptr palette_out, cb_dptr seeds, cb_dptr ptimes,
                      cb_dptr pals, float tstart, float tstep, int nrows_out,
                      cb_stream s) {
    CB_REQUIRE(nrows_out > 0, "no palette rows");
    k_interp_palette<<<nrows_out, 256, 0, cb_cs(s)>>>(
        cb_ptr<float4>(palette_out), cb_ptr<mwc_st>(seeds),
        cb_ptr<const float>(ptimes), cb_ptr<const float4>(pals), tstart, tstep);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

}  // extern "C"
