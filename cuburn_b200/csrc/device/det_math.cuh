// Deterministic elementary functions for the interpolation stage.
//
// The packed genome parameters must be reproducible bit-for-bit by a CPU
// restatement, so this stage avoids libdevice / fast-math transcendentals.
// Every function below is a fixed sequence of IEEE-754 binary64 add / mul / div
// operations (issued through the __d*_rn intrinsics so the compiler can never
// contract them into FMAs) followed by one rounding to binary32.  The same
// sequences are written in numpy float64 in oracle/flame_ref.py.
#pragma once

#define DM(a, b) __dmul_rn((a), (b))
#define DA(a, b) __dadd_rn((a), (b))
#define DS(a, b) __dsub_rn((a), (b))
#define DD(a, b) __ddiv_rn((a), (b))

// f32 helpers with explicit single rounding (never fused)
#define FM(a, b) __fmul_rn((a), (b))
#define FA(a, b) __fadd_rn((a), (b))
#define FS(a, b) __fsub_rn((a), (b))
#define FD(a, b) __fdiv_rn((a), (b))

// log2 of a positive normal float.
__device__ __forceinline__ float det_log2f(float x) {
    unsigned int bits = __float_as_uint(x);
    int e = (int)(bits >> 23) - 127;
    float m = __uint_as_float((bits & 0x007fffffu) | 0x3f800000u);   // [1, 2)
    if (m > 1.41421354f) { m = FM(m, 0.5f); e += 1; }
    double md = (double)m;
    double s = DD(DS(md, 1.0), DA(md, 1.0));
    double z = DM(s, s);
    double p = 1.0 / 23.0;
    p = DA(DM(p, z), 1.0 / 21.0);
    p = DA(DM(p, z), 1.0 / 19.0);
    p = DA(DM(p, z), 1.0 / 17.0);
    p = DA(DM(p, z), 1.0 / 15.0);
    p = DA(DM(p, z), 1.0 / 13.0);
    p = DA(DM(p, z), 1.0 / 11.0);
    p = DA(DM(p, z), 1.0 / 9.0);
    p = DA(DM(p, z), 1.0 / 7.0);
    p = DA(DM(p, z), 1.0 / 5.0);
    p = DA(DM(p, z), 1.0 / 3.0);
    p = DA(DM(p, z), 1.0);
    double ln_m = DM(DM(2.0, s), p);
    double r = DA((double)e, DM(ln_m, 1.4426950408889634));
    return (float)r;
}

// 2^v for float v (result rounded to float; saturates outside 2^+-1000).
__device__ __forceinline__ float det_exp2f(float v) {
    double vd = (double)v;
    double n = floor(DA(vd, 0.5));
    double f = DS(vd, n);
    double t = DM(f, 0.6931471805599453);
    double p = 1.0 / 6227020800.0;          // 1/13!
    p = DA(DM(p, t), 1.0 / 479001600.0);
    p = DA(DM(p, t), 1.0 / 39916800.0);
    p = DA(DM(p, t), 1.0 / 3628800.0);
    p = DA(DM(p, t), 1.0 / 362880.0);
    p = DA(DM(p, t), 1.0 / 40320.0);
    p = DA(DM(p, t), 1.0 / 5040.0);
    p = DA(DM(p, t), 1.0 / 720.0);
    p = DA(DM(p, t), 1.0 / 120.0);
    p = DA(DM(p, t), 1.0 / 24.0);
    p = DA(DM(p, t), 1.0 / 6.0);
    p = DA(DM(p, t), 0.5);
    p = DA(DM(p, t), 1.0);
    p = DA(DM(p, t), 1.0);
    if (!(n == n)) return __int_as_float(0x7fc00000);            // NaN in -> NaN out
    double nc = fmin(fmax(n, -1000.0), 1000.0);
    long long ebits = ((long long)nc + 1023LL) << 52;
    double r = DM(p, __longlong_as_double(ebits));
    return (float)r;
}

// sin and cos of a float angle in radians.
__device__ __forceinline__ void det_sincosf(float x, float *sn, float *cs) {
    double xd = (double)x;
    double k = floor(DA(DM(xd, 0.6366197723675814), 0.5));
    double r = DS(DS(xd, DM(k, 1.57079632673412561417e+00)),
                  DM(k, 6.07710050650619224932e-11));
    double z = DM(r, r);
    double ps = 1.0 / 355687428096000.0;     // 1/17!
    ps = DS(DM(ps, z), 1.0 / 1307674368000.0);   // -1/15!
    ps = DA(DM(ps, z), 1.0 / 6227020800.0);      // +1/13!
    ps = DS(DM(ps, z), 1.0 / 39916800.0);        // -1/11!
    ps = DA(DM(ps, z), 1.0 / 362880.0);          // +1/9!
    ps = DS(DM(ps, z), 1.0 / 5040.0);            // -1/7!
    ps = DA(DM(ps, z), 1.0 / 120.0);             // +1/5!
    ps = DS(DM(ps, z), 1.0 / 6.0);               // -1/3!
    ps = DA(DM(ps, z), 1.0);
    double s = DM(r, ps);
    double pc = 1.0 / 20922789888000.0;      // 1/16!
    pc = DS(DM(pc, z), 1.0 / 87178291200.0);     // -1/14!
    pc = DA(DM(pc, z), 1.0 / 479001600.0);       // +1/12!
    pc = DS(DM(pc, z), 1.0 / 3628800.0);         // -1/10!
    pc = DA(DM(pc, z), 1.0 / 40320.0);           // +1/8!
    pc = DS(DM(pc, z), 1.0 / 720.0);             // -1/6!
    pc = DA(DM(pc, z), 1.0 / 24.0);              // +1/4!
    pc = DS(DM(pc, z), 0.5);                     // -1/2!
    pc = DA(DM(pc, z), 1.0);
    double c = pc;
    long long q = ((long long)k) & 3LL;
    double so, co;
    if (q == 0)      { so = s;  co = c;  }
    else if (q == 1) { so = c;  co = -s; }
    else if (q == 2) { so = -s; co = -c; }
    else             { so = -c; co = s;  }
    *sn = (float)so;
    *cs = (float)co;
}
