/* b200pt_detmath.h — elementary functions with ONE bit pattern per input on every machine.
 *
 * The reference evaluates sin / cos / tan / asin / acos / atan / pow / log / exp in GLSL, whose results are whatever the
 * GPU's special-function unit returns (shaders/random.glsl, transform.glsl, raytrace.rgen).  A CUDA port would call
 * CUDA's libm and a CPU checker glibc's: both are within 1-2 ulp, but not of each other, and one different last bit in
 * a sampled direction decides stochastic branches further down a path (rejection loops, Fresnel coin, an ill-conditioned
 * silhouette test), so 0.5-15 % of the pixels of a same-seed render used to differ between the kernels and the oracle.
 *
 * These versions use only IEEE-754 double add / sub / mul / div / sqrt and integer operations on the bits — no fused
 * multiply-add (compile with contraction off: nvcc -fmad=false, gcc -ffp-contract=off), no libm — in a fixed order, and
 * round the double result to float once.  They are therefore (a) bit-identical between the sm_100a kernels and any
 * host compiler, and (b) correctly rounded floats except for ~1e-6 of the inputs (double evaluation error 1e-15
 * against the float rounding boundary), i.e. at least as faithful to the mathematical function as any libm.
 * B200 issues FP64 at half the FP32 rate, so a 30-operation double kernel costs about what CUDA's accurate float
 * sinf / powf expansions cost.  tests/test_detmath.py measures the error against mpmath-grade references.
 *
 * Included by the device code (rtx-pathtracer_b200/csrc/device_math.cuh) and by the CPU oracle (oracle/): the same
 * source, so the two sides of every parity test agree on these functions by construction, and disagree only where
 * the path-tracing logic itself differs. */
#ifndef B200PT_DETMATH_H
#define B200PT_DETMATH_H
#include <stdint.h>
#include <string.h>
#include <math.h>

#ifdef __CUDACC__
#define DM_HD __host__ __device__ inline
#else
#define DM_HD inline
#endif

namespace b200pt_dm {

DM_HD uint64_t dBits(double d) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b; memcpy(&b, &d, 8); return b;
#endif
}
DM_HD double bitsD(uint64_t b) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double d; memcpy(&d, &b, 8); return d;
#endif
}
/* round to nearest integer, ties to even, for |x| < 2^51 (the classic 1.5 * 2^52 trick is exact IEEE arithmetic) */
DM_HD double rne(double x) { const double big = 6755399441055744.0; return (x + big) - big; }

/* sin and cos of x (radians), |x| up to ~1e6 (angles here are within a few turns) */
DM_HD void sincosD(double x, double *s, double *c) {
    if (!(fabs(x) < 1.0e6)) { const double n = x - x; *s = n; *c = n; return; }      /* inf / NaN / out of range -> NaN (0 for finite huge) */
    const double kd = rne(x * 0.6366197723675814);
    const double r = (x - kd * 1.5707963267341256) - kd * 6.077100506506192e-11;     /* pi/2 in two pieces; kd * hi is exact */
    const double r2 = r * r;
    const double sp = -0.16666666666666666 + r2 * (0.008333333333333333 + r2 * (-0.0001984126984126984 + r2 * (2.7557319223985893e-06 + r2 * (-2.505210838544172e-08 +
                      r2 * (1.6059043836821613e-10 + r2 * (-7.647163731819816e-13 + r2 * 2.8114572543455206e-15))))));
    const double sr = r + r * r2 * sp;
    const double cp = -0.5 + r2 * (0.041666666666666664 + r2 * (-0.001388888888888889 + r2 * (2.48015873015873e-05 + r2 * (-2.755731922398589e-07 +
                      r2 * (2.08767569878681e-09 + r2 * (-1.1470745597729725e-11 + r2 * (4.779477332387385e-14 + r2 * -1.5619206968586225e-16)))))));
    const double cr = 1.0 + r2 * cp;
    const long long k = (long long)kd;
    switch ((int)(k & 3)) {
        case 0: *s = sr; *c = cr; break;
        case 1: *s = cr; *c = -sr; break;
        case 2: *s = -sr; *c = -cr; break;
        default: *s = -cr; *c = sr; break;
    }
}

/* atan for x >= 0 */
DM_HD double atanPos(double ax) {
    const bool inv = ax > 1.0;
    if (inv) ax = 1.0 / ax;
    ax = ax / (1.0 + sqrt(1.0 + ax * ax));       /* two angle halvings: argument <= tan(pi/16) */
    ax = ax / (1.0 + sqrt(1.0 + ax * ax));
    const double t2 = ax * ax;
    const double p = -0.3333333333333333 + t2 * (0.2 + t2 * (-0.14285714285714285 + t2 * (0.1111111111111111 + t2 * (-0.09090909090909091 + t2 * (0.07692307692307693 +
                     t2 * (-0.06666666666666667 + t2 * (0.058823529411764705 + t2 * (-0.05263157894736842 + t2 * (0.047619047619047616 + t2 * (-0.043478260869565216 +
                     t2 * (0.04 + t2 * -0.037037037037037035)))))))))));
    const double a = 4.0 * (ax + ax * t2 * p);
    return inv ? 1.5707963267948966 - a : a;
}
DM_HD double atanD(double x) {
    if (x != x) return x;
    const double a = atanPos(fabs(x));
    return x < 0.0 ? -a : a;
}
DM_HD double atan2D(double y, double x) {
    if (x != x || y != y) return x + y;
    if (x == 0.0) return y > 0.0 ? 1.5707963267948966 : (y < 0.0 ? -1.5707963267948966 : 0.0);
    const double a = atanPos(fabs(y / x));                     /* in [0, pi/2] */
    const double q = x > 0.0 ? a : 3.141592653589793 - a;      /* angle of (|y|, x) */
    return y < 0.0 ? -q : q;
}

/* natural logarithm of a positive finite double that came from a float (so never subnormal as a double) */
DM_HD double logD(double x) {
    if (x != x || x < 0.0) return x - x + (x != x ? x : bitsD(0x7ff8000000000000ull));
    if (x == 0.0) return -bitsD(0x7ff0000000000000ull);
    const uint64_t b = dBits(x);
    if ((b >> 52) == 0x7ffull) return x;                       /* +inf */
    int e = (int)((b >> 52) & 0x7ffull) - 1023;
    double m = bitsD((b & 0x000fffffffffffffull) | 0x3ff0000000000000ull);     /* [1, 2) */
    if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
    const double s = (m - 1.0) / (m + 1.0);
    const double s2 = s * s;
    const double p = 0.3333333333333333 + s2 * (0.2 + s2 * (0.14285714285714285 + s2 * (0.1111111111111111 + s2 * (0.09090909090909091 + s2 * (0.07692307692307693 +
                     s2 * (0.06666666666666667 + s2 * (0.058823529411764705 + s2 * (0.05263157894736842 + s2 * (0.047619047619047616 + s2 * 0.043478260869565216)))))))));
    const double lm = 2.0 * (s + s * s2 * p);
    const double ed = (double)e;
    return ed * 0.6931471801362932 + (ed * 4.236521365809284e-10 + lm);
}

DM_HD double expD(double x) {
    if (x != x) return x;
    if (x > 709.0) return bitsD(0x7ff0000000000000ull);
    if (x < -745.0) return 0.0;
    const double kd = rne(x * 1.4426950408889634);
    const double r = (x - kd * 0.6931471801362932) - kd * 4.236521365809284e-10;
    const double p = 0.5 + r * (0.16666666666666666 + r * (0.041666666666666664 + r * (0.008333333333333333 + r * (0.001388888888888889 + r * (0.0001984126984126984 +
                     r * (2.48015873015873e-05 + r * (2.7557319223985893e-06 + r * (2.755731922398589e-07 + r * (2.505210838544172e-08 + r * (2.08767569878681e-09 +
                     r * (1.6059043836821613e-10 + r * 1.1470745597729725e-11)))))))))));
    const double er = 1.0 + (r + r * r * p);
    long long k = (long long)kd;
    /* 2^k in two factors so that results in the subnormal range of double (never reached from float use) stay finite */
    const long long k1 = k / 2, k2 = k - k1;
    return er * bitsD((uint64_t)(k1 + 1023) << 52) * bitsD((uint64_t)(k2 + 1023) << 52);
}

/* ---- float front ends: one rounding at the end ---- */
DM_HD float sinF(float x) { double s, c; sincosD((double)x, &s, &c); return (float)s; }
DM_HD float cosF(float x) { double s, c; sincosD((double)x, &s, &c); return (float)c; }
DM_HD float tanF(float x) { double s, c; sincosD((double)x, &s, &c); return (float)(s / c); }
DM_HD float atanF(float x) { return (float)atanD((double)x); }
DM_HD float atan2F(float y, float x) { return (float)atan2D((double)y, (double)x); }
DM_HD float asinF(float x) {
    const double d = (double)x;
    if (!(fabs(d) <= 1.0)) return (float)(d - d + bitsD(0x7ff8000000000000ull));
    return (float)atan2D(d, sqrt((1.0 - d) * (1.0 + d)));
}
DM_HD float acosF(float x) {
    const double d = (double)x;
    if (!(fabs(d) <= 1.0)) return (float)(d - d + bitsD(0x7ff8000000000000ull));
    return (float)atan2D(sqrt((1.0 - d) * (1.0 + d)), d);
}
DM_HD float logF(float x) { return (float)logD((double)x); }
DM_HD float expF(float x) { return (float)expD((double)x); }
/* GLSL pow(x, y) is defined for x > 0, and for x == 0 with y > 0; the rest follows C powf: pow(x, 0) = 1, pow(0, y < 0) = inf,
 * negative base with an integral exponent = +-|x|^y, otherwise NaN */
DM_HD float powF(float x, float y) {
    const double dx = (double)x, dy = (double)y;
    if (dx != dx || dy != dy) return (float)(dx + dy);
    if (dy == 0.0) return 1.0f;
    if (dx == 0.0) return dy > 0.0 ? 0.0f : (float)bitsD(0x7ff0000000000000ull);
    if (dx < 0.0) {      /* like C powf: a negative base is defined for integral exponents only */
        if (!(fabs(dy) < 9.0e15) || rne(dy) != dy) return (float)bitsD(0x7ff8000000000000ull);
        const double r = expD(dy * logD(-dx));
        const double half = dy * 0.5;
        return (float)(rne(half) != half ? -r : r);
    }
    return (float)expD(dy * logD(dx));
}

}  /* namespace b200pt_dm */
#endif /* B200PT_DETMATH_H */
