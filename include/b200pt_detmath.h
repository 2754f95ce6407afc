/* b200pt_detmath.h — elementary functions with ONE bit pattern per input on every machine.
 *
 * The reference evaluates sin / cos / tan / asin / acos / atan / pow / log / exp in GLSL, whose results are whatever the
 * GPU's special-function unit returns (shaders/random.glsl, transform.glsl, raytrace.rgen).  A CUDA port would call
 * CUDA's libm and a CPU checker glibc's: both are within 1-2 ulp, but not of each other, and one different last bit in
 * a sampled direction decides stochastic branches further down a path (rejection loops, Fresnel coin, an ill-conditioned
 * silhouette test), so 0.5-15 % of the pixels of a same-seed render used to differ between the kernels and the oracle.
 *
 * These versions use only IEEE-754 add / sub / mul / div / sqrt and integer operations on the bits — no fused
 * multiply-add (compile with contraction off: nvcc -fmad=false, gcc -ffp-contract=off), no libm — in a fixed order.
 * They are therefore bit-identical between the sm_100a kernels and any host compiler.  The hot functions (sin, cos,
 * tan, asin, acos, atan, atan2, log, exp) are single-precision kernels of ~20 operations, accurate to 1-3 ulp; pow is
 * exp(y log x) of those, like GLSL's.  A first version evaluated everything in double: bit-identical too, but B200's
 * FP64 pipe made sponzaXML frames 1.4-2x slower (profiles/r02_detmath.txt); the only double code left is the range
 * reduction of sin / cos for |x| >= 1e4, which the tracer never reaches.
 * tests/test_detmath.py measures the errors against float64 references.
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

/* ---- single-precision kernels: the hot functions (one per bounce and per vMF lobe) ------------------------------------
 * Same rules — IEEE float add / sub / mul / div / sqrt and bit operations only, fixed order — at about 20 FP32 operations
 * each (CUDA's accurate sinf / expf / logf are 40-100, the double kernels above ~100 issue slots).  Accuracy 1-3 ulp
 * (tests/test_detmath.py), which is what GLSL implementations deliver at best; bit-identity between device and host is
 * what matters here and holds by construction. */
DM_HD uint32_t fBits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t b; memcpy(&b, &f, 4); return b;
#endif
}
DM_HD float bitsF(uint32_t b) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}
DM_HD float rnef(float x) { const float big = 12582912.0f; return (x + big) - big; }      /* |x| < 2^22 */

DM_HD void sincosF(float x, float *s, float *c) {
    if (!(fabsf(x) < 1.0e4f)) { double ds, dc; sincosD((double)x, &ds, &dc); *s = (float)ds; *c = (float)dc; return; }
    const float kf = rnef(x * 0.636619747f);
    /* pi/2 = 1.5703125 + 4.837512969970703125e-4 + 7.54978995489188216e-8 (Cody-Waite: the first two products are exact) */
    float r = x - kf * 1.5703125f;
    r = r - kf * 4.837512969970703125e-4f;
    r = r - kf * 7.54978995489188216e-8f;
    const float r2 = r * r;
    const float sp = -0.166666672f + r2 * (0.00833333377f + r2 * (-0.000198412701f + r2 * 2.75573188e-06f));
    const float sr = r + r * r2 * sp;
    const float cp = 0.0416666679f + r2 * (-0.00138888892f + r2 * (2.48015876e-05f + r2 * -2.75573199e-07f));
    const float cr = (1.0f - 0.5f * r2) + r2 * r2 * cp;
    switch (((int)kf) & 3) {
        case 0: *s = sr; *c = cr; break;
        case 1: *s = cr; *c = -sr; break;
        case 2: *s = -sr; *c = -cr; break;
        default: *s = -cr; *c = sr; break;
    }
}
DM_HD float sinF(float x) { float s, c; sincosF(x, &s, &c); return s; }
DM_HD float cosF(float x) { float s, c; sincosF(x, &s, &c); return c; }
DM_HD float tanF(float x) { float s, c; sincosF(x, &s, &c); return s / c; }

/* atan on [0, inf): reduce to |t| <= tan(pi/8) with one division (x > tan(3pi/8): -1/x, x > tan(pi/8): (x-1)/(x+1)), then
 * t + t^3 P(t^2), P = degree-4 Chebyshev fit of (atan(t)/t - 1)/t^2 on [0, tan^2(pi/8)] (own fit, |error| 1.6e-8) */
DM_HD float atanPosF(float ax) {
    float y0, t;
    if (ax > 2.41421366f) { y0 = 1.57079637f; t = -1.0f / ax; }
    else if (ax > 0.414213568f) { y0 = 0.785398185f; t = (ax - 1.0f) / (ax + 1.0f); }
    else { y0 = 0.0f; t = ax; }
    const float z = t * t;
    const float p = -0.333333318f + z * (0.199995399f + z * (-0.142639562f + z * (0.107437313f + z * -0.0645192862f)));
    return y0 + (t + t * z * p);
}
DM_HD float atanF(float x) {
    if (x != x) return x;
    const float a = atanPosF(fabsf(x));
    return x < 0.0f ? -a : a;
}
DM_HD float atan2F(float y, float x) {
    if (x != x || y != y) return x + y;
    if (x == 0.0f) return y > 0.0f ? 1.57079637f : (y < 0.0f ? -1.57079637f : 0.0f);
    const float a = atanPosF(fabsf(y / x));
    const float q = x > 0.0f ? a : 3.14159274f - a;
    return y < 0.0f ? -q : q;
}
DM_HD float asinF(float x) {
    if (!(fabsf(x) <= 1.0f)) return bitsF(0x7fc00000u);
    return atan2F(x, sqrtf((1.0f - x) * (1.0f + x)));
}
DM_HD float acosF(float x) {
    if (!(fabsf(x) <= 1.0f)) return bitsF(0x7fc00000u);
    return atan2F(sqrtf((1.0f - x) * (1.0f + x)), x);
}

DM_HD float logF(float x) {
    if (x != x || x < 0.0f) return bitsF(0x7fc00000u);
    if (x == 0.0f) return bitsF(0xff800000u);
    uint32_t b = fBits(x);
    if ((b >> 23) == 0xffu) return x;                           /* +inf */
    int e = 0;
    if ((b >> 23) == 0u) { x = x * 8388608.0f; b = fBits(x); e = -23; }      /* subnormal */
    e += (int)(b >> 23) - 127;
    float m = bitsF((b & 0x007fffffu) | 0x3f800000u);           /* [1, 2) */
    if (m > 1.41421354f) { m = m * 0.5f; e += 1; }
    const float s = (m - 1.0f) / (m + 1.0f);
    const float s2 = s * s;
    const float p = s2 * (0.333333343f + s2 * (0.2f + s2 * (0.142857149f + s2 * 0.111111112f)));
    const float lm = 2.0f * (s + s * p);
    const float ef = (float)e;
    return ef * 0.693359375f + (ef * -2.12194440e-4f + lm);      /* ln 2 in two pieces; ef * hi is exact */
}

DM_HD float expF(float x) {
    if (!(fabsf(x) <= 87.0f)) {                                 /* NaN, overflow, underflow and subnormal results: off the fast path */
        if (x != x) return x;
        if (x > 88.8f) return bitsF(0x7f800000u);
        if (x < -104.0f) return 0.0f;
    }
    const float kf = rnef(x * 1.44269502f);
    const float r = (x - kf * 0.693359375f) - kf * -2.12194440e-4f;
    const float p = r * r * (0.5f + r * (0.166666672f + r * (0.0416666679f + r * (0.00833333377f + r * (0.00138888892f + r * 0.000198412701f)))));
    const float er = 1.0f + (r + p);                            /* in (0.70, 1.42) */
    const int k = (int)kf;
    if (k >= -125 && k <= 126) return bitsF(fBits(er) + ((uint32_t)k << 23));      /* er * 2^k: exact, the result is normal */
    const int k1 = k / 2, k2 = k - k1;                          /* 2^k in two normal factors (the result may be subnormal) */
    return er * bitsF((uint32_t)(k1 + 127) << 23) * bitsF((uint32_t)(k2 + 127) << 23);
}

/* ---- pow ------------------------------------------------------------------------------------------------------------------
 * exp(y log x) in single precision, which is how GLSL defines pow (exp2(y * log2(x)), GLSL 4.60 spec 8.2) and what the
 * reference's GPU evaluates: the relative error grows with |y ln x| (about (1 + |y ln x|) ulp: 1e-6 for a Phong lobe
 * of exponent 100 at 25 degrees).  More accurate versions were measured and dropped: log x carried as two floats (<= 4 ulp for
 * exponents up to 2000) cost sponzaXML — one Phong material — 9 % of its frame rate, a double-precision kernel 30-50 %
 * (B200's FP64 pipe); bit-identity between device and host, the point of this header, holds for all three. */
DM_HD float powPosF(float x, float y) { return expF(y * logF(x)); }
/* GLSL pow(x, y) is defined for x > 0, and for x == 0 with y > 0; the rest follows C powf: pow(x, 0) = 1, pow(0, y < 0) = inf,
 * negative base with an integral exponent = +-|x|^y, otherwise NaN */
DM_HD float powF(float x, float y) {
    if (x != x || y != y) return x + y;
    if (y == 0.0f) return 1.0f;
    if (x == 0.0f) return y > 0.0f ? 0.0f : bitsF(0x7f800000u);
    if (fabsf(y) > 3.0e38f) {                                   /* y = +-inf */
        const float ax = fabsf(x);
        if (ax == 1.0f) return 1.0f;
        return ((ax > 1.0f) == (y > 0.0f)) ? bitsF(0x7f800000u) : 0.0f;
    }
    if (fabsf(x) > 3.0e38f) return (y > 0.0f) ? (x > 0.0f ? x : bitsF(0x7f800000u)) : 0.0f;      /* x = +-inf (sign of odd powers ignored) */
    if (x < 0.0f) {      /* like C powf: a negative base is defined for integral exponents only */
        if (fabsf(y) >= 16777216.0f) return powPosF(-x, y);     /* every float that large is an even integer */
        if (rnef(y) != y) return bitsF(0x7fc00000u);
        const float r = powPosF(-x, y);
        const float half = y * 0.5f;
        return rnef(half) != half ? -r : r;
    }
    return powPosF(x, y);
}

}  /* namespace b200pt_dm */
#endif /* B200PT_DETMATH_H */
