// oracle_math.h — TEST INFRASTRUCTURE (see sift_oracle.cpp header). Deterministic float32
// transcendentals of the arithmetic spec (DESIGN.md §"Arithmetic spec").
//
// The reference evaluates exp / atan2 / sin / cos / pow with Metal built-ins under Metal's
// default fast-math (SIFTGradient.metal:36, SIFTOrientation.metal:115, SIFTDescriptor.metal:
// 147-158,211), whose rounding is unspecified and unreproducible off Apple hardware. The spec
// therefore fixes each of them as a short sequence of IEEE-754 binary32 +, −, ×, ÷, fma — the
// classic Cephes single-precision kernels — so that a CPU and a GPU evaluate bit-identical
// values (a 1-ulp wobble in atan2 would flip nearest-bin votes of the 36-bin histogram).
// tests/test_oracle_math.py checks every function against glibc to ≤ 4 ulp.
//
// Every operation below is written out; the file must be compiled with -ffp-contract=off so
// that only the fmaf() calls fuse.
#ifndef ORACLE_MATH_H
#define ORACLE_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

// e^x for x in [-87, 88]; 0 below. Cephes expf: n = rint(x·log2e), r = x − n·ln2 (two-piece),
// degree-5 polynomial in r, scale by 2^n through the exponent field.
static inline float om_expf(float x) {
    if (x < -87.0f) return 0.0f;
    float t = fmaf(x, 1.44269504088896341f, 12582912.0f);  // + 1.5·2^23 rounds to an integer
    float n = t - 12582912.0f;
    float r = fmaf(n, -0.693359375f, x);
    r = fmaf(n, 2.12194440e-4f, r);
    float z = r * r;
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    p = fmaf(p, z, r);
    p = p + 1.0f;
    int32_t ni = (int32_t)n;
    uint32_t bits = (uint32_t)(ni + 127) << 23;
    float s;
    memcpy(&s, &bits, 4);
    return p * s;
}

// 2^x as e^(x·ln2).
static inline float om_exp2f(float x) { return om_expf(x * 0.693147180559945309f); }

// atan2(y, x) in (−π, π]; atan2(0, 0) = 0. One division: a = min/max, folded to
// (min − max)/(min + max) + π/4 above tan(π/8); Cephes atanf polynomial.
static inline float om_atan2f(float y, float x) {
    float ax = fabsf(x), ay = fabsf(y);
    float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    if (mx == 0.0f) return 0.0f;
    float a, off;
    if (mn > mx * 0.414213562373095049f) {
        a = (mn - mx) / (mn + mx);
        off = 0.785398163397448310f;
    } else {
        a = mn / mx;
        off = 0.0f;
    }
    float z = a * a;
    float p = 8.05374449538e-2f;
    p = fmaf(p, z, -1.38776856032e-1f);
    p = fmaf(p, z, 1.99777106478e-1f);
    p = fmaf(p, z, -3.33329491539e-1f);
    p = p * z;
    float r = fmaf(p, a, a);
    r = r + off;
    if (ay > ax) r = 1.57079632679489662f - r;
    if (x < 0.0f) r = 3.14159265358979324f - r;
    if (y < 0.0f) r = -r;
    return r;
}

// sin and cos for |x| ≤ 8192. Cephes sinf/cosf: octant j (made even), three-piece π/4
// reduction, degree-3 polynomials in r².
static inline void om_sincosf(float x, float* s, float* c) {
    float ax = fabsf(x);
    int32_t j = (int32_t)(ax * 1.27323954473516f);
    j = (j + 1) & ~1;
    float y = (float)j;
    float r = fmaf(y, -0.78515625f, ax);
    r = fmaf(y, -2.4187564849853515625e-4f, r);
    r = fmaf(y, -3.77489497744594108e-8f, r);
    float z = r * r;
    float ps = -1.9515295891e-4f;
    ps = fmaf(ps, z, 8.3321608736e-3f);
    ps = fmaf(ps, z, -1.6666654611e-1f);
    float sp = fmaf(ps * z, r, r);
    float pc = 2.443315711809948e-5f;
    pc = fmaf(pc, z, -1.388731625493765e-3f);
    pc = fmaf(pc, z, 4.166664568298827e-2f);
    float cp = fmaf(pc * z, z, fmaf(-0.5f, z, 1.0f));
    int32_t q = (j >> 1) & 3;
    float sv, cv;
    if (q == 0) { sv = sp; cv = cp; }
    else if (q == 1) { sv = cp; cv = -sp; }
    else if (q == 2) { sv = -sp; cv = -cp; }
    else { sv = -cp; cv = sp; }
    if (x < 0.0f) sv = -sv;
    *s = sv;
    *c = cv;
}

#endif  // ORACLE_MATH_H
