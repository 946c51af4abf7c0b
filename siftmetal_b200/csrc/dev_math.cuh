// dev_math.cuh — device transcendentals of the arithmetic spec (DESIGN.md §"Arithmetic spec").
//
// Fixed sequences of IEEE binary32 operations (Cephes single-precision kernels) replacing the
// Metal built-ins the reference calls under fast-math (SIFTGradient.metal:36 atan2,
// SIFTOrientation.metal:115 / SIFTDescriptor.metal:211 exp, SIFTDescriptor.metal:147-148
// cos/sin, :158 pow). Every operation is an explicit round-to-nearest intrinsic, so the result
// does not depend on -fmad and is bit-identical to the CPU oracle's restatement of the same
// sequences (checked on the GPU by tests/test_gpu_parity.py::test_device_math).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sift {

__device__ __forceinline__ float dm_expf(float x) {
    if (x < -87.0f) return 0.0f;
    float t = __fmaf_rn(x, 1.44269504088896341f, 12582912.0f);
    float n = __fsub_rn(t, 12582912.0f);
    float r = __fmaf_rn(n, -0.693359375f, x);
    r = __fmaf_rn(n, 2.12194440e-4f, r);
    float z = __fmul_rn(r, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    p = __fmaf_rn(p, z, r);
    p = __fadd_rn(p, 1.0f);
    int ni = (int)n;
    float s = __int_as_float((ni + 127) << 23);
    return __fmul_rn(p, s);
}

__device__ __forceinline__ float dm_exp2f(float x) {
    return dm_expf(__fmul_rn(x, 0.693147180559945309f));
}

// Branch-free: the two range-reduction forms share one division (selected operands), and the
// zero case is a final select — same operations on the same values as the oracle's branches.
__device__ __forceinline__ float dm_atan2f(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const bool big = mn > __fmul_rn(mx, 0.414213562373095049f);
    const float num = big ? __fsub_rn(mn, mx) : mn;
    const float den = big ? __fadd_rn(mn, mx) : mx;
    const float off = big ? 0.785398163397448310f : 0.0f;
    const float a = __fdiv_rn(num, den);
    const float z = __fmul_rn(a, a);
    float p = 8.05374449538e-2f;
    p = __fmaf_rn(p, z, -1.38776856032e-1f);
    p = __fmaf_rn(p, z, 1.99777106478e-1f);
    p = __fmaf_rn(p, z, -3.33329491539e-1f);
    p = __fmul_rn(p, z);
    float r = __fmaf_rn(p, a, a);
    r = __fadd_rn(r, off);
    if (ay > ax) r = __fsub_rn(1.57079632679489662f, r);
    if (x < 0.0f) r = __fsub_rn(3.14159265358979324f, r);
    if (y < 0.0f) r = -r;
    return mx == 0.0f ? 0.0f : r;
}

__device__ __forceinline__ void dm_sincosf(float x, float* s, float* c) {
    float ax = fabsf(x);
    int j = (int)__fmul_rn(ax, 1.27323954473516f);
    j = (j + 1) & ~1;
    float y = (float)j;
    float r = __fmaf_rn(y, -0.78515625f, ax);
    r = __fmaf_rn(y, -2.4187564849853515625e-4f, r);
    r = __fmaf_rn(y, -3.77489497744594108e-8f, r);
    float z = __fmul_rn(r, r);
    float ps = -1.9515295891e-4f;
    ps = __fmaf_rn(ps, z, 8.3321608736e-3f);
    ps = __fmaf_rn(ps, z, -1.6666654611e-1f);
    float sp = __fmaf_rn(__fmul_rn(ps, z), r, r);
    float pc = 2.443315711809948e-5f;
    pc = __fmaf_rn(pc, z, -1.388731625493765e-3f);
    pc = __fmaf_rn(pc, z, 4.166664568298827e-2f);
    float cp = __fmaf_rn(__fmul_rn(pc, z), z, __fmaf_rn(-0.5f, z, 1.0f));
    int q = (j >> 1) & 3;
    float sv, cv;
    if (q == 0) { sv = sp; cv = cp; }
    else if (q == 1) { sv = cp; cv = -sp; }
    else if (q == 2) { sv = -sp; cv = -cp; }
    else { sv = -cp; cv = sp; }
    if (x < 0.0f) sv = -sv;
    *s = sv;
    *c = cv;
}

}  // namespace sift
