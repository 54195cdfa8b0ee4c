// trig.cuh -- sin / cos / asin / acos of the BSDF samplers (reference model/brdf.py:28-29,50-51, utils/ops.py:32-44) as explicit fp32
// polynomial code, one IEEE operation per step (never contracted, never reassociated).  The reference takes these from the libm of
// whatever device its tensors live on (Sleef in torch on a CPU, libdevice on a GPU): ~1 ulp apart, which moves a secondary hit
// point by ~1e-7 -- visible through a hash grid with 4e-5 cells.  Defining them here, with the same operation sequence as the
// CPU checker, makes the sampled directions (and so every secondary ray) reproducible bit for bit; it is also cheaper than
// libdevice's sincosf/asinf/acosf (no slow-path range reduction: theta in [0, pi/2], phi in [0, 2 pi]).
// Cody-Waite reduction by pi/2 in three parts; single-precision minimax coefficients (Cephes sinf / cosf / asinf).
// Accuracy: <= 1.6 ulp (sin, cos, acos), <= 2.5 ulp (asin) on those ranges.
#pragma once
#include "common.cuh"

__device__ __forceinline__ void iris_sincosf(float x, float &s, float &c) {
    const float k = rintf(__fmul_rn(x, 0.636619772367581343f));
    float r = __fmaf_rn(-k, 1.5703125f, x);
    r = __fmaf_rn(-k, 4.837512969970703125e-4f, r);
    r = __fmaf_rn(-k, 7.54978995489188216e-8f, r);
    const float z = __fmul_rn(r, r);
    float p = __fmaf_rn(-1.9515295891e-4f, z, 8.3321608736e-3f);
    p = __fmaf_rn(p, z, -1.6666654611e-1f);
    const float sn = __fmaf_rn(p, __fmul_rn(z, r), r);
    float q = __fmaf_rn(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    q = __fmaf_rn(q, z, 4.166664568298827e-2f);
    const float cs = __fmaf_rn(q, __fmul_rn(z, z), __fmaf_rn(-0.5f, z, 1.0f));
    const int n = (int)k & 3;
    const float a = (n & 1) ? cs : sn, b = (n & 1) ? sn : cs;
    s = (n & 2) ? -a : a;
    c = ((n + 1) & 2) ? -b : b;
}

__device__ __forceinline__ float iris_asin_poly(float z) {
    float p = __fmaf_rn(4.2163199048e-2f, z, 2.4181311049e-2f);
    p = __fmaf_rn(p, z, 4.5470025998e-2f);
    p = __fmaf_rn(p, z, 7.4953002686e-2f);
    return __fmaf_rn(p, z, 1.6666752422e-1f);
}

__device__ __forceinline__ float iris_asinf(float x) {
    const float a = fabsf(x);
    if (!(a <= 1.0f)) return __int_as_float(0x7fc00000);
    float r;
    if (a > 0.5f) {
        const float z = __fmul_rn(0.5f, __fsub_rn(1.0f, a));
        const float s = __fsqrt_rn(z);
        const float t = __fmaf_rn(__fmul_rn(s, z), iris_asin_poly(z), s);
        r = __fmaf_rn(-2.0f, t, 1.57079632679489662f);
    } else {
        const float z = __fmul_rn(a, a);
        r = __fmaf_rn(__fmul_rn(a, z), iris_asin_poly(z), a);
    }
    return x < 0.0f ? -r : r;
}

__device__ __forceinline__ float iris_acosf(float x) {
    if (!(fabsf(x) <= 1.0f)) return __int_as_float(0x7fc00000);
    if (x > 0.5f) {
        const float z = __fmul_rn(0.5f, __fsub_rn(1.0f, x));
        const float s = __fsqrt_rn(z);
        return __fmul_rn(2.0f, __fmaf_rn(__fmul_rn(s, z), iris_asin_poly(z), s));
    }
    if (x < -0.5f) {
        const float z = __fmul_rn(0.5f, __fadd_rn(1.0f, x));
        const float s = __fsqrt_rn(z);
        return __fmaf_rn(-2.0f, __fmaf_rn(__fmul_rn(s, z), iris_asin_poly(z), s), 3.14159265358979324f);
    }
    const float z = __fmul_rn(x, x);
    return __fsub_rn(1.57079632679489662f, __fmaf_rn(__fmul_rn(x, z), iris_asin_poly(z), x));
}
