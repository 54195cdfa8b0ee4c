/* oracle/trig.c -- TEST INFRASTRUCTURE, not product code.
 *
 * The transcendentals of the BSDF samplers (reference model/brdf.py:28-29,50-51 asin / acos of the sampled
 * polar angle, utils/ops.py:32-44 sin / cos of (theta, phi)) as explicit fp32 polynomial code: every step is a
 * single IEEE operation (mul, add, fmaf, sqrtf), so this file (gcc -ffp-contract=off, fmaf from libm) and
 * iris_b200/csrc/trig.cuh (__fmaf_rn / __fmul_rn / __fadd_rn) produce the same bits for the same input.
 *
 * Why: the reference takes these from whatever libm its tensors live on (Sleef inside torch on a CPU,
 * libdevice on a GPU); those agree to ~1 ulp, not bit for bit, and a 1-ulp change of a sampled direction moves
 * the secondary hit point by ~1e-7 -- visible through a hash grid whose finest cells are 4e-5 wide.  With one
 * shared definition the oracle and the CUDA path trace bit-identical secondary rays.
 *
 * Argument reduction is Cody-Waite in three parts of pi/2; the minimax coefficients are the classic
 * single-precision ones (Cephes sinf / cosf / asinf, Moshier).  Accuracy <= 1.6 ulp (sin, cos, acos) and <= 2.5 ulp (asin) on the ranges the samplers
 * use (theta in [0, pi/2], phi in [0, 2 pi], asin / acos arguments in [0, 1]); tests/test_oracle_cpu.py checks
 * it against libm in double precision.
 */
#include <math.h>
#include <stdint.h>

#define PIO2_1 1.5703125f                    /* pi/2 = PIO2_1 + PIO2_2 + PIO2_3 (+ 2^-48) */
#define PIO2_2 4.837512969970703125e-4f
#define PIO2_3 7.54978995489188216e-8f
#define TWO_OVER_PI 0.636619772367581343f
#define PIO2_F 1.57079632679489662f
#define PI_F 3.14159265358979324f

static inline void sincos_f32(float x, float *s, float *c) {
    const float k = rintf(x * TWO_OVER_PI);
    float r = fmaf(-k, PIO2_1, x);
    r = fmaf(-k, PIO2_2, r);
    r = fmaf(-k, PIO2_3, r);
    const float z = r * r;
    float p = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
    p = fmaf(p, z, -1.6666654611e-1f);
    const float sn = fmaf(p, z * r, r);
    float q = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    q = fmaf(q, z, 4.166664568298827e-2f);
    const float cs = fmaf(q, z * z, fmaf(-0.5f, z, 1.0f));
    const int n = (int)k & 3;
    const float a = (n & 1) ? cs : sn, b = (n & 1) ? sn : cs;
    *s = (n & 2) ? -a : a;
    *c = ((n + 1) & 2) ? -b : b;
}

static inline float asin_poly(float z) {
    float p = fmaf(4.2163199048e-2f, z, 2.4181311049e-2f);
    p = fmaf(p, z, 4.5470025998e-2f);
    p = fmaf(p, z, 7.4953002686e-2f);
    p = fmaf(p, z, 1.6666752422e-1f);
    return p;
}

static inline float asin_f32(float x) {
    const float a = fabsf(x);
    float r;
    if (!(a <= 1.0f)) return NAN;
    if (a > 0.5f) {
        const float z = 0.5f * (1.0f - a);
        const float s = sqrtf(z);
        const float t = fmaf(s * z, asin_poly(z), s);
        r = fmaf(-2.0f, t, PIO2_F);
    } else {
        const float z = a * a;
        r = fmaf(a * z, asin_poly(z), a);
    }
    return x < 0.0f ? -r : r;
}

static inline float acos_f32(float x) {
    if (!(fabsf(x) <= 1.0f)) return NAN;
    if (x > 0.5f) {
        const float z = 0.5f * (1.0f - x);
        const float s = sqrtf(z);
        return 2.0f * fmaf(s * z, asin_poly(z), s);
    }
    if (x < -0.5f) {
        const float z = 0.5f * (1.0f + x);
        const float s = sqrtf(z);
        return fmaf(-2.0f, fmaf(s * z, asin_poly(z), s), PI_F);
    }
    const float z = x * x;
    return PIO2_F - fmaf(x * z, asin_poly(z), x);
}

/* IEEE (correctly rounded) square root.  torch's CPU sqrt kernel is NOT correctly rounded (AVX-512 build of torch 2.11: 0.6 % of
 * fp32 inputs come back one ulp off; measured against sqrt in double precision), CUDA's sqrtf / torch.sqrt on a GPU and C's
 * sqrtf are.  The samplers take sqrt(u) before asin / acos and for the emitter barycentrics, so the checker uses this one. */
void oracle_sqrt(const float *x, int64_t n, float *y) {
    for (int64_t i = 0; i < n; ++i) y[i] = sqrtf(x[i]);
}
void oracle_sincos(const float *x, int64_t n, float *s, float *c) {
    for (int64_t i = 0; i < n; ++i) sincos_f32(x[i], s + i, c + i);
}
void oracle_asin(const float *x, int64_t n, float *y) {
    for (int64_t i = 0; i < n; ++i) y[i] = asin_f32(x[i]);
}
void oracle_acos(const float *x, int64_t n, float *y) {
    for (int64_t i = 0; i < n; ++i) y[i] = acos_f32(x[i]);
}
