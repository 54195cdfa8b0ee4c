// shading.cuh -- per-lane shading math: frames, BSDF samplers / evaluation (+ analytic Jacobian), emitter sampling
// and evaluation, nearest-voxel SLF lookup, and the uniform-sample source.  Restates, for one lane, what the reference
// expresses as chains of ATen kernels: model/brdf.py:20-59,78-210, utils/ops.py:12-82, model/emitter.py:180-255,
// model/slf.py:41-70 (paths relative to the reference tree).
#pragma once
#include "../../include/iris_b200.h"
#include "common.cuh"
#include "traverse.cuh"
#include "trig.cuh"

// ------------------------------------------------------------------------------------------------ samples
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Four uniforms in [0,1) with 24 random bits (same lattice as torch.rand for fp32): columns 4*block .. 4*block+3.
__device__ __forceinline__ float4 sample4(const IrisSampler &s, int64_t lane, int block) {
    if (s.U != nullptr) {
        const float *p = s.U + lane * (int64_t)s.stride + 4 * block;
        float4 r;
        r.x = 4 * block + 0 < s.stride ? p[0] : 0.f;
        r.y = 4 * block + 1 < s.stride ? p[1] : 0.f;
        r.z = 4 * block + 2 < s.stride ? p[2] : 0.f;
        r.w = 4 * block + 3 < s.stride ? p[3] : 0.f;
        return r;
    }
    const uint64_t l = (uint64_t)lane + s.lane_offset;
    uint32_t o[4];
    philox4x32_10((uint32_t)l, (uint32_t)(l >> 32), (uint32_t)block, 0u, (uint32_t)s.seed, (uint32_t)(s.seed >> 32), o);
    const float k = 5.9604644775390625e-08f;   // 2^-24
    return make_float4((float)(o[0] >> 8) * k, (float)(o[1] >> 8) * k, (float)(o[2] >> 8) * k, (float)(o[3] >> 8) * k);
}

// ------------------------------------------------------------------------------------------------ frames + samplers
// Everything between the uniforms and a sampled direction is one IEEE rounding per operation, in the order of the reference's
// ATen chain (products rounded before they are summed; the 1x3 @ 3x3 frame change accumulates left to right), with the
// transcendentals of trig.cuh: the CPU checker restates the same sequence, so secondary rays agree bit for bit.
// torch.cross evaluates a1*b2 - a2*b1 inside ONE kernel, where the compiler contracts it to fma(a1, b2, -round(a2*b1)) -- on the CPU
// build (checked bit for bit) and, by nvcc's own contraction rule, on the CUDA build the reference runs on.  Spelled out here.
__device__ __forceinline__ f3 cross_torch(f3 a, f3 b) {
    return mk3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)), __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}
// utils/ops.py:12-30
__device__ __forceinline__ void normal_space(f3 n, f3 &t, f3 &b) {
    const f3 a = fabsf(n.x) <= 0.1f ? mk3(1.f, 0.f, 0.f) : mk3(0.f, 1.f, 0.f);
    t = normalize_nf(cross_torch(a, n));
    b = cross_torch(n, t);
}
// utils/ops.py:32-44 followed by the frame change of model/brdf.py:32-33
__device__ __noinline__ f3 sphere_to_world(float theta, float phi, f3 n) {
    float st, ct, sp, cp;
    iris_sincosf(theta, st, ct);
    iris_sincosf(phi, sp, cp);
    const f3 l = normalize_nf(mk3(xmul(st, cp), xmul(st, sp), ct));
    f3 t, b;
    normal_space(n, t, b);
    return mk3(xadd(xadd(xmul(l.x, t.x), xmul(l.y, b.x)), xmul(l.z, n.x)), xadd(xadd(xmul(l.x, t.y), xmul(l.y, b.y)), xmul(l.z, n.y)),
               xadd(xadd(xmul(l.x, t.z), xmul(l.y, b.z)), xmul(l.z, n.z)));
}
// model/brdf.py:20-34
__device__ __forceinline__ f3 diffuse_sampler(float u0, float u1, f3 n) {
    return sphere_to_world(iris_asinf(__fsqrt_rn(u0)), xmul(IRIS_PI * 2.f, u1), n);
}
// model/brdf.py:36-59
__device__ __forceinline__ f3 specular_sampler(float u0, float u1, float roughness, f3 wo, f3 n) {
    // (1-u0)/(u0*(alpha^2-1)+1) cancels catastrophically for small roughness: one rounding per op, in the reference's
    // order, or the sampled lobe direction moves by whole quanta of acos near 1
    const float alpha = xmul(roughness, roughness);
    const float c2 = __fdiv_rn(xsub(1.f, u0), xadd(xmul(u0, xsub(xmul(alpha, alpha), 1.f)), 1.f));
    const f3 wh = sphere_to_world(iris_acosf(__fsqrt_rn(c2)), xmul(2.f * IRIS_PI, u1), n);
    const float d2 = xmul(2.f, xdot(wo, wh));
    return normalize_nf(mk3(xsub(xmul(d2, wh.x), wo.x), xsub(xmul(d2, wh.y), wo.y), xsub(xmul(d2, wh.z), wo.z)));
}
// origin of a secondary ray: x + RayEpsilon * wi, product rounded before the sum (utils/path_tracing.py:97,260,286,365,391,443,471)
__device__ __forceinline__ float4 ray4(f3 o, float w) { return make_float4(o.x, o.y, o.z, w); }
__device__ __forceinline__ f3 ray_origin(f3 x, f3 wi) {
    return mk3(xadd(x.x, xmul(IRIS_RAY_EPSILON, wi.x)), xadd(x.y, xmul(IRIS_RAY_EPSILON, wi.y)), xadd(x.z, xmul(IRIS_RAY_EPSILON, wi.z)));
}

// ------------------------------------------------------------------------------------------------ BSDF
struct Mat {
    f3 a;       // albedo
    float r;    // roughness in [0.02,1]
    float m;    // metallic
};

struct Angles {
    float NoL, NoV, VoH, NoH;
};
__device__ __forceinline__ Angles angles(f3 wi, f3 wo, f3 n) {
    // one rounding per op (the ATen chain of model/brdf.py:151-155): NoH feeds the ill-conditioned GGX denominator
    const f3 h = normalize_nf(mk3(xadd(wi.x, wo.x), xadd(wi.y, wo.y), xadd(wi.z, wo.z)));
    Angles a;
    a.NoL = fmaxf(xdot(wi, n), 0.f);
    a.NoV = fmaxf(xdot(wo, n), 0.f);
    a.VoH = fmaxf(xdot(wo, h), 0.f);
    a.NoH = fmaxf(xdot(n, h), 0.f);
    return a;
}
__device__ __forceinline__ float pow5(float x) { const float x2 = x * x; return x2 * x2 * x; }

// d(brdf_c)/d(albedo_c), d(brdf_c)/d(roughness), d(brdf_c)/d(metallic)  -- pdf carries no gradient (D.data, model/brdf.py:160)
struct BrdfJac {
    f3 da, dr, dm;
};

// model/brdf.py:138-175.  JAC: also the analytic Jacobian used by the adjoint.
template <bool JAC>
__device__ __forceinline__ void eval_brdf(f3 wi, f3 wo, f3 n, const Mat &mat, f3 &f, float &pdf, BrdfJac *J) {
    const Angles g = angles(wi, wo, n);
    const float r = mat.r;
    // utils/ops.py:74-82 in its own op order, one rounding each: den = NoH^2 (a2-1) + 1 cancels for small roughness
    const float al = xmul(r, r);
    const float a2 = xmul(al, al);
    const float den = xadd(xmul(xmul(g.NoH, g.NoH), xsub(a2, 1.f)), 1.f);
    const float D = __fdiv_rn(a2, xmul(xmul(IRIS_PI, den), den));
    pdf = 0.5f * (D / (4.f * fmaxf(g.VoH, 1e-4f)) * g.NoH) + 0.5f * (g.NoL / IRIS_PI);
    const float om = 1.f - mat.m;
    const f3 kd = mat.a * om;
    const f3 ks = mk3(0.04f * om + mat.a.x * mat.m, 0.04f * om + mat.a.y * mat.m, 0.04f * om + mat.a.z * mat.m);
    float k = r + 1.f;
    k = k * k / 8.f;
    const float gl = 1.f / (g.NoL * (1.f - k) + k), gv = 1.f / (g.NoV * (1.f - k) + k);
    const float G = gl * gv;
    const float x5 = pow5(1.f - g.VoH);
    const f3 F = mk3(ks.x + (1.f - ks.x) * x5, ks.y + (1.f - ks.y) * x5, ks.z + (1.f - ks.z) * x5);
    const float spec = D * G / 4.f * g.NoL;
    f = mk3(kd.x / IRIS_PI * g.NoL + spec * F.x, kd.y / IRIS_PI * g.NoL + spec * F.y, kd.z / IRIS_PI * g.NoL + spec * F.z);
    if (JAC) {
        const float dFdks = 1.f - x5;
        const float da = (om / IRIS_PI) * g.NoL + spec * dFdks * mat.m;
        J->da = mk3(da, da, da);
        J->dm = mk3((-mat.a.x / IRIS_PI) * g.NoL + spec * dFdks * (mat.a.x - 0.04f),
                    (-mat.a.y / IRIS_PI) * g.NoL + spec * dFdks * (mat.a.y - 0.04f),
                    (-mat.a.z / IRIS_PI) * g.NoL + spec * dFdks * (mat.a.z - 0.04f));
        // dD/dr = dD/da2 * 4 r^3 ; dG/dr = dG/dk * (r+1)/4
        const float dD = (den - 2.f * a2 * g.NoH * g.NoH) / (IRIS_PI * den * den * den) * (4.f * r * r * r);
        const float dG = -G * ((1.f - g.NoL) * gl + (1.f - g.NoV) * gv) * ((r + 1.f) * 0.25f);
        const float ds = (dD * G + D * dG) / 4.f * g.NoL;
        J->dr = mk3(ds * F.x, ds * F.y, ds * F.z);
    }
}

// model/brdf.py:177-210: 50/50 lobe pick, weight = brdf/pdf (0 where pdf == 0, NaN -> 0)
template <bool JAC>
__device__ __forceinline__ void sample_brdf(float u1, float u2x, float u2y, f3 wo, f3 n, const Mat &mat, f3 &wi, float &pdf, f3 &w,
                                            BrdfJac *J) {
    wi = u1 > 0.5f ? diffuse_sampler(u2x, u2y, n) : specular_sampler(u2x, u2y, mat.r, wo, n);
    f3 f;
    eval_brdf<JAC>(wi, wo, n, mat, f, pdf, J);
    const bool ok = pdf > 0.f;
    const float ip = ok ? 1.f / pdf : 0.f;
    w = mk3(f.x * ip, f.y * ip, f.z * ip);
    bool bad = !ok;
    if (w.x != w.x) { w.x = 0.f; bad = true; }
    if (w.y != w.y) { w.y = 0.f; bad = true; }
    if (w.z != w.z) { w.z = 0.f; bad = true; }
    if (JAC) {
        // d(w)/d(mat) = d(f)/d(mat) / pdf; lanes with pdf == 0 carry zero gradient (defined deviation)
        const float s = bad ? 0.f : ip;
        J->da = J->da * s; J->dr = J->dr * s; J->dm = J->dm * s;
    }
}

// model/brdf.py:112-136 weights of the specular bake: fac = G*VoH*NoL/max(NoH,1e-4); (1-x5)*fac, x5*fac
__device__ __forceinline__ void specular_weights(f3 wi, f3 wo, f3 n, float r, float &w0, float &w1) {
    const Angles g = angles(wi, wo, n);
    float k = r + 1.f;
    k = k * k / 8.f;
    const float G = (1.f / (g.NoL * (1.f - k) + k)) * (1.f / (g.NoV * (1.f - k) + k));
    const float x5 = pow5(1.f - g.VoH);
    const float fac = G * g.VoH * g.NoL / fmaxf(g.NoH, 1e-4f);
    w0 = (1.f - x5) * fac;
    w1 = x5 * fac;
}

// ------------------------------------------------------------------------------------------------ SLF + emitters
// model/slf.py:41-70: nearest voxel; index arithmetic is exact (one rounding per op) so the voxel matches the oracle.
__device__ __forceinline__ f3 slf_lookup(const IrisShadeParams &P, f3 x) {
    const float H = (float)P.slf_H;
    int gx = __float2int_rz(xmul(__fdiv_rn(xsub(x.x, P.slf_vmin), P.slf_range), H));
    int gy = __float2int_rz(xmul(__fdiv_rn(xsub(x.y, P.slf_vmin), P.slf_range), H));
    int gz = __float2int_rz(xmul(__fdiv_rn(xsub(x.z, P.slf_vmin), P.slf_range), H));
    gx = min(max(gx, 0), P.slf_H - 1);
    gy = min(max(gy, 0), P.slf_H - 1);
    gz = min(max(gz, 0), P.slf_H - 1);
    const int32_t idx = __ldg(P.slf_inds + ((int64_t)gz * P.slf_H + gy) * P.slf_H + gx);
    if (idx < 0) return mk3(0.f, 0.f, 0.f);
    return mk3(__ldg(P.slf_radiance + 3 * (int64_t)idx), __ldg(P.slf_radiance + 3 * (int64_t)idx + 1), __ldg(P.slf_radiance + 3 * (int64_t)idx + 2));
}

__device__ __forceinline__ int32_t emitter_of(const IrisShadeParams &P, int32_t prim) {
    return prim >= 0 ? __ldg(P.emitter_of_face + prim) : -1;
}
__device__ __forceinline__ f3 emitter_radiance(const IrisShadeParams &P, int32_t e) {
    return mk3(__ldg(P.radiance + 3 * (int64_t)e), __ldg(P.radiance + 3 * (int64_t)e + 1), __ldg(P.radiance + 3 * (int64_t)e + 2));
}
__device__ __forceinline__ float emitter_pdf_area(const IrisShadeParams &P, int32_t e) {
    return __ldg(P.emitter_pdf + e) / fmaxf(__ldg(P.emitter_area + e), 1e-12f);
}

// The sampled emitter triangle as the BVH stores it (v0, e1 = v1 - v0, e2 = v2 - v0 rounded once; zero-area -> never hit)
__device__ __forceinline__ void emitter_triangle(const IrisShadeParams &P, int32_t e, f3 &v0, f3 &e1, f3 &e2) {
    const float *V = P.emitter_vertices + 9 * (int64_t)e;
    v0 = mk3(__ldg(V + 0), __ldg(V + 1), __ldg(V + 2));
    e1 = mk3(xsub(__ldg(V + 3), v0.x), xsub(__ldg(V + 4), v0.y), xsub(__ldg(V + 5), v0.z));
    e2 = mk3(xsub(__ldg(V + 6), v0.x), xsub(__ldg(V + 7), v0.y), xsub(__ldg(V + 8), v0.z));
    const f3 c = xcross(e1, e2);
    if (c.x == 0.f && c.y == 0.f && c.z == 0.f) { e1 = mk3(0.f, 0.f, 0.f); e2 = e1; }
}

// model/emitter.py:224-255
__device__ __forceinline__ void sample_emitter(const IrisShadeParams &P, float u1, float u2x, float u2y, f3 x, f3 &wi, float &pdf,
                                               int32_t &e, int32_t &face) {
    const float u = fmaxf(u1, 1e-12f);
    int lo = 0, hi = P.n_emitters;   // first index with cdf >= u  (torch.searchsorted, right=False)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(P.emitter_cdf + mid) < u) lo = mid + 1; else hi = mid;
    }
    e = min(lo, P.n_emitters - 1);
    const float xi = __fsqrt_rn(u2x);
    const float bu = xsub(1.f, xi), bv = xmul(xi, u2y), bw = xsub(xsub(1.f, bu), bv);
    const float *V = P.emitter_vertices + 9 * (int64_t)e;
    // p1 = v0*u + v1*v + v2*w, products rounded, summed left to right (model/emitter.py:246-247)
    const f3 p1 = mk3(xadd(xadd(xmul(__ldg(V + 0), bu), xmul(__ldg(V + 3), bv)), xmul(__ldg(V + 6), bw)),
                      xadd(xadd(xmul(__ldg(V + 1), bu), xmul(__ldg(V + 4), bv)), xmul(__ldg(V + 7), bw)),
                      xadd(xadd(xmul(__ldg(V + 2), bu), xmul(__ldg(V + 5), bv)), xmul(__ldg(V + 8), bw)));
    wi = normalize_nf(mk3(xsub(p1.x, x.x), xsub(p1.y, x.y), xsub(p1.z, x.z)));
    pdf = emitter_pdf_area(P, e);
    face = __ldg(P.face_of_emitter + e);
}
