// shade_maps.cuh -- the training-step shading of train_brdf_crf.py:193-206 with utils/ops.py:99-119 (lerp_specular):
//   kd = albedo (1 - metallic)        ks = 0.04 (1 - metallic) + albedo metallic
//   L  = kd * diffuse + ks * lerp(specular0, roughness) + lerp(specular1, roughness)
// where lerp interpolates the R baked roughness levels at r = (roughness - 0.02) / 0.98 * (R - 1) between floor(r) and ceil(r).
// One lane per pixel; every operation is a single IEEE rounding in the order of the ATen chain (parity: rel 1e-3, in practice
// bit-identical on L).  HBM-bound: 20 + 12 + 24 R bytes in, 12 out per pixel (176 B at R = 6).
// Adjoint: d_mat += J^T dL for (albedo rgb, roughness, metallic); the maps are data and receive no gradient.
#pragma once
#include "common.cuh"

struct SpecLerp {
    f3 s0, s1;     // levels floor(r), ceil(r)
    float w;       // r - floor(r)
};

__device__ __forceinline__ SpecLerp spec_levels(const float *__restrict__ spec, int64_t i, int R, float roughness, int &r0, int &r1, float &w) {
    // utils/ops.py:108-115
    const float r = xmul(__fdiv_rn(xsub(roughness, 0.02f), 0.98f), (float)(R - 1));
    const float fl = floorf(r), ce = ceilf(r);
    r0 = min(max((int)fl, 0), R - 1);
    r1 = min(max((int)ce, 0), R - 1);
    w = xsub(r, fl);
    SpecLerp o;
    o.s0 = ld3(spec, i * R + r0);
    o.s1 = ld3(spec, i * R + r1);
    o.w = w;
    return o;
}

__device__ __forceinline__ f3 spec_lerp(const SpecLerp &s) {
    // utils/ops.py:118   s0*(1-r_) + s1*r_
    const float a = xsub(1.0f, s.w);
    return mk3(xadd(xmul(s.s0.x, a), xmul(s.s1.x, s.w)), xadd(xmul(s.s0.y, a), xmul(s.s1.y, s.w)), xadd(xmul(s.s0.z, a), xmul(s.s1.z, s.w)));
}

__global__ void k_brdf_shading_forward(const float *__restrict__ mat, const float *__restrict__ diffuse, const float *__restrict__ spec0,
                                       const float *__restrict__ spec1, int R, int64_t n, float *__restrict__ L) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const f3 a = mk3(mat[5 * i], mat[5 * i + 1], mat[5 * i + 2]);
    const float rough = mat[5 * i + 3], m = mat[5 * i + 4];
    const float om = xsub(1.0f, m), ksd = xmul(0.04f, om);
    int r0, r1;
    float w;
    const f3 l0 = spec_lerp(spec_levels(spec0, i, R, rough, r0, r1, w));
    const f3 l1 = spec_lerp(spec_levels(spec1, i, R, rough, r0, r1, w));
    const f3 d = ld3(diffuse, i);
    const float av[3] = {a.x, a.y, a.z}, dv[3] = {d.x, d.y, d.z}, l0v[3] = {l0.x, l0.y, l0.z}, l1v[3] = {l1.x, l1.y, l1.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float kd = xmul(av[c], om), ks = xadd(ksd, xmul(av[c], m));
        L[3 * i + c] = xadd(xmul(kd, dv[c]), xadd(xmul(ks, l0v[c]), l1v[c]));
    }
}

__global__ void k_brdf_shading_backward(const float *__restrict__ mat, const float *__restrict__ diffuse, const float *__restrict__ spec0,
                                        const float *__restrict__ spec1, int R, int64_t n, const float *__restrict__ dL, float *__restrict__ d_mat) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float av[3] = {mat[5 * i], mat[5 * i + 1], mat[5 * i + 2]};
    const float rough = mat[5 * i + 3], m = mat[5 * i + 4];
    const float om = 1.0f - m;
    int r0, r1;
    float w;
    const SpecLerp A = spec_levels(spec0, i, R, rough, r0, r1, w), B = spec_levels(spec1, i, R, rough, r0, r1, w);
    const f3 l0 = spec_lerp(A);
    const f3 d = ld3(diffuse, i);
    const float dv[3] = {d.x, d.y, d.z}, l0v[3] = {l0.x, l0.y, l0.z};
    const float dA[3] = {A.s1.x - A.s0.x, A.s1.y - A.s0.y, A.s1.z - A.s0.z}, dB[3] = {B.s1.x - B.s0.x, B.s1.y - B.s0.y, B.s1.z - B.s0.z};
    float d_m = 0.f, d_w = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float g = dL[3 * i + c];
        const float d_kd = g * dv[c], d_ks = g * l0v[c];
        const float ks = 0.04f * om + av[c] * m;
        d_mat[5 * i + c] += d_kd * om + d_ks * m;
        d_m += d_ks * (av[c] - 0.04f) - d_kd * av[c];
        d_w += g * (ks * dA[c] + dB[c]);
    }
    d_mat[5 * i + 3] += d_w * (float)(R - 1) / 0.98f;          // r_ = r - floor(r): d r_/d roughness = (R-1)/0.98 almost everywhere
    d_mat[5 * i + 4] += d_m;
}
