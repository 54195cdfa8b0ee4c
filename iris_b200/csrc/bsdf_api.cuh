// bsdf_api.cuh -- the per-lane BSDF samplers as stand-alone entry points: BaseBRDF.sample_diffuse (MODE 0), .sample_specular
// (MODE 1) and .sample_brdf (MODE 2) of the reference's model/brdf.py:78-88,112-136,177-210.  The fused estimators inline the
// same device functions (shading.cuh); these launches serve callers that drive the loop themselves (bake_shading.py:113-114,
// 173-177) and the parity test that pins the sampled directions bit for bit.
#pragma once
#include "shading.cuh"

// u (n,3): MODE 0/1 read columns 0,1 (sample2); MODE 2 reads column 0 as sample1 and 1,2 as sample2.
// mat (n,5) = albedo rgb, roughness, metallic (MODE 2); roughness: scalar level (MODE 1 when mat == NULL, else mat[:,3]).
// wi (n,3), pdf (n), w0 (n,3) [MODE 0: ones, MODE 1: F0*fac broadcast, MODE 2: brdf/pdf], w1 (n,3) [MODE 1: F1*fac broadcast]
template <int MODE>
__global__ void __launch_bounds__(IRIS_BLOCK) k_bsdf_sample(const float *__restrict__ u, int u_stride, const float *__restrict__ wo_in,
                                                             const float *__restrict__ normal, const float *__restrict__ mat_in, float roughness,
                                                             int64_t n, float *wi_out, float *pdf_out, float *w0_out, float *w1_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const f3 nr = ld3(normal, i);
    const float *ui = u + i * (int64_t)u_stride;
    f3 wi, w0 = mk3(1.f, 1.f, 1.f), w1 = mk3(0.f, 0.f, 0.f);
    float pdf;
    if (MODE == 0) {
        wi = diffuse_sampler(ui[0], ui[1], nr);
        pdf = fmaxf(xdot(nr, wi), 0.f) / IRIS_PI;
    } else if (MODE == 1) {
        const f3 wo = ld3(wo_in, i);
        const float r = mat_in ? mat_in[5 * i + 3] : roughness;
        wi = specular_sampler(ui[0], ui[1], r, wo, nr);
        const Angles g = angles(wi, wo, nr);
        const float al = xmul(r, r), a2 = xmul(al, al);
        const float den = xadd(xmul(xmul(g.NoH, g.NoH), xsub(a2, 1.f)), 1.f);
        const float D = __fdiv_rn(a2, xmul(xmul(IRIS_PI, den), den));
        pdf = D / (4.f * fmaxf(g.VoH, 1e-4f)) * g.NoH;
        float a, b;
        specular_weights(wi, wo, nr, r, a, b);
        w0 = mk3(a, a, a);
        w1 = mk3(b, b, b);
    } else {
        const f3 wo = ld3(wo_in, i);
        Mat mat;
        mat.a = mk3(mat_in[5 * i], mat_in[5 * i + 1], mat_in[5 * i + 2]);
        mat.r = mat_in[5 * i + 3];
        mat.m = mat_in[5 * i + 4];
        sample_brdf<false>(ui[0], ui[1], ui[2], wo, nr, mat, wi, pdf, w0, nullptr);
    }
    st3(wi_out, i, wi);
    if (pdf_out) pdf_out[i] = pdf;
    if (w0_out) st3(w0_out, i, w0);
    if (MODE == 1 && w1_out) st3(w1_out, i, w1);
}
