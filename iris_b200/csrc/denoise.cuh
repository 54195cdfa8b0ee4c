// denoise.cuh -- edge-avoiding a-trous wavelet filter for the baked shading maps (Dammertz, Sewtz, Hanika, Lensch 2010).
//
// Where it sits: bake_shading.py:81,126-131,190-203 passes every baked map through `mitsuba.OptixDenoiser` before writing it.
// That denoiser is a learned network inside OptiX (absent third-party, no weights, no golden vectors): PARITY WITH IT IS UNPINNED
// BY NATURE.  What is provided instead is a deterministic, documented filter with the same call surface -- one (H, W, 3) map in,
// one out -- that uses what the bake already has on the device: the primary-hit normals and positions as edge-stopping guides.
// oracle/denoise.py restates it in numpy; the GPU test compares the two.
//
// One launch per level i = 0 .. iterations-1, ping-pong between two buffers: for pixel p the 5 x 5 taps q = p + 2^i (dx, dy),
//   w(q) = h(dx) h(dy) * exp(-|c_p - c_q|^2 / (sigma_c 2^-i)^2) * max(0, n_p . n_q)^sigma_n * exp(-|x_p - x_q|^2 / sigma_x^2)
// with the B3-spline h = (1/16, 1/4, 3/8, 1/4, 1/16); out_p = sum w c_q / sum w.  A pixel whose guide normal is all zero (no primary
// hit: the bake leaves those pixels black) is neither filtered nor used as a tap.  Guides may be NULL (colour term only).
#pragma once
#include "common.cuh"

__global__ void __launch_bounds__(256) k_denoise_atrous(const float *__restrict__ src, const float *__restrict__ normal, const float *__restrict__ position,
                                                        int H, int W, int step, float inv_sc2, float sigma_n, float inv_sx2, float *__restrict__ dst) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const int64_t p = (int64_t)y * W + x;
    const f3 cp = ld3(src, p);
    f3 np_ = mk3(0.f, 0.f, 1.f), xp = mk3(0.f, 0.f, 0.f);
    if (normal) {
        np_ = ld3(normal, p);
        if (is_zero3(np_)) { st3(dst, p, cp); return; }
    }
    if (position) xp = ld3(position, p);
    const float h[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
    f3 acc = mk3(0.f, 0.f, 0.f);
    float wsum = 0.f;
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy) {
        const int yy = y + dy * step;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int dx = -2; dx <= 2; ++dx) {
            const int xx = x + dx * step;
            if (xx < 0 || xx >= W) continue;
            const int64_t q = (int64_t)yy * W + xx;
            const f3 cq = ld3(src, q);
            const f3 dc = cq - cp;
            float w = h[dx + 2] * h[dy + 2] * expf(-dot(dc, dc) * inv_sc2);
            if (normal) {
                const f3 nq = ld3(normal, q);
                if (is_zero3(nq)) continue;
                w *= powf(fmaxf(dot(np_, nq), 0.f), sigma_n);
            }
            if (position) {
                const f3 dxp = ld3(position, q) - xp;
                w *= expf(-dot(dxp, dxp) * inv_sx2);
            }
            acc = acc + cq * w;
            wsum += w;
        }
    }
    st3(dst, p, wsum > 0.f ? acc * (1.f / wsum) : cp);
}
