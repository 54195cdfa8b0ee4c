// crf.cuh -- EmorCRF forward / adjoint (reference crf/model_crf.py:68-86), the step right after the estimator in every trainer
// (train_emitter.py:191-193): ldr_c = lerp(crf_c, clip(hdr_c * exposure, 0, 1)) on a regular grid of n_bins samples over [0,1].
// The reference goes through torch_interpolations.RegularGridInterpolator (absent third-party); linear interpolation on a
// regular grid is restated directly: bin i = min(floor(x*(n-1)), n-2), weight = x*(n-1) - i.
// Adjoint: d_hdr = d_ldr * (crf[i+1]-crf[i]) * (n-1) * exposure inside the clip range (0 outside), and d_crf gets the two
// interpolation weights -- accumulated in a per-block shared-memory copy of the (3, n_bins) table, one atomic per bin per block.
#pragma once
#include "common.cuh"

__global__ void k_crf_forward(const float *__restrict__ hdr, const float *__restrict__ exposure, int exp_stride, const float *__restrict__ crf,
                              int n_bins, int64_t n, float *__restrict__ ldr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * n) return;
    const int64_t row = i / 3;
    const int c = (int)(i - 3 * row);
    const float x = fminf(fmaxf(hdr[i] * exposure[row * exp_stride], 0.f), 1.f);
    const float s = x * (float)(n_bins - 1);
    const int b = min((int)s, n_bins - 2);
    const float w = s - (float)b;
    const float *t = crf + c * n_bins;
    ldr[i] = t[b] * (1.f - w) + t[b + 1] * w;
}

__global__ void k_crf_backward(const float *__restrict__ hdr, const float *__restrict__ exposure, int exp_stride, const float *__restrict__ crf,
                               int n_bins, const float *__restrict__ d_ldr, int64_t n, float *__restrict__ d_hdr, float *__restrict__ d_crf) {
    extern __shared__ float acc[];                      // (3, n_bins)
    for (int k = threadIdx.x; k < 3 * n_bins; k += blockDim.x) acc[k] = 0.f;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 3 * n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / 3;
        const int c = (int)(i - 3 * row);
        const float e = exposure[row * exp_stride];
        const float raw = hdr[i] * e;
        const float x = fminf(fmaxf(raw, 0.f), 1.f);
        const float s = x * (float)(n_bins - 1);
        const int b = min((int)s, n_bins - 2);
        const float w = s - (float)b;
        const float g = d_ldr[i];
        const float *t = crf + c * n_bins;
        if (d_hdr) d_hdr[i] = (raw >= 0.f && raw <= 1.f) ? g * (t[b + 1] - t[b]) * (float)(n_bins - 1) * e : 0.f;
        if (d_crf) {
            atomicAdd(&acc[c * n_bins + b], g * (1.f - w));
            atomicAdd(&acc[c * n_bins + b + 1], g * w);
        }
    }
    if (d_crf) {
        __syncthreads();
        for (int k = threadIdx.x; k < 3 * n_bins; k += blockDim.x)
            if (acc[k] != 0.f) atomicAdd(d_crf + k, acc[k]);
    }
}
