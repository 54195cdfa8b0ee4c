// emitter_extract.cuh -- emitter extraction of extract_emitter_ldr.py:76-110 on the device: per-triangle sum and count of the
// radiance seen at the triangles hit by the training rays (torch_scatter 'sum' over ray_intersect's triangle index), then
// is_emitter = max_c(sum_c / max(count, 1)) > threshold, and vertices / area / unit normal of the selected triangles
// (cross product and normalisation in the ATen operation order: one rounding per multiply/subtract, fma-chain vector norm).
// Atomic-bound integer/float scatter: 16 B in and 4 atomics per valid ray.
#pragma once
#include "common.cuh"

__global__ void k_tri_accumulate(const int32_t *__restrict__ prim, const uint8_t *__restrict__ valid, const float *__restrict__ radiance, int64_t n,
                                 int64_t n_faces, float *__restrict__ tri_sum, int32_t *__restrict__ tri_count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (valid && !valid[i])) return;
    const int32_t f = prim[i];
    if (f < 0 || f >= n_faces) return;
    atomicAdd(tri_sum + 3 * (int64_t)f, radiance[3 * i]);
    atomicAdd(tri_sum + 3 * (int64_t)f + 1, radiance[3 * i + 1]);
    atomicAdd(tri_sum + 3 * (int64_t)f + 2, radiance[3 * i + 2]);
    atomicAdd(tri_count + f, 1);
}

__global__ void k_emitter_classify(const float *__restrict__ tri_sum, const int32_t *__restrict__ tri_count, int64_t n_faces, float threshold,
                                   uint8_t *__restrict__ is_emitter) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    const float c = (float)max(tri_count[f], 1);
    const float m = fmaxf(fmaxf(__fdiv_rn(tri_sum[3 * f], c), __fdiv_rn(tri_sum[3 * f + 1], c)), __fdiv_rn(tri_sum[3 * f + 2], c));
    is_emitter[f] = m > threshold ? 1 : 0;
}

// emitter_faces: the K selected face indices in increasing order (the order of boolean-mask indexing in the reference)
__global__ void k_emitter_geometry(const float *__restrict__ verts, const int32_t *__restrict__ faces, const int64_t *__restrict__ emitter_faces, int64_t K,
                                   float *__restrict__ out_vertices, float *__restrict__ out_area, float *__restrict__ out_normal) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K) return;
    const int64_t f = emitter_faces[e];
    f3 v[3];
    for (int k = 0; k < 3; ++k) {
        v[k] = ld3(verts, faces[3 * f + k]);
        st3(out_vertices, 3 * e + k, v[k]);
    }
    const f3 a = mk3(xsub(v[1].x, v[0].x), xsub(v[1].y, v[0].y), xsub(v[1].z, v[0].z));
    const f3 b = mk3(xsub(v[2].x, v[0].x), xsub(v[2].y, v[0].y), xsub(v[2].z, v[0].z));
    const f3 c = xcross(a, b);
    const float nrm = sqrtf(__fmaf_rn(c.z, c.z, __fmaf_rn(c.y, c.y, xmul(c.x, c.x))));     // torch CPU vector_norm
    const float d = fmaxf(nrm, 1e-12f);                                                    // NF.normalize eps
    st3(out_normal, e, mk3(__fdiv_rn(c.x, d), __fdiv_rn(c.y, d), __fdiv_rn(c.z, d)));
    out_area[e] = __fdiv_rn(nrm, 2.0f);
}
