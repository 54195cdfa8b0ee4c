// slf_bake.cuh -- the voxel surface-light-field bake of slf_bake.py:70-145 / model/slf.py:16-61 on the device:
//   bounds of the hit points -> occupancy of the H^3 grid (SpatialHist > 0) -> compact table index in raster order
//   (torch.where(mask) order = exclusive prefix sum over z,y,x) -> scatter-add of radiance and count -> mean.
// Voxel arithmetic is the reference's, one rounding per op: g = clamp(trunc((x - vmin) / (vmax - vmin) * H), 0, H-1).
// All of it is HBM/atomic-bound integer work: 12 B in per point for marking, 24 B in + 4 atomics per point for accumulation.
#pragma once
#include "common.cuh"

__device__ __forceinline__ int64_t slf_cell(f3 x, float vmin, float vrange, int H) {
    const float Hf = (float)H;
    int gx = __float2int_rz(xmul(__fdiv_rn(xsub(x.x, vmin), vrange), Hf));
    int gy = __float2int_rz(xmul(__fdiv_rn(xsub(x.y, vmin), vrange), Hf));
    int gz = __float2int_rz(xmul(__fdiv_rn(xsub(x.z, vmin), vrange), Hf));
    gx = min(max(gx, 0), H - 1);
    gy = min(max(gy, 0), H - 1);
    gz = min(max(gz, 0), H - 1);
    return ((int64_t)gz * H + gy) * H + gx;
}

// order-preserving float <-> uint32 map, so that atomicMin / atomicMax on the encoded value order like the floats
__device__ __forceinline__ uint32_t slf_float_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float slf_key_float(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k); }

// slf_bake.py:73-85: min and max over every coordinate of the valid points.  keys[0] = min key, keys[1] = max key.
__global__ void k_slf_bounds(const float *__restrict__ pos, const uint8_t *__restrict__ valid, int64_t n, uint32_t *keys) {
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (valid && !valid[i]) continue;
        const f3 p = ld3(pos, i);
        const uint32_t kx = slf_float_key(p.x), ky = slf_float_key(p.y), kz = slf_float_key(p.z);
        lo = min(lo, min(kx, min(ky, kz)));
        hi = max(hi, max(kx, max(ky, kz)));
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) {
        if (lo != 0xFFFFFFFFu) atomicMin(keys, lo);
        if (hi != 0u) atomicMax(keys + 1, hi);
    }
}
__global__ void k_slf_bounds_decode(const uint32_t *keys, float *minmax) {
    minmax[0] = slf_key_float(keys[0]);
    minmax[1] = slf_key_float(keys[1]);
}

// slf_bake.py:96-113: occupancy.  A plain store of 1 is enough (the histogram is only compared with 0).
__global__ void k_slf_mark(const float *__restrict__ pos, const uint8_t *__restrict__ valid, int64_t n, float vmin, float vrange, int H,
                           int32_t *__restrict__ occupancy) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (valid && !valid[i])) return;
    occupancy[slf_cell(ld3(pos, i), vmin, vrange, H)] = 1;
}

// model/slf.py:29-32: inds = -1 where empty, else the rank of the voxel among the occupied ones in raster order
__global__ void k_slf_index(const int32_t *__restrict__ occupancy, const int32_t *__restrict__ rank, int64_t cells, int32_t *__restrict__ inds) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cells) inds[i] = occupancy[i] ? rank[i] : -1;
}

// model/slf.py:56-61.  Like the reference, a point that falls into an empty voxel (inds = -1) must not happen during the bake
// (the mask comes from the same points); such points are skipped here where torch would index row -1.
__global__ void k_slf_accumulate(const float *__restrict__ pos, const uint8_t *__restrict__ valid, const float *__restrict__ radiance, int64_t n,
                                 float vmin, float vrange, int H, const int32_t *__restrict__ inds, float *__restrict__ sum, int32_t *__restrict__ count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (valid && !valid[i])) return;
    const int32_t idx = __ldg(inds + slf_cell(ld3(pos, i), vmin, vrange, H));
    if (idx < 0) return;
    atomicAdd(sum + 3 * (int64_t)idx, radiance[3 * i]);
    atomicAdd(sum + 3 * (int64_t)idx + 1, radiance[3 * i + 1]);
    atomicAdd(sum + 3 * (int64_t)idx + 2, radiance[3 * i + 2]);
    atomicAdd(count + idx, 1);
}

// slf_bake.py:138: radiance / clamp_min(count, 1)
__global__ void k_slf_finalize(float *__restrict__ sum, const int32_t *__restrict__ count, int64_t n_cells) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * n_cells) return;
    sum[i] = __fdiv_rn(sum[i], (float)max(count[i / 3], 1));
}
