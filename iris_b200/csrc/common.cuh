// common.cuh -- small device helpers shared by every kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define IRIS_BLOCK 128
#define IRIS_PI 3.14159265358979323846f
#define IRIS_RAY_EPSILON 8.940696716308594e-05f   // 1500 * 2^-24 = mitsuba.math.RayEpsilon (fp32 variants)

struct f3 {
    float x, y, z;
};
__host__ __device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ __forceinline__ bool is_zero3(f3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }   // false for NaN
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// NF.normalize: v / max(|v|, 1e-12).  |v|^2 is accumulated as fma(z,z,fma(y,y,x*x)) -- the order torch's CPU vector_norm uses for a
// 3-vector (checked bit for bit) -- and every step is a single IEEE rounding, so directions match the reference's CPU run.
#ifdef IRIS_HOST_EMULATION
__device__ __forceinline__ f3 normalize_nf(f3 a) {
#else
__device__ __noinline__ f3 normalize_nf(f3 a) {   // out of line: ~55 SASS instructions (IEEE sqrt + 3 IEEE divisions) x a dozen call sites
#endif
    const float l = fmaxf(__fsqrt_rn(__fmaf_rn(a.z, a.z, __fmaf_rn(a.y, a.y, __fmul_rn(a.x, a.x)))), 1e-12f);
    return mk3(__fdiv_rn(a.x, l), __fdiv_rn(a.y, l), __fdiv_rn(a.z, l));
}
__device__ __forceinline__ f3 ld3(const float *p, int64_t i) { return mk3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
__device__ __forceinline__ void st3(float *p, int64_t i, f3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

// "Exact" arithmetic: one IEEE rounding per operation, never contracted into FMA.  Everything that decides a hit
// index, a hit point or a voxel index goes through these so that the CUDA kernels and oracle/intersect.c (built with
// -ffp-contract=off) agree bit for bit.
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdot(f3 a, f3 b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }
__device__ __forceinline__ f3 xcross(f3 a, f3 b) {
    return mk3(xsub(xmul(a.y, b.z), xmul(a.z, b.y)), xsub(xmul(a.z, b.x), xmul(a.x, b.z)), xsub(xmul(a.x, b.y), xmul(a.y, b.x)));
}

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

#ifndef IRIS_HOST_EMULATION
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif
