// Host emulation of the device traversal (iris_b200/csrc/traverse.cuh) over the BVH produced by the host builder, so the
// builder + traversal logic is testable without a GPU.  CUDA intrinsics are shimmed with their IEEE host equivalents;
// compile with -ffp-contract=off.  Exposes one C function for ctypes.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

#undef __device__
#undef __host__
#undef __forceinline__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
static inline float __uint_as_float(uint32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t v; std::memcpy(&v, &f, 4); return v; }
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline float4 __ldg(const float4 *p) { return *p; }
static inline float __shfl_xor_sync(unsigned, float v, int) { return v; }
#define IRIS_HOST_EMULATION 1
#include "../../iris_b200/csrc/traverse.cuh"

extern "C" int host_trace(const float *verts, int64_t nv, const int32_t *faces, int64_t nf, const float *o, const float *d, int64_t n,
                          float *t, int32_t *prim, float *uv, float *p, float *nrm, int64_t *n_nodes) {
    HostBvh hb;
    if (host_bvh_build(verts, nv, faces, nf, &hb) != 0) return -1;
    SceneView S;
    S.nodes = reinterpret_cast<const float4 *>(hb.nodes);
    S.tris = reinterpret_cast<const float4 *>(hb.tris);
    S.n_tris = hb.n_tris;
    n_nodes[0] = hb.n_nodes;
    n_nodes[1] = hb.max_depth;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) {
        f3 ro = mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), rd = mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        Hit h = trace_closest(S, ro, rd);
        f3 hp, hn;
        hit_surface(S, h, rd, hp, hn);
        t[i] = h.t; prim[i] = h.prim;
        uv[2 * i] = h.prim >= 0 ? h.u : 0.f; uv[2 * i + 1] = h.prim >= 0 ? h.v : 0.f;
        p[3 * i] = hp.x; p[3 * i + 1] = hp.y; p[3 * i + 2] = hp.z;
        nrm[3 * i] = hn.x; nrm[3 * i + 1] = hn.y; nrm[3 * i + 2] = hn.z;
    }
    host_bvh_free(&hb);
    return 0;
}
