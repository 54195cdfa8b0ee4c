// traverse.cuh -- closest-hit traversal of the 8-wide compressed BVH (bvh8.h) + the exact triangle test.
//
// Replaces scene.ray_intersect_preliminary (OptiX) behind ray_intersect, reference utils/path_tracing.py:30-35.
// One ray per lane.  Per visited node the lane issues five 16-byte read-only loads (the whole 80-byte node), decodes
// eight quantised child boxes with byte arithmetic, and keeps (node group, triangle group) bit masks so that the
// traversal stack holds at most one 8-byte entry per level (Ylitie, Karras, Laine 2017).  Hit selection is the
// oracle's: minimum t, ties to the lowest prim index; box tests use `<=` so equal-t triangles are never culled.
#pragma once
#include "bvh8.h"
#include "common.cuh"

#include <limits.h>
#ifndef IRIS_SORT_BLOCK
#define IRIS_SORT_BLOCK 256   // block size of the kernels that re-order their rays
#endif
#ifndef IRIS_POSTPONE_NUM
#define IRIS_POSTPONE_NUM 1      // postpone the remaining triangles when fewer than NUM/DEN of the lanes that
#define IRIS_POSTPONE_DEN 5      // entered the triangle phase are still in it
#endif
#define IRIS_STACK 48   // one node group + one postponed triangle group per level: 2 * depth <= IRIS_STACK

struct SceneView {
    const float4 *nodes;   // 5 per node
    const float4 *tris;    // 3 per triangle record
    int64_t n_tris;
};

struct Hit {
    float t, u, v;
    int32_t prim;   // caller's face index, -1 = miss
    int32_t slot;   // record index in SceneView::tris
};

#ifndef IRIS_HOST_EMULATION
__constant__ uint32_t c_byte_magic = 0x4B000000u;
__constant__ uint32_t c_half_magic = 0x64646464u;
#endif
__device__ __forceinline__ uint32_t sign_extend_s8x4(uint32_t x) {
#ifdef IRIS_HOST_EMULATION   // tests/host/traverse_host.cpp
    return ((x >> 7) & 0x01010101u) * 0xFFu;
#else
    uint32_t r;
    asm("prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"(r) : "r"(x));
    return r;
#endif
}
// byte j of w as a float, exactly, without the conversion unit: 0x4B0000bb is the float 2^23 + bb, so one PRMT (ALU pipe)
// and one FADD (FMA pipe) replace an I2F.U8 that would otherwise saturate the XU pipe (48 conversions per node visit)
__device__ __forceinline__ float byte_f(uint32_t w, int j) {
#ifdef IRIS_HOST_EMULATION
    return (float)((w >> (8 * j)) & 0xFFu);
#else
    // the 2^23 pattern comes from constant memory so that it stays a register/constant operand and the byte selector is the
    // immediate; with both literal, ptxas keeps 0x4B000000 as the immediate and moves the selector into a register per PRMT
    uint32_t r;
    const uint32_t magic = c_byte_magic;
    switch (j) {   // j is a literal after unrolling
    case 0: asm("prmt.b32 %0, %1, %2, 0x7440;" : "=r"(r) : "r"(w), "r"(magic)); break;
    case 1: asm("prmt.b32 %0, %1, %2, 0x7441;" : "=r"(r) : "r"(w), "r"(magic)); break;
    case 2: asm("prmt.b32 %0, %1, %2, 0x7442;" : "=r"(r) : "r"(w), "r"(magic)); break;
    default: asm("prmt.b32 %0, %1, %2, 0x7443;" : "=r"(r) : "r"(w), "r"(magic)); break;
    }
    return __uint_as_float(r) - 8388608.0f;
#endif
}

#ifndef IRIS_HOST_EMULATION
// bytes j and j+1 of w (j = 0 or 2) -> out[k] = fma(1024 + q_k, a, c), as one PRMT, two HADD2.F32 and one FFMA2
__device__ __forceinline__ void slab_pair(uint32_t w, int j, float a, float c, float (&out)[2]) {
    const uint32_t magic = c_half_magic;
    uint32_t h2;
    if (j == 0) asm("prmt.b32 %0, %1, %2, 0x4140;" : "=r"(h2) : "r"(w), "r"(magic));
    else asm("prmt.b32 %0, %1, %2, 0x4342;" : "=r"(h2) : "r"(w), "r"(magic));
    float v0, v1;
    asm("{ .reg .b16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h; }" : "=f"(v0), "=f"(v1) : "r"(h2));
    unsigned long long v, aa, cc, r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(v0), "f"(v1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(v), "l"(aa), "l"(cc));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(out[0]), "=f"(out[1]) : "l"(r));
}
#endif

#if IRIS_NODE_FP16
// one 32-bit word = the fp16 plane coordinates of two children -> out[k] = fma(q_k, a, b)
__device__ __forceinline__ void slab_pair_h(uint32_t w, float a, float b, float (&out)[2]) {
#ifdef IRIS_HOST_EMULATION
    out[0] = fmaf((float)bvh8_decode_q((bvh8_q_t)(w & 0xFFFFu)), a, b);
    out[1] = fmaf((float)bvh8_decode_q((bvh8_q_t)(w >> 16)), a, b);
#else
    float v0, v1;
    asm("{ .reg .b16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h; }" : "=f"(v0), "=f"(v1) : "r"(w));
    unsigned long long v, aa, bb, r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(v0), "f"(v1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(v), "l"(aa), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(out[0]), "=f"(out[1]) : "l"(r));
#endif
}
// 16-byte word `w` (a per-ray register: the near or far plane word of an axis) of the node at np.  The address is formed with one
// IMAD.WIDE on the FMA pipe, so picking near/far by the ray's sign costs no ALU-pipe work (SELs on the loaded words would cost 24).
__device__ __forceinline__ uint4 load_plane_word(const float4 *np, uint32_t w) {
#ifdef IRIS_HOST_EMULATION
    const float4 a = ldg4(np + w);
    return make_uint4(__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w));
#else
    unsigned long long addr;
    asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(addr) : "r"(w), "l"(np));
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(addr));
    return r;
#endif
}
#endif

// Moller-Trumbore barycentrics + projected t, operation order of oracle/intersect.c:tri_test (rdd = 1/(d.d))
__device__ __forceinline__ bool tri_test(f3 o, f3 d, float rdd, f3 v0, f3 e1, f3 e2, float &t, float &u, float &v) {
    f3 p = xcross(d, e2);
    float det = xdot(e1, p);
    float mag = xadd(xadd(fabsf(xmul(e1.x, p.x)), fabsf(xmul(e1.y, p.y))), fabsf(xmul(e1.z, p.z)));
    if (!(fabsf(det) > xmul(1e-6f, mag))) return false;
    float inv = __fdiv_rn(1.0f, det);
    f3 s = mk3(xsub(o.x, v0.x), xsub(o.y, v0.y), xsub(o.z, v0.z));
    float uu = xmul(xdot(s, p), inv);
    if (!(uu >= 0.0f && uu <= 1.0f)) return false;
    f3 q = xcross(s, e1);
    float vv = xmul(xdot(d, q), inv);
    if (!(vv >= 0.0f && xadd(uu, vv) <= 1.0f)) return false;
    f3 h = mk3(xsub(__fmaf_rn(vv, e2.x, __fmaf_rn(uu, e1.x, v0.x)), o.x), xsub(__fmaf_rn(vv, e2.y, __fmaf_rn(uu, e1.y, v0.y)), o.y),
               xsub(__fmaf_rn(vv, e2.z, __fmaf_rn(uu, e1.z, v0.z)), o.z));
    float tt = xmul(xdot(h, d), rdd);
    if (!(tt > 0.0f && tt < __int_as_float(0x7f800000))) return false;
    t = tt; u = uu; v = vv;
    return true;
}

__device__ __forceinline__ void load_tri(const SceneView &S, int32_t slot, f3 &v0, f3 &e1, f3 &e2, int32_t &prim) {
    const float4 a = ldg4(S.tris + 3 * (int64_t)slot), b = ldg4(S.tris + 3 * (int64_t)slot + 1), c = ldg4(S.tris + 3 * (int64_t)slot + 2);
    v0 = mk3(a.x, a.y, a.z);
    e1 = mk3(a.w, b.x, b.y);
    e2 = mk3(b.z, b.w, c.x);
    prim = __float_as_int(c.y);
}

// ANYHIT = false: closest hit (minimum t, ties to the lowest prim index).
// ANYHIT = true : occlusion query against a known candidate (t_limit, prim_limit): returns as soon as some triangle beats the
//                 candidate under the same ordering (t < t_limit, or t == t_limit and prim < prim_limit); best.slot >= 0 then.
// IRIS_TRACE_NOINLINE: kernels that cast more than one ray per lane (k_bounce_single, the wavefront kernels) call ONE out-of-line
// copy of the traversal loop instead of inlining it at every site: the loop is ~14 KB of SASS, and two or three copies of it plus
// the shading code overflow the 32 KB instruction cache (ncu: stall_no_instruction was the top stall reason of k_bounce_single).
template <bool ANYHIT>
__device__ __forceinline__ Hit trace_ray_body(const SceneView &S, f3 o, f3 d, float t_limit, int32_t prim_limit, int anyhit_rt);

#ifndef IRIS_HOST_EMULATION
__device__ __noinline__ Hit trace_ray_shared(const float4 *nodes, const float4 *tris, float ox, float oy, float oz, float dx, float dy, float dz,
                                             float t_limit, int32_t prim_limit, int anyhit);
#endif

template <bool ANYHIT>
__device__ __forceinline__ Hit trace_ray(const SceneView &S, f3 o, f3 d, float t_limit, int32_t prim_limit) {
#if defined(IRIS_TRACE_NOINLINE) && !defined(IRIS_HOST_EMULATION)
    return trace_ray_shared(S.nodes, S.tris, o.x, o.y, o.z, d.x, d.y, d.z, t_limit, prim_limit, ANYHIT ? 1 : 0);
#else
    return trace_ray_body<ANYHIT>(S, o, d, t_limit, prim_limit, 0);
#endif
}

// Traversal as explicit state + step, so that the same code serves the one-ray-per-lane loop (trace_ray_body) and the persistent
// kernel that refills finished lanes with new rays (k_trace_persistent).
struct TravState {
    f3 o, d;
#if IRIS_NODE_FP16
    uint32_t nwx, nwy, nwz;    // node word holding the NEAR planes of each axis (2,3,4 for positive directions, 7,6,5 for negative)
    float idx, idy, idz, rdd;
#else
    float sx, sy, sz, idx, idy, idz, rdd;
#endif
    uint32_t octinv;
    uint2 ngroup, tgroup;
    int sp;
    int anyhit;
    bool done;
    Hit best;
};

__device__ __forceinline__ void trav_init(TravState &T, f3 o, f3 d, float t_limit, int32_t prim_limit, int anyhit) {
    T.o = o;
    T.d = d;
    T.best.t = t_limit;
    T.best.u = T.best.v = 0.f;
    T.best.prim = prim_limit;
    T.best.slot = -1;
    T.sp = 0;
    T.anyhit = anyhit;
    T.done = false;
    const float sx = fabsf(d.x) < 1e-30f ? copysignf(1e-30f, d.x) : d.x;
    const float sy = fabsf(d.y) < 1e-30f ? copysignf(1e-30f, d.y) : d.y;
    const float sz = fabsf(d.z) < 1e-30f ? copysignf(1e-30f, d.z) : d.z;
    T.idx = 1.0f / sx;
    T.idy = 1.0f / sy;
    T.idz = 1.0f / sz;
    T.octinv = (sx < 0.f ? 0u : 4u) | (sy < 0.f ? 0u : 2u) | (sz < 0.f ? 0u : 1u);
#if IRIS_NODE_FP16
    T.nwx = sx < 0.f ? 7u : 2u;
    T.nwy = sy < 0.f ? 6u : 3u;
    T.nwz = sz < 0.f ? 5u : 4u;
#else
    T.sx = sx; T.sy = sy; T.sz = sz;
#endif
    T.rdd = __fdiv_rn(1.0f, xdot(d, d));
    T.ngroup = make_uint2(0u, 0x80000000u);
    T.tgroup = make_uint2(0u, 0u);
}

// One iteration: at most one node, then this lane's pending triangles (with postponing), then a pop.  Sets T.done.
__device__ __forceinline__ void trav_step(const SceneView &S, TravState &T, uint2 *stack) {
    const uint32_t octinv4 = T.octinv * 0x01010101u;
    if (T.ngroup.y > 0x00FFFFFFu) {
        const uint32_t hits = T.ngroup.y;
        const uint32_t imask = T.ngroup.y;
        const uint32_t bit = 31u - __clz(hits);
        const uint32_t base = T.ngroup.x;
        T.ngroup.y &= ~(1u << bit);
        if (T.ngroup.y > 0x00FFFFFFu) {
            stack[T.sp++] = T.ngroup;   // depth <= IRIS_STACK is checked when the scene is built
        }
        const uint32_t slot = (bit - 24u) ^ T.octinv;
        const uint32_t rel = __popc(imask & ~(0xFFFFFFFFu << slot) & 0xFFu);
#if IRIS_NODE_FP16
        const float4 *np = S.nodes + BVH8_NODE_F4 * (int64_t)(base + rel);
        const float4 n0 = ldg4(np), n1 = ldg4(np + 1);
        // near / far plane words per axis (8 fp16 each = children 0..7): words lo_x lo_y lo_z hi_z hi_y hi_x = 2..7, near + far = 9
        const uint4 nx = load_plane_word(np, T.nwx), fx = load_plane_word(np, 9u - T.nwx);
        const uint4 ny = load_plane_word(np, T.nwy), fy = load_plane_word(np, 9u - T.nwy);
        const uint4 nz = load_plane_word(np, T.nwz), fz = load_plane_word(np, 9u - T.nwz);
        const uint32_t ew = __float_as_uint(n0.w);
        T.ngroup.x = __float_as_uint(n1.x);
        T.tgroup.x = __float_as_uint(n1.y);
        const float ax = __uint_as_float((ew & 0xFFu) << 23) * T.idx;
        const float ay = __uint_as_float(((ew >> 8) & 0xFFu) << 23) * T.idy;
        const float az = __uint_as_float(((ew >> 16) & 0xFFu) << 23) * T.idz;
        const float bx = (n0.x - T.o.x) * T.idx, by = (n0.y - T.o.y) * T.idy, bz = (n0.z - T.o.z) * T.idz;
        const uint32_t nxw[4] = {nx.x, nx.y, nx.z, nx.w}, fxw[4] = {fx.x, fx.y, fx.z, fx.w};
        const uint32_t nyw[4] = {ny.x, ny.y, ny.z, ny.w}, fyw[4] = {fy.x, fy.y, fy.z, fy.w};
        const uint32_t nzw[4] = {nz.x, nz.y, nz.z, nz.w}, fzw[4] = {fz.x, fz.y, fz.z, fz.w};
        uint32_t hitmask = 0u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t meta4 = __float_as_uint(h ? n1.w : n1.z);
            const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
            const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
            const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
            const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {               // two children per word: HADD2.F32 x2 + one FFMA2 per plane pair
                const int wd = 2 * h + jj;
                float t0x[2], t1x[2], t0y[2], t1y[2], t0z[2], t1z[2];
                slab_pair_h(nxw[wd], ax, bx, t0x); slab_pair_h(fxw[wd], ax, bx, t1x);
                slab_pair_h(nyw[wd], ay, by, t0y); slab_pair_h(fyw[wd], ay, by, t1y);
                slab_pair_h(nzw[wd], az, bz, t0z); slab_pair_h(fzw[wd], az, bz, t1z);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float tn = fmaxf(fmaxf(t0x[k], t0y[k]), fmaxf(t0z[k], 0.0f));
                    const float tf = fminf(fminf(t1x[k], t1y[k]), fminf(t1z[k], T.best.t));
                    if (tn <= tf) {
                        const uint32_t cb = (child_bits4 >> (8 * (2 * jj + k))) & 0xFFu;
                        const uint32_t bi = (bit_index4 >> (8 * (2 * jj + k))) & 0xFFu;
                        hitmask |= cb << bi;
                    }
                }
            }
        }
#else
        const float4 *np = S.nodes + BVH8_NODE_F4 * (int64_t)(base + rel);
        const float4 n0 = ldg4(np), n1 = ldg4(np + 1), n2 = ldg4(np + 2), n3 = ldg4(np + 3), n4 = ldg4(np + 4);
        const uint32_t ew = __float_as_uint(n0.w);
        T.ngroup.x = __float_as_uint(n1.x);
        T.tgroup.x = __float_as_uint(n1.y);
        const float ax = __uint_as_float((ew & 0xFFu) << 23) * T.idx;
        const float ay = __uint_as_float(((ew >> 8) & 0xFFu) << 23) * T.idy;
        const float az = __uint_as_float(((ew >> 16) & 0xFFu) << 23) * T.idz;
        const float bx = (n0.x - T.o.x) * T.idx, by = (n0.y - T.o.y) * T.idy, bz = (n0.z - T.o.z) * T.idz;
        uint32_t hitmask = 0u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t meta4 = __float_as_uint(h ? n1.w : n1.z);
            const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
            const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
            const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
            const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
            const uint32_t qlx = __float_as_uint(h ? n2.y : n2.x), qly = __float_as_uint(h ? n2.w : n2.z), qlz = __float_as_uint(h ? n3.y : n3.x);
            const uint32_t qhx = __float_as_uint(h ? n3.w : n3.z), qhy = __float_as_uint(h ? n4.y : n4.x), qhz = __float_as_uint(h ? n4.w : n4.z);
            const uint32_t nx = T.sx < 0.f ? qhx : qlx, fx = T.sx < 0.f ? qlx : qhx;
            const uint32_t ny = T.sy < 0.f ? qhy : qly, fy = T.sy < 0.f ? qly : qhy;
            const uint32_t nz = T.sz < 0.f ? qhz : qlz, fz = T.sz < 0.f ? qlz : qhz;
#if defined(IRIS_HOST_EMULATION) || defined(IRIS_NODE_SCALAR)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float t0x = fmaf(byte_f(nx, j), ax, bx), t1x = fmaf(byte_f(fx, j), ax, bx);
                const float t0y = fmaf(byte_f(ny, j), ay, by), t1y = fmaf(byte_f(fy, j), ay, by);
                const float t0z = fmaf(byte_f(nz, j), az, bz), t1z = fmaf(byte_f(fz, j), az, bz);
                const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));
                const float tf = fminf(fminf(t1x, t1y), fminf(t1z, T.best.t));
                if (tn <= tf) {
                    const uint32_t cb = (child_bits4 >> (8 * j)) & 0xFFu;
                    const uint32_t bi = (bit_index4 >> (8 * j)) & 0xFFu;
                    hitmask |= cb << bi;
                }
            }
#else
            // Two children per step.  One PRMT turns two quantised bytes into the half2 (1024 + q_j, 1024 + q_j+1) (0x64 is the
            // high byte of fp16 1024, whose ulp is 1), the two f16 -> f32 conversions and the packed FFMA2 run on the FMA pipe, and
            // the 1024 is folded into the plane offset (b - 1024 a, one rounding of ~2^-14 quantisation steps).  Against one PRMT +
            // FADD + FFMA per byte this halves the ALU-pipe work of the slab test, the pipe this loop is bound by.
            const float cx = fmaf(-1024.0f, ax, bx), cy = fmaf(-1024.0f, ay, by), cz = fmaf(-1024.0f, az, bz);
#pragma unroll
            for (int j = 0; j < 4; j += 2) {
                float t0x[2], t1x[2], t0y[2], t1y[2], t0z[2], t1z[2];
                slab_pair(nx, j, ax, cx, t0x); slab_pair(fx, j, ax, cx, t1x);
                slab_pair(ny, j, ay, cy, t0y); slab_pair(fy, j, ay, cy, t1y);
                slab_pair(nz, j, az, cz, t0z); slab_pair(fz, j, az, cz, t1z);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float tn = fmaxf(fmaxf(t0x[k], t0y[k]), fmaxf(t0z[k], 0.0f));
                    const float tf = fminf(fminf(t1x[k], t1y[k]), fminf(t1z[k], T.best.t));
                    if (tn <= tf) {
                        const uint32_t cb = (child_bits4 >> (8 * (j + k))) & 0xFFu;
                        const uint32_t bi = (bit_index4 >> (8 * (j + k))) & 0xFFu;
                        hitmask |= cb << bi;
                    }
                }
            }
#endif
        }
#endif
        T.ngroup.y = (hitmask & 0xFF000000u) | (ew >> 24);
        T.tgroup.y = hitmask & 0x00FFFFFFu;
    } else {
        T.tgroup = T.ngroup;
        T.ngroup = make_uint2(0u, 0u);
    }
#if !defined(IRIS_HOST_EMULATION) && !defined(IRIS_NO_POSTPONE)
    // triangle postponing (Ylitie et al. 2017): when fewer than ~20% of the lanes that entered the triangle phase are
    // still testing triangles, the stragglers park their remaining triangles on the stack and go back to node work
    const int tri_lanes = __popc(__activemask());
#endif
    while (T.tgroup.y != 0u) {
#if !defined(IRIS_HOST_EMULATION) && !defined(IRIS_NO_POSTPONE)
        if (__popc(__activemask()) * IRIS_POSTPONE_DEN < tri_lanes * IRIS_POSTPONE_NUM && T.sp < IRIS_STACK - 1) {
            stack[T.sp++] = T.tgroup;
            break;
        }
#endif
        const uint32_t ti = 31u - __clz(T.tgroup.y);
        T.tgroup.y &= ~(1u << ti);
        const int32_t slot = (int32_t)(T.tgroup.x + ti);
        f3 v0, e1, e2;
        int32_t prim;
        load_tri(S, slot, v0, e1, e2, prim);
        float t, u, v;
        if (tri_test(T.o, T.d, T.rdd, v0, e1, e2, t, u, v)) {
            if (t < T.best.t || (t == T.best.t && prim < T.best.prim)) {
                T.best.t = t; T.best.u = u; T.best.v = v; T.best.prim = prim; T.best.slot = slot;
                if (T.anyhit) { T.done = true; break; }          // (no early return: one exit from the step)
            }
        }
    }
    if (!T.done && T.ngroup.y <= 0x00FFFFFFu) {
        if (T.sp == 0) T.done = true;
        else T.ngroup = stack[--T.sp];
    }
}

template <bool ANYHIT>
__device__ __forceinline__ Hit trace_ray_body(const SceneView &S, f3 o, f3 d, float t_limit, int32_t prim_limit, int anyhit_rt) {
    TravState T;
    uint2 stack[IRIS_STACK];
    trav_init(T, o, d, t_limit, prim_limit, (ANYHIT || anyhit_rt) ? 1 : 0);
    while (!T.done) trav_step(S, T, stack);
    return T.best;
}

#ifndef IRIS_HOST_EMULATION
__device__ __noinline__ Hit trace_ray_shared(const float4 *nodes, const float4 *tris, float ox, float oy, float oz, float dx, float dy, float dz,
                                             float t_limit, int32_t prim_limit, int anyhit) {
    SceneView S;
    S.nodes = nodes;
    S.tris = tris;
    S.n_tris = 0;
    // one loop body for both query kinds: the any-hit early exit is a run-time test on `anyhit`
    Hit h = trace_ray_body<false>(S, mk3(ox, oy, oz), mk3(dx, dy, dz), t_limit, prim_limit, anyhit);
    return h;
}
#endif

__device__ __forceinline__ Hit trace_closest(const SceneView &S, f3 o, f3 d) {
    return trace_ray<false>(S, o, d, __int_as_float(0x7f800000), -1);
}
__device__ __forceinline__ bool trace_occluded(const SceneView &S, f3 o, f3 d, float t_limit, int32_t prim_limit) {
    return trace_ray<true>(S, o, d, t_limit, prim_limit).slot >= 0;
}
#ifndef IRIS_HOST_EMULATION
// out-of-line variants for kernels that cast several rays per lane (one copy of the loop in the instruction cache)
__device__ __forceinline__ Hit trace_closest_shared(const SceneView &S, f3 o, f3 d) {
    return trace_ray_shared(S.nodes, S.tris, o.x, o.y, o.z, d.x, d.y, d.z, __int_as_float(0x7f800000), -1, 0);
}
__device__ __forceinline__ bool trace_occluded_shared(const SceneView &S, f3 o, f3 d, float t_limit, int32_t prim_limit) {
    return trace_ray_shared(S.nodes, S.tris, o.x, o.y, o.z, d.x, d.y, d.z, t_limit, prim_limit, 1).slot >= 0;
}
#endif

#ifndef IRIS_HOST_EMULATION
// ------------------------------------------------------------------------------------------------------------------------
// Block-cooperative ray re-ordering.  The lanes of a block hold the spp samples of a few neighbouring pixels: origins are
// close together but directions cover the hemisphere, and a warp of such rays diverges after a couple of BVH levels (11-14
// of 32 lanes active).  Before tracing, the block counting-sorts its rays by a 6-bit direction key (octant + which third of
// the octant), each thread traces the ray that landed in ITS slot, and the hit goes back to the owner through shared memory.
// Tracing is a pure function of the ray, so results are unchanged bit for bit; warps now hold rays inside a narrow cone.
// Every thread of the block must call this (inactive lanes pass valid = false).
// ------------------------------------------------------------------------------------------------------------------------
struct SortSmem {
    float ray[6][IRIS_SORT_BLOCK];      // o.xyz d.xyz of slot s
    float lim[IRIS_SORT_BLOCK];         // t_limit
    int plim[IRIS_SORT_BLOCK];          // prim_limit, or INT_MIN for an empty slot
    float hit[3][IRIS_SORT_BLOCK];      // t u v
    int hprim[IRIS_SORT_BLOCK], hslot[IRIS_SORT_BLOCK];
    int hist[64], base[64];
};

__device__ __forceinline__ unsigned dir_key(f3 d) {
    const unsigned oct = (d.x < 0.f ? 4u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 1u : 0u);
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    const unsigned dom = ax >= ay ? (ax >= az ? 0u : 2u) : (ay >= az ? 1u : 2u);          // dominant axis
    const float m = fmaxf(ax, fmaxf(ay, az)), lo = dom == 0u ? fminf(ay, az) : dom == 1u ? fminf(ax, az) : fminf(ax, ay);
    const unsigned steep = lo * 2.f > m ? 1u : 0u;                                          // near the octant diagonal or near the axis
    return (oct << 3) | (dom << 1) | steep;
}

template <bool ANYHIT>
__device__ __forceinline__ Hit block_sorted_trace(const SceneView &S, SortSmem &sm, f3 o, f3 d, bool valid, float t_limit, int32_t prim_limit) {
    const int tid = threadIdx.x;
    if (tid < 64) sm.hist[tid] = 0;
    __syncthreads();
    const unsigned key = dir_key(d);
    int rank = 0;
    if (valid) rank = atomicAdd(&sm.hist[key], 1);
    sm.plim[tid] = INT_MIN;
    __syncthreads();
    if (tid < 32) {                                    // exclusive scan of the 64 bins by one warp
        const int a = sm.hist[2 * tid], b = sm.hist[2 * tid + 1];
        int s = a + b;
#pragma unroll
        for (int o2 = 1; o2 < 32; o2 <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, s, o2);
            if (tid >= o2) s += v;
        }
        sm.base[2 * tid] = s - a - b;
        sm.base[2 * tid + 1] = s - b;
    }
    __syncthreads();
    int slot = -1;
    if (valid) {
        slot = sm.base[key] + rank;
        sm.ray[0][slot] = o.x; sm.ray[1][slot] = o.y; sm.ray[2][slot] = o.z;
        sm.ray[3][slot] = d.x; sm.ray[4][slot] = d.y; sm.ray[5][slot] = d.z;
        sm.lim[slot] = t_limit;
        sm.plim[slot] = prim_limit;
    }
    __syncthreads();
    if (sm.plim[tid] != INT_MIN) {
        const f3 ro = mk3(sm.ray[0][tid], sm.ray[1][tid], sm.ray[2][tid]), rd = mk3(sm.ray[3][tid], sm.ray[4][tid], sm.ray[5][tid]);
        const Hit h = trace_ray<ANYHIT>(S, ro, rd, sm.lim[tid], sm.plim[tid]);
        sm.hit[0][tid] = h.t; sm.hit[1][tid] = h.u; sm.hit[2][tid] = h.v;
        sm.hprim[tid] = h.prim; sm.hslot[tid] = h.slot;
    }
    __syncthreads();
    Hit out;
    out.t = t_limit; out.u = out.v = 0.f; out.prim = prim_limit; out.slot = -1;
    if (valid) {
        out.t = sm.hit[0][slot]; out.u = sm.hit[1][slot]; out.v = sm.hit[2][slot];
        out.prim = sm.hprim[slot]; out.slot = sm.hslot[slot];
    }
    __syncthreads();                                   // the arrays are reused by the next call
    return out;
}
#endif

// Surface record of a hit: p = fma(v,e2,fma(u,e1,v0)), n = normalize(e1 x e2) flipped toward -d (oracle finish()).
__device__ __forceinline__ void surface_from_triangle(const Hit &h, f3 d, f3 v0, f3 e1, f3 e2, f3 &p, f3 &n) {
    p = mk3(__fmaf_rn(h.v, e2.x, __fmaf_rn(h.u, e1.x, v0.x)), __fmaf_rn(h.v, e2.y, __fmaf_rn(h.u, e1.y, v0.y)),
            __fmaf_rn(h.v, e2.z, __fmaf_rn(h.u, e1.z, v0.z)));
    f3 c = xcross(e1, e2);
    float len = __fsqrt_rn(xdot(c, c));
    n = mk3(__fdiv_rn(c.x, len), __fdiv_rn(c.y, len), __fdiv_rn(c.z, len));
    if (xdot(n, d) > 0.0f) n = -n;
}
__device__ __forceinline__ void hit_surface(const SceneView &S, const Hit &h, f3 d, f3 &p, f3 &n) {
    if (h.prim < 0) { p = mk3(0.f, 0.f, 0.f); n = mk3(0.f, 0.f, 0.f); return; }
    f3 v0, e1, e2;
    int32_t prim;
    load_tri(S, h.slot, v0, e1, e2, prim);
    surface_from_triangle(h, d, v0, e1, e2, p, n);
}
