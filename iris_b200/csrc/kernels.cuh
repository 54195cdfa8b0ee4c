// kernels.cuh -- the estimator kernels: ray cast, shading-map bake, path_tracing_single forward (primary + bounce) and its
// replay adjoint.  One lane = one path sample; lane i belongs to pixel i / spp (the reference's repeat_interleave order,
// utils/path_tracing.py:340).
#pragma once
#include "shading.cuh"



// ------------------------------------------------------------------------------------------------ ray_intersect
// utils/path_tracing.py:17-48
__global__ void __launch_bounds__(IRIS_BLOCK) k_intersect(SceneView S, const float *__restrict__ o, const float *__restrict__ d, int64_t n,
                                                           float *t, int32_t *prim, float *uv, float *p, float *nrm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const f3 ro = ld3(o, i), rd = ld3(d, i);
    const Hit h = trace_closest(S, ro, rd);
    if (t) t[i] = h.t;
    if (prim) prim[i] = h.prim;
    if (uv) { uv[2 * i] = h.prim >= 0 ? h.u : 0.f; uv[2 * i + 1] = h.prim >= 0 ? h.v : 0.f; }
    if (p || nrm) {
        f3 hp, hn;
        hit_surface(S, h, rd, hp, hn);
        if (p) st3(p, i, hp);
        if (nrm) st3(nrm, i, hn);
    }
}

// Persistent variant: a fixed grid of warps pulls rays from a global counter; whenever fewer than IRIS_PERSIST_THRESH lanes of a
// warp still traverse, the finished lanes write their hit and fetch the next rays (Aila & Laine 2009, "dynamic fetch").
#ifndef IRIS_PERSIST_THRESH
#define IRIS_PERSIST_THRESH 20
#endif
// fetch counters of the persistent kernels that have no caller workspace (ray_intersect, bake): a ring, one slot per launch, so that
// launches in flight on different streams never share a counter
#define IRIS_COUNTER_RING 1024
__device__ unsigned long long g_ray_counters[IRIS_COUNTER_RING];

__global__ void __launch_bounds__(IRIS_BLOCK) k_intersect_persistent(SceneView S, const float *__restrict__ o, const float *__restrict__ d, int64_t n,
                                                                      float *t, int32_t *prim, float *uv, float *p, float *nrm, unsigned long long *counter) {
    uint2 stack[IRIS_STACK];
    TravState T;
    T.done = true;
    int64_t ray = -1;
    const unsigned lane = threadIdx.x & 31u;
    bool exhausted = false;
    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, T.done);
        if (T.done && ray >= 0) {
            const Hit h = T.best;
            if (t) t[ray] = h.t;
            if (prim) prim[ray] = h.prim;
            if (uv) { uv[2 * ray] = h.prim >= 0 ? h.u : 0.f; uv[2 * ray + 1] = h.prim >= 0 ? h.v : 0.f; }
            if (p || nrm) {
                f3 hp, hn;
                hit_surface(S, h, T.d, hp, hn);
                if (p) st3(p, ray, hp);
                if (nrm) st3(nrm, ray, hn);
            }
            ray = -1;
        }
        if (need && !exhausted) {
            const int cnt = __popc(need), leader = __ffs(need) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)cnt);
            base = __shfl_sync(0xffffffffu, base, leader);
            if (T.done) {
                const int64_t r = (int64_t)base + __popc(need & ((1u << lane) - 1u));
                if (r < n) {
                    ray = r;
                    trav_init(T, ld3(o, r), ld3(d, r), __int_as_float(0x7f800000), -1, 0);
                }
            }
            if ((int64_t)base + cnt >= n) exhausted = true;
        }
        unsigned active = __ballot_sync(0xffffffffu, !T.done);
        if (active == 0u) break;
        do {
            if (!T.done) trav_step(S, T, stack);
            active = __ballot_sync(0xffffffffu, !T.done);
        } while (active != 0u && (exhausted || __popc(active) >= IRIS_PERSIST_THRESH));
    }
}

// ------------------------------------------------------------------------------------------------ per-pixel mean over spp
// Lanes of one pixel are consecutive.  Segmented inclusive scan inside the warp, then the last lane of every segment
// adds the segment sum (already scaled by 1/spp) to the pixel: one atomic per (warp, pixel) pair.
__device__ __forceinline__ void pixel_accumulate(float *out, int64_t pixel, bool in_range, f3 v, float inv_spp) {
    const unsigned lane = threadIdx.x & 31u;
    const long long key = in_range ? (long long)pixel : -1 - (long long)lane;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float ux = __shfl_up_sync(0xffffffffu, v.x, o), uy = __shfl_up_sync(0xffffffffu, v.y, o), uz = __shfl_up_sync(0xffffffffu, v.z, o);
        const long long uk = __shfl_up_sync(0xffffffffu, key, o);
        if (lane >= (unsigned)o && uk == key) { v.x += ux; v.y += uy; v.z += uz; }
    }
    const long long nk = __shfl_down_sync(0xffffffffu, key, 1);
    if (in_range && (lane == 31u || nk != key)) {
        atomicAdd(out + 3 * pixel, v.x * inv_spp);
        atomicAdd(out + 3 * pixel + 1, v.y * inv_spp);
        atomicAdd(out + 3 * pixel + 2, v.z * inv_spp);
    }
}

// Radiance seen along a secondary ray when every non-emissive hit terminates in the SLF (trace_roughness == 0:
// bake_shading.py:121-122 and path_tracing_single, utils/path_tracing.py:395).  model/emitter.py:180-221.
// e_hit: emitter row if the hit triangle is an emitter, else -1.  valid_next as the reference defines it.
__device__ __forceinline__ f3 radiance_at_hit(const IrisShadeParams &P, const Hit &h, f3 p, int32_t &e_hit, float &emit_pdf, bool &valid_next) {
    e_hit = emitter_of(P, h.prim);
    emit_pdf = 0.f;
    valid_next = false;
    if (h.prim < 0) return mk3(0.f, 0.f, 0.f);
    if (e_hit >= 0) { emit_pdf = emitter_pdf_area(P, e_hit); return emitter_radiance(P, e_hit); }
    const f3 S = slf_lookup(P, p);
    valid_next = !(S.x + S.y + S.z > 0.f);
    return S;
}

// ------------------------------------------------------------------------------------------------ bake
// bake_shading.py:108-123 (MODE 0) and :168-188 (MODE 1)
template <int MODE>
__global__ void __launch_bounds__(IRIS_SORT_BLOCK) k_bake(SceneView S, IrisShadeParams P, IrisSampler smp, float roughness,
                                                           const float *__restrict__ position, const float *__restrict__ normal,
                                                           const float *__restrict__ wo_in, int64_t n_pixels, int spp, float *out0, float *out1) {
    __shared__ SortSmem sort;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = n_pixels * spp;
    const bool in_range = i < n;
    const int64_t pix = in_range ? i / spp : 0;
    f3 L0 = mk3(0.f, 0.f, 0.f), L1 = mk3(0.f, 0.f, 0.f);
    f3 wi = mk3(0.f, 0.f, 1.f), org = mk3(0.f, 0.f, 0.f);
    float w0 = 1.f, w1 = 0.f;
    if (in_range) {
        const f3 x = ld3(position, pix), nr = ld3(normal, pix);
        const float4 u = sample4(smp, i, 0);
        if (MODE == 0) {
            wi = diffuse_sampler(u.x, u.y, nr);
        } else {
            const f3 wo = ld3(wo_in, pix);
            wi = specular_sampler(u.x, u.y, roughness, wo, nr);
            specular_weights(wi, wo, nr, roughness, w0, w1);
        }
        org = ray_origin(x, wi);
    }
    const Hit h = block_sorted_trace<false>(S, sort, org, wi, in_range, __int_as_float(0x7f800000), -1);
    if (in_range) {
        f3 hp, hn;
        hit_surface(S, h, wi, hp, hn);
        int32_t e;
        float epdf;
        bool vn;
        const f3 Le = radiance_at_hit(P, h, hp, e, epdf, vn);
        L0 = Le * w0;
        L1 = Le * w1;
    }
    const float inv = 1.f / (float)spp;
    pixel_accumulate(out0, pix, in_range, L0, inv);
    if (MODE == 1) pixel_accumulate(out1, pix, in_range, L1, inv);
}

// Persistent form of the bake: a fixed grid of warps pulls (pixel, sample) lanes from a global counter, generates the ray in
// place, traverses, and a finished lane shades its hit, adds it to its pixel and pulls the next sample (same dynamic fetch as
// k_trace_queue, without the queue: the generator and the radiance lookup are a few dozen instructions against ~300 warp
// instructions of traversal per ray, so running them at partial occupancy costs little).
template <int MODE>
__global__ void __launch_bounds__(IRIS_BLOCK, 8) k_bake_persistent(SceneView S, IrisShadeParams P, IrisSampler smp, float roughness,
                                                                   const float *__restrict__ position, const float *__restrict__ normal,
                                                                   const float *__restrict__ wo_in, int64_t n_pixels, int spp, float *out0, float *out1,
                                                                   unsigned long long *counter) {
    uint2 stack[IRIS_STACK];
    TravState T;
    T.done = true;
    int64_t ray = -1;
    float w0 = 1.f, w1 = 0.f;
    const int64_t n = n_pixels * spp;
    const float inv = 1.f / (float)spp;
    const unsigned lane = threadIdx.x & 31u;
    bool exhausted = false;
    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, T.done);
        if (T.done && ray >= 0) {
            f3 hp, hn;
            hit_surface(S, T.best, T.d, hp, hn);
            int32_t e;
            float epdf;
            bool vn;
            const f3 Le = radiance_at_hit(P, T.best, hp, e, epdf, vn);
            const int64_t pix = ray / spp;
            if (Le.x != 0.f || Le.y != 0.f || Le.z != 0.f) {
                atomicAdd(out0 + 3 * pix, Le.x * w0 * inv);
                atomicAdd(out0 + 3 * pix + 1, Le.y * w0 * inv);
                atomicAdd(out0 + 3 * pix + 2, Le.z * w0 * inv);
                if (MODE == 1) {
                    atomicAdd(out1 + 3 * pix, Le.x * w1 * inv);
                    atomicAdd(out1 + 3 * pix + 1, Le.y * w1 * inv);
                    atomicAdd(out1 + 3 * pix + 2, Le.z * w1 * inv);
                }
            }
            ray = -1;
        }
        if (need && !exhausted) {
            const int cnt = __popc(need), leader = __ffs(need) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)cnt);
            base = __shfl_sync(0xffffffffu, base, leader);
            if (T.done) {
                const int64_t r = (int64_t)base + __popc(need & ((1u << lane) - 1u));
                if (r < n) {
                    ray = r;
                    const int64_t pix = r / spp;
                    const f3 x = ld3(position, pix), nr = ld3(normal, pix);
                    const float4 u = sample4(smp, r, 0);
                    f3 wi;
                    if (MODE == 0) {
                        wi = diffuse_sampler(u.x, u.y, nr);
                    } else {
                        const f3 wo = ld3(wo_in, pix);
                        wi = specular_sampler(u.x, u.y, roughness, wo, nr);
                        specular_weights(wi, wo, nr, roughness, w0, w1);
                    }
                    trav_init(T, ray_origin(x, wi), wi,
                              __int_as_float(0x7f800000), -1, 0);
#ifndef IRIS_NO_RAY_SKIP
                    // a specular sample below the horizon has both weights exactly zero: whatever it hits adds nothing -> not traced
                    if (MODE == 1 && w0 == 0.f && w1 == 0.f) { T.done = true; ray = -1; }
#endif
                }
            }
            if ((int64_t)base + cnt >= n) exhausted = true;
        }
        unsigned active = __ballot_sync(0xffffffffu, !T.done);
        if (active == 0u) {
            if (exhausted) break;
            continue;
        }
        do {
            if (!T.done) trav_step(S, T, stack);
            active = __ballot_sync(0xffffffffu, !T.done);
        } while (active != 0u && (exhausted || __popc(active) >= IRIS_PERSIST_THRESH));
    }
}

// ------------------------------------------------------------------------------------------------ path_tracing_single
// Workspace per lane: three float4.
//   w0 = (x0.xyz, code)   code: -2 = continue (valid_next), -1 = miss, >= 0 = emitter row hit by the primary ray
//   w1 = (n0.xyz, metallic)        w2 = (albedo.rgb, roughness)      [w1.w, w2 written by the field kernel]
__device__ __forceinline__ f3 camera_dir(const float *__restrict__ rays, int64_t pix, float u0, float u1) {
    // utils/path_tracing.py:338-339, one rounding per op like the ATen chain
    const float *r = rays + 12 * pix;
    const float du = xsub(u0, 0.5f), dv = xsub(u1, 0.5f);
    const f3 v = mk3(xadd(xadd(r[3], xmul(r[6], du)), xmul(r[9], dv)), xadd(xadd(r[4], xmul(r[7], du)), xmul(r[10], dv)),
                     xadd(xadd(r[5], xmul(r[8], du)), xmul(r[11], dv)));
    return normalize_nf(v);
}

__global__ void __launch_bounds__(IRIS_BLOCK) k_primary(SceneView S, IrisShadeParams P, IrisSampler smp, const float *__restrict__ rays,
                                                         int64_t n_pixels, int spp, float4 *__restrict__ w0, float4 *__restrict__ w1) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels * spp) return;
    const int64_t pix = i / spp;
    const float4 u = sample4(smp, i, 0);
    const f3 o = mk3(rays[12 * pix], rays[12 * pix + 1], rays[12 * pix + 2]);
    const f3 wi = camera_dir(rays, pix, u.x, u.y);
    const Hit h = trace_closest(S, o, wi);
    f3 hp, hn;
    hit_surface(S, h, wi, hp, hn);
    int code = -1;
    if (h.prim >= 0) {
        const int32_t e = emitter_of(P, h.prim);
        code = e >= 0 ? e : -2;
    }
    w0[i] = make_float4(hp.x, hp.y, hp.z, __int_as_float(code));
    w1[i] = make_float4(hn.x, hn.y, hn.z, 0.f);
}

// Backward record per lane: six float4 (SoA by word so that lanes coalesce).
//   r0 = (e0, e_nee, e_b, -)        emitter rows of the three radiance gathers (-1 = none), as int bits
//   r1 = (c_nee.xyz, c_b.x)         dL/dradiance[e_nee] = g * c_nee,  dL/dradiance[e_b] = g * c_b,  dL/dradiance[e0] = g
//   r2 = (c_b.yz, Ja.xy)   r3 = (Ja.z, Jr.xyz)   r4 = (Jm.xyz, -)     J* = d(L_c)/d(albedo_c | roughness | metallic)
//   r5 = (x0.xyz, code)
#ifndef IRIS_BOUNCE_MINBLOCKS
#define IRIS_BOUNCE_MINBLOCKS 5
#endif
// ------------------------------------------------------------------------------------------------ ray queue
// Wavefront form of the secondary bounce: a generator kernel writes rays, ONE small persistent kernel traces them with dynamic
// fetch (finished lanes pull the next ray, so warps stay full on incoherent rays and the loop fits the instruction cache), and a
// shading kernel consumes the hits.  A ray is two float4: (origin, t_limit) and (direction, prim_limit); t_limit < 0 marks an
// empty slot.  Rays [0, n_anyhit) are occlusion queries against the candidate (t_limit, prim_limit), the rest closest-hit.
// A hit is one float4 (t, u, v, slot); slot < 0 = miss / unoccluded.
#ifndef IRIS_QUEUE_MINBLOCKS
#define IRIS_QUEUE_MINBLOCKS 8
#endif
__global__ void __launch_bounds__(IRIS_BLOCK, IRIS_QUEUE_MINBLOCKS) k_trace_queue(SceneView S, const float4 *__restrict__ ro, const float4 *__restrict__ rd, int64_t n_rays,
                                                             int64_t n_anyhit, float4 *__restrict__ hit, unsigned long long *counter,
                                                             const unsigned long long *__restrict__ n_lanes_dev, int rays_per_lane) {
    // n_lanes_dev != NULL: the queue holds rays_per_lane x (*n_lanes_dev) rays (the live lanes of a wavefront bounce, counted on the
    // device); with two rays per lane the first half are the occlusion queries
    if (n_lanes_dev != nullptr) {
        const int64_t c = (int64_t)*n_lanes_dev;
        n_rays = rays_per_lane * c;
        n_anyhit = rays_per_lane == 2 ? c : 0;
    }
    uint2 stack[IRIS_STACK];
    TravState T;
    T.done = true;
    int64_t ray = -1;
    const unsigned lane = threadIdx.x & 31u;
    bool exhausted = false;
    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, T.done);
        if (T.done && ray >= 0) {
            hit[ray] = make_float4(T.best.t, T.best.u, T.best.v, __int_as_float(T.best.slot));
            ray = -1;
        }
        if (need && !exhausted) {
            const int cnt = __popc(need), leader = __ffs(need) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)cnt);
            base = __shfl_sync(0xffffffffu, base, leader);
            if (T.done) {
                const int64_t r = (int64_t)base + __popc(need & ((1u << lane) - 1u));
                if (r < n_rays) {
                    // (no branch around trav_init: an empty slot is initialised like a ray and retired at once -- branching around the
                    // inlined initialisation costs the compiler's code generation dearly, see k_bake_persistent)
                    const float4 a = ro[r], b = rd[r];
                    ray = r;
                    trav_init(T, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), a.w, __float_as_int(b.w), r < n_anyhit ? 1 : 0);
                    if (!(a.w >= 0.f)) {
                        T.done = true;
                        ray = -1;
                        hit[r] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
                    }
                }
            }
            if ((int64_t)base + cnt >= n_rays) exhausted = true;
        }
        unsigned active = __ballot_sync(0xffffffffu, !T.done);
        if (active == 0u) {
            if (exhausted) break;
            continue;
        }
        do {
            if (!T.done) trav_step(S, T, stack);
            active = __ballot_sync(0xffffffffu, !T.done);
        } while (active != 0u && (exhausted || __popc(active) >= IRIS_PERSIST_THRESH));
    }
}

// Queue hit -> Hit + surface point / normal (one triangle record load).
__device__ __forceinline__ Hit queue_hit_surface(const SceneView &S, float4 q, f3 d, f3 &p, f3 &n) {
    Hit h;
    h.t = q.x; h.u = q.y; h.v = q.z;
    h.slot = __float_as_int(q.w);
    h.prim = -1;
    p = mk3(0.f, 0.f, 0.f);
    n = p;
    if (h.slot >= 0) {
        f3 v0, e1, e2;
        load_tri(S, h.slot, v0, e1, e2, h.prim);
        surface_from_triangle(h, d, v0, e1, e2, p, n);
    }
    return h;
}

// bake through the ray queue (bake_shading.py:108-123 / :168-188): sample a direction per lane, trace, weight the radiance found.
template <int MODE>
__global__ void __launch_bounds__(IRIS_BLOCK) k_bake_gen(IrisSampler smp, float roughness, const float *__restrict__ position, const float *__restrict__ normal,
                                                          const float *__restrict__ wo_in, int64_t i0, int64_t nc, int spp, float4 *__restrict__ ro,
                                                          float4 *__restrict__ rd) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nc) return;
    const int64_t i = i0 + j, pix = i / spp;
    const f3 x = ld3(position, pix), nr = ld3(normal, pix);
    const float4 u = sample4(smp, i, 0);
    const f3 wi = MODE == 0 ? diffuse_sampler(u.x, u.y, nr) : specular_sampler(u.x, u.y, roughness, ld3(wo_in, pix), nr);
    ro[j] = ray4(ray_origin(x, wi), __int_as_float(0x7f800000));
    rd[j] = make_float4(wi.x, wi.y, wi.z, __int_as_float(-1));
}

template <int MODE>
__global__ void __launch_bounds__(IRIS_BLOCK) k_bake_shade(SceneView S, IrisShadeParams P, float roughness, const float *__restrict__ normal,
                                                            const float *__restrict__ wo_in, int64_t i0, int64_t nc, int spp, const float4 *__restrict__ rd,
                                                            const float4 *__restrict__ hit, float *out0, float *out1) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = j < nc;
    const int64_t i = i0 + j, pix = in_range ? i / spp : 0;
    f3 L0 = mk3(0.f, 0.f, 0.f), L1 = L0;
    if (in_range) {
        const float4 dq = rd[j];
        const f3 wi = mk3(dq.x, dq.y, dq.z);
        float w0 = 1.f, w1 = 0.f;
        if (MODE == 1) specular_weights(wi, ld3(wo_in, pix), ld3(normal, pix), roughness, w0, w1);
        f3 hp, hn;
        const Hit h = queue_hit_surface(S, hit[j], wi, hp, hn);
        int32_t e;
        float epdf;
        bool vn;
        const f3 Le = radiance_at_hit(P, h, hp, e, epdf, vn);
        L0 = Le * w0;
        L1 = Le * w1;
    }
    const float inv = 1.f / (float)spp;
    pixel_accumulate(out0, pix, in_range, L0, inv);
    if (MODE == 1) pixel_accumulate(out1, pix, in_range, L1, inv);
}

// path_tracing_single, generator half: everything of utils/path_tracing.py:359-404 that does not depend on a ray cast.
// Per-sample state streams (float4, chunk-local index j, stride nc):
//   s0 = (L_nee.rgb if unoccluded, pdf_b)   s1 = (wb.rgb, e_nee)                      e_nee < 0: no shadow ray
//   RECORD: s2 = (c_nee.rgb, JaN.x) s3 = (JaN.yz, JrN.xy) s4 = (JrN.z, JmN.xyz) s5 = (Jb.da, Jb.dr.x) s6 = (Jb.dr.yz, Jb.dm.xy) s7 = (Jb.dm.z,..)
#define IRIS_SINGLE_STATE_STREAMS 8
// REC: 0 = inference, 1 = record for the emitter gradient only (no BRDF Jacobians: train_emitter.py), 2 = full record
template <int REC>
__global__ void __launch_bounds__(IRIS_BLOCK) k_single_gen(IrisShadeParams P, IrisSampler smp, const float *__restrict__ rays, int64_t i0, int64_t nc, int spp,
                                                            const float4 *__restrict__ w0, const float4 *__restrict__ w1, const float4 *__restrict__ w2,
                                                            float4 *__restrict__ ro, float4 *__restrict__ rd, float4 *__restrict__ st) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nc) return;
    const int64_t i = i0 + j;
    const float4 a = w0[i];
    float4 ro_s = make_float4(0.f, 0.f, 0.f, -1.f), rd_s = ro_s, ro_b = ro_s, rd_b = ro_s;
    if (__float_as_int(a.w) == -2) {
        const float4 b = w1[i], c = w2[i];
        const f3 x0 = mk3(a.x, a.y, a.z), n0 = mk3(b.x, b.y, b.z);
        Mat mat;
        mat.a = mk3(c.x, c.y, c.z);
        mat.r = c.w;
        mat.m = b.w;
        const float4 ua = sample4(smp, i, 0), ub = sample4(smp, i, 1);
        const f3 wo = -camera_dir(rays, i / spp, ua.x, ua.y);
        int32_t e_nee = -1;
        f3 L_nee = mk3(0.f, 0.f, 0.f), c_nee = L_nee, JaN = L_nee, JrN = L_nee, JmN = L_nee;
        {   // emitter sample: the contribution if the shadow ray comes back unoccluded (see k_bounce_single for the equivalence)
            f3 wi;
            float pdf_e;
            int32_t e, face;
            sample_emitter(P, ua.z, ua.w, ub.x, x0, wi, pdf_e, e, face);
            const f3 org = ray_origin(x0, wi);
            f3 v0, e1, e2;
            emitter_triangle(P, e, v0, e1, e2);
            float tl, bu, bv;
            if (tri_test(org, wi, __fdiv_rn(1.0f, xdot(wi, wi)), v0, e1, e2, tl, bu, bv)) {
                ro_s = make_float4(org.x, org.y, org.z, tl);
                rd_s = make_float4(wi.x, wi.y, wi.z, __int_as_float(face));
                Hit h;
                h.t = tl; h.u = bu; h.v = bv; h.prim = face; h.slot = -1;
                f3 hp, hn;
                surface_from_triangle(h, wi, v0, e1, e2, hp, hn);
                const f3 dlt = hp - x0;
                const float G = fabsf(-(wi.x * hn.x) - (wi.y * hn.y) - (wi.z * hn.z)) / fmaxf(dot(dlt, dlt), 1e-6f);
                f3 f;
                float pdf_b;
                BrdfJac J;
                eval_brdf<REC == 2>(wi, wo, n0, mat, f, pdf_b, &J);
                pdf_b *= G;
                float w = (pdf_e > 0.f && !isinf(pdf_b)) ? pdf_e * pdf_e / fmaxf(pdf_e * pdf_e + pdf_b * pdf_b, 1e-6f) : 0.f;
                if (isinf(pdf_e) || pdf_b == 0.f) w = 1.f;
                const float s = G / fmaxf(pdf_e, 1e-6f) * w;
                const f3 W = emitter_radiance(P, e) * s;
                L_nee = f * W;
                e_nee = e;
                if (REC >= 1) c_nee = f * s;
                if (REC == 2) { JaN = J.da * W; JrN = J.dr * W; JmN = J.dm * W; }
#ifndef IRIS_NO_RAY_SKIP
                // an emitter below the horizon of x0 (NoL clamps to 0): the contribution, its coefficient and its Jacobian are all exactly
                // zero whatever the shadow ray returns -> no shadow ray (a NaN anywhere fails the test and keeps the ray)
                bool dead = is_zero3(L_nee);
                if (REC >= 1) dead = dead && is_zero3(c_nee);
                if (REC == 2) dead = dead && is_zero3(JaN) && is_zero3(JrN) && is_zero3(JmN);
                if (dead) { ro_s = make_float4(0.f, 0.f, 0.f, -1.f); rd_s = ro_s; e_nee = -1; }
#endif
            }
        }
        f3 wi, wb;
        float pdf_b;
        BrdfJac J;
        sample_brdf<REC == 2>(ub.y, ub.z, ub.w, wo, n0, mat, wi, pdf_b, wb, &J);
        ro_b = ray4(ray_origin(x0, wi), __int_as_float(0x7f800000));
        rd_b = make_float4(wi.x, wi.y, wi.z, __int_as_float(-1));
#ifndef IRIS_NO_RAY_SKIP
        {   // a BSDF sample with zero weight (direction below the horizon, or pdf == 0) and zero Jacobian adds nothing to L or to any
            // gradient whatever it hits -> no ray; k_single_shade sees a miss (Le = 0) for it
            bool dead = is_zero3(wb);
            if (REC == 2) dead = dead && is_zero3(J.da) && is_zero3(J.dr) && is_zero3(J.dm);
            if (dead) ro_b = make_float4(0.f, 0.f, 0.f, -1.f);
        }
#endif
        st[j] = make_float4(L_nee.x, L_nee.y, L_nee.z, pdf_b);
        st[nc + j] = make_float4(wb.x, wb.y, wb.z, __int_as_float(e_nee));
        if (REC >= 1) st[2 * nc + j] = make_float4(c_nee.x, c_nee.y, c_nee.z, JaN.x);
        if (REC == 2) {
            st[3 * nc + j] = make_float4(JaN.y, JaN.z, JrN.x, JrN.y);
            st[4 * nc + j] = make_float4(JrN.z, JmN.x, JmN.y, JmN.z);
            st[5 * nc + j] = make_float4(J.da.x, J.da.y, J.da.z, J.dr.x);
            st[6 * nc + j] = make_float4(J.dr.y, J.dr.z, J.dm.x, J.dm.y);
            st[7 * nc + j] = make_float4(J.dm.z, 0.f, 0.f, 0.f);
        }
    }
    ro[j] = ro_s; rd[j] = rd_s;
    ro[nc + j] = ro_b; rd[nc + j] = rd_b;
}

// path_tracing_single, shading half: visibility, emitter / SLF radiance at the BSDF hit, MIS, the adjoint record, pixel mean.
template <int REC>
__global__ void __launch_bounds__(IRIS_BLOCK) k_single_shade(SceneView S, IrisShadeParams P, int64_t i0, int64_t nc, int64_t n, int spp,
                                                              const float4 *__restrict__ w0, const float4 *__restrict__ rd, const float4 *__restrict__ hit,
                                                              const float4 *__restrict__ st, float *L_out, float4 *__restrict__ rec) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = j < nc;
    const int64_t i = i0 + j;
    const int64_t pix = in_range ? i / spp : 0;
    f3 L = mk3(0.f, 0.f, 0.f);
    if (in_range) {
        const float4 a = w0[i];
        const int code = __float_as_int(a.w);
        int32_t e0 = -1, e_nee = -1, e_b = -1;
        f3 c_nee = mk3(0.f, 0.f, 0.f), c_b = mk3(0.f, 0.f, 0.f);
        f3 Ja = c_nee, Jr = c_nee, Jm = c_nee;
        if (code >= 0) {
            e0 = code;
            L = emitter_radiance(P, e0);
        } else if (code == -2) {
            const f3 x0 = mk3(a.x, a.y, a.z);
            const float4 s0 = st[j], s1 = st[nc + j];
            if (__float_as_int(s1.w) >= 0 && __float_as_int(hit[j].w) < 0) {
                L = mk3(s0.x, s0.y, s0.z);
                if (REC >= 1) {
                    const float4 s2 = st[2 * nc + j];
                    e_nee = __float_as_int(s1.w);
                    c_nee = mk3(s2.x, s2.y, s2.z);
                    if (REC == 2) {
                        const float4 s3 = st[3 * nc + j], s4 = st[4 * nc + j];
                        Ja = mk3(s2.w, s3.x, s3.y); Jr = mk3(s3.z, s3.w, s4.x); Jm = mk3(s4.y, s4.z, s4.w);
                    }
                }
            }
            const float4 dq = rd[nc + j];
            const f3 wi = mk3(dq.x, dq.y, dq.z), wb = mk3(s1.x, s1.y, s1.z);
            float pdf_b = s0.w;
            f3 hp, hn;
            const Hit h = queue_hit_surface(S, hit[nc + j], wi, hp, hn);
            int32_t eh;
            float pdf_e;
            bool vn;
            const f3 Le = radiance_at_hit(P, h, hp, eh, pdf_e, vn);
            float G = 1.f;
            if (vn) {
                const f3 dlt = x0 - hp;
                G = fabsf(-(hn.x * wi.x) - (hn.y * wi.y) - (hn.z * wi.z)) / fmaxf(dot(dlt, dlt), 1e-6f);
            }
            pdf_b *= G;
            float w = (pdf_b > 0.f && !isinf(pdf_e)) ? pdf_b * pdf_b / (pdf_e * pdf_e + pdf_b * pdf_b) : 0.f;
            if (isinf(pdf_b) || pdf_e == 0.f) w = 1.f;
            L = L + wb * Le * w;
            if (REC >= 1 && eh >= 0) { e_b = eh; c_b = wb * w; }
            if (REC == 2) {
                const float4 s5 = st[5 * nc + j], s6 = st[6 * nc + j], s7 = st[7 * nc + j];
                const f3 Lw = Le * w;
                Ja = Ja + mk3(s5.x, s5.y, s5.z) * Lw; Jr = Jr + mk3(s5.w, s6.x, s6.y) * Lw; Jm = Jm + mk3(s6.z, s6.w, s7.x) * Lw;
            }
        }
        if (REC >= 1) {
            rec[i] = make_float4(__int_as_float(e0), __int_as_float(e_nee), __int_as_float(e_b), 0.f);
            rec[n + i] = make_float4(c_nee.x, c_nee.y, c_nee.z, c_b.x);
            rec[2 * n + i] = make_float4(c_b.y, c_b.z, Ja.x, Ja.y);
            if (REC == 2) {        // the Jacobian words are only written (and only read back) when the field gradient is wanted
                rec[3 * n + i] = make_float4(Ja.z, Jr.x, Jr.y, Jr.z);
                rec[4 * n + i] = make_float4(Jm.x, Jm.y, Jm.z, 0.f);
            }
            rec[5 * n + i] = a;
        }
    }
    pixel_accumulate(L_out, pix, in_range, L, 1.f / (float)spp);
}

template <bool RECORD>
__global__ void __launch_bounds__(IRIS_BLOCK, IRIS_BOUNCE_MINBLOCKS) k_bounce_single(SceneView S, IrisShadeParams P, IrisSampler smp, const float *__restrict__ rays,
                                                               int64_t n_pixels, int spp, const float4 *__restrict__ w0,
                                                               const float4 *__restrict__ w1, const float4 *__restrict__ w2, float *L_out,
                                                               float4 *__restrict__ rec) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = n_pixels * spp;
    const bool in_range = i < n;
    const int64_t pix = in_range ? i / spp : 0;
    f3 L = mk3(0.f, 0.f, 0.f);
    if (in_range) {
        const float4 a = w0[i];
        const int code = __float_as_int(a.w);
        int32_t e0 = -1, e_nee = -1, e_b = -1;
        f3 c_nee = mk3(0.f, 0.f, 0.f), c_b = mk3(0.f, 0.f, 0.f);
        f3 Ja = c_nee, Jr = c_nee, Jm = c_nee;
        if (code >= 0) {
            e0 = code;
            L = emitter_radiance(P, e0);
        } else if (code == -2) {
            const float4 b = w1[i], c = w2[i];
            const f3 x0 = mk3(a.x, a.y, a.z), n0 = mk3(b.x, b.y, b.z);
            Mat mat;
            mat.a = mk3(c.x, c.y, c.z);
            mat.r = c.w;
            mat.m = b.w;
            const float4 ua = sample4(smp, i, 0), ub = sample4(smp, i, 1);
            const f3 wo = -camera_dir(rays, pix, ua.x, ua.y);
            // ---- emitter sampling + shadow ray + MIS (utils/path_tracing.py:359-382)
            // The reference casts a closest-hit ray and keeps the sample only if the hit triangle IS the sampled emitter
            // triangle (a miss contributes Le = 0).  Equivalent and cheaper: intersect the sampled triangle directly, then
            // ask whether anything beats it under the closest-hit ordering (any-hit query bounded by its t).
            {
                f3 wi;
                float pdf_e;
                int32_t e, face;
                sample_emitter(P, ua.z, ua.w, ub.x, x0, wi, pdf_e, e, face);
                const f3 org = ray_origin(x0, wi);
                f3 v0, e1, e2;
                emitter_triangle(P, e, v0, e1, e2);
                float tl, bu, bv;
                if (tri_test(org, wi, __fdiv_rn(1.0f, xdot(wi, wi)), v0, e1, e2, tl, bu, bv) && !trace_occluded_shared(S, org, wi, tl, face)) {
                    Hit h;
                    h.t = tl; h.u = bu; h.v = bv; h.prim = face; h.slot = -1;
                    f3 hp, hn;
                    surface_from_triangle(h, wi, v0, e1, e2, hp, hn);
                    const f3 dlt = hp - x0;
                    const float G = fabsf(-(wi.x * hn.x) - (wi.y * hn.y) - (wi.z * hn.z)) / fmaxf(dot(dlt, dlt), 1e-6f);
                    f3 f;
                    float pdf_b;
                    BrdfJac J;
                    eval_brdf<RECORD>(wi, wo, n0, mat, f, pdf_b, &J);
                    pdf_b *= G;
                    float w = (pdf_e > 0.f && !isinf(pdf_b)) ? pdf_e * pdf_e / fmaxf(pdf_e * pdf_e + pdf_b * pdf_b, 1e-6f) : 0.f;
                    if (isinf(pdf_e) || pdf_b == 0.f) w = 1.f;
                    const float s = G / fmaxf(pdf_e, 1e-6f) * w;
                    const f3 W = emitter_radiance(P, e) * s;
                    L = L + f * W;
                    if (RECORD) {
                        const f3 cn = f * s, ja = J.da * W, jr = J.dr * W, jm = J.dm * W;
                        bool dead = false;
#ifndef IRIS_NO_RAY_SKIP
                        dead = is_zero3(f * W) && is_zero3(cn) && is_zero3(ja) && is_zero3(jr) && is_zero3(jm);   // k_single_gen drops such samples
#endif
                        if (!dead) { e_nee = e; c_nee = cn; }
                        Ja = Ja + ja; Jr = Jr + jr; Jm = Jm + jm;
                    }
                }
            }
            // ---- BSDF sampling + ray + emitter / SLF radiance + MIS (utils/path_tracing.py:385-404)
            {
                f3 wi, wb;
                float pdf_b;
                BrdfJac J;
                sample_brdf<RECORD>(ub.y, ub.z, ub.w, wo, n0, mat, wi, pdf_b, wb, &J);
                const f3 org = ray_origin(x0, wi);
                bool dead = false;
#ifndef IRIS_NO_RAY_SKIP
                dead = is_zero3(wb);                                   // zero weight (and Jacobian): no ray, as in k_single_gen
                if (RECORD) dead = dead && is_zero3(J.da) && is_zero3(J.dr) && is_zero3(J.dm);
#endif
                Hit h;
                h.t = __int_as_float(0x7f800000); h.u = h.v = 0.f; h.prim = -1; h.slot = -1;
                if (!dead) h = trace_closest_shared(S, org, wi);
                f3 hp, hn;
                hit_surface(S, h, wi, hp, hn);
                int32_t eh;
                float pdf_e;
                bool vn;
                const f3 Le = radiance_at_hit(P, h, hp, eh, pdf_e, vn);
                float G = 1.f;
                if (vn) {
                    const f3 dlt = x0 - hp;
                    G = fabsf(-(hn.x * wi.x) - (hn.y * wi.y) - (hn.z * wi.z)) / fmaxf(dot(dlt, dlt), 1e-6f);
                }
                pdf_b *= G;
                float w = (pdf_b > 0.f && !isinf(pdf_e)) ? pdf_b * pdf_b / (pdf_e * pdf_e + pdf_b * pdf_b) : 0.f;
                if (isinf(pdf_b) || pdf_e == 0.f) w = 1.f;
                L = L + wb * Le * w;
                if (RECORD) {
                    if (eh >= 0) { e_b = eh; c_b = wb * w; }
                    const f3 Lw = Le * w;
                    Ja = Ja + J.da * Lw; Jr = Jr + J.dr * Lw; Jm = Jm + J.dm * Lw;
                }
            }
        }
        if (RECORD) {
            rec[i] = make_float4(__int_as_float(e0), __int_as_float(e_nee), __int_as_float(e_b), 0.f);
            rec[n + i] = make_float4(c_nee.x, c_nee.y, c_nee.z, c_b.x);
            rec[2 * n + i] = make_float4(c_b.y, c_b.z, Ja.x, Ja.y);
            rec[3 * n + i] = make_float4(Ja.z, Jr.x, Jr.y, Jr.z);
            rec[4 * n + i] = make_float4(Jm.x, Jm.y, Jm.z, 0.f);
            rec[5 * n + i] = a;
        }
    }
    pixel_accumulate(L_out, pix, in_range, L, 1.f / (float)spp);
}

// ------------------------------------------------------------------------------------------------ adjoint
// Replays the record: g = dL[pixel]/spp per lane.  Emitter-radiance gradients are reduced in per-thread shared-memory
// accumulators over a grid-stride loop (no atomics, no bank conflicts), then across the block, and leave the block as
// ONE atomic per emitter row and channel.  d_mat (n,5) = J^T g feeds the field adjoint.
#define IRIS_BWD_KMAX 96   // 96 rows * 3 * 128 threads * 4 B = 147 KB dynamic shared memory
__global__ void __launch_bounds__(IRIS_BLOCK) k_single_backward(const float *__restrict__ dL, int64_t n_pixels, int spp,
                                                                 const float4 *__restrict__ rec, int K, float *d_radiance, float *d_mat) {
    extern __shared__ float acc[];   // [K*3][IRIS_BLOCK] when K <= IRIS_BWD_KMAX
    const bool priv = K <= IRIS_BWD_KMAX && d_radiance != nullptr;
    const int tid = threadIdx.x;
    if (priv)
        for (int r = 0; r < 3 * K; ++r) acc[r * IRIS_BLOCK + tid] = 0.f;
    const int64_t n = n_pixels * spp;
    const float inv = 1.f / (float)spp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pix = i / spp;
        const f3 g = mk3(dL[3 * pix] * inv, dL[3 * pix + 1] * inv, dL[3 * pix + 2] * inv);
        const float4 r0 = rec[i], r1 = rec[n + i], r2 = rec[2 * n + i];
        const int e0 = __float_as_int(r0.x), en = __float_as_int(r0.y), eb = __float_as_int(r0.z);
        if (d_radiance) {
            if (priv) {
                if (e0 >= 0) { acc[(3 * e0) * IRIS_BLOCK + tid] += g.x; acc[(3 * e0 + 1) * IRIS_BLOCK + tid] += g.y; acc[(3 * e0 + 2) * IRIS_BLOCK + tid] += g.z; }
                if (en >= 0) { acc[(3 * en) * IRIS_BLOCK + tid] += g.x * r1.x; acc[(3 * en + 1) * IRIS_BLOCK + tid] += g.y * r1.y; acc[(3 * en + 2) * IRIS_BLOCK + tid] += g.z * r1.z; }
                if (eb >= 0) { acc[(3 * eb) * IRIS_BLOCK + tid] += g.x * r1.w; acc[(3 * eb + 1) * IRIS_BLOCK + tid] += g.y * r2.x; acc[(3 * eb + 2) * IRIS_BLOCK + tid] += g.z * r2.y; }
            } else {
                if (e0 >= 0) { atomicAdd(d_radiance + 3 * e0, g.x); atomicAdd(d_radiance + 3 * e0 + 1, g.y); atomicAdd(d_radiance + 3 * e0 + 2, g.z); }
                if (en >= 0) { atomicAdd(d_radiance + 3 * en, g.x * r1.x); atomicAdd(d_radiance + 3 * en + 1, g.y * r1.y); atomicAdd(d_radiance + 3 * en + 2, g.z * r1.z); }
                if (eb >= 0) { atomicAdd(d_radiance + 3 * eb, g.x * r1.w); atomicAdd(d_radiance + 3 * eb + 1, g.y * r2.x); atomicAdd(d_radiance + 3 * eb + 2, g.z * r2.y); }
            }
        }
        if (d_mat) {
            const float4 r3 = rec[3 * n + i], r4 = rec[4 * n + i];
            float *o = d_mat + 5 * i;
            o[0] = g.x * r2.z;
            o[1] = g.y * r2.w;
            o[2] = g.z * r3.x;
            o[3] = g.x * r3.y + g.y * r3.z + g.z * r3.w;
            o[4] = g.x * r4.x + g.y * r4.y + g.z * r4.z;
        }
    }
    if (priv && d_radiance) {
        __syncthreads();
        // row r is reduced by warp (r % 4): 128 partials -> 1
        const int warp = tid >> 5, lane = tid & 31;
        for (int r = warp; r < 3 * K; r += IRIS_BLOCK / 32) {
            float v = acc[r * IRIS_BLOCK + lane] + acc[r * IRIS_BLOCK + lane + 32] + acc[r * IRIS_BLOCK + lane + 64] + acc[r * IRIS_BLOCK + lane + 96];
            v = warp_sum(v);
            if (lane == 0 && v != 0.f) atomicAdd(d_radiance + r, v);
        }
    }
}
