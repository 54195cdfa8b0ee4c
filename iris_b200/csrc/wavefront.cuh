// wavefront.cuh -- the forward-only estimators that need the BRDF field at EVERY hit: path_tracing with indirect bounces
// (reference utils/path_tracing.py:214-318), trace_indirect (:409-502), path_tracing_det_diff / _det_spec (:50-212).
// The reference decides at each BSDF hit whether to stop in the surface light field from the roughness of the material
// THERE (emitter.py:210-219, threshold 0.6), so the field (tensor-core MLP kernel, field.cuh) has to run between "trace" and
// "shade": each bounce is bounce_a (NEE + BSDF sample + closest hit) -> k_field_forward on the hit points -> bounce_b
// (emitter / SLF radiance, MIS, state update).  Lane state lives in HBM as float4 streams (208 B per lane).
//
// Lane compaction (the reference shrinks its lane set at every depth with boolean masks, utils/path_tracing.py:347-352,492-501):
// the ray-queue form keeps a list of the lanes that are still alive (`idx`, `cnt` on the device; warp-aggregated appends in
// k_wave_init_* / k_wave_bounce_b).  Per bounce, thread j of the generator / resolve / shade kernels works on lane idx[j], the ray
// queue, the hit records and the material scratch are indexed by j (dense), the persistent trace kernel and the field kernel read
// the live count from device memory, and blocks past the count exit at once -- no host synchronisation, no work on dead lanes.
// Per-lane results do not depend on the order of the list (uniforms are keyed by the lane id), so outputs are unchanged bit for bit.
#pragma once
#include "kernels.cuh"

struct WaveState {
    float4 *S0;   // x.xyz | code (-2 = active)
    float4 *S1;   // n.xyz | metallic
    float4 *S2;   // albedo.rgb | roughness
    float4 *S3;   // wo.xyz
    float4 *S4;   // throughput.rgb
    float4 *S5;   // radiance gathered by the indirect loop
    float4 *S6;   // first-bounce radiance
    float4 *S7;   // first BSDF weight (w0)
    float4 *S8;   // second specular weight (w1, det_spec)
    float4 *H0;   // next hit: p.xyz | code (-2 = needs material)
    float4 *H1;   // next hit: n.xyz | bsdf pdf
    float4 *H2;   // sampled wi.xyz | hit prim (int bits)
    float4 *M1;   // material at next hit:  - | metallic
    float4 *M2;   // material at next hit: albedo.rgb | roughness
    // ray queue of one bounce (k_wave_gen -> k_trace_queue -> k_wave_resolve): lane i owns shadow ray i and closest-hit ray n + i
    float4 *RO;   // 2n: origin | t_limit (< 0 = empty slot)
    float4 *RD;   // 2n: direction | prim_limit
    float4 *HIT;  // 2n: t u v | slot
    float4 *PN;   // n: emitter-sample contribution to add if the shadow ray is unoccluded | bsdf pdf
    int32_t *idxA, *idxB;         // live-lane lists (ping-pong between bounces)
    unsigned long long *counter;  // dynamic-fetch counter of k_trace_queue
    unsigned long long *cnt;      // cnt[k] = number of live lanes entering bounce k (WAVE_MAX_BOUNCES + 1 entries)
};
#define WAVE_STREAMS 22          // 14 state + 7 queue streams per lane + one stream that holds the two live-lane lists
#define WAVE_TAIL_BYTES 1024     // fetch counter + live counts, after the streams
#define WAVE_MAX_BOUNCES 64
__host__ __device__ inline WaveState wave_carve(void *base, int64_t n) {
    float4 *p = reinterpret_cast<float4 *>(base);
    WaveState w;
    w.S0 = p; w.S1 = p + n; w.S2 = p + 2 * n; w.S3 = p + 3 * n; w.S4 = p + 4 * n; w.S5 = p + 5 * n; w.S6 = p + 6 * n; w.S7 = p + 7 * n;
    w.S8 = p + 8 * n; w.H0 = p + 9 * n; w.H1 = p + 10 * n; w.H2 = p + 11 * n; w.M1 = p + 12 * n; w.M2 = p + 13 * n;
    w.RO = p + 14 * n; w.RD = p + 16 * n; w.HIT = p + 18 * n; w.PN = p + 20 * n;
    w.idxA = reinterpret_cast<int32_t *>(p + 21 * n);
    w.idxB = w.idxA + n;
    w.counter = reinterpret_cast<unsigned long long *>(p + 22 * n);
    w.cnt = w.counter + 8;
    return w;
}

// six consecutive sample columns starting at col0
__device__ __forceinline__ void sample_cols6(const IrisSampler &s, int64_t lane, int col0, float u[6]) {
    if (s.U != nullptr) {
        const float *p = s.U + lane * (int64_t)s.stride + col0;
#pragma unroll
        for (int k = 0; k < 6; ++k) u[k] = col0 + k < s.stride ? p[k] : 0.f;
        return;
    }
    const int b0 = col0 >> 2;
    float v[12];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const float4 q = sample4(s, lane, b0 + b);
        v[4 * b] = q.x; v[4 * b + 1] = q.y; v[4 * b + 2] = q.z; v[4 * b + 3] = q.w;
    }
    const int off = col0 & 3;
#pragma unroll
    for (int k = 0; k < 6; ++k) u[k] = v[off + k];
}
// warp-aggregated append of lane id `i` to a live-lane list (one atomic per warp; order inside the warp is kept)
__device__ __forceinline__ void wave_append(int32_t i, int32_t *idx, unsigned long long *cnt) {
    const unsigned act = __activemask(), lane = threadIdx.x & 31u;
    const int leader = __ffs(act) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(cnt, (unsigned long long)__popc(act));
    base = __shfl_sync(act, base, leader);
    idx[base + __popc(act & ((1u << lane) - 1u))] = i;
}
// thread j of a bounce kernel -> lane id (dense list, or the identity when no list is kept); false: nothing to do
__device__ __forceinline__ bool wave_lane(const int32_t *idx, const unsigned long long *cnt, int64_t n, int64_t &i, int64_t &c) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    c = idx ? (int64_t)*cnt : n;
    if (j >= c) return false;
    i = idx ? (int64_t)idx[j] : j;
    return true;
}
__device__ __forceinline__ float nan0(float x) { return x != x ? 0.f : x; }
__device__ __forceinline__ f3 nan0(f3 v) { return mk3(nan0(v.x), nan0(v.y), nan0(v.z)); }
__device__ __forceinline__ f3 ld4(const float4 *p, int64_t i) { const float4 v = p[i]; return mk3(v.x, v.y, v.z); }

// ---- init: camera rays (path_tracing :231-246)
__global__ void __launch_bounds__(IRIS_BLOCK) k_wave_init_camera(SceneView S, IrisShadeParams P, IrisSampler smp, const float *__restrict__ rays,
                                                                  int64_t n_pixels, int spp, WaveState W, int32_t *idx, unsigned long long *cnt) {
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
    f3 Le = mk3(0.f, 0.f, 0.f);
    if (h.prim >= 0) {
        const int32_t e = emitter_of(P, h.prim);
        if (e >= 0) Le = emitter_radiance(P, e); else code = -2;
    }
    W.S0[i] = make_float4(hp.x, hp.y, hp.z, __int_as_float(code));
    W.S1[i] = make_float4(hn.x, hn.y, hn.z, 0.f);
    W.S3[i] = make_float4(-wi.x, -wi.y, -wi.z, 0.f);
    W.S4[i] = make_float4(1.f, 1.f, 1.f, 0.f);
    W.S5[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    W.S6[i] = make_float4(Le.x, Le.y, Le.z, 0.f);
    W.S7[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    W.S8[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx && code == -2) wave_append((int32_t)i, idx, cnt);
}

// ---- init: explicit surface points.  PIXELS: per-pixel arrays repeated spp times (det_*), prim == -1 disables the pixel;
//      otherwise one lane per row (trace_indirect).
__global__ void __launch_bounds__(IRIS_BLOCK) k_wave_init_points(const float *__restrict__ position, const float *__restrict__ wo_or_wi, int negate_dir,
                                                                  const float *__restrict__ normal, const int32_t *__restrict__ prim, int64_t n_rows,
                                                                  int spp, WaveState W, int32_t *idx, unsigned long long *cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * spp) return;
    const int64_t r = i / spp;
    const f3 x = ld3(position, r), d = ld3(wo_or_wi, r), nr = ld3(normal, r);
    const bool ok = prim == nullptr || prim[r] != -1;
    const float sg = negate_dir ? -1.f : 1.f;
    W.S0[i] = make_float4(x.x, x.y, x.z, __int_as_float(ok ? -2 : -1));
    W.S1[i] = make_float4(nr.x, nr.y, nr.z, 0.f);
    W.S3[i] = make_float4(sg * d.x, sg * d.y, sg * d.z, 0.f);
    W.S4[i] = make_float4(1.f, 1.f, 1.f, 0.f);
    W.S5[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    W.S6[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    W.S7[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    W.S8[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx && ok) wave_append((int32_t)i, idx, cnt);
}

// ---- bounce_a: NEE + BSDF sample + closest hit.
// KIND 0: first bounce of path_tracing (:253-286: clamps 1e-6, un-clamped MIS, radiance into S6, weight into S7)
// KIND 1: one depth of trace_indirect (:435-471: clamps 1e-12, NaN -> 0, throughput in S4, radiance into S5)
// KIND 2: first sample of det_diff (:93-97)      KIND 3: first sample of det_spec (:174-178)   [no NEE, weights into S7/S8]
template <int KIND>
__global__ void __launch_bounds__(IRIS_BLOCK) k_wave_bounce_a(SceneView S, IrisShadeParams P, IrisSampler smp, int col0, float level, int64_t n, WaveState W) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 s0 = W.S0[i];
    if (__float_as_int(s0.w) != -2) { W.H0[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)); return; }
    const f3 x0 = mk3(s0.x, s0.y, s0.z), n0 = ld4(W.S1, i), wo = ld4(W.S3, i);
    float u[6];
    sample_cols6(smp, i, col0, u);
    f3 wi, bw = mk3(1.f, 1.f, 1.f);
    float bpdf = 0.f;
    if (KIND <= 1) {
        Mat mat;
        const float4 m2 = W.S2[i];
        mat.a = mk3(m2.x, m2.y, m2.z);
        mat.r = m2.w;
        mat.m = W.S1[i].w;
        const float clampv = KIND == 0 ? 1e-6f : 1e-12f;
        {   // emitter sampling; same any-hit formulation as k_bounce_single
            f3 wl;
            float pdf_e;
            int32_t e, face;
            sample_emitter(P, u[0], u[1], u[2], x0, wl, pdf_e, e, face);
            const f3 org = ray_origin(x0, wl);
            f3 v0, e1, e2;
            emitter_triangle(P, e, v0, e1, e2);
            float tl, bu, bv;
            if (tri_test(org, wl, __fdiv_rn(1.0f, xdot(wl, wl)), v0, e1, e2, tl, bu, bv) && !trace_occluded_shared(S, org, wl, tl, face)) {
                Hit h;
                h.t = tl; h.u = bu; h.v = bv; h.prim = face; h.slot = -1;
                f3 hp, hn;
                surface_from_triangle(h, wl, v0, e1, e2, hp, hn);
                const f3 dlt = hp - x0;
                const float G = fabsf(-(wl.x * hn.x) - (wl.y * hn.y) - (wl.z * hn.z)) / fmaxf(dot(dlt, dlt), clampv);
                f3 f;
                float pb;
                eval_brdf<false>(wl, wo, n0, mat, f, pb, nullptr);
                pb *= G;
                float w = (pdf_e > 0.f && !isinf(pb)) ? pdf_e * pdf_e / (pdf_e * pdf_e + pb * pb) : 0.f;
                if (isinf(pdf_e) || pb == 0.f) w = 1.f;
                const f3 c = f * (emitter_radiance(P, e) * (G / fmaxf(pdf_e, clampv) * w));
                if (KIND == 0) {
                    const f3 a = ld4(W.S6, i) + c;
                    W.S6[i] = make_float4(a.x, a.y, a.z, 0.f);
                } else {
                    const f3 a = ld4(W.S5, i) + nan0(ld4(W.S4, i) * c);
                    W.S5[i] = make_float4(a.x, a.y, a.z, 0.f);
                }
            }
        }
        sample_brdf<false>(u[3], u[4], u[5], wo, n0, mat, wi, bpdf, bw, nullptr);
        if (KIND == 0) {
            W.S7[i] = make_float4(bw.x, bw.y, bw.z, 0.f);
        } else {
            const f3 t = ld4(W.S4, i) * bw;
            W.S4[i] = make_float4(t.x, t.y, t.z, 0.f);
        }
    } else if (KIND == 2) {
        wi = diffuse_sampler(u[0], u[1], n0);
        W.S7[i] = make_float4(1.f, 1.f, 1.f, 0.f);
    } else {
        wi = specular_sampler(u[0], u[1], level, wo, n0);
        float w0, w1;
        specular_weights(wi, wo, n0, level, w0, w1);
        W.S7[i] = make_float4(w0, w0, w0, 0.f);
        W.S8[i] = make_float4(w1, w1, w1, 0.f);
    }
    const f3 org = ray_origin(x0, wi);
    const Hit h = trace_closest_shared(S, org, wi);
    f3 hp, hn;
    hit_surface(S, h, wi, hp, hn);
    W.H0[i] = make_float4(hp.x, hp.y, hp.z, __int_as_float(h.prim >= 0 ? -2 : -1));
    W.H1[i] = make_float4(hn.x, hn.y, hn.z, bpdf);
    W.H2[i] = make_float4(wi.x, wi.y, wi.z, __int_as_float(h.prim));
}

// ---- the same bounce through the ray queue: k_wave_gen writes the two rays of a lane and the emitter-sample contribution that is
// pending on the shadow ray, k_trace_queue (kernels.cuh) traces, k_wave_resolve applies the contribution and stores the next hit.
template <int KIND>
__global__ void __launch_bounds__(IRIS_BLOCK) k_wave_gen(IrisShadeParams P, IrisSampler smp, int col0, float level, int64_t n, WaveState W, const int32_t *__restrict__ idx,
                                                          const unsigned long long *__restrict__ cnt) {
    int64_t i, c;
    if (!wave_lane(idx, cnt, n, i, c)) return;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t off = KIND <= 1 ? c : 0;             // queue layout: shadow rays [0,c), closest-hit rays [c,2c); without NEE only [0,c)
    const float4 empty = make_float4(0.f, 0.f, 0.f, -1.f);
    const float4 s0 = W.S0[i];
    if (__float_as_int(s0.w) != -2) { if (KIND <= 1) W.RO[j] = empty; W.RO[off + j] = empty; return; }
    const f3 x0 = mk3(s0.x, s0.y, s0.z), n0 = ld4(W.S1, i), wo = ld4(W.S3, i);
    float u[6];
    sample_cols6(smp, i, col0, u);
    f3 wi, bw = mk3(1.f, 1.f, 1.f);
    float bpdf = 0.f;
    float4 ro_s = empty, rd_s = empty;
    f3 pend = mk3(0.f, 0.f, 0.f);
    if (KIND <= 1) {
        Mat mat;
        const float4 m2 = W.S2[i];
        mat.a = mk3(m2.x, m2.y, m2.z);
        mat.r = m2.w;
        mat.m = W.S1[i].w;
        const float clampv = KIND == 0 ? 1e-6f : 1e-12f;
        {
            f3 wl;
            float pdf_e;
            int32_t e, face;
            sample_emitter(P, u[0], u[1], u[2], x0, wl, pdf_e, e, face);
            const f3 org = ray_origin(x0, wl);
            f3 v0, e1, e2;
            emitter_triangle(P, e, v0, e1, e2);
            float tl, bu, bv;
            if (tri_test(org, wl, __fdiv_rn(1.0f, xdot(wl, wl)), v0, e1, e2, tl, bu, bv)) {
                ro_s = make_float4(org.x, org.y, org.z, tl);
                rd_s = make_float4(wl.x, wl.y, wl.z, __int_as_float(face));
                Hit h;
                h.t = tl; h.u = bu; h.v = bv; h.prim = face; h.slot = -1;
                f3 hp, hn;
                surface_from_triangle(h, wl, v0, e1, e2, hp, hn);
                const f3 dlt = hp - x0;
                const float G = fabsf(-(wl.x * hn.x) - (wl.y * hn.y) - (wl.z * hn.z)) / fmaxf(dot(dlt, dlt), clampv);
                f3 f;
                float pb;
                eval_brdf<false>(wl, wo, n0, mat, f, pb, nullptr);
                pb *= G;
                float w = (pdf_e > 0.f && !isinf(pb)) ? pdf_e * pdf_e / (pdf_e * pdf_e + pb * pb) : 0.f;
                if (isinf(pdf_e) || pb == 0.f) w = 1.f;
                const f3 c = f * (emitter_radiance(P, e) * (G / fmaxf(pdf_e, clampv) * w));
                pend = KIND == 0 ? c : nan0(ld4(W.S4, i) * c);
            }
        }
        sample_brdf<false>(u[3], u[4], u[5], wo, n0, mat, wi, bpdf, bw, nullptr);
        if (KIND == 0) {
            W.S7[i] = make_float4(bw.x, bw.y, bw.z, 0.f);
        } else {
            const f3 t = ld4(W.S4, i) * bw;
            W.S4[i] = make_float4(t.x, t.y, t.z, 0.f);
        }
    } else if (KIND == 2) {
        wi = diffuse_sampler(u[0], u[1], n0);
        W.S7[i] = make_float4(1.f, 1.f, 1.f, 0.f);
    } else {
        wi = specular_sampler(u[0], u[1], level, wo, n0);
        float w0, w1;
        specular_weights(wi, wo, n0, level, w0, w1);
        W.S7[i] = make_float4(w0, w0, w0, 0.f);
        W.S8[i] = make_float4(w1, w1, w1, 0.f);
    }
    float4 ro_b = ray4(ray_origin(x0, wi), __int_as_float(0x7f800000));
#ifndef IRIS_NO_RAY_SKIP
    // rays that cannot change the result are not cast (as in k_single_gen): a shadow ray whose pending contribution is exactly zero
    // (emitter below the horizon), and -- with NEE, where everything the path gathers from here on is multiplied by it -- a BSDF ray
    // whose weight is exactly zero: the resolve step sees a miss, which ends the path with nothing added
    if (KIND <= 1 && is_zero3(pend)) { ro_s = empty; rd_s = empty; }
    if (KIND <= 1 && is_zero3(bw)) ro_b = empty;
#endif
    if (KIND <= 1) {
        W.RO[j] = ro_s;
        W.RD[j] = rd_s;
    }
    W.RO[off + j] = ro_b;
    W.RD[off + j] = make_float4(wi.x, wi.y, wi.z, __int_as_float(-1));
    W.PN[j] = make_float4(pend.x, pend.y, pend.z, bpdf);
}

template <int KIND>
__global__ void __launch_bounds__(IRIS_BLOCK) k_wave_resolve(SceneView S, int64_t n, WaveState W, const int32_t *__restrict__ idx,
                                                              const unsigned long long *__restrict__ cnt) {
    int64_t i, c;
    if (!wave_lane(idx, cnt, n, i, c)) return;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t off = KIND <= 1 ? c : 0;
    if (__float_as_int(W.S0[i].w) != -2) { W.H0[j] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)); return; }
    const float4 pn = W.PN[j];
    if (KIND <= 1 && W.RO[j].w >= 0.f && __float_as_int(W.HIT[j].w) < 0) {      // shadow ray cast and unoccluded
        if (KIND == 0) {
            const f3 a = ld4(W.S6, i) + mk3(pn.x, pn.y, pn.z);
            W.S6[i] = make_float4(a.x, a.y, a.z, 0.f);
        } else {
            const f3 a = ld4(W.S5, i) + mk3(pn.x, pn.y, pn.z);
            W.S5[i] = make_float4(a.x, a.y, a.z, 0.f);
        }
    }
    const float4 dq = W.RD[off + j];
    const f3 wi = mk3(dq.x, dq.y, dq.z);
    f3 hp, hn;
    const Hit h = queue_hit_surface(S, W.HIT[off + j], wi, hp, hn);
    W.H0[j] = make_float4(hp.x, hp.y, hp.z, __int_as_float(h.prim >= 0 ? -2 : -1));
    W.H1[j] = make_float4(hn.x, hn.y, hn.z, pn.w);
    W.H2[j] = make_float4(wi.x, wi.y, wi.z, __int_as_float(h.prim));
}

// ---- bounce_b: radiance at the hit (emitter, or SLF when the material THERE is rougher than tau), MIS, state update.
// KIND as above; HAS_FIELD = false is the BaseBRDF case (roughness == 1 everywhere).
template <int KIND>
__global__ void __launch_bounds__(IRIS_BLOCK) k_wave_bounce_b(IrisShadeParams P, float tau, int has_field, int64_t n, WaveState W,
                                                               const int32_t *__restrict__ idx, const unsigned long long *__restrict__ cnt,
                                                               int32_t *__restrict__ idx_next, unsigned long long *__restrict__ cnt_next) {
    int64_t i, c;
    if (!wave_lane(idx, cnt, n, i, c)) return;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // hit records / material scratch are indexed by the list position
    const float4 s0 = W.S0[i];
    if (__float_as_int(s0.w) != -2) return;
    const f3 x0 = mk3(s0.x, s0.y, s0.z);
    const float4 h0 = W.H0[j], h1 = W.H1[j], h2 = W.H2[j];
    const f3 pn = mk3(h0.x, h0.y, h0.z), nn = mk3(h1.x, h1.y, h1.z), wi = mk3(h2.x, h2.y, h2.z);
    const int32_t prim = __float_as_int(h2.w);
    float bpdf = h1.w;
    const bool has_mat = has_field && prim >= 0;            // the field kernel only writes M1 / M2 for lanes whose ray hit something
    const float4 m2 = has_mat ? W.M2[j] : make_float4(0.f, 0.f, 0.f, 1.f);
    const float m_next = has_mat ? W.M1[j].w : 0.f;
    // model/emitter.py:180-221
    f3 Le = mk3(0.f, 0.f, 0.f);
    float epdf = 0.f;
    bool valid_next = false;
    if (prim >= 0) {
        const int32_t e = emitter_of(P, prim);
        if (e >= 0) {
            Le = emitter_radiance(P, e);
            epdf = emitter_pdf_area(P, e);
        } else {
            valid_next = true;
            if (m2.w > tau) {
                Le = slf_lookup(P, pn);
                if (Le.x + Le.y + Le.z > 0.f) valid_next = false;
            }
        }
    }
    if (KIND <= 1) {
        const float clampv = KIND == 0 ? 1e-6f : 1e-12f;
        float G = 1.f;
        if (valid_next) {
            const f3 dlt = x0 - pn;
            G = fabsf(-(nn.x * wi.x) - (nn.y * wi.y) - (nn.z * wi.z)) / fmaxf(dot(dlt, dlt), clampv);
        }
        bpdf *= G;
        float w = (bpdf > 0.f && !isinf(epdf)) ? bpdf * bpdf / (epdf * epdf + bpdf * bpdf) : 0.f;
        if (isinf(bpdf) || epdf == 0.f) w = 1.f;
        if (KIND == 0) {
            const f3 a = ld4(W.S6, i) + ld4(W.S7, i) * Le * w;
            W.S6[i] = make_float4(a.x, a.y, a.z, 0.f);
        } else {
            const f3 a = ld4(W.S5, i) + nan0(ld4(W.S4, i) * Le * w);
            W.S5[i] = make_float4(a.x, a.y, a.z, 0.f);
        }
    } else {
        W.S6[i] = make_float4(Le.x, Le.y, Le.z, 0.f);     // det_*: weighted by S7 / S8 at the end, no MIS
    }
    W.S0[i] = make_float4(pn.x, pn.y, pn.z, __int_as_float(valid_next ? -2 : -1));
    W.S1[i] = make_float4(nn.x, nn.y, nn.z, m_next);
    W.S2[i] = m2;
    W.S3[i] = make_float4(-wi.x, -wi.y, -wi.z, 0.f);
    if (idx_next && valid_next) wave_append((int32_t)i, idx_next, cnt_next);
}

// ---- finish.  MODE 0: per-pixel mean of S6 + S7*S5 (path_tracing).  MODE 1: det: out0 = mean S7*(S6+S5), out1 = mean S8*(S6+S5).
//               MODE 2: per-lane S5 (trace_indirect).
template <int MODE>
__global__ void __launch_bounds__(IRIS_BLOCK) k_wave_finish(int64_t n_rows, int spp, WaveState W, float *out0, float *out1) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = n_rows * spp;
    const bool in_range = i < n;
    const int64_t pix = in_range ? i / spp : 0;
    f3 a = mk3(0.f, 0.f, 0.f), b = a;
    if (in_range) {
        const f3 s5 = ld4(W.S5, i), s6 = ld4(W.S6, i), s7 = ld4(W.S7, i);
        if (MODE == 0) a = s6 + s7 * s5;
        if (MODE == 1) { a = s7 * (s6 + s5); b = ld4(W.S8, i) * (s6 + s5); }
        if (MODE == 2) { out0[3 * i] = s5.x; out0[3 * i + 1] = s5.y; out0[3 * i + 2] = s5.z; }
    }
    if (MODE == 2) return;
    const float inv = 1.f / (float)spp;
    pixel_accumulate(out0, pix, in_range, a, inv);
    if (MODE == 1 && out1) pixel_accumulate(out1, pix, in_range, b, inv);
}
