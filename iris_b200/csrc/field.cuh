// field.cuh -- the BRDF field NGPBRDF.forward (reference model/brdf.py:243-260): multiresolution hash-grid encoding
// (32 levels x 2 fp16 features) + 64-64-16 ReLU MLP + sigmoid, restating tiny-cuda-nn's HashGrid/FullyFusedMLP exactly
// as oracle/field.py defines it (rounding points included).
//
// One lane encodes one sample (256 half2 gathers, fp32 accumulate in the oracle's corner order); the 32 samples of a
// warp then go through the MLP as a 32x64 tile: activations staged in shared memory as fp16, weights resident in
// shared memory, fp32 accumulators in registers.
#pragma once
#include <cuda_fp16.h>

#include "shading.cuh"

#define FIELD_LEVELS 32
#define FIELD_WIDTH 64
#define FIELD_OUT 16
#define FIELD_LD 72   // padded leading dimension (halfs): conflict-free fragment loads

struct FieldLevel {
    float scale;
    uint32_t res;
    uint32_t size;     // entries in the level (each entry = 2 features)
    uint32_t offset;   // first entry
    uint32_t dense;    // 1: x + y*res + z*res^2, 0: spatial hash
};
__constant__ FieldLevel c_levels[FIELD_LEVELS];

// tcnn grid layout: scale_l = 16*1.3^l - 1, res = ceil(scale)+1, size = min(next_multiple(res^3,8), 2^19)
inline int64_t field_level_table(FieldLevel out[FIELD_LEVELS]) {
    uint64_t off = 0;
    for (int l = 0; l < FIELD_LEVELS; ++l) {
        // float64 evaluation rounded once: portable across libms (oracle/field.py:level_table does the same)
        const float scale = (float)(16.0 * pow((double)1.3f, (double)l) - 1.0);
        const uint64_t res = (uint64_t)ceilf(scale) + 1;
        uint64_t dense = res * res * res;
        uint64_t size = dense > 0x7FFFFFFFull ? 0x7FFFFFFFull : dense;
        size = (size + 7) / 8 * 8;
        if (size > (1ull << 19)) size = 1ull << 19;
        out[l].scale = scale;
        out[l].res = (uint32_t)res;
        out[l].size = (uint32_t)size;
        out[l].offset = (uint32_t)off;
        out[l].dense = dense <= size ? 1u : 0u;
        off += size;
    }
    return (int64_t)off;
}

__device__ __forceinline__ float field_coord(float p, float vmin, float range) {
    // (p - vmin) / (vmax - vmin) * 2 - 1, one rounding per op (model/brdf.py:253-255)
    return xsub(xmul(__fdiv_rn(xsub(p, vmin), range), 2.0f), 1.0f);
}

// The first FIELD_DENSE_LEVELS levels are dense (x + y*res + z*res^2, modulo the level size because the reference feeds
// negative coordinates that wrap in uint32); their res / size are compile-time constants of the fixed tcnn configuration, so the
// modulo is a multiply-shift and the loop bodies are branch-free (the gathers of several levels can then be batched).
#define FIELD_DENSE_LEVELS 7
__device__ __host__ constexpr uint32_t field_dense_res(int l) { return l == 0 ? 16u : l == 1 ? 21u : l == 2 ? 28u : l == 3 ? 36u : l == 4 ? 46u : l == 5 ? 60u : 78u; }
__device__ __host__ constexpr uint32_t field_dense_size(int l) { return (field_dense_res(l) * field_dense_res(l) * field_dense_res(l) + 7u) / 8u * 8u; }

template <int DENSE_L>   // DENSE_L >= 0: dense level with compile-time geometry; -1: hashed level (size == 2^19)
__device__ __forceinline__ void field_level_indices(const FieldLevel &L, f3 x, uint32_t idx[8], float w[8]) {
    const float px = __fmaf_rn(L.scale, x.x, 0.5f), py = __fmaf_rn(L.scale, x.y, 0.5f), pz = __fmaf_rn(L.scale, x.z, 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const float wx1 = xsub(px, fx), wy1 = xsub(py, fy), wz1 = xsub(pz, fz);
    const float wx0 = xsub(1.0f, wx1), wy0 = xsub(1.0f, wy1), wz0 = xsub(1.0f, wz1);
    const uint32_t cx = (uint32_t)__float2int_rz(fx), cy = (uint32_t)__float2int_rz(fy), cz = (uint32_t)__float2int_rz(fz);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint32_t ix = cx + (c & 1), iy = cy + ((c >> 1) & 1), iz = cz + ((c >> 2) & 1);
        if (DENSE_L >= 0) {
            constexpr uint32_t res = field_dense_res(DENSE_L >= 0 ? DENSE_L : 0), size = field_dense_size(DENSE_L >= 0 ? DENSE_L : 0);
            idx[c] = (ix + iy * res + iz * (res * res)) % size;
        } else {
            idx[c] = (ix ^ (iy * 2654435761u) ^ (iz * 805459861u)) & ((1u << 19) - 1u);
        }
        w[c] = xmul(xmul((c & 1) ? wx1 : wx0, (c & 2) ? wy1 : wy0), (c & 4) ? wz1 : wz0);
    }
}

// corner weights only (second pass of the grouped encoder: recomputed instead of kept live across the gathers)
__device__ __forceinline__ void field_level_weights(const FieldLevel &L, f3 x, float w[8]) {
    const float px = __fmaf_rn(L.scale, x.x, 0.5f), py = __fmaf_rn(L.scale, x.y, 0.5f), pz = __fmaf_rn(L.scale, x.z, 0.5f);
    const float wx1 = xsub(px, floorf(px)), wy1 = xsub(py, floorf(py)), wz1 = xsub(pz, floorf(pz));
    const float wx0 = xsub(1.0f, wx1), wy0 = xsub(1.0f, wy1), wz0 = xsub(1.0f, wz1);
#pragma unroll
    for (int c = 0; c < 8; ++c) w[c] = xmul(xmul((c & 1) ? wx1 : wx0, (c & 2) ? wy1 : wy0), (c & 4) ? wz1 : wz0);
}

template <int DENSE_L>
__device__ __forceinline__ void field_level_gather(const __half2 *__restrict__ grid, f3 x, int l, __half2 v[8], bool pair = true) {
    const FieldLevel L = c_levels[l];
    uint32_t idx[8];
    float w[8];
    field_level_indices<DENSE_L>(L, x, idx, w);
    // volatile asm keeps the gathers of a whole group in program order ahead of their consumers (ptxas otherwise sinks every load
    // next to its use to save registers, leaving one or two requests in flight per lane)
#ifndef FIELD_GATHER_NO_V2
    if (DENSE_L < 0 && pair && (idx[0] ^ idx[1]) == 1u) {
        // hashed level, even cell x: corners c and c+1 (x, x+1 at the same y, z) hash to idx and idx ^ 1, the two halves of one aligned
        // 8-byte slot -> four 8-byte loads instead of eight 4-byte ones (same sectors, half the L1 requests).  `pair` is a launch-uniform
        // switch: +20 % on incoherent positions (secondary hits, training pixels), -1.5 % on the spp-coherent primary hits of
        // path_tracing_single, whose lanes mostly share sectors anyway -- that caller turns it off
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
            uint32_t r0, r1;
            asm volatile("ld.global.nc.v2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "l"(grid + L.offset + (idx[c] & ~1u)));
            const bool swap = idx[c] & 1u;
            const uint32_t a = swap ? r1 : r0, b = swap ? r0 : r1;
            v[c] = *reinterpret_cast<const __half2 *>(&a);
            v[c + 1] = *reinterpret_cast<const __half2 *>(&b);
        }
        return;
    }
#endif
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint32_t r;
        asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(r) : "l"(grid + L.offset + idx[c]));
        v[c] = *reinterpret_cast<__half2 *>(&r);
    }
}

// Scheduling fence for a group of gathers.  ptxas sinks every load next to its first use to save registers, which leaves one or
// two requests in flight per lane in a latency-bound loop.  Making every gathered word depend on ALL words of the group (xor with
// `(w0^w1^...) & zero`, where `zero` is a run-time 0 the compiler cannot fold) forces the whole group to be issued first.
template <int N>
__device__ __forceinline__ void field_pin_group(__half2 (*v)[8], uint32_t zero) {
    uint32_t tok = 0;
#pragma unroll
    for (int k = 0; k < N; ++k)
#pragma unroll
        for (int c = 0; c < 8; ++c) tok ^= *reinterpret_cast<uint32_t *>(&v[k][c]);
    tok &= zero;
#pragma unroll
    for (int k = 0; k < N; ++k)
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint32_t *>(&v[k][c]) ^= tok;
}

template <class Put>
__device__ __forceinline__ void field_level_reduce(f3 x, int l, const __half2 v[8], Put &put) {
    float w[8];
    field_level_weights(c_levels[l], x, w);
    float f0 = 0.f, f1 = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float2 q = __half22float2(v[c]);
        f0 = xadd(f0, xmul(w[c], q.x));
        f1 = xadd(f1, xmul(w[c], q.y));
    }
    put(l, __floats2half2_rn(f0, f1));
}

// Encodes one sample: hands the 32 fp16 feature pairs (level l -> features 2l, 2l+1) to `put(l, pair)`.
// The gathers of a GROUP of levels (4 dense, or 5 hashed: 32-40 independent 4-byte loads) are issued before anything is consumed,
// so each lane keeps dozens of L2 requests in flight -- the encoder is latency-bound, not bandwidth-bound.
template <class Put>
__device__ __forceinline__ void field_encode_to(const __half2 *__restrict__ grid, f3 x, Put put, bool pair = true) {
    const uint32_t zero = c_levels[0].dense - 1u;      // 0 at run time (level 0 is dense), opaque to the compiler
    {
        __half2 v[4][8];
        field_level_gather<0>(grid, x, 0, v[0]);
        field_level_gather<1>(grid, x, 1, v[1]);
        field_level_gather<2>(grid, x, 2, v[2]);
        field_level_gather<3>(grid, x, 3, v[3]);
        field_pin_group<4>(v, zero);
#pragma unroll
        for (int k = 0; k < 4; ++k) field_level_reduce(x, k, v[k], put);
    }
    {
        __half2 v[3][8];
        field_level_gather<4>(grid, x, 4, v[0]);
        field_level_gather<5>(grid, x, 5, v[1]);
        field_level_gather<6>(grid, x, 6, v[2]);
        field_pin_group<3>(v, zero);
#pragma unroll
        for (int k = 0; k < 3; ++k) field_level_reduce(x, 4 + k, v[k], put);
    }
#pragma unroll 1
    for (int l0 = FIELD_DENSE_LEVELS; l0 < FIELD_LEVELS; l0 += 5) {
        __half2 v[5][8];
#pragma unroll
        for (int k = 0; k < 5; ++k) field_level_gather<-1>(grid, x, l0 + k, v[k], pair);
        field_pin_group<5>(v, zero);
#pragma unroll
        for (int k = 0; k < 5; ++k) field_level_reduce(x, l0 + k, v[k], put);
    }
}
// ... to a contiguous row dst[0..63]
__device__ __forceinline__ void field_encode(const __half2 *__restrict__ grid, f3 x, __half *dst, bool pair = true) {
    field_encode_to(grid, x, [dst](int l, __half2 v) { *reinterpret_cast<__half2 *>(dst + 2 * l) = v; }, pair);
}

__device__ __forceinline__ void mma16816(float c[4], const uint32_t a[4], const uint32_t b[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// acc[mt][nt][4] = X(32x64, fp16, ld 72) * W^T with W (NT*8 x 64, fp16, ld 72, row-major [out][in])
template <int NT>
__device__ __forceinline__ void warp_gemm(const __half *X, const __half *W, float acc[2][NT][4]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[mt][nt][k] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const int k0 = 16 * kk + 2 * t;
        uint32_t a[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const __half *r0 = X + (16 * mt + g) * FIELD_LD + k0, *r1 = r0 + 8 * FIELD_LD;
            a[mt][0] = *reinterpret_cast<const uint32_t *>(r0);
            a[mt][1] = *reinterpret_cast<const uint32_t *>(r1);
            a[mt][2] = *reinterpret_cast<const uint32_t *>(r0 + 8);
            a[mt][3] = *reinterpret_cast<const uint32_t *>(r1 + 8);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const __half *w = W + (8 * nt + g) * FIELD_LD + k0;
            uint32_t b[2];
            b[0] = *reinterpret_cast<const uint32_t *>(w);
            b[1] = *reinterpret_cast<const uint32_t *>(w + 8);
            mma16816(acc[0][nt], a[0], b);
            mma16816(acc[1][nt], a[1], b);
        }
    }
}

// relu + fp16 round + store back as the next layer's input
__device__ __forceinline__ void warp_store_relu(__half *X, float acc[2][8][4]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            __half *r0 = X + (16 * mt + g) * FIELD_LD + 8 * nt + 2 * t;
            *reinterpret_cast<__half2 *>(r0) = __floats2half2_rn(fmaxf(acc[mt][nt][0], 0.f), fmaxf(acc[mt][nt][1], 0.f));
            *reinterpret_cast<__half2 *>(r0 + 8 * FIELD_LD) = __floats2half2_rn(fmaxf(acc[mt][nt][2], 0.f), fmaxf(acc[mt][nt][3], 0.f));
        }
}

__device__ __forceinline__ void field_load_weights(const __half *__restrict__ mlp, __half *Wsm) {
    // [W1 64x64 | W2 64x64 | W3 16x64] row-major -> padded rows of FIELD_LD
    for (int i = threadIdx.x; i < (64 + 64 + 16) * 32; i += blockDim.x) {
        const int row = i >> 5, c2 = i & 31;
        *reinterpret_cast<__half2 *>(Wsm + row * FIELD_LD + 2 * c2) = *reinterpret_cast<const __half2 *>(mlp + row * 64 + 2 * c2);
    }
}

// sigmoid of the fp16 network output, rounded to fp16 again (the reference applies .sigmoid() to tcnn's half tensor)
__device__ __forceinline__ float sigmoid16(float y_acc) {
    const float y = __half2float(__float2half_rn(y_acc));
    const float s = 1.0f / (1.0f + expf(-y));
    return __half2float(__float2half_rn(s));
}

#define FIELD_SMEM_BYTES ((64 + 64 + 16) * FIELD_LD * 2 + (IRIS_BLOCK / 32) * 32 * FIELD_LD * 2)

// WS = true : positions and the continue flag come from the estimator workspace w0, results go to w2 / w1.w
// WS = false: NGPBRDF.forward on a (n,3) position array -> mat (n,5)
__device__ __forceinline__ void warp_tile_to_global(const __half *Xs, __half *dst, int64_t row0, int64_t n);
template <bool WS>
__global__ void __launch_bounds__(IRIS_BLOCK) k_field_forward(IrisShadeParams P, int64_t n, const float *__restrict__ position, float *__restrict__ mat,
                                                               const float4 *__restrict__ w0, float4 *__restrict__ w1, float4 *__restrict__ w2,
                                                               __half *__restrict__ x_save, int pair, const unsigned long long *__restrict__ n_dev) {
    if (n_dev != nullptr) n = min(n, (int64_t)*n_dev);         // rows counted on the device (live lanes of a wavefront bounce)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half *Wsm = reinterpret_cast<__half *>(smem_raw);
    __half *Xs = Wsm + (64 + 64 + 16) * FIELD_LD + (threadIdx.x >> 5) * 32 * FIELD_LD;
    field_load_weights(reinterpret_cast<const __half *>(P.mlp_f16), Wsm);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t n_tiles = (n + IRIS_BLOCK - 1) / IRIS_BLOCK;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = tile * IRIS_BLOCK + threadIdx.x;
        bool active = i < n;
        f3 p = mk3(0.f, 0.f, 0.f);
        if (active) {
            if (WS) {
                const float4 a = w0[i];
                active = __float_as_int(a.w) == -2;
                p = mk3(a.x, a.y, a.z);
            } else {
                p = ld3(position, i);
            }
        }
        __half *row = Xs + lane * FIELD_LD;
        if (active) {
            const f3 x = mk3(field_coord(p.x, P.field_vmin, P.field_range), field_coord(p.y, P.field_vmin, P.field_range),
                             field_coord(p.z, P.field_vmin, P.field_range));
            field_encode(reinterpret_cast<const __half2 *>(P.grid_f16), x, row, pair != 0);
        } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) *reinterpret_cast<__half2 *>(row + 2 * k) = __floats2half2_rn(0.f, 0.f);
        }
        __syncwarp();
        if (x_save) warp_tile_to_global(Xs, x_save, tile * IRIS_BLOCK + (threadIdx.x & ~31), n);   // encoded inputs, kept for the adjoint
        if (__any_sync(0xffffffffu, active)) {
            float acc[2][8][4];
            warp_gemm<8>(Xs, Wsm, acc);
            __syncwarp();
            warp_store_relu(Xs, acc);
            __syncwarp();
            warp_gemm<8>(Xs, Wsm + 64 * FIELD_LD, acc);
            __syncwarp();
            warp_store_relu(Xs, acc);
            __syncwarp();
            float out[2][2][4];
            warp_gemm<2>(Xs, Wsm + 128 * FIELD_LD, out);
            __syncwarp();
            // stage the 32x16 fp32 outputs so that every lane can read its own sample's row
            float *Ys = reinterpret_cast<float *>(Xs);
            const int g = lane >> 2, t = lane & 3;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    float *r0 = Ys + (16 * mt + g) * 17 + 8 * nt + 2 * t;
                    r0[0] = out[mt][nt][0]; r0[1] = out[mt][nt][1];
                    r0[8 * 17] = out[mt][nt][2]; r0[8 * 17 + 1] = out[mt][nt][3];
                }
            __syncwarp();
            if (active) {
                const float *y = Ys + lane * 17;
                const float a0 = sigmoid16(y[0]), a1 = sigmoid16(y[1]), a2 = sigmoid16(y[2]);
                const float r = sigmoid16(y[3]) * 0.98f + 0.02f, m = sigmoid16(y[4]);
                if (WS) {
                    w2[i] = make_float4(a0, a1, a2, r);
                    float4 b = w1[i];
                    b.w = m;
                    w1[i] = b;
                } else {
                    float *o = mat + 5 * i;
                    o[0] = a0; o[1] = a1; o[2] = a2; o[3] = r; o[4] = m;
                }
            }
            __syncwarp();
        }
    }
}

// =====================================================================================================================
// Adjoint of the field: d_mat (n,5) -> d_params ([W1|W2|W3 | grid], fp32, tcnn's flat layout).
//
// Kernel A (k_field_backward_dgrad), one warp per 32 samples: re-encode, forward with ReLU masks kept as register bit masks,
// dy = d_mat * d(sigmoid) normalised per sample to [-1,1] by a power of two s_i (fp16 range), dgrad through W3, W2, W1 on
// tensor cores (transposed weights resident in shared memory), scatter of s_i * w_corner * dx into the grid gradient with
// vectorised reductions (red.global.add.v2.f32).  Activations X,h1,h2 and the normalised dh2^,dh1^,dy^ are streamed to HBM
// with coalesced 16-byte stores.
// Kernel B (k_field_backward_wgrad): dW = sum_i s_i * dh^_i (x) h_i as a split-K TF32 GEMM over the samples, accumulators in
// registers for the whole chunk, then ONE atomic per weight per CTA.
// Gradient semantics = oracle/field.py: straight-through across every fp16 rounding, sigmoid' from the fp32 sigmoid.
// =====================================================================================================================
#define FIELD_LD3 24   // leading dimension (halfs) of the 16-wide operands (dy^, W3^T)
#define FIELD_BWD_WSM_HALFS ((64 + 64 + 16) * FIELD_LD + 2 * 64 * FIELD_LD + 64 * FIELD_LD3)
#ifndef FIELD_BWD_BLOCK
#define FIELD_BWD_BLOCK 256
#endif
#define FIELD_BWD_SMEM_BYTES (FIELD_BWD_WSM_HALFS * 2 + (FIELD_BWD_BLOCK / 32) * 32 * FIELD_LD * 2)
#define FIELD_ACT_BYTES_PER_SAMPLE (6 * 128 + 32 + 4)   // X h1 h2 dh2^ dh1^ dx^ (64 halfs each) | dy^ (16 halfs) | s (float)

struct FieldAct {   // SoA activation streams of one chunk of n samples
    __half *X, *h1, *h2, *dh2, *dh1, *dx, *dy;
    float *s;
};
__host__ __device__ inline FieldAct field_act_carve(void *base, int64_t n) {
    FieldAct a;
    __half *p = reinterpret_cast<__half *>(base);
    a.X = p; a.h1 = p + 64 * n; a.h2 = p + 128 * n; a.dh2 = p + 192 * n; a.dh1 = p + 256 * n; a.dx = p + 320 * n; a.dy = p + 384 * n;
    a.s = reinterpret_cast<float *>(p + 400 * n);
    return a;
}

// coalesced copy of a warp's 32 x 64 fp16 tile (shared, ld FIELD_LD) to rows [row0, row0+32) of a [n][64] global array
__device__ __forceinline__ void warp_tile_to_global(const __half *Xs, __half *dst, int64_t row0, int64_t n) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int v = it * 32 + lane;          // 256 vectors of 8 halfs
        const int r = v >> 3, c = (v & 7) * 8;
        if (row0 + r < n) *reinterpret_cast<uint4 *>(dst + (row0 + r) * 64 + c) = *reinterpret_cast<const uint4 *>(Xs + r * FIELD_LD + c);
    }
}

// the inverse: rows [row0, row0+32) of a [n][64] global array into the warp's shared tile (rows past n become zeros)
__device__ __forceinline__ void warp_tile_from_global(__half *Xs, const __half *src, int64_t row0, int64_t n) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int v = it * 32 + lane;
        const int r = v >> 3, c = (v & 7) * 8;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (row0 + r < n) q = __ldg(reinterpret_cast<const uint4 *>(src + (row0 + r) * 64 + c));
        *reinterpret_cast<uint4 *>(Xs + r * FIELD_LD + c) = q;
    }
}

__device__ __forceinline__ uint32_t relu_mask_bits(const float acc[8][4]) {
    uint32_t m = 0;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int k = 0; k < 4; ++k) m |= (acc[nt][k] > 0.f ? 1u : 0u) << (nt * 4 + k);
    return m;
}

// out[mt][nt][4] = A(32 x KT*16, fp16, ld lda) * B with B given TRANSPOSED in shared memory: BT[n][k] (ld ldb)
template <int NT, int KT>
__device__ __forceinline__ void warp_gemm_t(const __half *A, int lda, const __half *BT, int ldb, float acc[2][NT][4]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[mt][nt][k] = 0.f;
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
        const int k0 = 16 * kk + 2 * t;
        uint32_t a[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const __half *r0 = A + (16 * mt + g) * lda + k0, *r1 = r0 + 8 * lda;
            a[mt][0] = *reinterpret_cast<const uint32_t *>(r0);
            a[mt][1] = *reinterpret_cast<const uint32_t *>(r1);
            a[mt][2] = *reinterpret_cast<const uint32_t *>(r0 + 8);
            a[mt][3] = *reinterpret_cast<const uint32_t *>(r1 + 8);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const __half *w = BT + (8 * nt + g) * ldb + k0;
            uint32_t b[2];
            b[0] = *reinterpret_cast<const uint32_t *>(w);
            b[1] = *reinterpret_cast<const uint32_t *>(w + 8);
            mma16816(acc[0][nt], a[0], b);
            mma16816(acc[1][nt], a[1], b);
        }
    }
}

__device__ __forceinline__ void red_add_v2(float *addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
// 16-byte reduction (sm_90+): two adjacent grid entries at once; addr must be 16-byte aligned
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// WS = true: positions/flags from the estimator record word r5 (x0.xyz, code), d_mat from the workspace; WS = false: plain arrays
template <bool WS>
__global__ void __launch_bounds__(FIELD_BWD_BLOCK, 512 / FIELD_BWD_BLOCK) k_field_backward_dgrad(IrisShadeParams P, int64_t n, const float *__restrict__ position,
                                                                      const float4 *__restrict__ r5, const float *__restrict__ d_mat,
                                                                      FieldAct act, int x_saved) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half *Wsm = reinterpret_cast<__half *>(smem_raw);              // W1 | W2 | W3 (forward, [out][in])
    __half *W1T = Wsm + (64 + 64 + 16) * FIELD_LD;                     // [in][out] copies for dgrad
    __half *W2T = W1T + 64 * FIELD_LD;
    __half *W3T = W2T + 64 * FIELD_LD;                                 // [in=64][out=16], ld FIELD_LD3
    __half *Xs = Wsm + FIELD_BWD_WSM_HALFS + (threadIdx.x >> 5) * 32 * FIELD_LD;
    const __half *mlp = reinterpret_cast<const __half *>(P.mlp_f16);
    field_load_weights(mlp, Wsm);
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
        const int o = i >> 6, in = i & 63;
        W1T[in * FIELD_LD + o] = mlp[i];
        W2T[in * FIELD_LD + o] = mlp[4096 + i];
    }
    for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) {
        const int o = i >> 6, in = i & 63;
        W3T[in * FIELD_LD3 + o] = mlp[8192 + i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const __half2 *grid = reinterpret_cast<const __half2 *>(P.grid_f16);
    const int64_t n_tiles = (n + FIELD_BWD_BLOCK - 1) / FIELD_BWD_BLOCK;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * FIELD_BWD_BLOCK + (threadIdx.x & ~31);
        const int64_t i = row0 + lane;
        bool active = i < n;
        f3 p = mk3(0.f, 0.f, 0.f);
        float dm[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (active) {
            if (WS) {
                const float4 a = r5[i];
                active = __float_as_int(a.w) == -2;
                p = mk3(a.x, a.y, a.z);
            } else {
                p = ld3(position, i);
            }
#pragma unroll
            for (int k = 0; k < 5; ++k) dm[k] = d_mat[5 * i + k];
            active = active && (dm[0] != 0.f || dm[1] != 0.f || dm[2] != 0.f || dm[3] != 0.f || dm[4] != 0.f);
        }
        if (!__any_sync(0xffffffffu, active)) {
            // nothing to do for these 32 samples: the wgrad kernel must still see zeros
            if (i < n) {
                act.s[i] = 0.f;
                *reinterpret_cast<uint4 *>(act.dy + 16 * i) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4 *>(act.dy + 16 * i + 8) = make_uint4(0, 0, 0, 0);
            }
            uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int v = it * 32 + lane, r = v >> 3, c = (v & 7) * 8;
                if (row0 + r < n) {
                    if (!x_saved) *reinterpret_cast<uint4 *>(act.X + (row0 + r) * 64 + c) = z;
                    *reinterpret_cast<uint4 *>(act.h1 + (row0 + r) * 64 + c) = z;
                    *reinterpret_cast<uint4 *>(act.h2 + (row0 + r) * 64 + c) = z;
                    *reinterpret_cast<uint4 *>(act.dh2 + (row0 + r) * 64 + c) = z;
                    *reinterpret_cast<uint4 *>(act.dh1 + (row0 + r) * 64 + c) = z;
                    *reinterpret_cast<uint4 *>(act.dx + (row0 + r) * 64 + c) = z;
                }
            }
            continue;
        }
        if (x_saved) {
            // act.X already holds the encoded inputs (written by the forward field kernel into the adjoint record)
            warp_tile_from_global(Xs, act.X, row0, n);
            __syncwarp();
        } else {
            const f3 x = mk3(field_coord(p.x, P.field_vmin, P.field_range), field_coord(p.y, P.field_vmin, P.field_range),
                             field_coord(p.z, P.field_vmin, P.field_range));
            __half *row = Xs + lane * FIELD_LD;
            if (active) {
                field_encode(grid, x, row);
            } else {
#pragma unroll
                for (int k = 0; k < 32; ++k) *reinterpret_cast<__half2 *>(row + 2 * k) = __floats2half2_rn(0.f, 0.f);
            }
            __syncwarp();
            warp_tile_to_global(Xs, act.X, row0, n);
        }
        // ---- forward, keeping the ReLU masks
        float acc[2][8][4];
        uint32_t m1[2], m2[2];
        warp_gemm<8>(Xs, Wsm, acc);
        m1[0] = relu_mask_bits(acc[0]); m1[1] = relu_mask_bits(acc[1]);
        __syncwarp();
        warp_store_relu(Xs, acc);
        __syncwarp();
        warp_tile_to_global(Xs, act.h1, row0, n);
        warp_gemm<8>(Xs, Wsm + 64 * FIELD_LD, acc);
        m2[0] = relu_mask_bits(acc[0]); m2[1] = relu_mask_bits(acc[1]);
        __syncwarp();
        warp_store_relu(Xs, acc);
        __syncwarp();
        warp_tile_to_global(Xs, act.h2, row0, n);
        float out[2][2][4];
        warp_gemm<2>(Xs, Wsm + 128 * FIELD_LD, out);
        __syncwarp();
        float *Ys = reinterpret_cast<float *>(Xs);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                float *r0 = Ys + (16 * mt + g) * 17 + 8 * nt + 2 * t;
                r0[0] = out[mt][nt][0]; r0[1] = out[mt][nt][1];
                r0[8 * 17] = out[mt][nt][2]; r0[8 * 17 + 1] = out[mt][nt][3];
            }
        __syncwarp();
        // ---- dy = d_mat * d(mat)/dy, normalised per sample by a power of two
        float dy[5];
        float sc = 0.f;
        {
            const float *y = Ys + lane * 17;
            float mx = 0.f;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const float yk = __half2float(__float2half_rn(y[k]));
                const float s = 1.0f / (1.0f + expf(-yk));
                dy[k] = active ? dm[k] * (k == 3 ? 0.98f : 1.0f) * s * (1.0f - s) : 0.f;
                mx = fmaxf(mx, fabsf(dy[k]));
            }
            if (mx > 0.f && mx < __int_as_float(0x7f800000)) {
                int e;
                frexpf(mx, &e);               // mx = f * 2^e, f in [0.5,1)
                sc = ldexpf(1.0f, e);         // mx / sc in [0.5,1)
            }
        }
        __syncwarp();
        __half *Dy = Xs + 32 * 36;            // after the Ys region (32*17 floats = 1088 halfs) -> halfs [1152, 1152+768)
        {
            const float inv = sc > 0.f ? 1.0f / sc : 0.f;
            __half *d = Dy + lane * FIELD_LD3;
#pragma unroll
            for (int k = 0; k < 16; ++k) d[k] = __float2half_rn(k < 5 ? dy[k] * inv : 0.f);
            if (i < n) {
                act.s[i] = sc;
                *reinterpret_cast<uint4 *>(act.dy + 16 * i) = *reinterpret_cast<const uint4 *>(d);
                *reinterpret_cast<uint4 *>(act.dy + 16 * i + 8) = *reinterpret_cast<const uint4 *>(d + 8);
            }
        }
        __syncwarp();
        // ---- dgrad: dh2^ = dy^ W3 (.) mask2 ; dh1^ = dh2^ W2 (.) mask1 ; dx^ = dh1^ W1
        warp_gemm_t<8, 1>(Dy, FIELD_LD3, W3T, FIELD_LD3, acc);
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (!((m2[mt] >> (nt * 4 + k)) & 1u)) acc[mt][nt][k] = 0.f;
        {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    __half *r0 = Xs + (16 * mt + g) * FIELD_LD + 8 * nt + 2 * t;
                    *reinterpret_cast<__half2 *>(r0) = __floats2half2_rn(acc[mt][nt][0], acc[mt][nt][1]);
                    *reinterpret_cast<__half2 *>(r0 + 8 * FIELD_LD) = __floats2half2_rn(acc[mt][nt][2], acc[mt][nt][3]);
                }
        }
        __syncwarp();
        warp_tile_to_global(Xs, act.dh2, row0, n);
        warp_gemm_t<8, 4>(Xs, FIELD_LD, W2T, FIELD_LD, acc);
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (!((m1[mt] >> (nt * 4 + k)) & 1u)) acc[mt][nt][k] = 0.f;
                __half *r0 = Xs + (16 * mt + g) * FIELD_LD + 8 * nt + 2 * t;
                *reinterpret_cast<__half2 *>(r0) = __floats2half2_rn(acc[mt][nt][0], acc[mt][nt][1]);
                *reinterpret_cast<__half2 *>(r0 + 8 * FIELD_LD) = __floats2half2_rn(acc[mt][nt][2], acc[mt][nt][3]);
            }
        __syncwarp();
        warp_tile_to_global(Xs, act.dh1, row0, n);
        warp_gemm_t<8, 4>(Xs, FIELD_LD, W1T, FIELD_LD, acc);
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                __half *r0 = Xs + (16 * mt + g) * FIELD_LD + 8 * nt + 2 * t;
                *reinterpret_cast<__half2 *>(r0) = __floats2half2_rn(acc[mt][nt][0], acc[mt][nt][1]);
                *reinterpret_cast<__half2 *>(r0 + 8 * FIELD_LD) = __floats2half2_rn(acc[mt][nt][2], acc[mt][nt][3]);
            }
        __syncwarp();
        warp_tile_to_global(Xs, act.dx, row0, n);      // dx^ (normalised): scattered into the grid by k_field_backward_scatter
        __syncwarp();
    }
}

// Kernel A2: one lane per sample, no gathers: recompute the encoder's indices / weights and add s_i * w_corner * dx^ to the grid
// gradient with 8-byte vector reductions.  Consecutive lanes are the spp samples of one pixel, so on the coarse and middle
// levels the whole warp (or a few groups of lanes) falls into the SAME grid cell: those lanes are reduced with shuffles first and
// one lane per corner issues the reduction -- "reduce per warp, then one atomic per parameter tile" -- which removes the
// same-address serialisation at the L2 atomic units.  Lanes in many different cells (fine levels) go straight to the reductions.
#ifndef FIELD_SCATTER_MAX_GROUPS
#define FIELD_SCATTER_MAX_GROUPS 8
#endif
template <bool WS>
__global__ void __launch_bounds__(256) k_field_backward_scatter(IrisShadeParams P, int64_t n, const float *__restrict__ position,
                                                                 const float4 *__restrict__ r5, FieldAct act, float *__restrict__ d_grid) {
    // grid-stride over 256-sample blocks: the launch may cap the grid (scatter_ctas_per_sm) so that the kernel leaves room on every SM
    // for a kernel of another stream (the next tile's ray casts are issue-bound, this kernel is bound by the reduction rate of the LSU)
    const unsigned lane = threadIdx.x & 31u;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i - threadIdx.x < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float sc = i < n ? act.s[i] : 0.f;
    const bool live = sc > 0.f;
    if (!__any_sync(0xffffffffu, live)) continue;
    f3 p = mk3(0.f, 0.f, 0.f);
    if (live) {
        if (WS) { const float4 a = r5[i]; p = mk3(a.x, a.y, a.z); } else p = ld3(position, i);
    }
    const f3 x = mk3(field_coord(p.x, P.field_vmin, P.field_range), field_coord(p.y, P.field_vmin, P.field_range),
                     field_coord(p.z, P.field_vmin, P.field_range));
    const uint4 *row = reinterpret_cast<const uint4 *>(act.dx + 64 * (live ? i : 0));
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {              // 8 halfs = 4 levels per 16-byte load
        uint4 v = make_uint4(0, 0, 0, 0);
        if (live) v = __ldg(row + q);
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int l = 4 * q + j;
            const float2 dx = __half22float2(*reinterpret_cast<const __half2 *>(&w4[j]));
            const float gx = dx.x * sc, gy = dx.y * sc;
            const bool has = live && (gx != 0.f || gy != 0.f);
            const FieldLevel L = c_levels[l];
            const float px = __fmaf_rn(L.scale, x.x, 0.5f), py = __fmaf_rn(L.scale, x.y, 0.5f), pz = __fmaf_rn(L.scale, x.z, 0.5f);
            const uint32_t cx = (uint32_t)__float2int_rz(floorf(px)), cy = (uint32_t)__float2int_rz(floorf(py)), cz = (uint32_t)__float2int_rz(floorf(pz));
            // exact cell key: coordinates of one level span < 2^17, so 21 bits per axis are injective
            const unsigned long long key = has ? ((unsigned long long)(cx & 0x1FFFFFu) | ((unsigned long long)(cy & 0x1FFFFFu) << 21) |
                                                  ((unsigned long long)(cz & 0x1FFFFFu) << 42))
                                               : 0xFFFFFFFFFFFFFFFFull;
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            const bool leader = (unsigned)(__ffs(peers) - 1) == lane;
            const unsigned leaders = __ballot_sync(0xffffffffu, leader && has);
            float wc[8];
            uint32_t idxc[8];
            if (l < FIELD_DENSE_LEVELS) {      // generic (runtime) dense indexing: 7 of 32 levels
                const float wx1 = xsub(px, floorf(px)), wy1 = xsub(py, floorf(py)), wz1 = xsub(pz, floorf(pz));
                const float wx0 = xsub(1.0f, wx1), wy0 = xsub(1.0f, wy1), wz0 = xsub(1.0f, wz1);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t ix = cx + (c & 1), iy = cy + ((c >> 1) & 1), iz = cz + ((c >> 2) & 1);
                    idxc[c] = (ix + iy * L.res + iz * L.res * L.res) % L.size;
                    wc[c] = xmul(xmul((c & 1) ? wx1 : wx0, (c & 2) ? wy1 : wy0), (c & 4) ? wz1 : wz0);
                }
            } else {
                field_level_indices<-1>(L, x, idxc, wc);
            }
            if (__popc(leaders) <= FIELD_SCATTER_MAX_GROUPS) {
                unsigned todo = leaders;
                while (todo) {                                     // warp-uniform loop over the (few) distinct cells
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const unsigned long long k0 = __shfl_sync(0xffffffffu, key, src);
                    const bool mine = has && key == k0;
                    float sx[8], sy[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        sx[c] = mine ? wc[c] * gx : 0.f;
                        sy[c] = mine ? wc[c] * gy : 0.f;
                    }
                    // Reduce the 8 corner pairs over the warp by recursive halving: at offsets 16, 8, 4 a lane keeps half of its
                    // items and hands the other half to its partner (8 + 4 + 2 shuffles), then two butterfly steps on the one
                    // item left: 18 shuffles instead of 80.  Lanes 4c .. 4c+3 end up with the warp sum of corner c.
                    const bool b4 = lane & 16u, b3 = lane & 8u, b2 = lane & 4u;
                    float ax[4], ay[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float keepx = b4 ? sx[k + 4] : sx[k], keepy = b4 ? sy[k + 4] : sy[k];
                        const float sendx = b4 ? sx[k] : sx[k + 4], sendy = b4 ? sy[k] : sy[k + 4];
                        ax[k] = keepx + __shfl_xor_sync(0xffffffffu, sendx, 16);
                        ay[k] = keepy + __shfl_xor_sync(0xffffffffu, sendy, 16);
                    }
                    float bx[2], by[2];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const float keepx = b3 ? ax[k + 2] : ax[k], keepy = b3 ? ay[k + 2] : ay[k];
                        const float sendx = b3 ? ax[k] : ax[k + 2], sendy = b3 ? ay[k] : ay[k + 2];
                        bx[k] = keepx + __shfl_xor_sync(0xffffffffu, sendx, 8);
                        by[k] = keepy + __shfl_xor_sync(0xffffffffu, sendy, 8);
                    }
                    float rx = (b2 ? bx[1] : bx[0]) + __shfl_xor_sync(0xffffffffu, b2 ? bx[0] : bx[1], 4);
                    float ry = (b2 ? by[1] : by[0]) + __shfl_xor_sync(0xffffffffu, b2 ? by[0] : by[1], 4);
                    rx += __shfl_xor_sync(0xffffffffu, rx, 2);
                    ry += __shfl_xor_sync(0xffffffffu, ry, 2);
                    rx += __shfl_xor_sync(0xffffffffu, rx, 1);
                    ry += __shfl_xor_sync(0xffffffffu, ry, 1);
                    // corner c = lane >> 2 with the group leader's index for it (static select tree, no dynamic register indexing)
                    uint32_t id8[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) id8[c] = __shfl_sync(0xffffffffu, idxc[c], src);
                    const uint32_t i4a = b4 ? id8[4] : id8[0], i4b = b4 ? id8[5] : id8[1], i4c = b4 ? id8[6] : id8[2], i4d = b4 ? id8[7] : id8[3];
                    const uint32_t i2a = b3 ? i4c : i4a, i2b = b3 ? i4d : i4b;
                    const uint32_t id = b2 ? i2b : i2a;
                    if ((lane & 3u) == 0u) red_add_v2(d_grid + 2 * (int64_t)(L.offset + id), rx, ry);
                }
            } else if (has) {
#ifndef FIELD_SCATTER_NO_V4
                if (l >= FIELD_DENSE_LEVELS && (cx & 1u) == 0u) {
                    // hashed level, even cell x: corners c and c+1 (x and x+1, same y z) hash to idx and idx ^ 1 -- the two halves of one
                    // aligned 16-byte slot of the gradient table -- so four 16-byte reductions replace eight 8-byte ones
#pragma unroll
                    for (int c = 0; c < 8; c += 2) {
                        const bool swap = idxc[c] & 1u;                                 // the even-x corner sits in the odd entry
                        const float ax = wc[c] * gx, ay = wc[c] * gy, bx = wc[c + 1] * gx, by = wc[c + 1] * gy;
                        red_add_v4(d_grid + 2 * (int64_t)(L.offset + (idxc[c] & ~1u)), swap ? bx : ax, swap ? by : ay, swap ? ax : bx, swap ? ay : by);
                    }
                } else
#endif
                {
#pragma unroll
                    for (int c = 0; c < 8; ++c) red_add_v2(d_grid + 2 * (int64_t)(L.offset + idxc[c]), wc[c] * gx, wc[c] * gy);
                }
            }
        }
    }
    }
}

__device__ __forceinline__ void mma1688_tf32(float c[4], const uint32_t a[4], const uint32_t b[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// dW[o][i] += sum_s s_s * D[s][o] * H[s][i] for one 32-sample tile: warp w owns rows [16w,16w+16) (MT16 = number of 16-row
// slices that exist: 4 for the 64-row matrices, 1 for W3)
// acc (16 x 64 slice of dW, rows mrow0..mrow0+15) += sum over the tile's 32 samples of s * D[sample][row] * H[sample][col], on
// mma.sync m16n8k8 TF32.  The MMA's row / column labels are free as long as the A/B fragments and the accumulator agree, so they are
// chosen for wide shared-memory loads: MMA row g <-> feature mrow0 + 2g, row g+8 <-> mrow0 + 2g + 1 (one 4-byte load gives both),
// MMA column g of n-tile nt <-> feature 8g + nt (one 16-byte load gives the lane its column of all eight n-tiles).  Against scalar
// 2-byte loads this is 16 instead of 80 shared-memory loads per lane and call; wgrad_store() undoes the permutation.
#ifndef FIELD_WGRAD_SCALAR_LOADS
__device__ __forceinline__ void wgrad_tile(const __half *D, int ldd, const __half *H, const float *sc, int mrow0, float acc[8][4]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int k0 = 8 * ks;
        const float s0 = sc[k0 + t], s1 = sc[k0 + t + 4];
        const float2 d0 = __half22float2(*reinterpret_cast<const __half2 *>(D + (k0 + t) * ldd + mrow0 + 2 * g));
        const float2 d1 = __half22float2(*reinterpret_cast<const __half2 *>(D + (k0 + t + 4) * ldd + mrow0 + 2 * g));
        uint32_t a[4];
        a[0] = to_tf32(d0.x * s0);
        a[1] = to_tf32(d0.y * s0);
        a[2] = to_tf32(d1.x * s1);
        a[3] = to_tf32(d1.y * s1);
        const uint4 h0 = *reinterpret_cast<const uint4 *>(H + (k0 + t) * FIELD_LD + 8 * g);
        const uint4 h1 = *reinterpret_cast<const uint4 *>(H + (k0 + t + 4) * FIELD_LD + 8 * g);
        const uint32_t w0[4] = {h0.x, h0.y, h0.z, h0.w}, w1[4] = {h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float2 p0 = __half22float2(*reinterpret_cast<const __half2 *>(&w0[nt >> 1]));
            const float2 p1 = __half22float2(*reinterpret_cast<const __half2 *>(&w1[nt >> 1]));
            uint32_t b[2];
            b[0] = __float_as_uint((nt & 1) ? p0.y : p0.x);
            b[1] = __float_as_uint((nt & 1) ? p1.y : p1.x);
            mma1688_tf32(acc[nt], a, b);
        }
    }
}
// adds a warp's accumulators to dW (row-major [out][64]): accumulator (nt, k) holds MMA row g (+8 for k >= 2), MMA column 2t + (k & 1)
__device__ __forceinline__ void wgrad_store(float *dW, int mrow0, const float acc[8][4]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int k = 0; k < 4; ++k) atomicAdd(dW + (mrow0 + 2 * g + (k >> 1)) * 64 + 8 * (2 * t + (k & 1)) + nt, acc[nt][k]);
}
#else
__device__ __forceinline__ void wgrad_tile(const __half *D, int ldd, const __half *H, const float *sc, int mrow0, float acc[8][4]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int k0 = 8 * ks;
        const float s0 = sc[k0 + t], s1 = sc[k0 + t + 4];
        uint32_t a[4];
        a[0] = to_tf32(__half2float(D[(k0 + t) * ldd + mrow0 + g]) * s0);
        a[1] = to_tf32(__half2float(D[(k0 + t) * ldd + mrow0 + g + 8]) * s0);
        a[2] = to_tf32(__half2float(D[(k0 + t + 4) * ldd + mrow0 + g]) * s1);
        a[3] = to_tf32(__half2float(D[(k0 + t + 4) * ldd + mrow0 + g + 8]) * s1);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            uint32_t b[2];
            b[0] = __float_as_uint(__half2float(H[(k0 + t) * FIELD_LD + 8 * nt + g]));
            b[1] = __float_as_uint(__half2float(H[(k0 + t + 4) * FIELD_LD + 8 * nt + g]));
            mma1688_tf32(acc[nt], a, b);
        }
    }
}

__device__ __forceinline__ void wgrad_store(float *dW, int mrow0, const float acc[8][4]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int k = 0; k < 4; ++k) atomicAdd(dW + (mrow0 + g + 8 * (k >> 1)) * 64 + 8 * nt + 2 * t + (k & 1), acc[nt][k]);
}
#endif

// Shared-memory image of one 32-sample tile: X | h1 | h2 | dh2^ | dh1^ (32 x 64 fp16, ld 72) | dy^ (32 x 16, ld 24) | s (32 floats)
#define FIELD_WGRAD_TILE_BYTES (5 * 32 * FIELD_LD * 2 + 32 * FIELD_LD3 * 2 + 32 * 4)
#define FIELD_WGRAD_STAGES 3
#define FIELD_WGRAD_SMEM_BYTES (FIELD_WGRAD_STAGES * FIELD_WGRAD_TILE_BYTES)

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;     // src-size 0 -> the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}

// issue the asynchronous copies of one tile (no registers staged, LDGSTS)
__device__ __forceinline__ void wgrad_stage_tile(const FieldAct &act, int64_t row0, int64_t n, unsigned char *buf) {
    __half *sX = reinterpret_cast<__half *>(buf);
    for (int v = threadIdx.x; v < 5 * 256; v += IRIS_BLOCK) {
        const int which = v >> 8, r = (v & 255) >> 3, c = (v & 7) * 8;
        const __half *src = which == 0 ? act.X : which == 1 ? act.h1 : which == 2 ? act.h2 : which == 3 ? act.dh2 : act.dh1;
        const bool ok = row0 + r < n;
        cp_async16(sX + which * 32 * FIELD_LD + r * FIELD_LD + c, src + (ok ? (row0 + r) * 64 + c : 0), ok);
    }
    __half *sDy = sX + 5 * 32 * FIELD_LD;
    float *sS = reinterpret_cast<float *>(sDy + 32 * FIELD_LD3);
    if (threadIdx.x < 64) {
        const int r = threadIdx.x >> 1, c = (threadIdx.x & 1) * 8;
        const bool ok = row0 + r < n;
        cp_async16(sDy + r * FIELD_LD3 + c, act.dy + (ok ? (row0 + r) * 16 + c : 0), ok);
    } else if (threadIdx.x < 96) {
        const int r = threadIdx.x - 64;
        const bool ok = row0 + r < n;
        cp_async4(sS + r, act.s + (ok ? row0 + r : 0), ok);
    }
}

__global__ void __launch_bounds__(IRIS_BLOCK) k_field_backward_wgrad(FieldAct act, int64_t n, float *__restrict__ d_mlp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float a1[8][4], a2[8][4], a3[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int k = 0; k < 4; ++k) a1[nt][k] = a2[nt][k] = a3[nt][k] = 0.f;
    const int64_t n_tiles = (n + 31) / 32;
    // software pipeline: FIELD_WGRAD_STAGES tiles in flight per CTA
    int64_t next = blockIdx.x;
#pragma unroll
    for (int s = 0; s < FIELD_WGRAD_STAGES - 1; ++s) {
        if (next < n_tiles) wgrad_stage_tile(act, next * 32, n, smem_raw + s * FIELD_WGRAD_TILE_BYTES);
        asm volatile("cp.async.commit_group;" ::: "memory");
        next += gridDim.x;
    }
    int stage = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        {   // prefetch tile + (STAGES-1) into the buffer freed by the previous iteration
            const int ps = (stage + FIELD_WGRAD_STAGES - 1) % FIELD_WGRAD_STAGES;
            if (next < n_tiles) wgrad_stage_tile(act, next * 32, n, smem_raw + ps * FIELD_WGRAD_TILE_BYTES);
            asm volatile("cp.async.commit_group;" ::: "memory");
            next += gridDim.x;
        }
        asm volatile("cp.async.wait_group %0;" ::"n"(FIELD_WGRAD_STAGES - 1) : "memory");
        __syncthreads();
        const __half *sX = reinterpret_cast<const __half *>(smem_raw + stage * FIELD_WGRAD_TILE_BYTES);
        const __half *sH1 = sX + 32 * FIELD_LD, *sH2 = sH1 + 32 * FIELD_LD, *sD2 = sH2 + 32 * FIELD_LD, *sD1 = sD2 + 32 * FIELD_LD;
        const __half *sDy = sD1 + 32 * FIELD_LD;
        const float *sS = reinterpret_cast<const float *>(sDy + 32 * FIELD_LD3);
        wgrad_tile(sD1, FIELD_LD, sX, sS, 16 * warp, a1);     // dW1 rows [16w,16w+16)
        wgrad_tile(sD2, FIELD_LD, sH1, sS, 16 * warp, a2);    // dW2
        if (warp == 0) wgrad_tile(sDy, FIELD_LD3, sH2, sS, 0, a3);   // dW3 (16 rows)
        __syncthreads();                                      // everyone is done with this buffer before it is refilled
        stage = (stage + 1) % FIELD_WGRAD_STAGES;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // one atomic per weight per CTA
    wgrad_store(d_mlp, 16 * warp, a1);
    wgrad_store(d_mlp + 4096, 16 * warp, a2);
    if (warp == 0) wgrad_store(d_mlp + 8192, 0, a3);
}
