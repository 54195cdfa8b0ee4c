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

// Encodes one sample: writes 64 fp16 features to dst[0..63].
__device__ __forceinline__ void field_encode(const __half2 *__restrict__ grid, f3 x, __half *dst) {
#pragma unroll 2
    for (int l = 0; l < FIELD_LEVELS; ++l) {
        const FieldLevel L = c_levels[l];
        const float px = __fmaf_rn(L.scale, x.x, 0.5f), py = __fmaf_rn(L.scale, x.y, 0.5f), pz = __fmaf_rn(L.scale, x.z, 0.5f);
        const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
        const float wx1 = xsub(px, fx), wy1 = xsub(py, fy), wz1 = xsub(pz, fz);
        const float wx0 = xsub(1.0f, wx1), wy0 = xsub(1.0f, wy1), wz0 = xsub(1.0f, wz1);
        const uint32_t cx = (uint32_t)__float2int_rz(fx), cy = (uint32_t)__float2int_rz(fy), cz = (uint32_t)__float2int_rz(fz);
        __half2 v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint32_t ix = cx + (c & 1), iy = cy + ((c >> 1) & 1), iz = cz + ((c >> 2) & 1);
            uint32_t idx;
            if (L.dense) idx = (ix + iy * L.res + iz * L.res * L.res) % L.size;
            else idx = (ix ^ (iy * 2654435761u) ^ (iz * 805459861u)) & (L.size - 1u);   // size == 2^19
            v[c] = __ldg(grid + L.offset + idx);
        }
        float f0 = 0.f, f1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float w = xmul(xmul((c & 1) ? wx1 : wx0, (c & 2) ? wy1 : wy0), (c & 4) ? wz1 : wz0);
            const float2 q = __half22float2(v[c]);
            f0 = xadd(f0, xmul(w, q.x));
            f1 = xadd(f1, xmul(w, q.y));
        }
        *reinterpret_cast<__half2 *>(dst + 2 * l) = __floats2half2_rn(f0, f1);
    }
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
template <bool WS>
__global__ void __launch_bounds__(IRIS_BLOCK) k_field_forward(IrisShadeParams P, int64_t n, const float *__restrict__ position, float *__restrict__ mat,
                                                               const float4 *__restrict__ w0, float4 *__restrict__ w1, float4 *__restrict__ w2) {
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
            field_encode(reinterpret_cast<const __half2 *>(P.grid_f16), x, row);
        } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) *reinterpret_cast<__half2 *>(row + 2 * k) = __floats2half2_rn(0.f, 0.f);
        }
        __syncwarp();
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
