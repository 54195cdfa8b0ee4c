// field_tc5.cuh -- the BRDF field forward on Blackwell's 5th-generation tensor cores (tcgen05 + TMEM).
//
// One CTA = 128 threads = one 128-row MMA tile: thread i owns sample i of the tile from the hash-grid gathers to the sigmoid.
//   * each thread writes its 64 fp16 features as eight 16-byte chunks into shared memory in the canonical K-major,
//     no-swizzle UMMA layout (8x8 core matrices of 128 contiguous bytes; LBO = distance between K-adjacent core matrices,
//     SBO = distance between 8-row groups);
//   * one elected thread issues four tcgen05.mma (M=128, N=64, K=16, fp16 x fp16 -> fp32) per layer against the weights,
//     resident in shared memory in the same layout, accumulating in 64 TMEM columns, and commits to an mbarrier;
//   * tcgen05.ld 32x32b hands every thread ITS row of the accumulator (TMEM lane i <-> thread i): ReLU, round to fp16, and the
//     row goes back to shared memory as the next layer's A operand.  No fragment shuffles, no register transposes.
// Rounding points are those of oracle/field.py (fp16 activations, fp32 accumulate, fp16 output, fp16 sigmoid).
#pragma once
#include "field.cuh"

#define TC5_ROWS 128
#define TC5_A_BYTES (TC5_ROWS * 128)                 // 128 rows x 64 halfs
#define TC5_A_LBO 2048                                // 16 row groups x 128 B
#define TC5_W_LBO 1024                                // 8 row groups x 128 B   (64 output rows)
#define TC5_W3_LBO 256                                // 2 row groups x 128 B   (16 output rows)
#define TC5_SBO 128
#define TC5_SMEM_BYTES (TC5_A_BYTES + 2 * 8192 + 2048 + 64)
#define TC5_TMEM_COLS 64

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    // SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout_type=0 (no swizzle) [61,64)
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) {
    // InstrDescriptor: c_format F32 (1) [4,6) | a_format F16 (0) [7,10) | b_format F16 (0) [10,13) | K-major A,B | N>>3 [17,23) | M>>4 [24,29)
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spins > (1u << 24)) __trap();        // a lost commit must fail loudly, never hang the device
    }
}
#define TC5_LD16(r, taddr)                                                                                                        \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"        \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),     \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                        \
                 : "r"(taddr))

// weights W[rows][64] (fp16, row-major, global) -> canonical K-major core-matrix layout in shared memory
__device__ __forceinline__ void tc5_stage_weights(const __half *__restrict__ W, int rows, uint32_t lbo, unsigned char *dst) {
    for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {          // one 16-byte chunk (8 halfs of one row) per step
        const int n = i >> 3, kc = i & 7;
        const uint4 v = *reinterpret_cast<const uint4 *>(W + n * 64 + kc * 8);
        *reinterpret_cast<uint4 *>(dst + kc * lbo + (n >> 3) * TC5_SBO + (n & 7) * 16) = v;
    }
}

__device__ int g_tc5_debug = 0;   // 0 normal, 1 skip the encode, 2 skip the MLP (profiling experiments only)

template <bool WS>
__global__ void __launch_bounds__(TC5_ROWS) k_field_forward_tc5(IrisShadeParams P, int64_t n, const float *__restrict__ position, float *__restrict__ mat,
                                                                 const float4 *__restrict__ w0, float4 *__restrict__ w1, float4 *__restrict__ w2,
                                                                 __half *__restrict__ x_save, int pair, const unsigned long long *__restrict__ n_dev) {
    if (n_dev != nullptr) n = min(n, (int64_t)*n_dev);         // rows counted on the device (live lanes of a wavefront bounce)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *sA = smem_raw;
    unsigned char *sW1 = sA + TC5_A_BYTES, *sW2 = sW1 + 8192, *sW3 = sW2 + 8192;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sW3 + 2048);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    const __half *mlp = reinterpret_cast<const __half *>(P.mlp_f16);
    tc5_stage_weights(mlp, 64, TC5_W_LBO, sW1);
    tc5_stage_weights(mlp + 4096, 64, TC5_W_LBO, sW2);
    tc5_stage_weights(mlp + 8192, 16, TC5_W3_LBO, sW3);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TC5_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t bar = smem_u32(mbar);
    const uint32_t aA = smem_u32(sA), aW1 = smem_u32(sW1), aW2 = smem_u32(sW2), aW3 = smem_u32(sW3);
    const uint32_t idesc64 = umma_idesc_f16(128, 64), idesc16 = umma_idesc_f16(128, 16);
    const uint32_t row_off = (tid >> 3) * TC5_SBO + (tid & 7) * 16;        // this thread's row inside every K chunk
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);              // TMEM lanes of this warp
    uint32_t phase = 0;
    const __half2 *grid = reinterpret_cast<const __half2 *>(P.grid_f16);
    const int64_t n_tiles = (n + TC5_ROWS - 1) / TC5_ROWS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = tile * TC5_ROWS + tid;
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
        // ---- encode: feature pair l lands in K chunk l/4 of this thread's row
        const int dbg = g_tc5_debug;
        if (active && dbg != 1) {
            const f3 x = mk3(field_coord(p.x, P.field_vmin, P.field_range), field_coord(p.y, P.field_vmin, P.field_range),
                             field_coord(p.z, P.field_vmin, P.field_range));
            unsigned char *rowp = sA + row_off;
            field_encode_to(grid, x, [rowp](int l, __half2 v) { *reinterpret_cast<__half2 *>(rowp + (l >> 2) * TC5_A_LBO + (l & 3) * 4) = v; }, pair != 0);
        } else {
#pragma unroll
            for (int kc = 0; kc < 8; ++kc) *reinterpret_cast<uint4 *>(sA + kc * TC5_A_LBO + row_off) = make_uint4(0, 0, 0, 0);
        }
#pragma unroll 1
        for (int layer = 0; layer < (dbg == 2 ? 0 : 3); ++layer) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes of A -> visible to the tensor core
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t aW = layer == 0 ? aW1 : layer == 1 ? aW2 : aW3;
                const uint32_t wl = layer == 2 ? TC5_W3_LBO : TC5_W_LBO;
                const uint32_t id = layer == 2 ? idesc16 : idesc64;
#pragma unroll
                for (int k = 0; k < 4; ++k)                                     // K = 64 = 4 x 16: two K chunks per instruction
                    umma_f16(tmem, umma_desc(aA + 2 * k * TC5_A_LBO, TC5_A_LBO, TC5_SBO), umma_desc(aW + 2 * k * wl, wl, TC5_SBO), id, k > 0 ? 1u : 0u);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
            }
            if (layer == 0 && x_save) {
                // keep the encoded inputs for the adjoint (row-major [n][64] fp16), while the tensor core reads the same tile:
                // lane v of a warp copies 16-byte chunk v & 7 of row v >> 3, so every store instruction covers 4 full rows
                const int wrow = tid & ~31, lane = tid & 31;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int v = it * 32 + lane, r = wrow + (v >> 3), c = v & 7;
                    const int64_t grow = tile * TC5_ROWS + r;
                    if (grow < n) *reinterpret_cast<uint4 *>(x_save + grow * 64 + c * 8) = *reinterpret_cast<const uint4 *>(sA + c * TC5_A_LBO + r * 16);
                }
                __syncwarp();      // the rows read above belong to other lanes of this warp, which overwrite them in their epilogue
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (layer < 2) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {                                   // 16 accumulator columns at a time
                    uint32_t r[16];
                    TC5_LD16(r, taddr + 16 * q);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    __align__(16) __half h[16];
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        *reinterpret_cast<__half2 *>(h + 2 * k) = __floats2half2_rn(fmaxf(__uint_as_float(r[2 * k]), 0.f), fmaxf(__uint_as_float(r[2 * k + 1]), 0.f));
                    *reinterpret_cast<uint4 *>(sA + (2 * q) * TC5_A_LBO + row_off) = *reinterpret_cast<const uint4 *>(h);
                    *reinterpret_cast<uint4 *>(sA + (2 * q + 1) * TC5_A_LBO + row_off) = *reinterpret_cast<const uint4 *>(h + 8);
                }
            } else {
                uint32_t r[16];
                TC5_LD16(r, taddr);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (active) {
                    const float a0 = sigmoid16(__uint_as_float(r[0])), a1 = sigmoid16(__uint_as_float(r[1])), a2 = sigmoid16(__uint_as_float(r[2]));
                    const float rr = sigmoid16(__uint_as_float(r[3])) * 0.98f + 0.02f, mm = sigmoid16(__uint_as_float(r[4]));
                    if (WS) {
                        w2[i] = make_float4(a0, a1, a2, rr);
                        float4 b = w1[i];
                        b.w = mm;
                        w1[i] = b;
                    } else {
                        float *o = mat + 5 * i;
                        o[0] = a0; o[1] = a1; o[2] = a2; o[3] = rr; o[4] = mm;
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                                        // TMEM and A are free for the next tile
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC5_TMEM_COLS) : "memory");
}
