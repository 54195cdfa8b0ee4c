// field_bwd_tc5.cuh -- the field adjoint's dgrad AND wgrad as ONE tcgen05 kernel (default whenever the forward kept the encoded inputs;
// iris_set_option("field_backward_impl", 0) selects the two-kernel form of field.cuh), so that the activations h1, h2 and the back-propagated dh2^, dh1^ never leave the SM (DESIGN.md section 9, item 1:
// the two-kernel form moves 1.6 KB/sample through HBM, this one 0.26 KB).  The grid scatter stays k_field_backward_scatter.
//
// One CTA = 128 threads = one 128-sample tile, thread i owns sample i (as in field_tc5.cuh).  Per tile, six MMA rounds:
//   forward   h1 = relu(X W1^T), h2 = relu(h1 W2^T), y = h2 W3^T                       (A = activation tile, B = weight tile, both K-major)
//   dgrad     dh2^ = (dy^ W3) . relu'(h2), dh1^ = (dh2^ W2) . relu'(h1), dx^ = dh1^ W1   (B = the SAME weight tiles read MN-major)
//   wgrad     dW3^T += h2^T dy~, dW2 += dh2~^T h1, dW1 += dh1~^T X                     (both operands = the activation tiles read MN-major,
//                                                                                       K = samples, M = 64; accumulators stay in TMEM
//                                                                                       over all tiles of the CTA)
// dh^ is normalised per sample by a power of two s_i (the grid gradient needs the small samples' relative precision); dh~ = dh^ *
// s_i / S is the wgrad operand, S = one power of two per CTA that bounds every |dy| of the CTA's samples (from max|d_mat| / 4), so
// dh~ fits fp16 and the accumulators are rescaled by S exactly before they leave.  Operand mechanics verified by
// tools/probe/umma_mn_probe.cu; tests/test_gpu_parity.py::test_fused_tcgen05_adjoint_equals_two_kernel_adjoint checks it against the
// two-kernel form (1e-6 of max), and the oracle / golden gradient tests run through it.
#pragma once
#include "field_tc5.cuh"

#define BT5_TILE_BYTES 16384
#define BT5_SMEM_BYTES (5 * BT5_TILE_BYTES + 2 * 8192 + 2048 + 64)
#define BT5_TMEM_COLS 256            // [0,64) forward / dgrad accumulator, [64,128) dW1, [128,192) dW2, [192,208) dW3^T

__device__ __forceinline__ uint32_t umma_idesc_f16_major(int M, int N, bool a_mn, bool b_mn) {
    return umma_idesc_f16(M, N) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
}
// MN-major view of a tile stored K-major with chunk stride `chunk_stride`: LBO = 128 (8-row groups), SBO = the chunk stride
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t chunk_stride) { return umma_desc(saddr, TC5_SBO, chunk_stride); }

// this thread's 16 accumulator columns [16q, 16q+16) of its TMEM lane
#define BT5_LD16(r, taddr, q)                      \
    do {                                           \
        TC5_LD16(r, (taddr) + 16 * (q));           \
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); \
    } while (0)

template <bool WS>
__global__ void __launch_bounds__(TC5_ROWS) k_field_backward_tc5(IrisShadeParams P, int64_t n, const float4 *__restrict__ r5, const float *__restrict__ d_mat,
                                                                  const __half *__restrict__ x_enc, __half *__restrict__ dx_out, float *__restrict__ s_out,
                                                                  float *__restrict__ d_mlp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *sX = smem_raw, *sH1 = sX + BT5_TILE_BYTES, *sH2 = sH1 + BT5_TILE_BYTES, *sD = sH2 + BT5_TILE_BYTES, *sDw = sD + BT5_TILE_BYTES;
    unsigned char *sW1 = sDw + BT5_TILE_BYTES, *sW2 = sW1 + 8192, *sW3 = sW2 + 8192;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sW3 + 2048);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 1);
    float *red = reinterpret_cast<float *>(tmem_slot + 1);                      // 4 floats: per-warp maxima
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const __half *mlp = reinterpret_cast<const __half *>(P.mlp_f16);
    tc5_stage_weights(mlp, 64, TC5_W_LBO, sW1);
    tc5_stage_weights(mlp + 4096, 64, TC5_W_LBO, sW2);
    tc5_stage_weights(mlp + 8192, 16, TC5_W3_LBO, sW3);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BT5_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- S: one power of two per CTA with |dy| <= max|d_mat| / 4 <= S for every sample this CTA will see
    const int64_t n_tiles = (n + TC5_ROWS - 1) / TC5_ROWS;
    float mx = 0.f;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = tile * TC5_ROWS + tid;
        if (i < n)
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const float v = fabsf(d_mat[5 * i + k]);
                if (v < __int_as_float(0x7f800000)) mx = fmaxf(mx, v);
            }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    float S = 1.f;
    if (mx > 0.f) {
        int e;
        frexpf(mx, &e);                                                          // mx < 2^e
        S = ldexpf(1.0f, e - 2);                                                 // |dy| <= mx / 4 < S
    }
    const float invS = 1.0f / S;
    const uint32_t tmem = *tmem_slot;
    const uint32_t bar = smem_u32(mbar);
    const uint32_t aX = smem_u32(sX), aH1 = smem_u32(sH1), aH2 = smem_u32(sH2), aD = smem_u32(sD), aDw = smem_u32(sDw);
    const uint32_t aW1 = smem_u32(sW1), aW2 = smem_u32(sW2), aW3 = smem_u32(sW3);
    const uint32_t id_f64 = umma_idesc_f16(128, 64), id_f16 = umma_idesc_f16(128, 16);                 // forward
    const uint32_t id_d64 = umma_idesc_f16_major(128, 64, false, true);                                // dgrad: B MN-major
    const uint32_t id_w64 = umma_idesc_f16_major(64, 64, true, true), id_w16 = umma_idesc_f16_major(64, 16, true, true);   // wgrad
    const uint32_t row_off = (tid >> 3) * TC5_SBO + (tid & 7) * 16;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0;
    bool first = true;                                                           // first tile of this CTA: dW accumulators start from zero
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = tile * TC5_ROWS + tid;
        bool active = i < n;
        float dm[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (active) {
            if (WS) active = __float_as_int(r5[i].w) == -2;
#pragma unroll
            for (int k = 0; k < 5; ++k) dm[k] = d_mat[5 * i + k];
            active = active && (dm[0] != 0.f || dm[1] != 0.f || dm[2] != 0.f || dm[3] != 0.f || dm[4] != 0.f);
        }
        // ---- X tile: this thread's row of the encoded inputs into the K-major layout
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (i < n) v = __ldg(reinterpret_cast<const uint4 *>(x_enc + i * 64 + kc * 8));
            *reinterpret_cast<uint4 *>(sX + kc * TC5_A_LBO + row_off) = v;
        }
        uint32_t m1[2] = {0u, 0u}, m2[2] = {0u, 0u};
        float sc = 0.f;
        // six rounds: 0,1,2 forward layers; 3,4,5 dgrad layers, each together with the wgrad GEMM whose operands are ready by then
#pragma unroll 1
        for (int round = 0; round < 6; ++round) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t accw = first ? 0u : 1u;
                if (round == 0) {
                    for (int k = 0; k < 4; ++k) umma_f16(tmem, umma_desc(aX + 2 * k * TC5_A_LBO, TC5_A_LBO, TC5_SBO), umma_desc(aW1 + 2 * k * TC5_W_LBO, TC5_W_LBO, TC5_SBO), id_f64, k > 0);
                } else if (round == 1) {
                    for (int k = 0; k < 4; ++k) umma_f16(tmem, umma_desc(aH1 + 2 * k * TC5_A_LBO, TC5_A_LBO, TC5_SBO), umma_desc(aW2 + 2 * k * TC5_W_LBO, TC5_W_LBO, TC5_SBO), id_f64, k > 0);
                } else if (round == 2) {
                    for (int k = 0; k < 4; ++k) umma_f16(tmem, umma_desc(aH2 + 2 * k * TC5_A_LBO, TC5_A_LBO, TC5_SBO), umma_desc(aW3 + 2 * k * TC5_W3_LBO, TC5_W3_LBO, TC5_SBO), id_f16, k > 0);
                } else if (round == 3) {
                    // dh2 = dy^ W3 : A = sD chunks 0,1 (K = 16 outputs), B = W3 tile [16 out][64 in] read MN-major (chunk stride 256)
                    umma_f16(tmem, umma_desc(aD, TC5_A_LBO, TC5_SBO), umma_desc_mn(aW3, TC5_W3_LBO), id_d64, 0u);
                    // dW3^T (64 x 16) += h2^T dy~ : A = sH2 MN-major (M = 64 features), B = sDw chunks 0,1 MN-major (N = 16), K = 128 samples
                    for (int k = 0; k < 8; ++k) umma_f16(tmem + 192, umma_desc_mn(aH2 + k * 256, TC5_A_LBO), umma_desc_mn(aDw + k * 256, TC5_A_LBO), id_w16, k > 0 ? 1u : accw);
                } else if (round == 4) {
                    for (int k = 0; k < 4; ++k) umma_f16(tmem, umma_desc(aD + 2 * k * TC5_A_LBO, TC5_A_LBO, TC5_SBO), umma_desc_mn(aW2 + k * 256, TC5_W_LBO), id_d64, k > 0);
                    // dW2 (64 x 64) += dh2~^T h1
                    for (int k = 0; k < 8; ++k) umma_f16(tmem + 128, umma_desc_mn(aDw + k * 256, TC5_A_LBO), umma_desc_mn(aH1 + k * 256, TC5_A_LBO), id_w64, k > 0 ? 1u : accw);
                } else {
                    for (int k = 0; k < 4; ++k) umma_f16(tmem, umma_desc(aD + 2 * k * TC5_A_LBO, TC5_A_LBO, TC5_SBO), umma_desc_mn(aW1 + k * 256, TC5_W_LBO), id_d64, k > 0);
                    // dW1 (64 x 64) += dh1~^T X
                    for (int k = 0; k < 8; ++k) umma_f16(tmem + 64, umma_desc_mn(aDw + k * 256, TC5_A_LBO), umma_desc_mn(aX + k * 256, TC5_A_LBO), id_w64, k > 0 ? 1u : accw);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (round <= 1) {
                // ---- h = relu(acc) -> fp16 -> next layer's A tile; keep the ReLU mask of this row
                unsigned char *dst = round == 0 ? sH1 : sH2;
                uint32_t *msk = round == 0 ? m1 : m2;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t r[16];
                    BT5_LD16(r, taddr, q);
                    __align__(16) __half h[16];
                    uint32_t bits = 0;
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float v = __uint_as_float(r[k]);
                        bits |= (v > 0.f ? 1u : 0u) << k;
                        h[k] = __float2half_rn(fmaxf(v, 0.f));
                    }
                    msk[q >> 1] |= bits << (16 * (q & 1));
                    *reinterpret_cast<uint4 *>(dst + (2 * q) * TC5_A_LBO + row_off) = *reinterpret_cast<const uint4 *>(h);
                    *reinterpret_cast<uint4 *>(dst + (2 * q + 1) * TC5_A_LBO + row_off) = *reinterpret_cast<const uint4 *>(h + 8);
                }
            } else if (round == 2) {
                // ---- dy = d_mat * d(mat)/dy, per-sample power-of-two normalisation (field.cuh, k_field_backward_dgrad)
                uint32_t r[16];
                BT5_LD16(r, taddr, 0);
                float dy[5], mxy = 0.f;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const float yk = __half2float(__float2half_rn(__uint_as_float(r[k])));
                    const float s = 1.0f / (1.0f + expf(-yk));
                    dy[k] = active ? dm[k] * (k == 3 ? 0.98f : 1.0f) * s * (1.0f - s) : 0.f;
                    mxy = fmaxf(mxy, fabsf(dy[k]));
                }
                sc = 0.f;
                if (mxy > 0.f && mxy < __int_as_float(0x7f800000)) {
                    int e;
                    frexpf(mxy, &e);
                    sc = ldexpf(1.0f, e);
                }
                const float inv = sc > 0.f ? 1.0f / sc : 0.f;
                __align__(16) __half hd[16], hw[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    hd[k] = __float2half_rn(k < 5 ? dy[k] * inv : 0.f);
                    hw[k] = __float2half_rn(k < 5 ? dy[k] * invS : 0.f);
                }
                *reinterpret_cast<uint4 *>(sD + row_off) = *reinterpret_cast<const uint4 *>(hd);
                *reinterpret_cast<uint4 *>(sD + TC5_A_LBO + row_off) = *reinterpret_cast<const uint4 *>(hd + 8);
                *reinterpret_cast<uint4 *>(sDw + row_off) = *reinterpret_cast<const uint4 *>(hw);
                *reinterpret_cast<uint4 *>(sDw + TC5_A_LBO + row_off) = *reinterpret_cast<const uint4 *>(hw + 8);
                if (i < n) s_out[i] = sc;
            } else if (round == 3 || round == 4) {
                // ---- dh^ = acc . relu' -> fp16 -> sD (next dgrad A operand) and, rescaled by s_i / S, -> sDw (wgrad operand)
                const uint32_t *msk = round == 3 ? m2 : m1;
                const float ws = sc * invS;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t r[16];
                    BT5_LD16(r, taddr, q);
                    const uint32_t bits = msk[q >> 1] >> (16 * (q & 1));
                    __align__(16) __half hd[16], hw[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float v = ((bits >> k) & 1u) ? __uint_as_float(r[k]) : 0.f;
                        hd[k] = __float2half_rn(v);
                        hw[k] = __float2half_rn(__half2float(hd[k]) * ws);
                    }
                    *reinterpret_cast<uint4 *>(sD + (2 * q) * TC5_A_LBO + row_off) = *reinterpret_cast<const uint4 *>(hd);
                    *reinterpret_cast<uint4 *>(sD + (2 * q + 1) * TC5_A_LBO + row_off) = *reinterpret_cast<const uint4 *>(hd + 8);
                    *reinterpret_cast<uint4 *>(sDw + (2 * q) * TC5_A_LBO + row_off) = *reinterpret_cast<const uint4 *>(hw);
                    *reinterpret_cast<uint4 *>(sDw + (2 * q + 1) * TC5_A_LBO + row_off) = *reinterpret_cast<const uint4 *>(hw + 8);
                }
            } else if (round == 5) {
                // ---- dx^ (normalised) -> fp16 -> the stream k_field_backward_scatter reads
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t r[16];
                    BT5_LD16(r, taddr, q);
                    __align__(16) __half hd[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) hd[k] = __float2half_rn(__uint_as_float(r[k]));
                    if (i < n) {
                        *reinterpret_cast<uint4 *>(dx_out + i * 64 + 16 * q) = *reinterpret_cast<const uint4 *>(hd);
                        *reinterpret_cast<uint4 *>(dx_out + i * 64 + 16 * q + 8) = *reinterpret_cast<const uint4 *>(hd + 8);
                    }
                }
            }
        }
        first = false;
    }
    // ---- drain the weight-gradient accumulators: M = 64 rows live in TMEM lanes 32 (m / 16) + m % 16, i.e. lanes 0..15 of every warp
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!first) {
        const int m = 16 * warp + lane;                                          // feature row held by this lane (lanes < 16)
        for (int blk = 0; blk < 3; ++blk) {                                      // dW1, dW2, dW3^T
            const int ncol = blk == 2 ? 16 : 64;
            for (int q = 0; q < ncol / 16; ++q) {
                uint32_t r[16];
                BT5_LD16(r, taddr + 64 * (blk + 1), q);                          // all lanes take part in the load (warp-collective)
                if (lane < 16) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float v = __uint_as_float(r[k]) * S;
                        if (v != 0.f) {
                            if (blk == 0) atomicAdd(d_mlp + m * 64 + 16 * q + k, v);                       // dW1[out m][in]
                            else if (blk == 1) atomicAdd(d_mlp + 4096 + m * 64 + 16 * q + k, v);            // dW2[out m][in]
                            else atomicAdd(d_mlp + 8192 + (16 * q + k) * 64 + m, v);                       // dW3[out][in m] (accumulated transposed)
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BT5_TMEM_COLS) : "memory");
}
