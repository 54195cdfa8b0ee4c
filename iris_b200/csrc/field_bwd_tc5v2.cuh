// field_bwd_tc5v2.cuh -- second generation of the fused field adjoint (dgrad + wgrad in one tcgen05 kernel, field_bwd_tc5.cuh is
// the first; same mathematics, same rounding points, same TMEM accumulators for dW1, dW2, dW3^T).  What changed, driven by the ncu
// capture of the first kernel (profiles/r2a_k_field_backward_tc5.txt: 8 warps per SM, issue slots 23 % busy, tensor pipe 12 %,
// 6.8 M local-memory instructions per launch -- a latency chain of six serial "sync -> MMA -> wait -> epilogue" rounds per tile):
//   * TMA: the 16 KB X tile (this tile's 128 rows of the `encoded` array) arrives as eight cp.async.bulk.tensor boxes (8 halfs x 128
//     rows of a 2-D tensor map over the row-major [n][64] fp16 array), each landing as one 2 KB K-chunk of the canonical K-major
//     core-matrix layout; out-of-range rows are zero-filled by the hardware; completion is signalled on an mbarrier with expect_tx.
//     The next tile's X is requested as soon as the last MMA round of the current tile has consumed the buffer, so the load
//     overlaps the dx^ epilogue.  dx^ leaves the same way: the epilogue writes the tile into shared memory and eight
//     cp.async.bulk.tensor stores (one bulk group) send it to the [n][64] stream the grid scatter reads.
//   * 8 epilogue warps per tile instead of 4: TMEM lane quadrant = warp % 4, accumulator columns [32 (warp / 4), +32): one
//     tcgen05.ld.32x32b.x32 per thread and round, half the per-thread epilogue work, twice the warps to hide its latency.
//   * warp specialisation: a ninth warp issues every tcgen05.mma / TMA instruction (one elected lane, operand descriptors built once
//     before the tile loop, so a round costs it two adds per MMA instead of ~10 instructions of descriptor arithmetic on the critical
//     path); the hand-over is two mbarriers -- "operands written" (256 epilogue arrivals) and "accumulator ready" (tcgen05.commit) --
//     instead of a block-wide barrier per round, and no shared-memory word is ever passed between epilogue threads (both threads of a
//     row recompute the row's scale from the same accumulator columns).
//   * packed arithmetic, no local memory.  The ReLU derivative is relu'(acc) of the fp32 accumulator (the checker differentiates
//     fp16(relu(acc)) straight through the rounding), and a positive accumulator below 2^-25 rounds to an fp16 zero: so h is stored
//     as max(fp16x2(acc), -0) -- negative pre-activations become MINUS zero, which the tensor core reads as zero -- and the dgrad rounds
//     recover the mask from the sign bits of the resident h1 / h2 tiles (one 16-byte shared load per 8 columns, PRMT sign replicate,
//     AND-NOT).  The wgrad operand dh~ = dh^ * (s_i / S) is one HMUL2 by a power of two.
// Selected by iris_set_option("field_backward_impl", 2) (default); 1 = the first-generation kernel, 0 = the two-kernel mma.sync form.
#pragma once
#include <cuda.h>

#include "field_bwd_tc5.cuh"

#define BT6_EPI_THREADS 256
#define BT6_THREADS 288              // 8 epilogue warps + 1 MMA / TMA issue warp
#define BT6_SMEM_BYTES (5 * BT5_TILE_BYTES + 2 * 8192 + 2048 + 32 + 16 + 48)   // tiles | W1 W2 | W3 | 3 mbarriers (+pad) | tmem slot | per-warp maxima

#define TC5_LD32(r, taddr)                                                                                                                    \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22," \
                 "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"                                                                             \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),    \
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),       \
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),       \
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                                         \
                 : "r"(taddr))

// fp16x2(max(acc, -0)): relu with the sign of a negative pre-activation kept in the (numerically inert) sign bit of the zero
__device__ __forceinline__ uint32_t pack_relu_f16x2(float lo, float hi) {
    uint32_t r;
    asm("{ .reg .b32 t; cvt.rn.f16x2.f32 t, %1, %2; max.f16x2 %0, t, %3; }" : "=r"(r) : "f"(hi), "f"(lo), "r"(0x80008000u));
    return r;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// 0xFFFF for every half of h whose sign bit is clear (pre-activation not negative), 0 where it is set
__device__ __forceinline__ uint32_t relu_mask_f16x2(uint32_t h) {
    uint32_t neg;
    asm("prmt.b32 %0, %1, 0, 0xBB99;" : "=r"(neg) : "r"(h));       // selector nibbles 9 / B: replicate the MSB of byte 1 / byte 3
    return ~neg;
}
__device__ __forceinline__ uint32_t hmul2_u32(uint32_t a, __half2 b) {
    const __half2 r = __hmul2(*reinterpret_cast<const __half2 *>(&a), b);
    return *reinterpret_cast<const uint32_t *>(&r);
}
__device__ __forceinline__ uint64_t desc_add(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }     // start-address field: no carry out of its 14 bits
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{ .reg .pred q; elect.sync _|q, 0xffffffff; selp.u32 %0, 1, 0, q; }" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

// rows [row0, row0 + 128) of a row-major [n][64] fp16 array <-> a K-major tile: box kc = columns [8 kc, 8 kc + 8) = K chunk kc
__device__ __forceinline__ void tma_load_tile(uint32_t smem_dst, const CUtensorMap *map, int32_t row0, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)BT5_TILE_BYTES) : "memory");
#pragma unroll
    for (int kc = 0; kc < 8; ++kc)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_dst + kc * TC5_A_LBO), "l"(reinterpret_cast<uint64_t>(map)), "r"(8 * kc), "r"(row0), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_tile(const CUtensorMap *map, int32_t row0, uint32_t smem_src) {
#pragma unroll
    for (int kc = 0; kc < 8; ++kc)
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                     ::"l"(reinterpret_cast<uint64_t>(map)), "r"(8 * kc), "r"(row0), "r"(smem_src + kc * TC5_A_LBO) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <bool WS>
__global__ void __launch_bounds__(BT6_THREADS) k_field_backward_tc5v2(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dx,
                                                                      IrisShadeParams P, int64_t n, const float4 *__restrict__ r5,
                                                                      const float *__restrict__ d_mat, float *__restrict__ s_out, float *__restrict__ d_mlp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *sX = smem_raw, *sH1 = sX + BT5_TILE_BYTES, *sH2 = sH1 + BT5_TILE_BYTES, *sD = sH2 + BT5_TILE_BYTES, *sDw = sD + BT5_TILE_BYTES;
    unsigned char *sW1 = sDw + BT5_TILE_BYTES, *sW2 = sW1 + 8192, *sW3 = sW2 + 8192;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sW3 + 2048);                  // [0] accumulator ready, [1] X tile landed, [2] operands written
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 4);
    float *red = reinterpret_cast<float *>(tmem_slot + 4);                      // 9 per-warp maxima
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_mma_warp = warp == 8;
    const int quad = warp & 3, half = (warp >> 2) & 1;                          // TMEM lane quadrant, column half (epilogue warps)
    const int row = quad * 32 + lane;                                           // tile row of this thread (two epilogue threads per row)
    const __half *mlp = reinterpret_cast<const __half *>(P.mlp_f16);
    tc5_stage_weights(mlp, 64, TC5_W_LBO, sW1);
    tc5_stage_weights(mlp + 4096, 64, TC5_W_LBO, sW2);
    tc5_stage_weights(mlp + 8192, 16, TC5_W3_LBO, sW3);
    const int64_t n_tiles = (n + TC5_ROWS - 1) / TC5_ROWS;
    const uint32_t bar_mma = smem_u32(mbar), bar_x = smem_u32(mbar + 1), bar_full = smem_u32(mbar + 2);
    const uint32_t aX = smem_u32(sX), aH1 = smem_u32(sH1), aH2 = smem_u32(sH2), aD = smem_u32(sD), aDw = smem_u32(sDw);
    const uint32_t aW1 = smem_u32(sW1), aW2 = smem_u32(sW2), aW3 = smem_u32(sW3);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_mma) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_x) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_full), "r"(BT6_EPI_THREADS) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BT5_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- S: one power of two per CTA with |dy| <= max|d_mat| / 4 <= S for every sample this CTA will see
    float mx = 0.f;
    if (!is_mma_warp) {
        // (a streaming pass over the CTA's share of d_mat: unrolled so that several tiles' loads are in flight)
        const int k0 = half ? 3 : 0;
#pragma unroll 4
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t i = tile * TC5_ROWS + row;
            if (i < n) {
                const float v0 = fabsf(__ldg(d_mat + 5 * i + k0)), v1 = fabsf(__ldg(d_mat + 5 * i + k0 + 1)), v2 = half ? 0.f : fabsf(__ldg(d_mat + 5 * i + 2));
                if (v0 < __int_as_float(0x7f800000)) mx = fmaxf(mx, v0);
                if (v1 < __int_as_float(0x7f800000)) mx = fmaxf(mx, v1);
                if (v2 < __int_as_float(0x7f800000)) mx = fmaxf(mx, v2);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mx = fmaxf(fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])), fmaxf(fmaxf(red[4], red[5]), fmaxf(red[6], red[7])));
    float S = 1.f;
    if (mx > 0.f) {
        int e;
        frexpf(mx, &e);                                                          // mx < 2^e
        S = ldexpf(1.0f, e - 2);                                                 // |dy| <= mx / 4 < S
    }
    const float invS = 1.0f / S;
    const uint32_t tmem = *tmem_slot;

    if (is_mma_warp) {
        // ================================================================ MMA / TMA issue warp
        const uint32_t id_f64 = umma_idesc_f16(128, 64), id_f16 = umma_idesc_f16(128, 16);                 // forward
        const uint32_t id_d64 = umma_idesc_f16_major(128, 64, false, true);                                // dgrad: B MN-major
        const uint32_t id_w64 = umma_idesc_f16_major(64, 64, true, true), id_w16 = umma_idesc_f16_major(64, 16, true, true);   // wgrad
        // operand descriptors, built once: K-major activation tiles (step 2 chunks = K 16), K-major weights, MN-major views (step 256 B)
        const uint64_t kX = umma_desc(aX, TC5_A_LBO, TC5_SBO), kH1 = umma_desc(aH1, TC5_A_LBO, TC5_SBO), kH2 = umma_desc(aH2, TC5_A_LBO, TC5_SBO);
        const uint64_t kD = umma_desc(aD, TC5_A_LBO, TC5_SBO);
        const uint64_t kW1 = umma_desc(aW1, TC5_W_LBO, TC5_SBO), kW2 = umma_desc(aW2, TC5_W_LBO, TC5_SBO), kW3 = umma_desc(aW3, TC5_W3_LBO, TC5_SBO);
        const uint64_t mW1 = umma_desc_mn(aW1, TC5_W_LBO), mW2 = umma_desc_mn(aW2, TC5_W_LBO), mW3 = umma_desc_mn(aW3, TC5_W3_LBO);
        const uint64_t mX = umma_desc_mn(aX, TC5_A_LBO), mH1 = umma_desc_mn(aH1, TC5_A_LBO), mH2 = umma_desc_mn(aH2, TC5_A_LBO), mDw = umma_desc_mn(aDw, TC5_A_LBO);
        const bool leader = elect_one();
        if (leader && (int64_t)blockIdx.x < n_tiles) tma_load_tile(aX, &tm_x, (int32_t)(blockIdx.x * TC5_ROWS), bar_x);      // first X tile
        uint32_t ph_full = 0, ph_x = 0, ph_mma = 0;
        bool first = true;                                                       // first tile of this CTA: dW accumulators start from zero
        int64_t prev_row0 = -1;                                                  // tile whose dx^ sits in sD, not yet stored
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const uint32_t accw = first ? 0u : 1u;
#pragma unroll 1
            for (int round = 0; round < 6; ++round) {
                mbar_wait(bar_full, ph_full);                                    // the epilogue of the previous round has written its operands
                ph_full ^= 1u;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (leader) {
                    if (round == 0) {
                        if (prev_row0 >= 0) tma_store_tile(&tm_dx, (int32_t)prev_row0, aD);     // the previous tile's dx^ is complete in sD
                        mbar_wait(bar_x, ph_x);                                  // this tile's X has landed
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16(tmem, desc_add(kX, 2 * k * TC5_A_LBO), desc_add(kW1, 2 * k * TC5_W_LBO), id_f64, k > 0);
                    } else if (round == 1) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16(tmem, desc_add(kH1, 2 * k * TC5_A_LBO), desc_add(kW2, 2 * k * TC5_W_LBO), id_f64, k > 0);
                    } else if (round == 2) {
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // sD (previous dx^) has been read out: this round's epilogue rewrites it
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16(tmem, desc_add(kH2, 2 * k * TC5_A_LBO), desc_add(kW3, 2 * k * TC5_W3_LBO), id_f16, k > 0);
                    } else if (round == 3) {
                        // dh2 = dy^ W3 : A = sD chunks 0,1 (K = 16 outputs), B = W3 tile [16 out][64 in] read MN-major (chunk stride 256)
                        umma_f16(tmem, kD, mW3, id_d64, 0u);
                        // dW3^T (64 x 16) += h2^T dy~ : A = sH2 MN-major (M = 64 features), B = sDw chunks 0,1 MN-major (N = 16), K = 128 samples
#pragma unroll
                        for (int k = 0; k < 8; ++k) umma_f16(tmem + 192, desc_add(mH2, k * 256), desc_add(mDw, k * 256), id_w16, k > 0 ? 1u : accw);
                    } else if (round == 4) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16(tmem, desc_add(kD, 2 * k * TC5_A_LBO), desc_add(mW2, k * 256), id_d64, k > 0);
                        // dW2 (64 x 64) += dh2~^T h1
#pragma unroll
                        for (int k = 0; k < 8; ++k) umma_f16(tmem + 128, desc_add(mDw, k * 256), desc_add(mH1, k * 256), id_w64, k > 0 ? 1u : accw);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16(tmem, desc_add(kD, 2 * k * TC5_A_LBO), desc_add(mW1, k * 256), id_d64, k > 0);
                        // dW1 (64 x 64) += dh1~^T X
#pragma unroll
                        for (int k = 0; k < 8; ++k) umma_f16(tmem + 64, desc_add(mDw, k * 256), desc_add(mX, k * 256), id_w64, k > 0 ? 1u : accw);
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_mma) : "memory");
                    if (round == 5) {
                        // once every MMA that reads sX is done, request the next tile's X: it lands under the dx^ epilogue
                        const int64_t nxt = tile + gridDim.x;
                        if (nxt < n_tiles) {
                            mbar_wait(bar_mma, ph_mma);
                            tma_load_tile(aX, &tm_x, (int32_t)(nxt * TC5_ROWS), bar_x);
                        }
                    }
                }
                if (round == 0) ph_x ^= 1u;
                ph_mma ^= 1u;
                __syncwarp();
            }
            prev_row0 = tile * TC5_ROWS;
            first = false;
        }
        mbar_wait(bar_full, ph_full);                                            // the last dx^ tile is in sD
        if (leader) {
            if (prev_row0 >= 0) tma_store_tile(&tm_dx, (int32_t)prev_row0, aD);
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");            // ... and has left shared memory before the CTA exits
        }
    } else {
        // ================================================================ epilogue warps
        const uint32_t row_off = (row >> 3) * TC5_SBO + (row & 7) * 16;
        const uint32_t taddr0 = tmem + ((uint32_t)(quad * 32) << 16);            // this warp's lanes, column 0
        const uint32_t taddr = taddr0 + 32 * half;                               // ... this half's columns
        const uint32_t chunk0 = 4 * half;                                        // first 16-byte K chunk of this thread's columns
        uint32_t phase = 0;
        // d_mat row (and the "sample is live" flag) of the tile after the current one are fetched a tile ahead
        float dm_n[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        int code_n = -1;                                                         // raw "lane state" word of the record (-2 = live): compared a tile later
        auto fetch = [&](int64_t tile) {
            const int64_t i = tile * TC5_ROWS + row;
            code_n = -1;
#pragma unroll
            for (int k = 0; k < 5; ++k) dm_n[k] = 0.f;
            if (tile < n_tiles && i < n) {
                code_n = WS ? __float_as_int(__ldg(reinterpret_cast<const float *>(r5 + i) + 3)) : -2;
#pragma unroll
                for (int k = 0; k < 5; ++k) dm_n[k] = __ldg(d_mat + 5 * i + k);
            }
        };
        fetch(blockIdx.x);
        mbar_arrive(bar_full);                                                   // nothing precedes the first round
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t i = tile * TC5_ROWS + row;
            float dm[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) dm[k] = dm_n[k];
            const bool active = code_n == -2 && (dm[0] != 0.f || dm[1] != 0.f || dm[2] != 0.f || dm[3] != 0.f || dm[4] != 0.f);
            fetch(tile + gridDim.x);
            float sc = 0.f;
#pragma unroll 1
            for (int round = 0; round < 6; ++round) {
                mbar_wait(bar_mma, phase);
                phase ^= 1u;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (round <= 1) {
                    // ---- h = relu(acc) -> fp16 -> next layer's A tile
                    unsigned char *dst = (round == 0 ? sH1 : sH2) + row_off;
                    uint32_t r[32];
                    TC5_LD32(r, taddr);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 v;
                        v.x = pack_relu_f16x2(__uint_as_float(r[8 * c + 0]), __uint_as_float(r[8 * c + 1]));
                        v.y = pack_relu_f16x2(__uint_as_float(r[8 * c + 2]), __uint_as_float(r[8 * c + 3]));
                        v.z = pack_relu_f16x2(__uint_as_float(r[8 * c + 4]), __uint_as_float(r[8 * c + 5]));
                        v.w = pack_relu_f16x2(__uint_as_float(r[8 * c + 6]), __uint_as_float(r[8 * c + 7]));
                        *reinterpret_cast<uint4 *>(dst + (chunk0 + c) * TC5_A_LBO) = v;
                    }
                } else if (round == 2) {
                    // ---- dy = d_mat * d(mat)/dy, per-sample power-of-two normalisation (field.cuh, k_field_backward_dgrad).  Both threads
                    //      of a row compute the row's scale (each needs it for its half of dh~); the first one writes the operands.
                    uint32_t r[16];
                    TC5_LD16(r, taddr0);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
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
                    if (half == 0) {
                        const float inv = sc > 0.f ? 1.0f / sc : 0.f;
                        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
                        uint4 hd, hw;
                        hd.x = pack_f16x2(dy[0] * inv, dy[1] * inv); hd.y = pack_f16x2(dy[2] * inv, dy[3] * inv); hd.z = pack_f16x2(dy[4] * inv, 0.f); hd.w = 0u;
                        hw.x = pack_f16x2(dy[0] * invS, dy[1] * invS); hw.y = pack_f16x2(dy[2] * invS, dy[3] * invS); hw.z = pack_f16x2(dy[4] * invS, 0.f); hw.w = 0u;
                        *reinterpret_cast<uint4 *>(sD + row_off) = hd;
                        *reinterpret_cast<uint4 *>(sD + TC5_A_LBO + row_off) = zero;
                        *reinterpret_cast<uint4 *>(sDw + row_off) = hw;
                        *reinterpret_cast<uint4 *>(sDw + TC5_A_LBO + row_off) = zero;
                        if (i < n) s_out[i] = sc;
                    }
                } else if (round == 3 || round == 4) {
                    // ---- dh^ = acc . relu'(h) -> fp16 -> sD (next dgrad A operand) and, rescaled by s_i / S, -> sDw (wgrad operand)
                    const unsigned char *hsrc = (round == 3 ? sH2 : sH1) + row_off;
                    const __half2 ws2 = __float2half2_rn(sc * invS);
                    uint32_t r[32];
                    TC5_LD32(r, taddr);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint4 h = *reinterpret_cast<const uint4 *>(hsrc + (chunk0 + c) * TC5_A_LBO);
                        uint4 hd, hw;
                        hd.x = pack_f16x2(__uint_as_float(r[8 * c + 0]), __uint_as_float(r[8 * c + 1])) & relu_mask_f16x2(h.x);
                        hd.y = pack_f16x2(__uint_as_float(r[8 * c + 2]), __uint_as_float(r[8 * c + 3])) & relu_mask_f16x2(h.y);
                        hd.z = pack_f16x2(__uint_as_float(r[8 * c + 4]), __uint_as_float(r[8 * c + 5])) & relu_mask_f16x2(h.z);
                        hd.w = pack_f16x2(__uint_as_float(r[8 * c + 6]), __uint_as_float(r[8 * c + 7])) & relu_mask_f16x2(h.w);
                        hw.x = hmul2_u32(hd.x, ws2); hw.y = hmul2_u32(hd.y, ws2); hw.z = hmul2_u32(hd.z, ws2); hw.w = hmul2_u32(hd.w, ws2);
                        *reinterpret_cast<uint4 *>(sD + (chunk0 + c) * TC5_A_LBO + row_off) = hd;
                        *reinterpret_cast<uint4 *>(sDw + (chunk0 + c) * TC5_A_LBO + row_off) = hw;
                    }
                } else {
                    // ---- dx^ (normalised) -> fp16 -> sD in the tile layout; the issue warp sends it with one bulk tensor store group
                    uint32_t r[32];
                    TC5_LD32(r, taddr);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 v;
                        v.x = pack_f16x2(__uint_as_float(r[8 * c + 0]), __uint_as_float(r[8 * c + 1]));
                        v.y = pack_f16x2(__uint_as_float(r[8 * c + 2]), __uint_as_float(r[8 * c + 3]));
                        v.z = pack_f16x2(__uint_as_float(r[8 * c + 4]), __uint_as_float(r[8 * c + 5]));
                        v.w = pack_f16x2(__uint_as_float(r[8 * c + 6]), __uint_as_float(r[8 * c + 7]));
                        *reinterpret_cast<uint4 *>(sD + (chunk0 + c) * TC5_A_LBO + row_off) = v;
                    }
                }
                // operands of the next round are written, the accumulator has been read: hand over to the issue warp
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(bar_full);
            }
        }
        // ---- drain the weight-gradient accumulators (every MMA has completed: the last round's "accumulator ready" was waited for):
        //      M = 64 rows live in TMEM lanes 32 (m / 16) + m % 16
        if (half == 0 && (int64_t)blockIdx.x < n_tiles) {
            const int m = 16 * quad + lane;                                      // feature row held by this lane (lanes < 16)
            for (int blk = 0; blk < 3; ++blk) {                                  // dW1, dW2, dW3^T
                const int ncol = blk == 2 ? 16 : 64;
                for (int q = 0; q < ncol / 16; ++q) {
                    uint32_t r[16];
                    BT5_LD16(r, taddr0 + 64 * (blk + 1), q);                     // all lanes take part in the load (warp-collective)
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
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BT5_TMEM_COLS) : "memory");
}
