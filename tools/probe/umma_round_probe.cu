// umma_round_probe.cu -- what does ONE dependent round of the fused field adjoint cost end to end?  128 epilogue threads + 1 issue warp:
//   issue warp : wait "operands written" -> tcgen05.fence -> 4 x tcgen05.mma (M128 N64 K16) -> commit
//   epilogue   : wait "accumulator ready" -> tcgen05.ld x32 (x2 halves) -> pack to fp16 -> st.shared (A operand of the next round)
//                -> [fence.proxy.async] -> tcgen05.fence -> arrive
// Variants: 0 = full round; 1 = no st.shared / no proxy fence (operands static); 2 = st.shared but no proxy fence (WRONG in general,
// timing only); 3 = no tcgen05.ld either (pure barrier ping-pong + MMAs).  Prints cycles per round.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o iris_b200/_lib/ab/umma_round_probe tools/probe/umma_round_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define LBO_A 2048
#define SBO 128
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spins > (1u << 22)) __trap();
    }
}
__global__ void __launch_bounds__(160) k_round(int variant, int rounds, long long *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sA = smem, *sB = smem + 16384;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + 32768);      // [0] accumulator ready, [1] operands written
    uint32_t *slot = reinterpret_cast<uint32_t *>(mbar + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 32768 / 4; i += 160) reinterpret_cast<uint32_t *>(smem)[i] = 0x2C002C00u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(smem_u32(mbar + 1)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot, bar_mma = smem_u32(mbar), bar_full = smem_u32(mbar + 1);
    if (warp == 4) {
        const uint32_t idesc = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        uint32_t ph = 0;
        const long long t0 = clock64();
        for (int r = 0; r < rounds; ++r) {
            mbar_wait(bar_full, ph);
            ph ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                for (int k = 0; k < 4; ++k) {
                    const uint64_t da = desc(smem_u32(sA) + 2 * k * LBO_A, LBO_A, SBO), db = desc(smem_u32(sB) + 2 * k * 1024, 1024, SBO);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(k > 0)) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_mma) : "memory");
            }
            __syncwarp();
        }
        mbar_wait(bar_full, ph);
        const long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) out[variant] = (t1 - t0) / rounds;
    } else {
        const uint32_t row_off = (tid >> 3) * SBO + (tid & 7) * 16;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        uint32_t ph = 0;
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_full) : "memory");
        for (int r = 0; r < rounds; ++r) {
            mbar_wait(bar_mma, ph);
            ph ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (variant != 3) {
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t v[32];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                                   "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                                   "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                 : "r"(taddr + 32 * hh));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (variant == 0 || variant == 2) {
                        for (int c = 0; c < 4; ++c) {
                            uint4 q;
                            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(q.x) : "f"(__uint_as_float(v[8 * c + 1]) * 1e-3f), "f"(__uint_as_float(v[8 * c]) * 1e-3f));
                            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(q.y) : "f"(__uint_as_float(v[8 * c + 3]) * 1e-3f), "f"(__uint_as_float(v[8 * c + 2]) * 1e-3f));
                            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(q.z) : "f"(__uint_as_float(v[8 * c + 5]) * 1e-3f), "f"(__uint_as_float(v[8 * c + 4]) * 1e-3f));
                            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(q.w) : "f"(__uint_as_float(v[8 * c + 7]) * 1e-3f), "f"(__uint_as_float(v[8 * c + 6]) * 1e-3f));
                            *reinterpret_cast<uint4 *>(sA + (4 * hh + c) * LBO_A + row_off) = q;
                        }
                    }
                }
            }
            if (variant == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_full) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}
int main() {
    long long *d, h[4];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(k_round, cudaFuncAttributeMaxDynamicSharedMemorySize, 33024);
    const char *names[4] = {"full round (ld, pack, st.shared, fence.proxy.async)", "tcgen05.ld only (static operands, no proxy fence)", "ld + st.shared, NO proxy fence (timing only)", "barrier ping-pong + MMAs only"};
    for (int v = 0; v < 4; ++v) k_round<<<148, 160, 33024>>>(v, 2000, d);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int v = 0; v < 4; ++v) printf("%-55s %5lld cycles per round\n", names[v], h[v]);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
