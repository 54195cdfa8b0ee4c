// umma_time_probe.cu -- how long do the tcgen05.mma shapes of the fused field adjoint take?  One CTA per SM; one thread issues a batch of
// MMAs (accumulating into one TMEM tile), commits to an mbarrier and waits; clock64 around the batch, median over repetitions.
// Shapes (all kind::f16, K = 16 per instruction, operands in the no-swizzle core-matrix layout of field_tc5.cuh):
//   fwd64   M128 N64, A K-major, B K-major          (forward layers 1, 2)
//   fwd16   M128 N16, A K-major, B K-major          (forward layer 3)
//   dgrad   M128 N64, A K-major, B MN-major         (dh W read without a transposed copy)
//   wgrad64 M64  N64, A MN-major, B MN-major        (dW1, dW2: K = samples)
//   wgrad16 M64  N16, A MN-major, B MN-major        (dW3^T)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o iris_b200/_lib/ab/umma_time_probe tools/probe/umma_time_probe.cu
#include <cstdio>
#include <cstdint>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>
#define LBO_A 2048
#define SBO 128
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// mode 0 fwd64, 1 fwd16, 2 dgrad, 3 wgrad64, 4 wgrad16
__global__ void __launch_bounds__(128) k_time(int mode, int batch, int reps, long long *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sA = smem, *sB = smem + 16384;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + 32768);
    uint32_t *slot = reinterpret_cast<uint32_t *>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3C003C00u;       // fp16 ones
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
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
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const int M = mode >= 3 ? 64 : 128, N = (mode == 1 || mode == 4) ? 16 : 64;
        const bool a_mn = mode >= 3, b_mn = mode >= 2;
        const uint32_t idesc = (1u << 4) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);
        const uint32_t lbo_b = N == 16 ? 256u : 1024u;
        uint32_t phase = 0;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
            for (int k = 0; k < batch; ++k) {
                const int kk = k & 3;           // stay inside the 16 KB tiles
                uint64_t da, db;
                if (a_mn) da = desc(aA + kk * 256, SBO, LBO_A); else da = desc(aA + 2 * kk * LBO_A, LBO_A, SBO);
                if (mode <= 1) db = desc(aB + 2 * kk * lbo_b, lbo_b, SBO);          // weights K-major
                else if (mode == 2) db = desc(aB + kk * 256, SBO, lbo_b);           // weights MN-major
                else db = desc(aB + kk * 256, SBO, LBO_A);                          // activation tile MN-major
                mma(tmem, da, db, idesc, k > 0);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
            const long long t1 = clock64();
            uint32_t done = 0;
            for (uint32_t spins = 0; !done; ++spins) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(mbar)), "r"(phase) : "memory");
                if (spins > (1u << 22)) __trap();
            }
            phase ^= 1u;
            const long long t2 = clock64();
            if (blockIdx.x == 0) { out[2 * r] = t1 - t0; out[2 * r + 1] = t2 - t0; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}
int main() {
    const int reps = 64;
    long long *d, h[2 * 64];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(k_time, cudaFuncAttributeMaxDynamicSharedMemorySize, 33024);
    const char *names[5] = {"fwd64   M128 N64 K/K  ", "fwd16   M128 N16 K/K  ", "dgrad   M128 N64 K/MN ", "wgrad64 M64  N64 MN/MN", "wgrad16 M64  N16 MN/MN"};
    for (int mode = 0; mode < 5; ++mode)
        for (int batch : {1, 4, 8, 64}) {
            k_time<<<148, 128, 33024>>>(mode, batch, reps, d);
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            std::vector<long long> iss, tot;
            for (int r = 8; r < reps; ++r) { iss.push_back(h[2 * r]); tot.push_back(h[2 * r + 1]); }
            std::sort(iss.begin(), iss.end()); std::sort(tot.begin(), tot.end());
            printf("%s batch %2d : issue %5lld cycles, issue + commit -> barrier %5lld cycles (%6.1f per MMA)\n", names[mode], batch, iss[iss.size() / 2], tot[tot.size() / 2],
                   (double)tot[tot.size() / 2] / batch);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
