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
#ifdef WARP_UNIFORM   // the whole warp runs the statement, one elected lane issues: no per-thread serialisation loop around UTCHMMA
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
#define ISSUER (warp == 0)
#define TIMER (tid == 0)
#define COMMIT "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n"
#else
#define ISSUER (tid == 0)
#define TIMER true
#define COMMIT "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
#endif
// mode 0 fwd64, 1 fwd16, 2 dgrad, 3 wgrad64, 4 wgrad16
__global__ void __launch_bounds__(128) k_time(int mode, int batch, int reps, long long *out, int n_acc) {
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
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (ISSUER) {
        const int M = mode >= 3 ? 64 : 128, N = (mode == 1 || mode == 4) ? 16 : 64;
        const bool a_mn = mode >= 3, b_mn = mode >= 2;
        const uint32_t idesc = (1u << 4) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);
        const uint32_t lbo_b = N == 16 ? 256u : 1024u;
        uint32_t phase = 0;
        uint64_t dA[4], dB[4];                  // descriptors built once (as the kernels do): the loop body is the bare instruction
        for (int kk = 0; kk < 4; ++kk) {
            dA[kk] = a_mn ? desc(aA + kk * 256, SBO, LBO_A) : desc(aA + 2 * kk * LBO_A, LBO_A, SBO);
            dB[kk] = mode <= 1 ? desc(aB + 2 * kk * lbo_b, lbo_b, SBO) : mode == 2 ? desc(aB + kk * 256, SBO, lbo_b) : desc(aB + kk * 256, SBO, LBO_A);
        }
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
            if (batch == 1) {
                mma(tmem, dA[0], dB[0], idesc, 0u);
            } else if (n_acc > 1) {             // the same instructions spread over n_acc independent accumulators (64 columns apart)
                for (int k = 0; k < batch; k += 4) {
                    mma(tmem + 64 * (0 % n_acc), dA[0], dB[0], idesc, k > 0); mma(tmem + 64 * (1 % n_acc), dA[1], dB[1], idesc, k > 0);
                    mma(tmem + 64 * (2 % n_acc), dA[2], dB[2], idesc, k > 0 || n_acc < 3); mma(tmem + 64 * (3 % n_acc), dA[3], dB[3], idesc, k > 0 || n_acc < 4);
                }
            } else {
                mma(tmem, dA[0], dB[0], idesc, 0u); mma(tmem, dA[1], dB[1], idesc, 1u); mma(tmem, dA[2], dB[2], idesc, 1u); mma(tmem, dA[3], dB[3], idesc, 1u);
                for (int k = 4; k < batch; k += 4) {
                    mma(tmem, dA[0], dB[0], idesc, 1u); mma(tmem, dA[1], dB[1], idesc, 1u); mma(tmem, dA[2], dB[2], idesc, 1u); mma(tmem, dA[3], dB[3], idesc, 1u);
                }
            }
            asm volatile(COMMIT ::"r"(smem_u32(mbar)) : "memory");
            const long long t1 = clock64();
            uint32_t done = 0;
            for (uint32_t spins = 0; !done; ++spins) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(mbar)), "r"(phase) : "memory");
                if (spins > (1u << 22)) __trap();
            }
            phase ^= 1u;
            const long long t2 = clock64();
            if (blockIdx.x == 0 && TIMER) { out[2 * r] = t1 - t0; out[2 * r + 1] = t2 - t0; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}
int main() {
    const int reps = 64;
    long long *d, h[2 * 64];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(k_time, cudaFuncAttributeMaxDynamicSharedMemorySize, 33024);
    const char *names[5] = {"fwd64   M128 N64 K/K  ", "fwd16   M128 N16 K/K  ", "dgrad   M128 N64 K/MN ", "wgrad64 M64  N64 MN/MN", "wgrad16 M64  N16 MN/MN"};
    for (int mode = 0; mode < 5; ++mode)
      for (int n_acc : {1, 2, 4})
        for (int batch : {1, 4, 8, 64}) {
            if (n_acc > 1 && (batch == 1 || (mode != 0 && mode != 3))) continue;
            k_time<<<148, 128, 33024>>>(mode, batch, reps, d, n_acc);
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            std::vector<long long> iss, tot;
            for (int r = 8; r < reps; ++r) { iss.push_back(h[2 * r]); tot.push_back(h[2 * r + 1]); }
            std::sort(iss.begin(), iss.end()); std::sort(tot.begin(), tot.end());
            printf("%s %d accumulator(s) batch %2d : issue %5lld cycles, issue + commit -> barrier %5lld cycles (%6.1f per MMA)\n", names[mode], n_acc, batch, iss[iss.size() / 2], tot[tot.size() / 2],
                   (double)tot[tot.size() / 2] / batch);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
