// ld_probe.cu -- how many random 32-byte sectors per second can one B200 pull from L2 into its SMs?  The hash-grid forward kernel
// (k_field_forward_tc5) misses L1 on 52 sectors per sample and runs at 3.2 G samples/s = 170 G sectors/s = 5.3 TB/s; this probe says how
// far that is from what the L1-miss path / L2 delivers for the same access pattern (one 4-byte or 8-byte word per random sector).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o iris_b200/_lib/ab/ld_probe tools/probe/ld_probe.cu && iris_b200/_lib/ab/ld_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// MLP independent loads in flight per thread and iteration; WORDS = 1: ld.b32, 2: ld.v2.b32
template <int MLP, int WORDS>
__global__ void __launch_bounds__(256) k_ld(const uint32_t *__restrict__ tab, uint32_t mask_sectors, int iters, uint32_t *sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = mix(tid + 1u), acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint32_t v[MLP][2];
#pragma unroll
        for (int k = 0; k < MLP; ++k) {
            s = s * 1664525u + 1013904223u;
            const uint32_t sec = mix(s) & mask_sectors;                  // one access per random 32-byte sector
            const uint32_t *p = tab + 8 * (size_t)sec + 2 * ((s >> 28) & 3u);
            if (WORDS == 1) asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v[k][0]) : "l"(p));
            else asm volatile("ld.global.nc.v2.b32 {%0, %1}, [%2];" : "=r"(v[k][0]), "=r"(v[k][1]) : "l"(p));
        }
#pragma unroll
        for (int k = 0; k < MLP; ++k) acc ^= v[k][0] ^ (WORDS == 2 ? v[k][1] : 0u);
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int MLP, int WORDS>
static void run(const uint32_t *tab, size_t bytes, int ctas_per_sm, uint32_t *sink) {
    const int iters = 2048 / MLP, blocks = 148 * ctas_per_sm, threads = 256;
    const uint32_t mask = (uint32_t)(bytes / 32) - 1u;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_ld<MLP, WORDS><<<blocks, threads>>>(tab, mask, 8, sink);
    cudaEventRecord(a);
    k_ld<MLP, WORDS><<<blocks, threads>>>(tab, mask, iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double n = (double)blocks * threads * iters * MLP;
    printf("table %6.1f MB  %d B/load  %2d loads in flight/thread  %2d CTAs/SM (%2d warps)  %7.3f ms  %6.1f G sectors/s = %5.2f TB/s of 32-byte sectors\n",
           bytes / 1e6, 4 * WORDS, MLP, ctas_per_sm, ctas_per_sm * 8, ms, n / ms / 1e6, n * 32 / ms / 1e9);
}

int main() {
    uint32_t *tab, *sink;
    const size_t big = (size_t)256 << 20;
    cudaMalloc(&tab, big);
    cudaMalloc(&sink, 64);
    cudaMemset(tab, 1, big);
    const size_t sizes[3] = {(size_t)2 << 20, (size_t)64 << 20, (size_t)256 << 20};    // one hashed level, the whole grid (L2-resident), beyond L2
    for (int k = 0; k < 3; ++k) {
        run<8, 1>(tab, sizes[k], 2, sink);
        run<8, 1>(tab, sizes[k], 8, sink);
        run<32, 1>(tab, sizes[k], 2, sink);
        run<32, 1>(tab, sizes[k], 8, sink);
        run<32, 2>(tab, sizes[k], 8, sink);
    }
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
