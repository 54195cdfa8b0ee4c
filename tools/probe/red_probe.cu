// red_probe.cu -- how fast does one B200 retire fp32 reductions to global memory?  Decides what the grid-gradient scatter
// (k_field_backward_scatter) can still gain: lane-ops/s for scalar / 8-byte / 16-byte `red.global.add.f32`, for spread and
// clustered addresses, for partially active warps, against shared-memory atomics.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/red_probe tools/probe/red_probe.cu && gpurun_out/red_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ void red1(float *p, float a) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory"); }
__device__ __forceinline__ void red2(float *p, float a, float b) { asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory"); }
__device__ __forceinline__ void red4(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// mode 0 scalar spread | 1 v2 spread | 2 v4 spread | 3 v2, lanes 2k/2k+1 in one 16-byte slot | 4 v2 all lanes one address
// 5 v2 spread, 8 lanes of 32 active | 6 v2, lanes 4k..4k+3 in one 32-byte sector | 7 v4 spread, 16 lanes active
// 8 shared-memory atomicAdd spread over 32 KB | 9 v2 spread in a 128-byte line per warp (32 lanes -> 16 slots)
template <int MODE>
__global__ void k_red(float *buf, uint32_t mask_entries, int iters, float *sink) {
    __shared__ float sm[8192];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    if (MODE == 8) { for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 0.f; __syncthreads(); }
    uint32_t s = mix(tid + 1u);
    for (int it = 0; it < iters; ++it) {
        s = s * 1664525u + 1013904223u;
        uint32_t e = mix(s) & mask_entries;          // entry = 8 bytes (2 floats)
        if (MODE == 0) red1(buf + 2 * (size_t)e, 1.f);
        if (MODE == 1) red2(buf + 2 * (size_t)e, 1.f, 2.f);
        if (MODE == 2) red4(buf + 2 * (size_t)(e & ~1u), 1.f, 2.f, 3.f, 4.f);
        if (MODE == 3) { uint32_t ee = (__shfl_sync(0xffffffffu, e, lane & ~1u) & ~1u) | (lane & 1u); red2(buf + 2 * (size_t)ee, 1.f, 2.f); }
        if (MODE == 4) { uint32_t ee = __shfl_sync(0xffffffffu, e, 0); red2(buf + 2 * (size_t)ee, 1.f, 2.f); }
        if (MODE == 5) { if ((lane & 3u) == 0u) red2(buf + 2 * (size_t)e, 1.f, 2.f); }
        if (MODE == 6) { uint32_t ee = (__shfl_sync(0xffffffffu, e, lane & ~3u) & ~3u) | (lane & 3u); red2(buf + 2 * (size_t)ee, 1.f, 2.f); }
        if (MODE == 7) { if ((lane & 1u) == 0u) red4(buf + 2 * (size_t)(e & ~1u), 1.f, 2.f, 3.f, 4.f); }
        if (MODE == 8) atomicAdd(&sm[e & 8191u], 1.f);
        if (MODE == 9) { uint32_t ee = (__shfl_sync(0xffffffffu, e, 0) & ~15u) | (mix(s ^ lane) & 15u); red2(buf + 2 * (size_t)ee, 1.f, 2.f); }
    }
    if (MODE == 8 && sink) { __syncthreads(); if (threadIdx.x == 0) sink[blockIdx.x] = sm[0]; }
}

template <int MODE>
static void run(const char *name, float *buf, uint32_t entries, float lanes_frac, float *sink) {
    const int iters = 256, blocks = 148 * 16, threads = 256;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_red<MODE><<<blocks, threads>>>(buf, entries - 1u, 16, sink);
    cudaEventRecord(a);
    k_red<MODE><<<blocks, threads>>>(buf, entries - 1u, iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double ops = (double)blocks * threads * iters * lanes_frac;
    printf("%-58s table %6.1f MB  %8.3f ms  %8.1f G lane-ops/s  %6.2f cyc/lane-op/SM @1.965GHz\n", name, entries * 8.0 / 1e6, ms, ops / ms / 1e6,
           148.0 * 1.965e9 / (ops / (ms * 1e-3)));
}

int main() {
    float *buf, *sink;
    const uint32_t big = 1u << 24;   // 16M entries x 8 B = 128 MB
    cudaMalloc(&buf, (size_t)big * 8);
    cudaMalloc(&sink, 1 << 20);
    cudaMemset(buf, 0, (size_t)big * 8);
    const uint32_t sizes[3] = {1u << 19, 1u << 22 | 0u, 1u << 24};   // 4 MB (one hashed level), 32 MB, 128 MB
    for (int k = 0; k < 3; ++k) {
        const uint32_t n = sizes[k];
        run<0>("scalar f32, spread", buf, n, 1.f, sink);
        run<1>("v2 (8 B), spread", buf, n, 1.f, sink);
        run<2>("v4 (16 B), spread", buf, n, 1.f, sink);
        run<3>("v2, lane pairs share a 16-byte slot", buf, n, 1.f, sink);
        run<6>("v2, lane quads share a 32-byte sector", buf, n, 1.f, sink);
        run<9>("v2, warp inside one 128-byte line", buf, n, 1.f, sink);
        run<4>("v2, whole warp one address", buf, n, 1.f, sink);
        run<5>("v2 spread, 8 of 32 lanes active", buf, n, 0.25f, sink);
        run<7>("v4 spread, 16 of 32 lanes active", buf, n, 0.5f, sink);
    }
    run<8>("shared-memory atomicAdd f32, spread over 32 KB", buf, 8192, 1.f, sink);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
