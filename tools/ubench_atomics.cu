// tools/ubench_atomics.cu -- micro-benchmark behind the statRead k-mer design (statistics kernel):
// what does one random increment of a 4^8-entry histogram cost on a B200, per memory space?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/ubench_atomics tools/ubench_atomics.cu
// Prints G increments/s for: 64-bit / 32-bit global reductions (REDG) on a 64 K-entry table, shared-memory atomics
// (ATOMS) on a 128 KB per-CTA table, a random 8-byte global load (the first-seen stamp check), a random shared-memory
// bitmap test, warp-private LDS+IADD+STS, and REDG + ATOMS interleaved.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t &x) { x = x * 1664525u + 1013904223u; return x >> 16; }

template <int MODE>
__global__ void __launch_bounds__(1024) bench(unsigned long long *t64, uint32_t *t32, int iters, uint32_t *sink) {
    extern __shared__ uint32_t sm[];
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    constexpr bool USES_SMEM = MODE == 2 || MODE == 3 || (MODE >= 5 && MODE <= 8) || MODE == 10;
    if (USES_SMEM) { for (int i = threadIdx.x; i < 32768; i += blockDim.x) sm[i] = 0; __syncthreads(); }
    uint32_t acc = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        const uint32_t idx = lcg(x);
        if (MODE == 0) atomicAdd(&t64[idx], 1ULL);
        else if (MODE == 1) atomicAdd(&t32[idx], 1u);
        else if (MODE == 2) atomicAdd(&sm[idx >> 1], 1u);
        else if (MODE == 3) atomicAdd(&sm[idx >> 1], (idx & 1u) ? 65536u : 1u);
        else if (MODE == 4) acc += (uint32_t)__ldcg(&t64[idx]);
        else if (MODE == 5) acc += (sm[idx >> 5] >> (idx & 31u)) & 1u;
        else if (MODE == 6) { uint32_t *p = &sm[warp * 1024 + (idx & 31u) * 32u + lane]; *p += 1u; }          // private, conflict-free
        else if (MODE == 7) { if (i & 1) atomicAdd(&t64[idx], 1ULL); else atomicAdd(&sm[idx >> 1], 1u); }
        else if (MODE == 8) { atomicAdd(&t64[idx], 1ULL); acc += (sm[idx >> 5] >> (idx & 31u)) & 1u; }           // RED + bitmap test
        else if (MODE == 10) {      // packed 16-bit halves, returning atomic, spill of 0x4000 at the crossing (the design of stat kernel v3)
            const uint32_t sh = (idx & 1u) << 4;
            const uint32_t old = atomicAdd(&sm[idx >> 1], 1u << sh);
            if (((old >> sh) & 0xFFFFu) == 0x3FFFu) { atomicSub(&sm[idx >> 1], 0x4000u << sh); atomicAdd(&t64[idx], 0x4000ULL); }
        }
        else if (MODE == 9) { atomicAdd(&t64[idx], 1ULL); acc += (uint32_t)__ldcg(&t64[65536 + idx]); }          // RED + stamp load
    }
    if (USES_SMEM) { __syncthreads(); for (int i = threadIdx.x; i < 32768; i += blockDim.x) acc += sm[i]; }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int MODE> int run(const char *name, unsigned long long *t64, uint32_t *t32, uint32_t *sink, int sms, int ctas_per_sm, int threads, size_t smem) {
    CK(cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 4096;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int grid = sms * ctas_per_sm;
    bench<MODE><<<grid, threads, smem>>>(t64, t32, iters, sink);
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(a));
        bench<MODE><<<grid, threads, smem>>>(t64, t32, iters, sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    const double ops = (double)grid * threads * iters;
    printf("{\"bench\": \"%s\", \"grid\": %d, \"threads\": %d, \"ms\": %.4f, \"G_ops_per_s\": %.2f}\n", name, grid, threads, best, ops / best / 1e6);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    unsigned long long *t64; uint32_t *t32, *sink;
    CK(cudaMalloc(&t64, 2 * 65536 * 8)); CK(cudaMalloc(&t32, 65536 * 4)); CK(cudaMalloc(&sink, 64));
    CK(cudaMemset(t64, 0, 2 * 65536 * 8)); CK(cudaMemset(t32, 0, 65536 * 4));
    const size_t S = 128 * 1024;
    if (run<0>("redg64_random_64k", t64, t32, sink, sms, 4, 512, 0)) return 1;
    if (run<1>("redg32_random_64k", t64, t32, sink, sms, 4, 512, 0)) return 1;
    if (run<2>("atoms32_random_128KB", t64, t32, sink, sms, 1, 512, S)) return 1;
    if (run<3>("atoms_packed16_random_128KB", t64, t32, sink, sms, 1, 512, S)) return 1;
    if (run<4>("ldg64_random_512KB", t64, t32, sink, sms, 4, 512, 0)) return 1;
    if (run<5>("lds_bitmap_random_8KB", t64, t32, sink, sms, 1, 512, S)) return 1;
    if (run<6>("lds_add_sts_private", t64, t32, sink, sms, 1, 512, S)) return 1;
    if (run<7>("redg64_atoms_interleaved", t64, t32, sink, sms, 1, 512, S)) return 1;
    if (run<8>("redg64_plus_bitmap_test", t64, t32, sink, sms, 1, 512, S)) return 1;
    if (run<9>("redg64_plus_stamp_load", t64, t32, sink, sms, 4, 512, 0)) return 1;
    if (run<10>("atoms_packed16_returning_spill", t64, t32, sink, sms, 1, 512, S)) return 1;
    if (run<10>("atoms_packed16_returning_spill_1024thr", t64, t32, sink, sms, 1, 1024, S)) return 1;
    if (run<0>("redg64_random_64k_1cta", t64, t32, sink, sms, 1, 512, 0)) return 1;
    return 0;
}
