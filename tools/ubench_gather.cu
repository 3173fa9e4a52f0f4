// tools/ubench_gather.cu -- what can the memory system deliver for INDEPENDENT random 32-byte gathers
// (the FM-index access pattern without its dependency chain), and how many DRAM bytes does each one cost?
// Variants of the load instruction / L2 fetch granularity; run under
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum  to get DRAM bytes per gather.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_gather tools/ubench_gather.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct __align__(32) Blk { uint32_t w[8]; };

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

#define LD8(Q, b, p) asm volatile("ld.global.nc" Q ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" \
    : "=r"(b.w[0]), "=r"(b.w[1]), "=r"(b.w[2]), "=r"(b.w[3]), "=r"(b.w[4]), "=r"(b.w[5]), "=r"(b.w[6]), "=r"(b.w[7]) : "l"(p))

template <int VARIANT>
__device__ __forceinline__ Blk load(const Blk* p) {
    Blk b;
    if (VARIANT == 0) LD8("", b, p);
    if (VARIANT == 1) LD8(".L2::64B", b, p);
    if (VARIANT == 2) {
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.w[0]), "=r"(b.w[1]), "=r"(b.w[2]), "=r"(b.w[3]) : "l"(p));
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(b.w[4]), "=r"(b.w[5]), "=r"(b.w[6]), "=r"(b.w[7]) : "l"(p));
    }
    if (VARIANT == 3) LD8(".L1::no_allocate", b, p);
    if (VARIANT == 4) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.w[0]), "=r"(b.w[1]), "=r"(b.w[2]), "=r"(b.w[3]) : "l"(p));
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(b.w[4]), "=r"(b.w[5]), "=r"(b.w[6]), "=r"(b.w[7]) : "l"(p));
    }
    return b;
}

template <int VARIANT, int PER_THREAD>
__global__ void gather(const Blk* __restrict__ tab, uint32_t n_blocks_mask, uint32_t* out, uint32_t seed) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) {
        const uint32_t idx = hash32(t * PER_THREAD + k + seed) & n_blocks_mask;
        const Blk b = load<VARIANT>(tab + idx);
        acc += b.w[0] ^ b.w[3] ^ b.w[7];
    }
    out[t] = acc;
}

// address-ordered variant: thread t gathers from a window of 2^win_lg blocks whose position grows with t (what a batch of
// queries SORTED by SA row would do); win_lg = table size -> fully random
template <int PER_THREAD>
__global__ void gather_windowed(const Blk* __restrict__ tab, uint32_t n_blocks_lg, uint32_t win_lg, uint32_t* out, uint32_t seed) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;        // 2^25 threads
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) {
        const uint32_t h = hash32(t * PER_THREAD + k + seed);
        const uint32_t base = n_blocks_lg >= 25 ? (t << (n_blocks_lg - 25)) : (t >> (25 - n_blocks_lg));
        const uint32_t idx = ((base >> win_lg) << win_lg) | (h & ((1u << win_lg) - 1u));
        const Blk b = load<1>(tab + idx);
        acc += b.w[0] ^ b.w[3] ^ b.w[7];
    }
    out[t] = acc;
}
void run_windowed(const Blk* tab, uint32_t n_blocks_lg, uint32_t win_lg, uint32_t* out) {
    const uint32_t threads = 1u << 25;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather_windowed<4><<<threads / 256, 256>>>(tab, n_blocks_lg, win_lg, out, 1u);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        gather_windowed<4><<<threads / 256, 256>>>(tab, n_blocks_lg, win_lg, out, 77u + r);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("ordered .L2::64B 16 GiB table, window %8.3f MiB  %7.3f ms  %7.1f G gathers/s\n", (double)(1ull << win_lg) * 32 / 1048576.0, best,
           (double)threads * 4 / (best * 1e-3) / 1e9);
}

template <int VARIANT, int PER_THREAD>
void run(const Blk* tab, uint32_t mask, uint32_t* out, const char* what) {
    const uint32_t threads = 1u << 25;                    // 33.5 M threads
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<VARIANT, PER_THREAD><<<threads / 256, 256>>>(tab, mask, out, 1u);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        gather<VARIANT, PER_THREAD><<<threads / 256, 256>>>(tab, mask, out, 77u + r);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double gathers = (double)threads * PER_THREAD;
    printf("%-44s %7.3f ms  %7.1f G gathers/s  %8.1f GB/s useful (32 B each)\n", what, best,
           gathers / (best * 1e-3) / 1e9, gathers * 32.0 / (best * 1e-3) / 1e9);
}

int main() {
    const uint32_t big_blocks = 1u << 26;                 // 2 GiB table
    const uint32_t small_blocks = 1u << 21;               // 64 MiB table (L2 resident)
    Blk* tab; uint32_t* out;
    cudaMalloc(&tab, (size_t)big_blocks * sizeof(Blk));
    cudaMemset(tab, 1, (size_t)big_blocks * sizeof(Blk));
    cudaMalloc(&out, sizeof(uint32_t) * (1u << 25));
    size_t gran = 0; cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity = %zu\n", gran);
    run<0, 4>(tab, big_blocks - 1, out, "DRAM  v8 (LDG.256)");
    run<1, 4>(tab, big_blocks - 1, out, "DRAM  v8 .L2::64B");
    run<2, 4>(tab, big_blocks - 1, out, "DRAM  2 x v4 (LDG.128)");
    run<3, 4>(tab, big_blocks - 1, out, "DRAM  v8 .L1::no_allocate");
    run<4, 4>(tab, big_blocks - 1, out, "DRAM  2 x v4 .L1::no_allocate");
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity = %zu\n", gran);
    run<0, 4>(tab, big_blocks - 1, out, "DRAM  v8 (LDG.256), fetch granularity 32");
    run<2, 4>(tab, big_blocks - 1, out, "DRAM  2 x v4, fetch granularity 32");
    run<4, 4>(tab, big_blocks - 1, out, "DRAM  2 x v4 no_allocate, granularity 32");
    run<0, 8>(tab, small_blocks - 1, out, "L2    v8 (LDG.256), 64 MiB table");
    run<2, 8>(tab, small_blocks - 1, out, "L2    2 x v4, 64 MiB table");
    // table-size sweep (.L2::64B loads): where does the rate fall off -- L2 capacity (126 MB) or TLB reach?
    cudaFree(tab);
    Blk* big; const uint32_t sweep_blocks = 1u << 29;         // 16 GiB
    if (cudaMalloc(&big, (size_t)sweep_blocks * sizeof(Blk)) == cudaSuccess) {
        cudaMemset(big, 1, (size_t)sweep_blocks * sizeof(Blk));
        for (uint32_t lg = 21; lg <= 29; ++lg) {
            char what[64]; snprintf(what, sizeof(what), "sweep .L2::64B, %6u MiB table", (1u << lg) / 32768u);
            run<1, 4>(big, (1u << lg) - 1, out, what);
        }
        // the same 134 M gathers over the 16 GiB table, but address-ordered at decreasing window sizes
        for (uint32_t w = 29; w >= 5; w -= 3) run_windowed(big, 29, w, out);
    }
    return 0;
}
