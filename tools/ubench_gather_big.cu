// tools/ubench_gather_big.cu -- the random-gather ceiling at the FOOTPRINT and in the SHAPE of the seed-match kernel:
//   (1) independent 16-byte gathers over tables of 8 .. 96 GB (does the rate of profiles/r01_ubench_gather_sweep.txt, measured up
//       to 16 GB, hold at the 80 GB of index the headline configuration touches -- TLB reach?), and
//   (2) DEPENDENT chains: every thread keeps exactly one gather in flight and needs the loaded value to form the next address
//       (what a backward search does), 2048 threads per SM, chain lengths 1 .. 6.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_gather_big tools/ubench_gather_big.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t hash64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }

__device__ __forceinline__ uint4 ld16(const uint4* p) {
    uint4 v; asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v;
}

template <int PER_THREAD>
__global__ void __launch_bounds__(256, 8) independent(const uint4* __restrict__ tab, uint64_t mask, uint32_t* out, uint64_t seed) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) acc += ld16(tab + (hash64(t * PER_THREAD + k + seed) & mask)).x;
    out[t] = acc;
}

__global__ void __launch_bounds__(256, 8) chained(const uint4* __restrict__ tab, uint64_t mask, uint32_t* out, uint64_t seed, int depth) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t idx = hash64(t + seed) & mask;
    uint32_t acc = 0;
    for (int k = 0; k < depth; ++k) {
        const uint4 v = ld16(tab + idx);
        acc += v.x;
        idx = hash64(idx + v.y + k) & mask;          // the next address needs the loaded value
    }
    out[t] = acc;
}

static float time_ms(void (*launch)(void*), void* ctx) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(ctx); float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); launch(ctx); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}

struct Ctx { const uint4* tab; uint64_t mask; uint32_t* out; int depth; uint64_t seed; };
static const uint32_t THREADS = 1u << 25;
static void l_ind(void* c_) { Ctx* c = (Ctx*)c_; independent<4><<<THREADS / 256, 256>>>(c->tab, c->mask, c->out, c->seed++); }
static void l_chn(void* c_) { Ctx* c = (Ctx*)c_; chained<<<THREADS / 256, 256>>>(c->tab, c->mask, c->out, c->seed++, c->depth); }

int main() {
    size_t free_b = 0, total_b = 0; cudaMemGetInfo(&free_b, &total_b);
    uint32_t* out; cudaMalloc(&out, (size_t)THREADS * 4);
    const uint64_t max_entries = 6ull << 30;                           // 6 G entries x 16 B = 96 GiB
    uint4* tab = nullptr;
    uint64_t entries = max_entries;
    while (entries >= (1ull << 29) && cudaMalloc(&tab, entries * 16) != cudaSuccess) { cudaGetLastError(); entries >>= 1; }
    if (!tab) { printf("allocation failed\n"); return 1; }
    cudaMemset(tab, 1, entries * 16);
    printf("table %.1f GiB (free %.1f of %.1f GiB)\n", entries * 16 / 1073741824.0, free_b / 1073741824.0, total_b / 1073741824.0);
    for (uint64_t e = 1ull << 29; e <= entries; e <<= 1) {             // 8 GiB, 16, 32, 64 (power-of-two masks; the last one may be 96 -> 64)
        Ctx c = { tab, e - 1, out, 0, 1 };
        const float ms = time_ms(l_ind, &c);
        printf("independent 16 B gathers, %5.0f GiB footprint: %7.3f ms  %6.1f G gathers/s\n", e * 16 / 1073741824.0, ms, THREADS * 4.0 / ms / 1e6);
    }
    // a 1.5x non-power-of-two footprint: top entries reached through a second range
    for (uint64_t e = 1ull << 29; e <= entries; e <<= 1)
        for (int depth = 1; depth <= 6; depth += (depth < 4 ? 1 : 2)) {
            Ctx c = { tab, e - 1, out, depth, 1 };
            const float ms = time_ms(l_chn, &c);
            printf("dependent chains, depth %d, 2048 threads/SM, %5.0f GiB footprint: %7.3f ms  %6.1f G gathers/s\n", depth, e * 16 / 1073741824.0, ms,
                   (double)THREADS * depth / ms / 1e6);
        }
    return 0;
}
