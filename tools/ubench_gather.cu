// tools/ubench_gather.cu -- what can the memory system deliver for INDEPENDENT random 32-byte gathers?
// (the FM-index access pattern without its dependency chain).  Prints useful GB/s (32 B per gather) for a
// table far larger than L2 and for one that fits in L2; pair with ncu for the DRAM-side byte count.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_gather tools/ubench_gather.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct __align__(32) Blk { uint32_t w[8]; };

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int PER_THREAD>
__global__ void gather(const Blk* __restrict__ tab, uint32_t n_blocks_mask, uint32_t* out, uint32_t seed) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) {
        const uint32_t idx = hash32(t * PER_THREAD + k + seed) & n_blocks_mask;
        Blk b;
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(b.w[0]), "=r"(b.w[1]), "=r"(b.w[2]), "=r"(b.w[3]), "=r"(b.w[4]), "=r"(b.w[5]), "=r"(b.w[6]), "=r"(b.w[7]) : "l"(tab + idx));
        acc += b.w[0] ^ b.w[3] ^ b.w[7];
    }
    out[t] = acc;
}

template <int PER_THREAD>
void run(const Blk* tab, uint32_t mask, uint32_t* out, const char* what) {
    const uint32_t threads = 1u << 25;                    // 33.5 M threads
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<PER_THREAD><<<threads / 256, 256>>>(tab, mask, out, 1u);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        gather<PER_THREAD><<<threads / 256, 256>>>(tab, mask, out, 77u + r);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double gathers = (double)threads * PER_THREAD;
    printf("%-34s %d gathers/thread: %7.3f ms  %8.1f G gathers/s  %8.1f GB/s useful (32 B each)\n", what, PER_THREAD, best,
           gathers / (best * 1e-3) / 1e9, gathers * 32.0 / (best * 1e-3) / 1e9);
}

int main() {
    const uint32_t big_blocks = 1u << 26;                 // 2 GiB table
    const uint32_t small_blocks = 1u << 21;               // 64 MiB table (L2 resident)
    Blk* tab; uint32_t* out;
    cudaMalloc(&tab, (size_t)big_blocks * sizeof(Blk));
    cudaMemset(tab, 1, (size_t)big_blocks * sizeof(Blk));
    cudaMalloc(&out, sizeof(uint32_t) * (1u << 25));
    run<1>(tab, big_blocks - 1, out, "2 GiB table (DRAM)");
    run<2>(tab, big_blocks - 1, out, "2 GiB table (DRAM)");
    run<4>(tab, big_blocks - 1, out, "2 GiB table (DRAM)");
    run<8>(tab, big_blocks - 1, out, "2 GiB table (DRAM)");
    run<4>(tab, small_blocks - 1, out, "64 MiB table (L2)");
    run<8>(tab, small_blocks - 1, out, "64 MiB table (L2)");
    return 0;
}
