// tools/ubench_dpx.cu -- issue-rate microbenchmark of the integer/DPX instructions the banded Gotoh kernel
// is built from, on whatever GPU it runs on.  Prints warp-instructions per clock per SM for each op.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_dpx tools/ubench_dpx.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 8

#define ASM3(str) { uint32_t d; asm volatile(str : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c) {
    if (OP == 0) ASM3("{.reg .b32 t; add.s16x2 t, %1, %2; max.s16x2 %0, t, %3;}")          // VIADDMNMX.S16x2
    if (OP == 1) ASM3("max.relu.s16x2 %0, %1, %2;")                                        // VIMNMX.S16x2.RELU
    if (OP == 2) ASM3("add.s16x2 %0, %1, %2;")                                             // VIADD.16x2
    if (OP == 3) ASM3("prmt.b32 %0, %1, %2, %3;")                                          // PRMT
    if (OP == 4) ASM3("{.reg .b32 t; max.u16x2 t, %1, %2; max.u16x2 %0, t, %3;}")          // VIMNMX3.U16x2
    if (OP == 5) ASM3("mad.lo.u32 %0, %1, 32, %2;")                                        // LEA / IMAD.SHL
    if (OP == 6) ASM3("mad.lo.u32 %0, %1, %3, %2;")                                        // IMAD, register multiplier
    if (OP == 7) ASM3("{.reg .s32 t; add.s32 t, %1, %2; max.s32 %0, t, %3;}")              // VIADDMNMX (s32)
    if (OP == 8) ASM3("add.u32 %0, %1, %2;")                                               // IADD3
    if (OP == 9) ASM3("max.s32 %0, %1, %2;")                                               // VIMNMX (s32)
    ASM3("xor.b32 %0, %1, %2;")                                                            // LOP3
}

template <int OP>
__global__ void bench(uint32_t* out, uint32_t seed, unsigned long long* cycles) {
    uint32_t r[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) r[k] = seed + threadIdx.x * 7 + k;
    uint32_t b = seed | 1, c = seed * 3 + 5;
    const unsigned long long t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) r[k] = op<OP>(r[k], b, c);
    }
    const unsigned long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) acc ^= r[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int sms) {
    const int threads = 1024, blocks = sms;                  // 32 warps per SM
    uint32_t* out; unsigned long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * threads * blocks);
    cudaMalloc(&cyc, sizeof(unsigned long long) * blocks);
    bench<OP><<<blocks, threads>>>(out, 12345u, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<blocks, threads>>>(out, 12345u, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[1024]; cudaMemcpy(h, cyc, sizeof(unsigned long long) * blocks, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < blocks; ++i) mean += (double)h[i]; mean /= blocks;
    const double warp_instr_per_sm = (double)ITERS * UNROLL * (threads / 32);
    printf("%-28s %7.3f warp-instr/clk/SM   (%.0f clk/SM, %.3f ms, %.2f GHz effective)\n", name, warp_instr_per_sm / mean, mean, ms,
           mean / (ms * 1e6));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device: %s, %d SMs\n", p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    run<0>("VIADDMNMX.S16x2", sms);
    run<1>("VIMNMX.S16x2.RELU", sms);
    run<2>("VIADD.16x2", sms);
    run<3>("PRMT", sms);
    run<4>("VIMNMX3.U16x2", sms);
    run<5>("LEA (a*32+b)", sms);
    run<6>("IMAD (reg multiplier)", sms);
    run<7>("VIADDMNMX (s32)", sms);
    run<8>("IADD3", sms);
    run<9>("VIMNMX (s32)", sms);
    run<10>("LOP3 (xor)", sms);
    return 0;
}
