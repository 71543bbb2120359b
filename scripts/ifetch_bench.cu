// Instruction-fetch roofline of an SM for code that does not fit the per-scheduler L0 instruction cache:
// a loop whose body is N independent integer instructions (8 accumulators), N from 128 to 16384.
// Reports warp-instructions per cycle per SM for several block counts. nvcc -arch=sm_100a -O3 ifetch_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int N> __global__ void k(unsigned *out, int iters, unsigned seed) {
    unsigned a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 + 11, a5 = a0 + 13, a6 = a0 ^ 17, a7 = a0 ^ 19;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < N / 8; ++i) {
            asm volatile("xor.b32 %0, %0, %8; xor.b32 %1, %1, %8; xor.b32 %2, %2, %8; xor.b32 %3, %3, %8;"
                         "xor.b32 %4, %4, %8; xor.b32 %5, %5, %8; xor.b32 %6, %6, %8; xor.b32 %7, %7, %8;"
                         : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(seed), "r"(iters));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}
template <int N> void run(unsigned *d, int sms, double mhz) {
    for (int warps_per_sm : {4, 8, 16, 32}) {
        int iters = (1 << 21) / N;
        int blocks = sms * warps_per_sm / 4;
        k<N><<<blocks, 128>>>(d, 8, 1); cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<N><<<blocks, 128>>>(d, iters, 1); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double instr = (double)blocks * 4 * iters * N;
        printf("{\"body_instructions\": %d, \"warps_per_sm\": %d, \"ipc_per_sm\": %.2f, \"ms\": %.3f}\n", N, warps_per_sm,
               instr / (ms * 1e-3 * mhz * 1e6) / sms, ms);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    unsigned *d; cudaMalloc(&d, 1 << 24);
    double mhz = clk / 1000.0;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f}\n", p.name, p.multiProcessorCount, mhz);
    run<256>(d, p.multiProcessorCount, mhz); run<512>(d, p.multiProcessorCount, mhz); run<768>(d, p.multiProcessorCount, mhz);
    run<1024>(d, p.multiProcessorCount, mhz); run<1536>(d, p.multiProcessorCount, mhz); run<2048>(d, p.multiProcessorCount, mhz);
    run<4096>(d, p.multiProcessorCount, mhz); run<8192>(d, p.multiProcessorCount, mhz); run<16384>(d, p.multiProcessorCount, mhz);
    return 0;
}
