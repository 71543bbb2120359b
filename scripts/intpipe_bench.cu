// Integer-pipe peak microbenchmark for the extension roofline ("the builder must measure
// it with a DPX/IMAD microbench"): dependent-free streams of the instruction mix one DP cell of
// update_column needs (add / max / compare-select on int32), plus DPX three-input min/max.
// Prints one JSON line; run on the GPU box, the result is recorded in profiles/.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(int *out, int iters, int seed) {
    int a0 = threadIdx.x + seed, a1 = a0 * 3 + 1, a2 = a0 ^ 0x55, a3 = a0 + 7;
    int b0 = blockIdx.x + 11, b1 = b0 * 5, b2 = b0 ^ 0x33, b3 = b0 - 3;
    const int go = -6 + (seed & 1), ge = -2 - (seed & 1);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) {          // IADD3-class: pure adds
                a0 += b0; a1 += b1; a2 += b2; a3 += b3;
                b0 += a1; b1 += a2; b2 += a3; b3 += a0;
            } else if (MODE == 1) {   // add + max pairs (the S/E/F recurrence)
                a0 = max(a0 + go, b0 + ge); a1 = max(a1 + go, b1 + ge);
                a2 = max(a2 + go, b2 + ge); a3 = max(a3 + go, b3 + ge);
                b0 = max(b0 + ge, a1); b1 = max(b1 + ge, a2); b2 = max(b2 + ge, a3); b3 = max(b3 + ge, a0);
            } else {                  // DPX: fused add+max (VIADDMNMX) / three-way max
                a0 = __viaddmax_s32(a0, go, b0); a1 = __viaddmax_s32(a1, go, b1);
                a2 = __viaddmax_s32(a2, go, b2); a3 = __viaddmax_s32(a3, go, b3);
                b0 = __vimax3_s32(b0, a1, a2); b1 = __vimax3_s32(b1, a2, a3);
                b2 = __vimax3_s32(b2, a3, a0); b3 = __vimax3_s32(b3, a0, a1);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + b0 + b1 + b2 + b3;
}

template <int MODE> double run(int sms, int *out, double ops_per_iter) {
    const int iters = 4096, blocks = sms * 8, threads = 512;
    k<MODE><<<blocks, threads>>>(out, 16, 1);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, threads>>>(out, iters, r);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = (double)blocks * threads * iters * 8 * ops_per_iter;
    return ops / (best * 1e-3) / 1e12;     // Tera int32-ops/s
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int *out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 512 * 4);
    double t_add = run<0>(p.multiProcessorCount, out, 8);          // 8 adds per unrolled step
    double t_addmax = run<1>(p.multiProcessorCount, out, 8 * 3 - 4);   // 8 max + 12 adds -> 20 ops
    double t_dpx = run<2>(p.multiProcessorCount, out, 8 * 2);      // 8 fused ops = 16 int32 ops
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz\": %d, \"iadd_tops\": %.2f, \"add_max_tops\": %.2f, "
           "\"dpx_fused_tops\": %.2f, \"nominal_lanes_x_clock_tops\": %.2f}\n",
           p.name, p.multiProcessorCount, p.clockRate / 1000, t_add, t_addmax, t_dpx,
           p.multiProcessorCount * 128.0 * p.clockRate * 1e3 / 1e12);
    return 0;
}
