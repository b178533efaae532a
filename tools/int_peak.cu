// Integer / DPX issue-rate micro-benchmark for the gapped-alignment roofline (not in MEASURED_PEAKS.json).
// Independent chains per thread of (a) __viaddmax_s32, (b) __vimax3_s32, (c) lop3+iadd mix, all SMs, 1024 threads per SM-slot.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/int_peak tools/int_peak.cu && tools/int_peak > profiles/int_peak.json
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void chain(int *out, int iters, int seed) {
    int a0 = threadIdx.x + seed, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3, a4 = a0 ^ 0x55, a5 = a0 ^ 0x33, a6 = a0 + 11, a7 = a0 + 13;
    const int b = seed | 1, c = seed ^ 0x1234;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (KIND == 0) { a0 = __viaddmax_s32(a0, b, c); a1 = __viaddmax_s32(a1, b, c); a2 = __viaddmax_s32(a2, b, c); a3 = __viaddmax_s32(a3, b, c);
                             a4 = __viaddmax_s32(a4, b, c); a5 = __viaddmax_s32(a5, b, c); a6 = __viaddmax_s32(a6, b, c); a7 = __viaddmax_s32(a7, b, c); }
            else if (KIND == 1) { a0 = __vimax3_s32(a0, b, a1); a1 = __vimax3_s32(a1, c, a2); a2 = __vimax3_s32(a2, b, a3); a3 = __vimax3_s32(a3, c, a4);
                                  a4 = __vimax3_s32(a4, b, a5); a5 = __vimax3_s32(a5, c, a6); a6 = __vimax3_s32(a6, b, a7); a7 = __vimax3_s32(a7, c, a0); }
            else { a0 = (a0 ^ b) + c; a1 = (a1 & b) + c; a2 = (a2 | b) - c; a3 = (a3 ^ c) + b; a4 = (a4 ^ b) + c; a5 = (a5 & c) + b; a6 = (a6 | c) - b; a7 = (a7 ^ c) + a0; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int KIND>
double run(int sms, int *out, double ops_per_iter) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    chain<KIND><<<sms * 2, 1024>>>(out, 100, 1);
    cudaEventRecord(e0);
    chain<KIND><<<sms * 2, 1024>>>(out, iters, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    return (double)sms * 2 * 1024 * iters * ops_per_iter / (ms * 1e-3);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int *out; cudaMalloc(&out, p.multiProcessorCount * 2 * 1024 * sizeof(int));
    const double viaddmax = run<0>(p.multiProcessorCount, out, 64), vimax3 = run<1>(p.multiProcessorCount, out, 64), mix = run<2>(p.multiProcessorCount, out, 128);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz_max\": %.0f, \"viaddmax_s32_ops_per_s\": %.4g, \"vimax3_s32_ops_per_s\": %.4g, \"lop_iadd_mix_ops_per_s\": %.4g, "
           "\"viaddmax_per_clk_per_sm\": %.1f, \"vimax3_per_clk_per_sm\": %.1f, \"lop_iadd_per_clk_per_sm\": %.1f, "
           "\"how\": \"8 independent chains per thread, 2 CTAs x 1024 threads per SM, 20000 x 8 unrolled iterations, CUDA events; per-clk figures use the max SM clock\"}\n",
           p.name, p.multiProcessorCount, clk / 1e3, viaddmax, vimax3, mix, viaddmax / (p.multiProcessorCount * (clk * 1e3)), vimax3 / (p.multiProcessorCount * (clk * 1e3)),
           mix / (p.multiProcessorCount * (clk * 1e3)));
    return 0;
}
