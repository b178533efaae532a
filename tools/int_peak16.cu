// Issue-rate micro-benchmark of the packed 16-bit DPX / video instructions (two alignments per register):
// VIADDMNMX.S16x2, VIMNMX3.S16x2, VIADD.16x2, plus IMAD and PRMT.  Same method as tools/int_peak.cu.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/int_peak16 tools/int_peak16.cu && tools/int_peak16
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void chain(unsigned *out, int iters, unsigned seed) {
    unsigned a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = threadIdx.x * (2 * u + 1) + seed;
    const unsigned b = seed | 0x00010001u, c = seed ^ 0x12341234u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (KIND == 0) a[u] = __viaddmax_s16x2(a[u], b, c);
                else if (KIND == 1) a[u] = __vimax3_s16x2(a[u], b, a[(u + 1) & 7]);
                else if (KIND == 2) a[u] = __vadd2(a[u], b);
                else if (KIND == 3) a[u] = a[u] * b + c;
                else if (KIND == 4) a[u] = __byte_perm(a[u], b, c);
                else if (KIND == 5) a[u] = __viaddmax_s16x2_relu(a[u], b, c);
                else a[u] = __vimax_s16x2_relu(a[u], b);
            }
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += a[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
double run(int sms, unsigned *out) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    chain<KIND><<<sms * 2, 1024>>>(out, 100, 1);
    cudaEventRecord(e0);
    chain<KIND><<<sms * 2, 1024>>>(out, iters, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    return (double)sms * 2 * 1024 * iters * 64.0 / (ms * 1e-3);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    unsigned *out; cudaMalloc(&out, p.multiProcessorCount * 2 * 1024 * sizeof(unsigned));
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double d = p.multiProcessorCount * (clk * 1e3);
    const double r0 = run<0>(p.multiProcessorCount, out), r1 = run<1>(p.multiProcessorCount, out), r2 = run<2>(p.multiProcessorCount, out),
                 r3 = run<3>(p.multiProcessorCount, out), r4 = run<4>(p.multiProcessorCount, out), r5 = run<5>(p.multiProcessorCount, out),
                 r6 = run<6>(p.multiProcessorCount, out);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz_max\": %.0f, \"per_clk_per_sm\": {\"viaddmax_s16x2\": %.1f, \"vimax3_s16x2\": %.1f, \"vadd2\": %.1f, "
           "\"imad\": %.1f, \"prmt\": %.1f, \"viaddmax_s16x2_relu\": %.1f, \"vimax_s16x2_relu\": %.1f}, "
           "\"how\": \"8 independent chains per thread, 2 CTAs x 1024 threads per SM, 20000 x 64 instructions, CUDA events; warp instructions x 32 lanes per max SM clock\"}\n",
           p.name, p.multiProcessorCount, clk / 1e3, r0 / d, r1 / d, r2 / d, r3 / d, r4 / d, r5 / d, r6 / d);
    return 0;
}
