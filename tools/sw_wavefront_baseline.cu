// Baseline kept on file for the design choice DESIGN.md section 4 makes: the north-star text describes the gapped aligner as
// ONE WARP per (read, primer) pair walking anti-diagonal wavefronts; the product uses ONE THREAD per pair with register-resident
// rows.  This tool measures the warp-per-pair form (score-only Glocal, fgbio's three matrices, DPX intrinsics, lanes = query
// rows, neighbours through warp shuffles, target staged in shared memory) on the product's problem shapes so that the two can
// be compared:   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/sw_wavefront_baseline tools/sw_wavefront_baseline.cu
//                tools/sw_wavefront_baseline > profiles/r02_sw_wavefront_baseline.json
// Not part of the library; nothing here is shipped.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

constexpr int kNeg = -(1 << 28);

// one warp per alignment; lane i owns query row i (n <= 32).  Step d: lane i computes cell (i, j = d - i).
// D = Diagonal, L = Left, U = Up as in oracle/idp_oracle.c; first row / column: all zero row 0 (free leading target gap),
// Up(i, 0) = open + i * extend.  Glocal end cell: best over the last row.
__global__ void __launch_bounds__(128) wavefront_kernel(const unsigned char *__restrict__ q, const unsigned char *__restrict__ t, int n, int m, int n_pairs,
                                                        int ma, int mi, int go, int ge, int *__restrict__ score) {
    extern __shared__ unsigned char tgt_all[];  // the targets of this block's warps, staged once
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, warps = blockDim.x >> 5;
    unsigned char *tgt = tgt_all + wib * ((m + 3) & ~3);
    const int oe = go + ge;
    for (int p = blockIdx.x * warps + wib; p < n_pairs; p += gridDim.x * warps) {
        for (int j = lane; j < m; j += 32) tgt[j] = t[(size_t)p * m + j];
        __syncwarp();
        const int qc = lane < n ? q[(size_t)p * n + lane] : 255;
        // this lane's row r = lane + 1 (1-based); column 0: only Up is real, Up(r, 0) = open + r * extend
        int Dc = kNeg, Lc = kNeg, Uc = go + (lane + 1) * ge, Mc = Uc;  // matrices at the last column this lane computed
        int Mprev = Mc;                                                 // M one column before that: the diagonal source of the row below
        int best = kNeg;
        for (int d = 0; d < n + m - 1; ++d) {
            const int j = d - lane;  // 0-based target column of this lane's cell in this step
            // row r-1: cell (r-1, j) was computed by lane-1 in the previous step, M(r-1, j-1) one step earlier
            int Dn = __shfl_up_sync(0xffffffffu, Dc, 1), Un = __shfl_up_sync(0xffffffffu, Uc, 1), Mdiag = __shfl_up_sync(0xffffffffu, Mprev, 1);
            if (lane == 0) { Dn = 0; Un = 0; Mdiag = 0; }  // row 0 is all zero in the three matrices (free leading target gap)
            if (lane < n && j >= 0 && j < m) {
                const int dn = Mdiag + (qc == tgt[j] ? ma : mi);
                const int un = __viaddmax_s32(Dn, oe, Un + ge);
                const int ln = __viaddmax_s32(Dc, oe, Lc + ge);
                const int mn = __vimax3_s32(dn, ln, un);
                Mprev = Mc;
                Dc = dn; Lc = ln; Uc = un; Mc = mn;
                if (lane == n - 1) best = max(best, mn);
            }
        }
        best = __shfl_sync(0xffffffffu, best, n - 1);
        if (lane == 0) score[p] = best;
        __syncwarp();
    }
}

// the product's form for the same work, minimal: one thread per pair, rows in registers (score only, n <= 32), as a yardstick
__global__ void __launch_bounds__(128) thread_per_pair_kernel(const unsigned char *__restrict__ q, const unsigned char *__restrict__ t, int n, int m, int n_pairs,
                                                              int ma, int mi, int go, int ge, int *__restrict__ score) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const int oe = go + ge;
    int D[32], L[32], M[32];
    unsigned char qq[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { qq[i] = i < n ? q[(size_t)p * n + i] : 255; D[i] = kNeg; L[i] = kNeg; M[i] = go + (i + 1) * ge; }
    int best = kNeg;
    for (int j = 0; j < m; ++j) {
        const unsigned char tc = t[(size_t)p * m + j];
        int mdiag = 0, d_up = 0, u_up = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (i < n) {
                const int dn = mdiag + (qq[i] == tc ? ma : mi);
                const int un = __viaddmax_s32(d_up, oe, u_up + ge);
                const int ln = __viaddmax_s32(D[i], oe, L[i] + ge);
                mdiag = M[i];
                const int mn = __vimax3_s32(dn, ln, un);
                D[i] = dn; L[i] = ln; M[i] = mn; d_up = dn; u_up = un;
                if (i == n - 1) best = max(best, mn);
            }
        }
    }
    score[p] = best;
}

static int cpu_glocal(const unsigned char *q, int n, const unsigned char *t, int m, int ma, int mi, int go, int ge) {
    std::vector<int> D((n + 1) * (m + 1)), L(D.size()), U(D.size());
    auto at = [&](int i, int j) { return i * (m + 1) + j; };
    for (int j = 0; j <= m; ++j) { D[at(0, j)] = L[at(0, j)] = U[at(0, j)] = 0; }
    for (int i = 1; i <= n; ++i) { U[at(i, 0)] = go + i * ge; D[at(i, 0)] = L[at(i, 0)] = kNeg; }
    int best = kNeg;
    for (int i = 1; i <= n; ++i)
        for (int j = 1; j <= m; ++j) {
            const int md = std::max({D[at(i - 1, j - 1)], L[at(i - 1, j - 1)], U[at(i - 1, j - 1)]});
            D[at(i, j)] = md + (q[i - 1] == t[j - 1] ? ma : mi);
            U[at(i, j)] = std::max(D[at(i - 1, j)] + go + ge, U[at(i - 1, j)] + ge);
            L[at(i, j)] = std::max(D[at(i, j - 1)] + go + ge, L[at(i, j - 1)] + ge);
            if (i == n) best = std::max({best, D[at(i, j)], L[at(i, j)], U[at(i, j)]});
        }
    return best;
}

int main() {
    const int ma = 1, mi = -4, go = -6, ge = -1;
    printf("{\n \"what\": \"score-only Glocal, fgbio three-matrix recurrence, 32-bit DPX; warp-per-pair anti-diagonal wavefront (north-star form) against thread-per-pair register rows (the product's form, minimal version)\",\n \"shapes\": [\n");
    const int shapes[2][2] = {{24, 35}, {24, 150}};
    for (int sidx = 0; sidx < 2; ++sidx) {
        const int n = shapes[sidx][0], m = shapes[sidx][1], n_pairs = 4 << 20;
        std::vector<unsigned char> hq((size_t)n_pairs * n), ht((size_t)n_pairs * m);
        srand(7 + sidx);
        for (auto &c : ht) c = "ACGT"[rand() & 3];
        for (int p = 0; p < n_pairs; ++p)
            for (int i = 0; i < n; ++i) hq[(size_t)p * n + i] = (p & 1) ? ht[(size_t)p * m + 3 + i] : "ACGT"[rand() & 3];  // half of the pairs: the query sits in the target
        unsigned char *dq, *dt; int *ds;
        cudaMalloc(&dq, hq.size()); cudaMalloc(&dt, ht.size()); cudaMalloc(&ds, (size_t)n_pairs * 4);
        cudaMemcpy(dq, hq.data(), hq.size(), cudaMemcpyHostToDevice); cudaMemcpy(dt, ht.data(), ht.size(), cudaMemcpyHostToDevice);
        cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
        double ms[2]; bool ok[2];
        for (int which = 0; which < 2; ++which) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            float best_ms = 1e30f;
            for (int it = 0; it < 4; ++it) {
                cudaEventRecord(e0);
                if (which == 0) wavefront_kernel<<<prop.multiProcessorCount * 16, 128, 4 * ((m + 3) & ~3)>>>(dq, dt, n, m, n_pairs, ma, mi, go, ge, ds);
                else thread_per_pair_kernel<<<(n_pairs + 127) / 128, 128>>>(dq, dt, n, m, n_pairs, ma, mi, go, ge, ds);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float t; cudaEventElapsedTime(&t, e0, e1);
                if (it) best_ms = std::min(best_ms, t);
            }
            std::vector<int> hs(n_pairs);
            cudaMemcpy(hs.data(), ds, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost);
            ok[which] = cudaGetLastError() == cudaSuccess;
            for (int p = 0; p < 2000 && ok[which]; ++p) ok[which] = hs[p] == cpu_glocal(&hq[(size_t)p * n], n, &ht[(size_t)p * m], m, ma, mi, go, ge);
            ms[which] = best_ms;
        }
        const double cells = (double)n_pairs * n * m;
        printf("  {\"query\": %d, \"target\": %d, \"pairs\": %d, \"warp_per_pair_wavefront\": {\"ms\": %.3f, \"gcups\": %.1f, \"matches_cpu\": %s, \"lane_efficiency\": %.2f},"
               " \"thread_per_pair_rows_in_registers\": {\"ms\": %.3f, \"gcups\": %.1f, \"matches_cpu\": %s}}%s\n",
               n, m, n_pairs, ms[0], cells / ms[0] / 1e6, ok[0] ? "true" : "false", (double)n * m / ((n + m - 1) * 32.0), ms[1], cells / ms[1] / 1e6,
               ok[1] ? "true" : "false", sidx == 0 ? "," : "");
        cudaFree(dq); cudaFree(dt); cudaFree(ds);
    }
    printf(" ]\n}\n");
    return 0;
}
