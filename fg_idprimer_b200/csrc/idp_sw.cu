// Launchers of the gapped stage (idp_kernels.cuh): candidate emission, Smith-Waterman, finalize, metrics.
#include <algorithm>

#include "idp_kernels.cuh"
#include "idp_launch.h"

namespace idp {

void launch_cand5(const LaunchCfg &c, cudaStream_t st, const DevReads &R, const unsigned *fall, Counters *ctr, unsigned *pair_work,
                  unsigned *pair_primer, unsigned pair_cap, unsigned *seg_off, int *seg_cnt, unsigned *spill) {
    const int words = (c.dp.n_primers + 31) / 32;
    const size_t smem = (size_t)8 * words * 4;
    cand5_thread_kernel<<<c.sm_count * 8, 256, 0, st>>>(c.dp, R, fall, ctr, pair_work, pair_primer, pair_cap, seg_off, seg_cnt, spill);
    cand5_kernel<<<c.sm_count * 4, 256, smem, st>>>(c.dp, R, fall, spill, ctr, pair_work, pair_primer, pair_cap, seg_off, seg_cnt);
}

void launch_cand3(const LaunchCfg &c, cudaStream_t st, const DevReads &R, ResultOut five, Counters *ctr, unsigned *all_list,
                  unsigned *pair_work, unsigned *pair_primer, unsigned pair_cap, unsigned *seg_off, int *seg_cnt) {
    cand3_kernel<<<(R.n + 255) / 256, 256, 0, st>>>(c.dp, R, five, ctr, all_list, pair_work, pair_primer, pair_cap, seg_off, seg_cnt);
}

// the pair kernel also produces the traceback-equivalent (start, end) when the packed register variant applies
static bool pairs_trace(const LaunchCfg &c) { return c.nq > 0 && c.dp.packed_trace_ok && c.dp.smat == nullptr; }

template <bool LOCAL, bool SMAT>
static void launch_sw_pairs_t(const LaunchCfg &c, cudaStream_t st, const DevReads &R, const unsigned *work_read, Counters *ctr,
                              const unsigned *pw, const unsigned *pp, int *ps, unsigned *pse, unsigned cap) {
    const int grid = c.sm_count * 8;
    if (!SMAT && pairs_trace(c)) {
        if (c.nq == 32) sw_pairs_kernel<32, LOCAL, false, true><<<grid, 128, 0, st>>>(c.dp, R, work_read, ctr, pw, pp, ps, pse, cap);
        else sw_pairs_kernel<64, LOCAL, false, true><<<grid, 128, 0, st>>>(c.dp, R, work_read, ctr, pw, pp, ps, pse, cap);
        return;
    }
    switch (c.nq) {
        case 32: sw_pairs_kernel<32, LOCAL, SMAT, false><<<grid, 128, 0, st>>>(c.dp, R, work_read, ctr, pw, pp, ps, pse, cap); break;
        case 64: sw_pairs_kernel<64, LOCAL, SMAT, false><<<grid, 128, 0, st>>>(c.dp, R, work_read, ctr, pw, pp, ps, pse, cap); break;
        default: sw_pairs_kernel<0, LOCAL, SMAT, false><<<grid, 128, 0, st>>>(c.dp, R, work_read, ctr, pw, pp, ps, pse, cap); break;
    }
}
void launch_sw_pairs(const LaunchCfg &c, cudaStream_t st, bool local, const DevReads &R, const unsigned *work_read, Counters *ctr,
                     const unsigned *pw, const unsigned *pp, int *ps, unsigned *pse, unsigned cap) {
    const bool smat = c.dp.smat != nullptr;
    if (local) { if (smat) launch_sw_pairs_t<true, true>(c, st, R, work_read, ctr, pw, pp, ps, pse, cap); else launch_sw_pairs_t<true, false>(c, st, R, work_read, ctr, pw, pp, ps, pse, cap); }
    else { if (smat) launch_sw_pairs_t<false, true>(c, st, R, work_read, ctr, pw, pp, ps, pse, cap); else launch_sw_pairs_t<false, false>(c, st, R, work_read, ctr, pw, pp, ps, pse, cap); }
}

template <bool LOCAL, bool SMAT>
static void launch_sw_all_t(const LaunchCfg &c, cudaStream_t st, const DevReads &R, const unsigned *work_read, const unsigned *n_work,
                            Fin *fin, Counters *ctr, int as5) {
    const int threads = c.dp.n_primers <= 32 ? 32 : c.dp.n_primers <= 64 ? 64 : 128;
    const int grid = c.sm_count * (threads == 128 ? 8 : threads == 64 ? 16 : 32);
    const bool packed = !SMAT && (LOCAL ? c.dp.fits16_local : c.dp.fits16_glocal);  // two alignments per register
    switch (c.nq) {
        case 32:
            if (packed) sw_all_kernel<32, LOCAL, SMAT, !SMAT><<<grid, threads, 0, st>>>(c.dp, R, work_read, n_work, fin, ctr, as5);
            else sw_all_kernel<32, LOCAL, SMAT, false><<<grid, threads, 0, st>>>(c.dp, R, work_read, n_work, fin, ctr, as5);
            break;
        case 64:
            if (packed) sw_all_kernel<64, LOCAL, SMAT, !SMAT><<<grid, threads, 0, st>>>(c.dp, R, work_read, n_work, fin, ctr, as5);
            else sw_all_kernel<64, LOCAL, SMAT, false><<<grid, threads, 0, st>>>(c.dp, R, work_read, n_work, fin, ctr, as5);
            break;
        default: sw_all_kernel<0, LOCAL, SMAT, false><<<grid, threads, 0, st>>>(c.dp, R, work_read, n_work, fin, ctr, as5); break;
    }
}
void launch_sw_all(const LaunchCfg &c, cudaStream_t st, bool local, const DevReads &R, const unsigned *work_read, const unsigned *n_work,
                   Fin *fin, Counters *ctr, int as5) {
    const bool smat = c.dp.smat != nullptr;
    if (local) { if (smat) launch_sw_all_t<true, true>(c, st, R, work_read, n_work, fin, ctr, as5); else launch_sw_all_t<true, false>(c, st, R, work_read, n_work, fin, ctr, as5); }
    else { if (smat) launch_sw_all_t<false, true>(c, st, R, work_read, n_work, fin, ctr, as5); else launch_sw_all_t<false, false>(c, st, R, work_read, n_work, fin, ctr, as5); }
}

void launch_finalize_pairs(const LaunchCfg &c, cudaStream_t st, const DevReads &R, const unsigned *work_read, const unsigned *n_work_ptr,
                           unsigned n_work_fixed, const unsigned *seg_off, const int *seg_cnt, const unsigned *pp, const int *ps,
                           const unsigned *pse, ResultOut out, uint8_t *rtype, Counters *ctr, int three) {
    const int grid = c.sm_count * 8;
    if (!pairs_trace(c)) pse = nullptr;
    if (c.dp.max_seq_len <= 64) finalize_pairs_kernel<64, 0><<<grid, 128, 0, st>>>(c.dp, R, work_read, n_work_ptr, n_work_fixed, seg_off, seg_cnt, pp, ps, pse, out, rtype, ctr, three);
    else finalize_pairs_kernel<kMaxPrimerBases, 0><<<grid, 128, 0, st>>>(c.dp, R, work_read, n_work_ptr, n_work_fixed, seg_off, seg_cnt, pp, ps, pse, out, rtype, ctr, three);
}
void launch_finalize_all(const LaunchCfg &c, cudaStream_t st, const DevReads &R, const unsigned *work_read, const unsigned *n_work_ptr,
                         const Fin *fin, ResultOut out, uint8_t *rtype, int three) {
    const int grid = c.sm_count * 8;
    if (c.nq == 32) finalize_all_kernel<64, 32><<<grid, 128, 0, st>>>(c.dp, R, work_read, n_work_ptr, fin, out, rtype, three);
    else if (c.nq == 64) finalize_all_kernel<64, 64><<<grid, 128, 0, st>>>(c.dp, R, work_read, n_work_ptr, fin, out, rtype, three);
    else finalize_all_kernel<kMaxPrimerBases, 0><<<grid, 128, 0, st>>>(c.dp, R, work_read, n_work_ptr, fin, out, rtype, three);
}

void launch_metrics(const LaunchCfg &c, cudaStream_t st, const DevReads &R, const uint8_t *rtype, Counters *ctr) {
    metrics_kernel<<<std::min<int>((R.n + 2047) / 2048, c.sm_count * 8), 256, 0, st>>>(R, rtype, ctr);
}

// ---- batched AsyncAligner.align (AA:37-54) ----
__global__ void align_kernel(DevPanel P, const uint32_t *__restrict__ qwords, int qw, const int32_t *__restrict__ qlen,
                             const uint8_t *__restrict__ tseq4, const uint64_t *__restrict__ toff, const int32_t *__restrict__ tlen,
                             int n, int mode, int32_t *__restrict__ score, int32_t *__restrict__ tstart, int32_t *__restrict__ tend) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Oriented T{tseq4 + toff[i], tlen[i], false};
    const Traced tr = sw_trace<kMaxPrimerBases>(P, qwords + (size_t)i * qw, qlen[i], T, tlen[i], mode);
    score[i] = tr.score; tstart[i] = tr.t_start; tend[i] = tr.t_end;
}

void launch_align(const LaunchCfg &c, cudaStream_t st, const uint32_t *qwords, int qw, const int32_t *qlen, const uint8_t *tseq4,
                  const uint64_t *toff, const int32_t *tlen, int n, int mode, int32_t *score, int32_t *tstart, int32_t *tend) {
    align_kernel<<<(n + 127) / 128, 128, 0, st>>>(c.dp, qwords, qw, qlen, tseq4, toff, tlen, n, mode, score, tstart, tend);
}

}  // namespace idp
