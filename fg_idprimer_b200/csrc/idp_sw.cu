// Launchers of the gapped stage (idp_kernels.cuh): candidate emission, Smith-Waterman, finalize, metrics.
#include <algorithm>
#include <cstdlib>

#include "idp_kernels.cuh"
#include "idp_launch.h"

namespace idp {

void launch_cand5(const LaunchCfg &c, cudaStream_t st, const DevReads &R, const unsigned *fall, Counters *ctr, unsigned *pair_work,
                  unsigned *pair_primer, unsigned pair_cap, unsigned *seg_off, int *seg_cnt, unsigned *spill) {
    const int words = (c.dp.n_primers + 31) / 32;
    const size_t smem = (size_t)8 * words * 4;
    const DevKmerIndex &ix = c.dp.kidx[1];
    const unsigned bm_words = ix.direct_bits != nullptr ? ix.direct_bits_words : 0u;  // + the 8 KB Bloom filter of the IUPAC k-mers
    const size_t bm_bytes = bm_words ? (size_t)bm_words * 4 + 2048 * 4 : 0;
    cudaFuncSetAttribute(cand5_thread_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(bm_bytes, 1024));
    // one block of 1024 threads per SM when the map takes 128 KB, two otherwise
    const int per_sm = bm_bytes > 100 * 1024 ? 1 : 2;
    cand5_thread_kernel<<<c.sm_count * per_sm, kCand5Threads, bm_bytes, st>>>(c.dp, R, fall, ctr, pair_work, pair_primer, pair_cap, seg_off, seg_cnt, spill, bm_words);
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
        // two pairs per thread in 16-bit halves when the scores allow it
        if (c.nq == 32 && (LOCAL ? c.dp.local16_ok : c.dp.glocal16_ok)) { sw_pairs_kernel<32, LOCAL, false, true, true><<<grid, 128, 0, st>>>(c.dp, R, work_read, ctr, pw, pp, ps, pse, cap); return; }
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
    const bool packed = !SMAT && (LOCAL ? c.dp.fits16_local : c.dp.fits16_glocal);  // two alignments per register
    // one block per read, one round of the panel if it fits: as many warps as the panel needs (two primers per thread when packed),
    // so that no warp of a resident block sits idle -- 192 primers: 96 threads, five blocks per SM instead of four with a warp idle
    const int per_thread = packed && c.nq > 0 ? 2 : 1;
    const int threads = std::min(128, std::max(32, ((c.dp.n_primers + per_thread - 1) / per_thread + 31) / 32 * 32));
    const int grid = c.sm_count * (1024 / threads);
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

void launch_classify(const LaunchCfg &c, cudaStream_t st, const uint16_t *flags, ResultOut five, int n, int min_insert, int max_insert,
                     uint8_t *cls_per_read, Counters *ctr) {
    classify_kernel<<<std::max(1, std::min((n + 255) / 256, c.sm_count * 8)), 256, 0, st>>>(c.dp, flags, five, n, min_insert, max_insert, cls_per_read, ctr);
}

void launch_metrics(const LaunchCfg &c, cudaStream_t st, const DevReads &R, const uint8_t *rtype, Counters *ctr) {
    metrics_kernel<<<std::min<int>((R.n + 2047) / 2048, c.sm_count * 8), 256, 0, st>>>(R, rtype, ctr);
}

// ---------------------------------------------------------------------------------------------------------------
// Batched AsyncAligner.align (AA:37-54).  The caller's ASCII arrives as is; align_pack_kernel turns it into the layouts
// the aligners read (targets: BAM nibbles, byte-aligned per alignment; queries: nibble words, bottom-aligned in NQ rows for
// the register-resident aligners or top-aligned for the generic one) and validates the letters on the way.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ascii_nib(uint32_t c) {  // upper-case IUPAC and '=' -> BAM code, anything else -> 255
    // "=ACMGRSVTWYHKDBN": the code is the letter's index
    switch (c) {
        case '=': return 0;  case 'A': return 1;  case 'C': return 2;  case 'M': return 3;  case 'G': return 4;  case 'R': return 5;
        case 'S': return 6;  case 'V': return 7;  case 'T': return 8;  case 'W': return 9;  case 'Y': return 10; case 'H': return 11;
        case 'K': return 12; case 'D': return 13; case 'B': return 14; case 'N': return 15;
        default: return 255;
    }
}

// one warp per alignment.  nq: 32 / 64 = bottom-aligned query rows (qwords has nq/8 words per alignment), 0 = top-aligned, qw words
__global__ void __launch_bounds__(256) align_pack_kernel(const char *__restrict__ q, const int64_t *__restrict__ q_off, const char *__restrict__ t,
                                                         const int64_t *__restrict__ t_off, int n, int nq, int qw, uint32_t *__restrict__ qwords,
                                                         uint8_t *__restrict__ tseq4, unsigned int *__restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    bool invalid = false;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const int64_t q0 = q_off[i], t0 = t_off[i];
        const int ql = (int)(q_off[i + 1] - q0), tl = (int)(t_off[i + 1] - t0);
        uint8_t *dst = tseq4 + (t0 >> 1) + i;  // monotone in i and never overlapping: see align_target_offset
        for (int b = lane; b < (tl + 1) / 2; b += 32) {
            const uint32_t hi = ascii_nib((unsigned char)t[t0 + 2 * b]), lo = 2 * b + 1 < tl ? ascii_nib((unsigned char)t[t0 + 2 * b + 1]) : 0u;
            invalid |= (hi | lo) > 15u;
            dst[b] = (uint8_t)(((hi & 15u) << 4) | (lo & 15u));
        }
        const int words = nq > 0 ? nq / 8 : qw, shift = nq > 0 ? nq - ql : 0;  // bottom-aligned: query base b sits in row shift + b
        for (int w = lane; w < words; w += 32) {
            uint32_t v = 0;
            for (int k = 0; k < 8; ++k) {
                const int b = 8 * w + k - shift;
                if (b >= 0 && b < ql) { const uint32_t c = ascii_nib((unsigned char)q[q0 + b]); invalid |= c > 15u; v |= (c & 15u) << (4 * k); }
            }
            qwords[(size_t)i * words + w] = v;
        }
    }
    if (__any_sync(0xffffffffu, invalid) && lane == 0) atomicOr(bad, 1u);
}
__host__ __device__ inline int64_t align_target_offset(int64_t t_off_i, int i) { return (t_off_i >> 1) + i; }

// register-resident, traceback-equivalent: one thread per alignment; a block takes spans of its 128 * tile alignments and deals
// them to its warps sorted by query length (as sw_pairs_kernel does), so a warp's 32 alignments have (nearly) the same rows
template <int NQ, bool LOCAL>
__global__ void __launch_bounds__(128, 4) align_fast_kernel(DevPanel P, const uint32_t *__restrict__ qbot, const int64_t *__restrict__ q_off,
                                                          const uint8_t *__restrict__ tseq4, const int64_t *__restrict__ t_off, int n,
                                                          int32_t *__restrict__ score, int32_t *__restrict__ tstart, int32_t *__restrict__ tend,
                                                          int32_t *__restrict__ qstart, int32_t *__restrict__ qend) {
    constexpr int kTile = 4, kBins = NQ + 2;
    __shared__ unsigned short order_s[128 * kTile];
    __shared__ int hist_s[kBins];
    const unsigned span = kTile * 128u;
    for (unsigned base = blockIdx.x * span; base < (unsigned)n; base += gridDim.x * span) {
        for (int i = threadIdx.x; i < kBins; i += blockDim.x) hist_s[i] = 0;
        __syncthreads();
        int bin[kTile], rank[kTile];
#pragma unroll
        for (int k = 0; k < kTile; ++k) {
            const unsigned a = base + (unsigned)k * 128u + threadIdx.x;
            bin[k] = a < (unsigned)n ? min((int)(q_off[a + 1] - q_off[a]), NQ) : kBins - 1;  // not an alignment: sorts to the end
            rank[k] = atomicAdd(&hist_s[bin[k]], 1);
        }
        __syncthreads();
        if (threadIdx.x < 32) {  // exclusive prefix sum over the bins, one warp
            int carry = 0;
            for (int b0 = 0; b0 < kBins; b0 += 32) {
                const int b = b0 + (int)threadIdx.x;
                const int v = b < kBins ? hist_s[b] : 0;
                int incl = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, d); if ((int)threadIdx.x >= d) incl += u; }
                if (b < kBins) hist_s[b] = carry + incl - v;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kTile; ++k) order_s[hist_s[bin[k]] + rank[k]] = (unsigned short)(k * 128 + threadIdx.x);
        __syncthreads();
        for (int k = 0; k < kTile; ++k) {
            const unsigned a = base + order_s[k * 128 + threadIdx.x];
            const bool active = a < (unsigned)n;
            int nrow = 0, m = 0;
            Oriented T{nullptr, 0, false};
            if (active) {
                nrow = (int)(q_off[a + 1] - q_off[a]);
                m = (int)(t_off[a + 1] - t_off[a]);
                T = Oriented{tseq4 + align_target_offset(t_off[a], (int)a), m, false};
            }
            const int n_warp_max = __reduce_max_sync(0xffffffffu, nrow);
            const int n_warp_min = __reduce_min_sync(0xffffffffu, active ? nrow : n_warp_max);
            if (active) {
                const TracedReg tr = sw_trace_reg<NQ, LOCAL>(P, qbot + (size_t)a * (NQ / 8), nrow, NQ - n_warp_max, NQ - n_warp_min, T, m);
                score[a] = tr.score; tstart[a] = tr.t_start; tend[a] = tr.t_end;
                if (qstart) { qstart[a] = 1; qend[a] = nrow; }  // Glocal: the whole query is aligned (Local with query coordinates takes the generic kernel)
            }
        }
        __syncthreads();  // order_s / hist_s are rewritten by the next span
    }
}

// generic: any query length up to kMaxPrimerBases, all three modes, query coordinates; rows in local memory
template <bool QS>
__global__ void __launch_bounds__(128) align_generic_kernel(DevPanel P, const uint32_t *__restrict__ qwords, int qw, const int64_t *__restrict__ q_off,
                                                            const uint8_t *__restrict__ tseq4, const int64_t *__restrict__ t_off, int n, int mode,
                                                            int32_t *__restrict__ score, int32_t *__restrict__ tstart, int32_t *__restrict__ tend,
                                                            int32_t *__restrict__ qstart, int32_t *__restrict__ qend) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int ql = (int)(q_off[i + 1] - q_off[i]), tl = (int)(t_off[i + 1] - t_off[i]);
    Oriented T{tseq4 + align_target_offset(t_off[i], i), tl, false};
    const Traced tr = sw_trace<kMaxPrimerBases, QS>(P, qwords + (size_t)i * qw, ql, T, tl, mode);
    score[i] = tr.score; tstart[i] = tr.t_start; tend[i] = tr.t_end;
    if (qstart) { qstart[i] = tr.q_start; qend[i] = tr.q_end; }
}

size_t align_target_bytes(int64_t total_target_bases, int n) { return (size_t)(total_target_bases / 2 + n + 8); }

// max_q: the longest query of the batch.  Returns the number of kernels launched.
int launch_align(const LaunchCfg &c, cudaStream_t st, int mode, const char *q, const int64_t *q_off, const char *t, const int64_t *t_off, int n,
                 int max_q, uint32_t *qwords, uint8_t *tseq4, unsigned *bad, int32_t *score, int32_t *tstart, int32_t *tend, int32_t *qstart,
                 int32_t *qend) {
    const bool want_q = qstart != nullptr && mode == IDP_MODE_LOCAL;
    const bool fast = max_q <= 64 && c.dp.packed_trace_ok && c.dp.smat == nullptr && mode != IDP_MODE_GLOBAL && !want_q;
    const int nq = fast ? (max_q <= 32 ? 32 : 64) : 0, qw = (max_q + 7) / 8;
    align_pack_kernel<<<std::max(1, std::min((n + 7) / 8, c.sm_count * 8)), 256, 0, st>>>(q, q_off, t, t_off, n, nq, qw, qwords, tseq4, bad);
    const int grid = std::max(1, std::min((n + 511) / 512, c.sm_count * 8));
    if (fast) {
        const bool local = mode == IDP_MODE_LOCAL;
        if (nq == 32) { if (local) align_fast_kernel<32, true><<<grid, 128, 0, st>>>(c.dp, qwords, q_off, tseq4, t_off, n, score, tstart, tend, qstart, qend);
                        else align_fast_kernel<32, false><<<grid, 128, 0, st>>>(c.dp, qwords, q_off, tseq4, t_off, n, score, tstart, tend, qstart, qend); }
        else { if (local) align_fast_kernel<64, true><<<grid, 128, 0, st>>>(c.dp, qwords, q_off, tseq4, t_off, n, score, tstart, tend, qstart, qend);
               else align_fast_kernel<64, false><<<grid, 128, 0, st>>>(c.dp, qwords, q_off, tseq4, t_off, n, score, tstart, tend, qstart, qend); }
    } else if (want_q) {
        align_generic_kernel<true><<<(n + 127) / 128, 128, 0, st>>>(c.dp, qwords, qw, q_off, tseq4, t_off, n, mode, score, tstart, tend, qstart, qend);
    } else {
        align_generic_kernel<false><<<(n + 127) / 128, 128, 0, st>>>(c.dp, qwords, qw, q_off, tseq4, t_off, n, mode, score, tstart, tend, qstart, qend);
    }
    return 2;
}

}  // namespace idp
