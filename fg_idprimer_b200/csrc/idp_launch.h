// Host-side launchers of the kernels, one translation unit per kernel family (idp_scan.cu, idp_sw.cu) so that the
// library builds in parallel; idp_ctx.cu strings them into the matcher chain.
#pragma once
#include "idp_device.cuh"

namespace idp {

// what a launcher needs to know of a context
struct LaunchCfg {
    DevPanel dp{};
    int sm_count = 148;
    int rw = 4;               // 32-bit words of read window the scan kernel keeps in registers (8 bases each)
    int nq = 32;              // register rows of the aligners: 32, 64, or 0 (generic local-memory variants)
    bool scan_stats = false;  // count base comparisons in the scan kernel (IDP_SCAN_STATS=1 when the context is created)
    // scan_fx_kernel's copies, laid out for the stored nibble order; NULL when primers are longer than 32 bases:
    const uint4 *pm4 = nullptr;      // primer masks padded to 4 words
    const uint2 *rule_fx = nullptr;  // mismatchAlign thresholds + k-mer rule table (FxLaunch::rule)
};

// kernel A+B: the exact-prefix pass (idp_scanfx.cuh) when the panel and the batch layout allow it, then the general scan
// (idp_scan.cuh) over the reads it parked -- or over the whole batch.  `parked`: room for R.n read indices; `masks`: 2 * ceil(R.n / 32) words.  The fall-through list
// `fall` and ctr->n_fall are complete when the last kernel enqueued here has run.  *launches is bumped per kernel.
int launch_scan(const LaunchCfg &c, cudaStream_t st, const DevReads &R, ResultOut five, uint8_t *rtype, unsigned *fall, unsigned *parked,
                uint32_t *masks, Counters *ctr, long *launches);

// gapped stage (idp_kernels.cuh)
void launch_cand5(const LaunchCfg &c, cudaStream_t st, const DevReads &R, const unsigned *fall, Counters *ctr, unsigned *pair_work,
                  unsigned *pair_primer, unsigned pair_cap, unsigned *seg_off, int *seg_cnt, unsigned *spill);
void launch_cand3(const LaunchCfg &c, cudaStream_t st, const DevReads &R, ResultOut five, Counters *ctr, unsigned *all_list,
                  unsigned *pair_work, unsigned *pair_primer, unsigned pair_cap, unsigned *seg_off, int *seg_cnt);
void launch_sw_pairs(const LaunchCfg &c, cudaStream_t st, bool local, const DevReads &R, const unsigned *work_read, Counters *ctr,
                     const unsigned *pw, const unsigned *pp, int *ps, unsigned *pse, unsigned cap);
void launch_sw_all(const LaunchCfg &c, cudaStream_t st, bool local, const DevReads &R, const unsigned *work_read, const unsigned *n_work,
                   Fin *fin, Counters *ctr, int as5);
void launch_finalize_pairs(const LaunchCfg &c, cudaStream_t st, const DevReads &R, const unsigned *work_read, const unsigned *n_work_ptr,
                           unsigned n_work_fixed, const unsigned *seg_off, const int *seg_cnt, const unsigned *pp, const int *ps,
                           const unsigned *pse, ResultOut out, uint8_t *rtype, Counters *ctr, int three);
void launch_finalize_all(const LaunchCfg &c, cudaStream_t st, const DevReads &R, const unsigned *work_read, const unsigned *n_work_ptr,
                         const Fin *fin, ResultOut out, uint8_t *rtype, int three);
void launch_metrics(const LaunchCfg &c, cudaStream_t st, const DevReads &R, const uint8_t *rtype, Counters *ctr);
void launch_classify(const LaunchCfg &c, cudaStream_t st, const uint16_t *flags, ResultOut five, int n, int min_insert, int max_insert,
                     uint8_t *cls_per_read, Counters *ctr);
// batched AsyncAligner.align: q / t are the caller's ASCII on the device, q_off / t_off their n + 1 offsets; qwords needs
// n * max(8, (max_q + 7) / 8) words, tseq4 align_target_bytes() bytes; *bad is OR-ed with 1 when a letter is not an upper-case
// IUPAC base or '='; qstart / qend may be NULL
size_t align_target_bytes(int64_t total_target_bases, int n);
int launch_align(const LaunchCfg &c, cudaStream_t st, int mode, const char *q, const int64_t *q_off, const char *t, const int64_t *t_off, int n,
                 int max_q, uint32_t *qwords, uint8_t *tseq4, unsigned *bad, int32_t *score, int32_t *tstart, int32_t *tend, int32_t *qstart,
                 int32_t *qend);

}  // namespace idp
