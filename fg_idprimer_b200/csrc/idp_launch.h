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
};

// kernel A+B (idp_scan.cuh)
int launch_scan(const LaunchCfg &c, cudaStream_t st, const DevReads &R, ResultOut five, uint8_t *rtype, unsigned *fall, Counters *ctr);

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
void launch_align(const LaunchCfg &c, cudaStream_t st, const uint32_t *qwords, int qw, const int32_t *qlen, const uint8_t *tseq4,
                  const uint64_t *toff, const int32_t *tlen, int n, int mode, int32_t *score, int32_t *tstart, int32_t *tend);

}  // namespace idp
