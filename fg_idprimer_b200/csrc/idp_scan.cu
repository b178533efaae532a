// Launchers of kernel A+B: scan_fx_kernel (idp_scanfx.cuh) and scan_kernel (idp_scan.cuh) -- shared-memory plans and
// template dispatch.
#include <algorithm>
#include <cstdlib>

#include "idp_launch.h"
#include "idp_scan.cuh"
#include "idp_scanfx.cuh"

namespace idp {

template <int RW, bool FAST>
static int launch_scan_t(const LaunchCfg &c, cudaStream_t st, const DevReads &R, ResultOut five, uint8_t *rtype, uint32_t *fall_mask, Counters *ctr,
                         const unsigned *list, uint32_t *defer_mask) {
    // shared-memory plan: the per-warp read tiles first (two buffers each), then as much of the panel side as fits in half
    // an SM (two CTAs per SM: the kernel holds 64 registers per thread)
    ScanLaunch S{};
    const DevKmerIndex &ix = c.dp.kidx[0];
    const size_t budget = (size_t)(224 * 1024 / kScanMinBlocks - 1024);
    size_t used = 0;
    S.list = list;
    S.fall_mask = fall_mask;
    S.defer_mask = FAST && list != nullptr && !c.scan_stats ? defer_mask : nullptr;
    S.stride = R.fixed_len > 0 ? (uint32_t)((R.fixed_len + 1) >> 1) : 0;
    // a warp tile is 32 reads: 32 * stride bytes, always a multiple of 16 as the TMA bulk copy requires
    S.stage_tile = list == nullptr && R.fixed_len > 0 && (reinterpret_cast<uintptr_t>(R.seq4) & 15) == 0;
    if (S.stage_tile) {
        // 32 reads, slack for the window words read past the last read's end, then the tile's 32 flag words
        S.flags_off = (uint32_t)((32 * (size_t)S.stride + 4 * RW + 16 + 15) & ~size_t(15));
        S.tile_smem_bytes = S.flags_off + 64;
        S.stage_flags = (reinterpret_cast<uintptr_t>(R.flags) & 15) == 0;
        used = (size_t)S.tile_smem_bytes * 2 * kScanWarps;
        if (used > budget - 16 * 1024) { S.stage_tile = 0; S.tile_smem_bytes = 0; used = 0; }
    }
    auto place = [&](size_t bytes, bool wanted, size_t align = 16) -> uint32_t {
        bytes = (bytes + 15) & ~size_t(15);
        const size_t at = (used + align - 1) & ~(align - 1);
        if (!wanted || bytes == 0 || at + bytes > budget) return 0xFFFFFFFFu;
        used = at + bytes;
        return (uint32_t)at;
    };
    S.direct_entries = ix.direct ? (1u << (2 * ix.k)) : 0;
    S.off_bsacc = place((size_t)c.dp.bs_groups * sizeof(uint2), FAST);
    S.off_bsmm = place((size_t)c.dp.bs_groups * kBsStride * 16 * 4, FAST, 128);  // 128-byte aligned: table offsets are OR-ed in
    S.off_seed = place((size_t)c.dp.seed_slots * 2 * 4, FAST && c.dp.seed_ok && c.dp.seed_slots <= 2048);
    S.off_direct = place((size_t)S.direct_entries * 4, !FAST && ix.direct != nullptr);
    S.off_diag = place((size_t)c.dp.n_primers * sizeof(uint2), FAST && c.dp.diag != nullptr);
    S.off_pinfo = place((size_t)c.dp.n_primers * sizeof(int4), true);
    S.off_pmask = place((size_t)c.dp.n_primers * c.dp.ww * 4, true);
    S.off_heads = place((size_t)ix.n_postings * 8, !FAST && ix.k > 0);
    const int n_tiles = (R.n + kScanThreads - 1) / kScanThreads;
    const int grid = std::max(1, std::min(n_tiles, kScanMinBlocks * c.sm_count));
    const bool onchip = FAST && (S.stage_tile || list != nullptr);  // the ONCHIP variant also serves list mode (it then loads from global)
    auto launch = [&](auto kernel) -> int {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget) != cudaSuccess) {
            set_error("cudaFuncSetAttribute(scan_kernel) failed: %s", cudaGetErrorString(cudaGetLastError()));
            return IDP_ERR_CUDA;
        }
        kernel<<<grid, kScanThreads, used, st>>>(c.dp, R, S, five, rtype, ctr);
        return IDP_OK;
    };
    if (c.scan_stats) {  // idp_timings.scan_base_compares wanted (IDP_SCAN_STATS=1): the counting variant of the kernel
        if (FAST && onchip) return launch(scan_kernel<RW, FAST, FAST, true>);
        return launch(scan_kernel<RW, FAST, false, true>);
    }
    if (FAST && onchip) return launch(scan_kernel<RW, FAST, FAST, false>);
    return launch(scan_kernel<RW, FAST, false, false>);
}
template <int RW>
static int launch_scan_rw(const LaunchCfg &c, cudaStream_t st, const DevReads &R, ResultOut five, uint8_t *rtype, uint32_t *fall_mask, Counters *ctr,
                          const unsigned *list, uint32_t *defer_mask) {
    if (RW <= 8 && c.dp.bs_ok && !getenv("IDP_SCAN_EXACT_WALK")) return launch_scan_t<(RW <= 8 ? RW : 8), true>(c, st, R, five, rtype, fall_mask, ctr, list, defer_mask);
    return launch_scan_t<RW, false>(c, st, R, five, rtype, fall_mask, ctr, list, defer_mask);
}
static int launch_scan_general(const LaunchCfg &c, cudaStream_t st, const DevReads &R, ResultOut five, uint8_t *rtype, uint32_t *fall_mask, Counters *ctr,
                               const unsigned *list, uint32_t *defer_mask) {
    switch (c.rw) {
        case 4: return launch_scan_rw<4>(c, st, R, five, rtype, fall_mask, ctr, list, defer_mask);
        case 5: return launch_scan_rw<5>(c, st, R, five, rtype, fall_mask, ctr, list, defer_mask);
        case 8: return launch_scan_rw<8>(c, st, R, five, rtype, fall_mask, ctr, list, defer_mask);
        case 16: return launch_scan_rw<16>(c, st, R, five, rtype, fall_mask, ctr, list, defer_mask);
        default: return launch_scan_rw<32>(c, st, R, five, rtype, fall_mask, ctr, list, defer_mask);
    }
}

// the exact-prefix pass applies when: every primer tolerates at most one mismatch and the panel has the table; primers fit 32
// bases (4 mask words, and then the k-mer rule table exists whenever -k is on); the batch is fixed-length with at least the
// 32 window bases per read, 16-byte aligned rows and flags
static bool fx_applies(const LaunchCfg &c, const DevReads &R) {
    const DevPanel &d = c.dp;
    return d.bs_ok && d.fx_tab != nullptr && c.pm4 != nullptr && c.rw == 4 && !c.scan_stats &&            c.rule_fx != nullptr && R.fixed_len >= 32 && R.n >= 32 && (reinterpret_cast<uintptr_t>(R.seq4) & 15) == 0 && (reinterpret_cast<uintptr_t>(R.flags) & 15) == 0 &&
           !getenv("IDP_SCAN_NO_FX") && !getenv("IDP_SCAN_EXACT_WALK");
}

static int launch_scan_fx(const LaunchCfg &c, cudaStream_t st, const DevReads &R, ResultOut five, uint8_t *rtype, uint32_t *fall_mask, uint32_t *park_mask) {
    FxLaunch S{};
    S.stride = (uint32_t)((R.fixed_len + 1) >> 1);
    S.flags_off = (uint32_t)((32 * (size_t)S.stride + 24 + 15) & ~size_t(15));  // slack: a lane reads five aligned words around its 16 bytes
    S.buf_bytes = S.flags_off + 64;
    S.nbuf = getenv("IDP_FX_NBUF") ? std::max(1, std::min(2, atoi(getenv("IDP_FX_NBUF")))) : 1;  // two buffers measured no faster (the shared-memory pipe binds)
    S.n_tiles = (uint32_t)(R.n / 32);
    S.park_mapped = R.ref_idx != nullptr && c.dp.n_loc[0] + c.dp.n_loc[1] > 0;
    S.k = c.dp.kidx[0].k;
    S.pm4 = c.pm4;
    S.rule = c.rule_fx;
    S.fall_mask = fall_mask;
    S.park_mask = park_mask;
    size_t used = (size_t)S.buf_bytes * S.nbuf * kFxWarps;
    const size_t budget = 72 * 1024;  // three CTAs per SM
    auto place = [&](size_t bytes, bool wanted) -> uint32_t {
        bytes = (bytes + 15) & ~size_t(15);
        if (!wanted || bytes == 0 || used + bytes > budget) return 0xFFFFFFFFu;
        const size_t at = used;
        used += bytes;
        return (uint32_t)at;
    };
    S.off_fx = place((size_t)c.dp.fx_slots * 2, true);
    S.off_pm4 = place((size_t)c.dp.n_primers * 16, true);
    S.off_rule = place((size_t)c.dp.n_primers * 8, true);
    if (cudaFuncSetAttribute(scan_fx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(used, budget)) != cudaSuccess) {
        set_error("cudaFuncSetAttribute(scan_fx_kernel) failed: %s", cudaGetErrorString(cudaGetLastError()));
        return IDP_ERR_CUDA;
    }
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_fx_kernel, kFxThreads, used) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 1; }
    const int grid = std::max(1, std::min<int>((int)((S.n_tiles + kFxWarps - 1) / kFxWarps), per_sm * c.sm_count));
    scan_fx_kernel<<<grid, kFxThreads, used, st>>>(c.dp, R, S, five, rtype);
    return IDP_OK;
}

static void launch_compact(const LaunchCfg &c, cudaStream_t st, const uint32_t *mask, int n_tiles, unsigned *list, unsigned *counter,
                           const unsigned *n_items_dev = nullptr) {
    const int grid = std::max(1, std::min((n_tiles + 255) / 256, c.sm_count * 8));
    compact_mask_kernel<<<grid, 256, 0, st>>>(mask, n_tiles, list, counter, n_items_dev);
}

template <int RW>
static void launch_member_check_t(const LaunchCfg &c, cudaStream_t st, const DevReads &R, const unsigned *parked, const unsigned *items, Counters *ctr,
                                  ResultOut five, uint8_t *rtype, uint32_t *fall_mask) {
    member_check_kernel<RW><<<c.sm_count * 2, 128, 0, st>>>(c.dp, R, parked, items, ctr, five, rtype, fall_mask);
}

// masks: 2 * ceil(R.n / 32) words of scratch (fall-through bits, then parked bits)
int launch_scan(const LaunchCfg &c, cudaStream_t st, const DevReads &R, ResultOut five, uint8_t *rtype, unsigned *fall, unsigned *parked,
                uint32_t *masks, Counters *ctr, long *launches) {
    int rc;
    const int n_tiles = (R.n + 31) / 32;
    uint32_t *fall_mask = masks, *park_mask = masks + n_tiles;
    if (fx_applies(c, R)) {
        if ((rc = launch_scan_fx(c, st, R, five, rtype, fall_mask, park_mask))) return rc;
        launch_compact(c, st, park_mask, n_tiles, parked, &ctr->n_parked);
        *launches += 2;
        // the parked bits are spent: their words now take the deferred bits, one per LIST item
        const bool defer = c.dp.kidx[0].k > 0 && c.dp.member_bits != nullptr && !getenv("IDP_SCAN_NO_DEFER");
        if ((rc = launch_scan_general(c, st, R, five, rtype, fall_mask, ctr, parked, defer ? park_mask : nullptr))) return rc;
        if (defer) {  // `fall` holds the deferred items until the last compaction below overwrites it
            launch_compact(c, st, park_mask, n_tiles, fall, &ctr->n_defer, &ctr->n_parked);
            launch_member_check_t<4>(c, st, R, parked, fall, ctr, five, rtype, fall_mask);  // fx_applies: four window words
            *launches += 2;
        }
    } else {
        rc = launch_scan_general(c, st, R, five, rtype, fall_mask, ctr, nullptr, nullptr);
    }
    if (rc) return rc;
    launch_compact(c, st, fall_mask, n_tiles, fall, &ctr->n_fall);
    *launches += 2;
    return IDP_OK;
}

}  // namespace idp
