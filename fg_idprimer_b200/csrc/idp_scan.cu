// Launcher of kernel A+B (idp_scan.cuh): shared-memory plan and template dispatch.
#include <algorithm>
#include <cstdlib>

#include "idp_launch.h"
#include "idp_scan.cuh"

namespace idp {

template <int RW, bool FAST>
static int launch_scan_t(const LaunchCfg &c, cudaStream_t st, const DevReads &R, ResultOut five, uint8_t *rtype, unsigned *fall, Counters *ctr) {
    // shared-memory plan: the per-warp read tiles first (two buffers each), the parking queue, then as much of the panel
    // side as fits in half an SM (two CTAs per SM: the kernel holds 64 registers per thread)
    ScanLaunch S{};
    const DevKmerIndex &ix = c.dp.kidx[0];
    const size_t budget = (size_t)(224 * 1024 / kScanMinBlocks - 1024);
    size_t used = 0;
    S.stride = R.fixed_len > 0 ? (uint32_t)((R.fixed_len + 1) >> 1) : 0;
    // a warp tile is 32 reads: 32 * stride bytes, always a multiple of 16 as the TMA bulk copy requires
    S.stage_tile = R.fixed_len > 0 && (reinterpret_cast<uintptr_t>(R.seq4) & 15) == 0;
    if (S.stage_tile) {
        // 32 reads, slack for the window words read past the last read's end, then the tile's 32 flag words
        S.flags_off = (uint32_t)((32 * (size_t)S.stride + 4 * RW + 16 + 15) & ~size_t(15));
        S.tile_smem_bytes = S.flags_off + 64;
        S.stage_flags = (reinterpret_cast<uintptr_t>(R.flags) & 15) == 0;
        used = (size_t)S.tile_smem_bytes * 2 * kScanWarps;
        if (used > budget - 24 * 1024) { S.stage_tile = 0; S.tile_smem_bytes = 0; used = 0; }
    }
    auto place = [&](size_t bytes, bool wanted, size_t align = 16) -> uint32_t {
        bytes = (bytes + 15) & ~size_t(15);
        const size_t at = (used + align - 1) & ~(align - 1);
        if (!wanted || bytes == 0 || at + bytes > budget) return 0xFFFFFFFFu;
        used = at + bytes;
        return (uint32_t)at;
    };
    S.direct_entries = ix.direct ? (1u << (2 * ix.k)) : 0;
    S.use_fx = FAST && c.dp.fx_tab != nullptr && !getenv("IDP_SCAN_NO_FX");
    S.off_queue = place((size_t)kScanWarps * kScanQueueCap * sizeof(uint2), S.use_fx != 0);
    S.off_bsacc = place((size_t)c.dp.bs_groups * sizeof(uint2), FAST);
    S.off_bsmm = place((size_t)c.dp.bs_groups * kBsStride * 16 * 4, FAST, 128);  // 128-byte aligned: table offsets are OR-ed in
    S.off_fx = place((size_t)c.dp.fx_slots * 2, S.use_fx != 0 && c.dp.fx_slots <= 16384);
    S.off_seed = place((size_t)c.dp.seed_slots * 2 * 4, FAST && c.dp.seed_ok && c.dp.seed_slots <= 2048);
    S.off_direct = place((size_t)S.direct_entries * 4, !FAST && ix.direct != nullptr);
    S.off_diag = place((size_t)c.dp.n_primers * sizeof(uint2), FAST && c.dp.diag != nullptr);
    S.off_pinfo = place((size_t)c.dp.n_primers * sizeof(int4), true);
    S.off_pmask = place((size_t)c.dp.n_primers * c.dp.ww * 4, true);
    S.off_heads = place((size_t)ix.n_postings * 8, !FAST && ix.k > 0);
    const int n_tiles = (R.n + kScanThreads - 1) / kScanThreads;
    const int grid = std::max(1, std::min(n_tiles, kScanMinBlocks * c.sm_count));
    const bool onchip = FAST && S.stage_tile;
    auto launch = [&](auto kernel) -> int {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget) != cudaSuccess) {
            set_error("cudaFuncSetAttribute(scan_kernel) failed: %s", cudaGetErrorString(cudaGetLastError()));
            return IDP_ERR_CUDA;
        }
        kernel<<<grid, kScanThreads, used, st>>>(c.dp, R, S, five, rtype, fall, ctr);
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
static int launch_scan_rw(const LaunchCfg &c, cudaStream_t st, const DevReads &R, ResultOut five, uint8_t *rtype, unsigned *fall, Counters *ctr) {
    if (RW <= 8 && c.dp.bs_ok && !getenv("IDP_SCAN_EXACT_WALK")) return launch_scan_t<(RW <= 8 ? RW : 8), true>(c, st, R, five, rtype, fall, ctr);
    return launch_scan_t<RW, false>(c, st, R, five, rtype, fall, ctr);
}

int launch_scan(const LaunchCfg &c, cudaStream_t st, const DevReads &R, ResultOut five, uint8_t *rtype, unsigned *fall, Counters *ctr) {
    switch (c.rw) {
        case 4: return launch_scan_rw<4>(c, st, R, five, rtype, fall, ctr);
        case 5: return launch_scan_rw<5>(c, st, R, five, rtype, fall, ctr);
        case 8: return launch_scan_rw<8>(c, st, R, five, rtype, fall, ctr);
        case 16: return launch_scan_rw<16>(c, st, R, five, rtype, fall, ctr);
        default: return launch_scan_rw<32>(c, st, R, five, rtype, fall, ctr);
    }
}

}  // namespace idp
