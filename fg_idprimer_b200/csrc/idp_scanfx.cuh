// Kernel A+B, common case: the exact-prefix pass of the ungapped scan.
//
//   basesInSequencingOrder (PM:39-45) . getPrimersForAlignmentTasks (PM:98-116) . mismatchAlign / numMismatches
//   (PM:128-147) . getBestUngappedAlignment (PM:237-248) -- for the reads whose answer one table lookup decides.
//
// When every primer tolerates at most one mismatch (DevPanel::bs_ok), a read whose first 12 bases spell an ISOLATED primer
// without a mismatch (DevPanel::fx_tab, built by idp_panel_create) has that primer as its only possible ungapped match: the
// whole candidate search is one hash, one 16-bit table load and one mismatchAlign.  If the primer is accepted, the k-mer
// rule (PM:108-116) is looked up in DevPanel::diag by the position of the (at most one) mismatch.  That settles ~90 % of
// the reads of an amplicon run in ~170 warp instructions per 32 reads.  Everything else -- a sequencing error or an
// ambiguity code in the first 12 bases, off-target templates, reads that need the location matcher or the reverse
// complement, short reads, the rare accepted primer whose k-mer rule the table cannot prove -- is PARKED: its index goes to
// a list (via a bit mask and compact_mask_kernel) that scan_kernel (idp_scan.cuh, list mode) works through afterwards with full lanes.  Exactness does not depend on
// what is parked, only on never deciding a read here that the general kernel would decide differently.
//
// Layout: persistent CTAs, one warp per 32-read tile; a tile's packed rows (any fixed stride: 18 B windowed rows or 75 B
// whole reads) and its 32 flag words arrive by TMA bulk copies on a per-warp mbarrier; the next tile is requested as soon
// as the lanes have lifted their windows into registers, so one buffer per warp is enough.
#pragma once
#include "idp_device.cuh"

namespace idp {

constexpr int kFxThreads = 512;
constexpr int kFxWarps = kFxThreads / 32;

struct FxLaunch {
    uint32_t stride;        // bytes per read
    uint32_t flags_off;     // byte offset of the tile's 32 flag words in a warp's buffer (multiple of 16)
    uint32_t buf_bytes;     // bytes of one tile buffer (multiple of 16)
    uint32_t nbuf;          // tile buffers per warp: 2 for small rows (one tile in flight per warp does not cover the HBM latency), else 1
    uint32_t n_tiles;       // full 32-read tiles; the reads after them are parked
    // byte offsets into dynamic shared memory; 0xFFFFFFFF: the table stays in global memory
    uint32_t off_fx, off_pm4, off_rule;
    int32_t park_mapped;    // 1: reads the location matcher could match (mapped read, mapped primers on the panel) are parked
    int32_t k;              // -k of the ungapped matcher (<= 0: off)
    const uint4 *pm4;       // [n] primer masks padded to 4 words (0xF beyond the primer), STORED nibble order; global copy
    // [n] per primer {Primer.length | cap << 10 | accept << 20 | (a k-run exists without a mismatch) << 30, k-mer rule bits}:
    // the thresholds of mismatchAlign (DevPanel::pinfo, cap clamped to 1023: the window has 32 positions) and the k-mer rule
    // table of DevPanel::diag with its bits laid out like `where` below -- one 8-byte load
    const uint2 *rule;
    uint32_t *fall_mask;    // [ceil(n / 32)] bit i of word t: read 32 t + i fell through to the gapped stage
    uint32_t *park_mask;    // [ceil(n / 32)] ... was left to scan_kernel
};

// Everything in this kernel works on the row words AS STORED (first base of a byte in its high nibble: base i of a word sits
// in nibble i ^ 1): the primer masks (FxLaunch::pm4), the table's hash (idp_panel_create) and the k-mer rule table
// (FxLaunch::rule) are laid out for that order on the host, so no nibble swap is ever executed.
// Mismatch flags of the 32 window positions in one word: nibble j of the result holds position j of word 0 in bit 0, of word
// 2 in bit 1, of word 1 in bit 2 and of word 3 in bit 3 (the order the two-level fold below produces).
constexpr uint32_t kFxFirst12 = 0x11115555u;  // the flags of word 0 and of the first four positions of word 1

__device__ __forceinline__ uint32_t fx_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ uint32_t lds16(uint32_t addr) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr)); return v; }
// 1 in bit 3 of some nibble iff some nibble of x is zero (exact as a boolean)
__device__ __forceinline__ uint32_t has_zero_nibble(uint32_t x) { return (x - 0x11111111u) & ~x & 0x88888888u; }

__global__ void __launch_bounds__(kFxThreads, 3) scan_fx_kernel(const __grid_constant__ DevPanel P, const __grid_constant__ DevReads R, const __grid_constant__ FxLaunch S,
                                                              const ResultOut five, uint8_t *__restrict__ rtype) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t tile_bar[kFxWarps][2];
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    // ---- panel-side tables into shared memory (once per persistent CTA) ----
    const uint16_t *fx_tab = P.fx_tab;
    const uint4 *pm4 = S.pm4;
    const uint2 *rule = S.rule;
    if (S.off_fx != 0xFFFFFFFFu) {
        uint32_t *d = reinterpret_cast<uint32_t *>(smem + S.off_fx);  // slot counts are powers of two >= 1024: whole words
        const uint32_t *src = reinterpret_cast<const uint32_t *>(P.fx_tab);
        for (int i = tid; i < P.fx_slots / 2; i += kFxThreads) d[i] = src[i];
        fx_tab = reinterpret_cast<const uint16_t *>(d);
    }
    if (S.off_pm4 != 0xFFFFFFFFu) {
        uint4 *d = reinterpret_cast<uint4 *>(smem + S.off_pm4);
        for (int i = tid; i < P.n_primers; i += kFxThreads) d[i] = S.pm4[i];
        pm4 = d;
    }
    if (S.off_rule != 0xFFFFFFFFu) {
        uint2 *d = reinterpret_cast<uint2 *>(smem + S.off_rule);
        for (int i = tid; i < P.n_primers; i += kFxThreads) d[i] = S.rule[i];
        rule = d;
    }
    const uint32_t bar0 = fx_smem_u32(&tile_bar[wib][0]);  // the second buffer's barrier is 8 bytes further on
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // the reads after the last full tile go to the general kernel
    if (blockIdx.x == 0 && tid == 0 && (R.n & 31)) {
        S.fall_mask[S.n_tiles] = 0u;
        S.park_mask[S.n_tiles] = (1u << (R.n & 31)) - 1u;
    }
    // ---- every warp walks its own tiles; no CTA-wide barrier from here on ----
    const uint32_t buf0 = fx_smem_u32(smem) + (uint32_t)wib * S.nbuf * S.buf_bytes;  // this warp's tile buffers (dynamic smem starts with them)
    const uint32_t tile_bytes = 32u * S.stride;
    const uint32_t warps_total = gridDim.x * kFxWarps;
    auto issue = [&](uint32_t wt, uint32_t b) {  // lane 0: the tile's rows and flags into buffer b, two bulk copies on its barrier
        const uint32_t buf = buf0 + b * S.buf_bytes, bar = bar0 + 8u * b;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tile_bytes + 64u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(buf),
                     "l"(R.seq4 + (size_t)wt * tile_bytes), "r"(tile_bytes), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(buf + S.flags_off),
                     "l"(R.flags + (size_t)wt * 32), "r"(64u), "r"(bar) : "memory");
    };
    uint32_t wt = blockIdx.x * kFxWarps + wib;
    const uint32_t ahead = S.nbuf * warps_total;  // the tile requested when a buffer is released
    if (lane == 0) {
        if (wt < S.n_tiles) issue(wt, 0u);
        if (S.nbuf > 1u && wt + warps_total < S.n_tiles) issue(wt + warps_total, 1u);
    }
    uint32_t parity = 0;  // bit b: parity of the next completion of buffer b
    uint32_t cur = 0;
    const uint32_t row_off = (uint32_t)lane * S.stride;
    const int len = R.fixed_len;
    for (; wt < S.n_tiles; wt += warps_total) {
        const uint32_t buf = buf0 + cur * S.buf_bytes;
        {   // wait for the tile
            const uint32_t bar = bar0 + 8u * cur, ph = (parity >> cur) & 1u;
            uint32_t done;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar), "r"(ph) : "memory");
            } while (!done);
            parity ^= 1u << cur;
        }
        const uint32_t row = buf + row_off;  // this lane's row
        const uint32_t row_w = row & ~3u;
        const int row_sh = (int)(row & 3u) * 8;
        // ---- the read's first 32 bases as stored (forward reads only; the others are parked) ----
        const uint32_t fl = lds16(buf + S.flags_off + 2u * lane);
        uint32_t w0, w1, w2, w3;
        {
            const uint32_t a0 = lds32(row_w), a1 = lds32(row_w + 4), a2 = lds32(row_w + 8), a3 = lds32(row_w + 12), a4 = lds32(row_w + 16);
            w0 = __funnelshift_r(a0, a1, row_sh);
            w1 = __funnelshift_r(a1, a2, row_sh);
            w2 = __funnelshift_r(a2, a3, row_sh);
            w3 = __funnelshift_r(a3, a4, row_sh);
        }
        __syncwarp();  // every lane has lifted its row: the buffer may be refilled while this tile is processed
        if (wt + ahead < S.n_tiles && lane == 0) issue(wt + ahead, cur);
        cur = S.nbuf > 1u ? cur ^ 1u : 0u;
        const int r = (int)(wt * 32u) + lane;
        // parked: the location matcher applies (PM:194), the window is the reverse complement of the read's end (PM:41), or
        // some code of the window is zero ('=' or past the read's end)
        bool park = (!(fl & IDP_F_UNMAPPED) && ((fl & IDP_F_REVERSE) || S.park_mapped)) ||
                    (has_zero_nibble(w0) | has_zero_nibble(w1) | has_zero_nibble(w2) | has_zero_nibble(w3)) != 0u;
        const uint32_t e = fx_tab[fx_hash(w0, w1, P.fx_mul, P.fx_shift)];
        park = park || e == 0u;
        const uint32_t p = e ? e - 1u : 0u;
        const uint4 pm = pm4[p];
        const uint2 ru = rule[p];
        const int p_len = (int)(ru.x & 1023u), p_cap = (int)((ru.x >> 10) & 1023u), p_acc = (int)((ru.x >> 20) & 1023u);
        // numMismatches (PM:138-147): read mask not a subset of primer mask; every nibble folds to one flag, the four words
        // are interleaved on the way down (layout: fx_where_bit)
        uint32_t where;
        {
            const uint32_t x0 = w0 & ~pm.x, x1 = w1 & ~pm.y, x2 = w2 & ~pm.z, x3 = w3 & ~pm.w;
            const uint32_t y0 = (x0 | (x0 >> 2)) & 0x33333333u, y1 = (x1 | (x1 >> 2)) & 0x33333333u;
            const uint32_t y2 = (x2 | (x2 >> 2)) & 0x33333333u, y3 = (x3 | (x3 >> 2)) & 0x33333333u;
            const uint32_t u = y0 | (y1 << 2), v = y2 | (y3 << 2);
            where = ((u | (u >> 1)) & 0x55555555u) | (((v | (v >> 1)) & 0x55555555u) << 1);
        }
        // the slot may belong to another spelling, and only a clean 12-base prefix proves that no other primer can be accepted
        park = park || (where & kFxFirst12) != 0u || len < p_len;  // (a read shorter than Primer.length takes the double-precision path)
        const int mm = min(__popc(where), p_cap);  // the while loop stops counting once count > maxMismatches (PM:142)
        const bool accept = mm <= p_acc;           // mismatchAlign (PM:130-134), thresholds precomputed on the host
        if (S.k > 0) {
            // the k-mer rule for the one accepted primer (PM:108-116): a run of k literally equal letters exists iff a run of k
            // non-degenerate primer letters avoids the mismatch position (the window spells every matched A/C/G/T letter literally)
            const bool proved = where ? (ru.y & where) != 0u : (ru.x >> 30) != 0u;  // an accepted primer has at most one flag set
            park = park || (accept && !proved);  // the general kernel walks the k-mers
        }
        const bool matched = accept && !park;
        if (!park) {
            if (five.compact) {
                uint4 o;
                o.x = matched ? ((p + 1u) | ((uint32_t)IDP_UNGAPPED << 24)) : 0u;
                o.y = matched ? ((uint32_t)mm | ((uint32_t)IDP_NEXT16_NA << 16)) : 0u;
                o.z = 0u; o.w = 0u;
                reinterpret_cast<uint4 *>(five.p)[r] = o;
            } else {
                uint4 *dst = reinterpret_cast<uint4 *>(five.p) + 2 * (size_t)r;
                dst[0] = matched ? make_uint4(p, (uint32_t)IDP_UNGAPPED, (uint32_t)mm, (uint32_t)IDP_NEXT_NA) : make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
                dst[1] = make_uint4(0u, 0u, 0u, 0u);
            }
            rtype[r] = matched ? (uint8_t)IDP_UNGAPPED : (uint8_t)IDP_NONE;
        }
        // fall-through reads (IP:444-450) and parked reads, one bit each; compact_mask_kernel turns the masks into lists
        const bool gap = !park && !matched;
        const unsigned bg = __ballot_sync(0xffffffffu, gap), bp = __ballot_sync(0xffffffffu, park);
        if (lane == 0) { S.fall_mask[wt] = bg; S.park_mask[wt] = bp; }
    }
}

// Bit masks (one bit per read, one word per 32-read tile) -> index list.  One thread per tile, a block-wide prefix sum and ONE
// atomic per 256 tiles reserve the room; inside a block the list keeps read order.
// n_items_dev != NULL: the mask covers *n_items_dev bits (a count an earlier kernel of the stream left on the device), n_tiles is
// then only the upper bound the grid was sized for.
__global__ void __launch_bounds__(256) compact_mask_kernel(const uint32_t *__restrict__ mask, int n_tiles, unsigned int *__restrict__ list,
                                                           unsigned int *__restrict__ counter, const unsigned int *__restrict__ n_items_dev) {
    if (n_items_dev != nullptr) n_tiles = (int)((*n_items_dev + 31u) / 32u);
    __shared__ unsigned warp_sum[8];
    __shared__ unsigned block_base;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int base = blockIdx.x * 256; base < n_tiles; base += gridDim.x * 256) {
        const int t = base + (int)threadIdx.x;
        uint32_t m = t < n_tiles ? mask[t] : 0u;
        const unsigned cnt = __popc(m);
        unsigned incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
        if (lane == 31) warp_sum[wib] = incl;
        __syncthreads();
        unsigned before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const unsigned v = warp_sum[w]; if (w < wib) before += v; total += v; }
        if (threadIdx.x == 0) block_base = total ? atomicAdd(counter, total) : 0u;
        __syncthreads();
        unsigned pos = block_base + before + incl - cnt;
        while (m) {
            list[pos++] = (unsigned)t * 32u + (unsigned)(__ffs((int)m) - 1);
            m &= m - 1u;
        }
        __syncthreads();  // warp_sum / block_base are rewritten by the next round
    }
}

}  // namespace idp
