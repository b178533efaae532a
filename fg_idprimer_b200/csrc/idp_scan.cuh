// Kernel A+B: k-mer prefilter + location + ungapped IUPAC scan (included at the end of idp_kernels.cuh).
//
//   basesInSequencingOrder (PM:39-45) . LocationBasedPrimerMatcher.find (PM:194-213) . getPrimersForAlignmentTasks
//   (PM:98-116) . mismatchAlign / numMismatches (PM:128-147) . getBestUngappedAlignment (PM:237-248)
//
// This is the GENERAL scan kernel: it takes the whole batch when the exact-prefix pass (idp_scanfx.cuh) does not apply, else the
// list of reads that pass parked (list mode).  Persistent CTAs; every warp walks its own tiles of 32 reads.  On a whole batch a
// tile's packed bases (2400 B for 150 bp reads) arrive in shared memory by one TMA bulk copy (cp.async.bulk + mbarrier),
// double-buffered per warp, so no CTA-wide barrier exists after the prologue; in list mode every lane fetches its own read.  The
// panel side sits in shared memory as far as it fits.  member_check_kernel at the end of the file finishes the reads whose one
// accepted primer still needs the exact k-mer membership test (ScanLaunch::defer_mask).
//
// Two ways to find the accepted primers of a read, both exact:
//  * bit-sliced (DevPanel::bs_ok: every primer tolerates at most one mismatch): the read's first 12 bases index precomputed
//    32-primer mismatch bit-vectors; two bit planes (seen one / seen two mismatches) decide 32 primers per instruction with no
//    divergence.  The few survivors get the full mismatchAlign.  The k-mer rule is then applied to what was ACCEPTED: no
//    accepted primer -> no match whatever the candidates were; one -> it only has to be a candidate (a run of k equal
//    letters on the diagonal proves it: looked up in DevPanel::diag by the position of the mismatch, or tested on the
//    letters; else an exact membership walk); two or more -> the exact walk below decides order.
//  * exact walk (any thresholds): read k-mers in order, posting lists in the reference's Set iteration order, an 8-byte
//    posting head rejects most candidates, survivors are evaluated after the loop in first-hit order.
#pragma once
#include "idp_device.cuh"

namespace idp {

constexpr int kScanThreads = 512;
constexpr int kScanMinBlocks = 2;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kBsStride = 16;     // positions per group in the table (host layout)
constexpr int kBsPositions = 12;  // positions the filter actually looks at

struct ScanLaunch {
    int32_t stage_tile;        // 1: reads are fixed-length and 16-byte aligned: tiles are staged into shared memory by TMA
    uint32_t stride;           // bytes per read when fixed-length
    uint32_t tile_smem_bytes;  // shared-memory bytes of one warp tile buffer (incl. slack), multiple of 16
    int32_t stage_flags;       // 1: the tile's flag words ride along (16-byte aligned flags array), at flags_off in the tile buffer
    uint32_t flags_off;
    // byte offsets into dynamic shared memory, or 0xFFFFFFFF when the structure stays in global memory
    uint32_t off_direct, off_heads, off_pmask, off_pinfo, off_bsmm, off_bsacc, off_seed, off_diag;
    uint32_t direct_entries;
    // list mode: the kernel works through `list[0 .. ctr->n_parked)` (the reads scan_fx_kernel could not decide) instead of
    // reads 0 .. R.n; nothing is staged then, every lane fetches its own read
    const unsigned int *list;
    uint32_t *fall_mask;       // [ceil(R.n / 32)] bit i of word t: read 32 t + i fell through to the gapped stage (IP:444-450)
    // list mode only, may be NULL: bit i of word t = list item 32 t + i has ONE accepted primer whose k-mer rule neither the table
    // nor the diagonal proves.  The kernel then records that primer provisionally and member_check_kernel does the exact walk
    // with one read per lane (inside this kernel the walk ran on a lane or two of a warp while the others waited)
    uint32_t *defer_mask;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!done);
}

struct Best2 {  // getBestUngappedAlignment (PM:237-248) as a running stable top-2 on key = mm / Primer.length
    int bp = -1, bmm = 0, sp = -1, smm = 0;
    double bkey = 0.0, skey = 0.0;
    __device__ __forceinline__ void offer(int p, int mm, double key) {
        if (bp < 0 || key < bkey) { sp = bp; smm = bmm; skey = bkey; bp = p; bmm = mm; bkey = key; }
        else if (sp < 0 || key < skey) { sp = p; smm = mm; skey = key; }
    }
};

// ST: count base comparisons (idp_timings.scan_base_compares).  Off in the product build of the kernel: the counter and
// its 64-bit per-warp total cost registers the hot loop does not have (3 % of the kernel, measured)
template <int RW, bool ST>
struct ScanCtx {
    uint32_t win[RW];  // read prefix in sequencing order, base i in nibble i, zero beyond the read
    int len;
    unsigned compares;
    __device__ __forceinline__ void count(unsigned n) { if constexpr (ST) compares += n; }
    const uint32_t *pmask;  // shared or global
    const int4 *pinfo;
};

// numMismatches + mismatchAlign (PM:128-147).  Returns true and mm when the rate test passes.
// where: the mismatch flags of the first four words folded into one word, bit 4*(pos%8) + pos/8 (see DevPanel::diag)
template <int RW, bool ST>
__device__ __forceinline__ bool mismatch_align(const DevPanel &P, ScanCtx<RW, ST> &c, int p, int &mm_out, int &plen_out, uint32_t *where = nullptr) {
    const int4 info = c.pinfo[p];  // {Primer.length, cap, accept, sequence.length}
    const uint32_t *pm = c.pmask + (size_t)p * P.ww;
    int total = 0;
    uint32_t folded = 0;
#pragma unroll
    for (int w = 0; w < RW; ++w)
        if (w < P.ww) {
            const uint32_t z = nib_nonzero(c.win[w] & ~pm[w]);  // read mask not a subset of primer mask (PM:143)
            total += __popc(z);
            if (w < 4) folded |= z << w;
        }
    if (where != nullptr) *where = folded;
    c.count((unsigned)min(c.len, info.w));
    plen_out = info.x;
    if (c.len >= info.x) {
        const int mm = min(total, info.y);  // the while loop stops counting once count > maxMismatches (PM:142)
        mm_out = mm;
        return mm <= info.z;
    }
    // short read: length = min(bases.length, primer.length), in the JVM's double arithmetic (PM:130-134)
    const int length = c.len;
    const double max_mismatches = (double)length * P.max_mismatch_rate;
    const double capd = floor(max_mismatches) + 1.0;
    const int mm = min(total, capd > 1e9 ? 1000000000 : (int)capd);
    mm_out = mm;
    return __ddiv_rn((double)mm, (double)length) <= P.max_mismatch_rate;
}

template <int RW, bool ST>
__device__ __noinline__ void offer_candidate(const DevPanel &P, ScanCtx<RW, ST> &c, Best2 &best, int p) {
    // .distinct (PM:114): a primer already holding a top-2 slot is skipped; one that lost before loses again (strict
    // comparisons), so no visited set is needed
    if (p == best.bp || p == best.sp) return;
    int mm, plen;
    if (mismatch_align<RW>(P, c, p, mm, plen)) best.offer(p, mm, __ddiv_rn((double)mm, (double)plen));
}

// Candidates that survive the 8-base head filter wait here and are evaluated after the k-mer loop, all lanes together
// (in first-hit order, which is the order the reference visits them in).
template <int RW, bool ST>
struct Pending {
    int q0 = -1, q1 = -1, q2 = -1, q3 = -1;
    __device__ __forceinline__ void flush(const DevPanel &P, ScanCtx<RW, ST> &c, Best2 &best) {
        if (q0 >= 0) offer_candidate<RW>(P, c, best, q0);
        if (q1 >= 0) offer_candidate<RW>(P, c, best, q1);
        if (q2 >= 0) offer_candidate<RW>(P, c, best, q2);
        if (q3 >= 0) offer_candidate<RW>(P, c, best, q3);
        q0 = q1 = q2 = q3 = -1;
    }
    __device__ __forceinline__ void push(const DevPanel &P, ScanCtx<RW, ST> &c, Best2 &best, int p) {
        if (p == q0 || p == q1 || p == q2 || p == q3) return;
        if (q3 >= 0) flush(P, c, best);
        if (q0 < 0) q0 = p; else if (q1 < 0) q1 = p; else if (q2 < 0) q2 = p; else q3 = p;
    }
};

// posting list walk: the 8-byte head rejects a candidate whose first 8 bases already exceed the accept threshold
// (exact: the full count can only be larger, and a short read's threshold can only be smaller)
template <int RW, bool ST>
__device__ __forceinline__ void walk_postings(const DevPanel &P, ScanCtx<RW, ST> &c, Best2 &best, Pending<RW, ST> &pend, const uint2 *heads,
                                              uint32_t off, uint32_t cnt) {
    for (uint32_t j = 0; j < cnt; ++j) {
        const uint2 h = heads[off + j];
        const uint32_t acc1 = h.x >> 24;
        const uint32_t partial = __popc(nib_nonzero(c.win[0] & ~h.y));
        if (partial < acc1 || acc1 == 255u) pend.push(P, c, best, (int)(h.x & 0xFFFFFu));
    }
    c.count(8u * cnt);
}

// The reference's candidate walk (PM:108-116) with the accept test applied on the fly: read k-mers left to right, each
// k-mer's primers in Set iteration order, first occurrence wins ties.
template <int RW, bool ST>
__device__ __noinline__ void kmer_scan_exact(const DevPanel &P, ScanCtx<RW, ST> &c, Best2 &best, const uint32_t *direct, const uint2 *heads) {
    const DevKmerIndex &ix = P.kidx[0];
    const int n_bases = min(ix.window, c.len);  // k-mers start at 0 .. n_bases - k
    Pending<RW, ST> pend;
    uint32_t sr[RW];  // shift register over the window: the next base is always the low nibble of sr[0]
#pragma unroll
    for (int w = 0; w < RW; ++w) sr[w] = c.win[w];
    const uint64_t kmask4 = ix.k >= 16 ? ~0ull : ((1ull << (4 * ix.k)) - 1ull);
    const uint32_t kmask2 = ix.k >= 16 ? ~0u : ((1u << (2 * ix.k)) - 1u);
    uint64_t key4 = 0;
    uint32_t key2 = 0;
    int bad = 0;  // > 0 while the current k-mer holds a nibble that is not exactly one of A/C/G/T
    for (int i = 0; i < n_bases; ++i) {
        const uint32_t nib = sr[0] & 15u;
#pragma unroll
        for (int w = 0; w < RW - 1; ++w) sr[w] = __funnelshift_r(sr[w], sr[w + 1], 4);
        sr[RW - 1] >>= 4;
        key4 = ((key4 << 4) | nib) & kmask4;
        key2 = ((key2 << 2) | ((0x30210u >> (2 * nib)) & 3u)) & kmask2;  // A,C,G,T (1,2,4,8) -> 0..3
        bad = max(bad - 1, 0);
        if (__popc(nib) != 1) bad = ix.k;
        if (i < ix.k - 1) continue;
        if (direct != nullptr && bad == 0) {
            const uint32_t e = direct[key2];
            if (e) walk_postings<RW>(P, c, best, pend, heads, e & 0xFFFFFu, e >> 20);
        } else {  // literal equality of IUPAC letters (PM:113): the full table, keyed by 4 bits per base
            OffCnt oc;
            if (kmer_lookup(ix, key4, oc)) walk_postings<RW>(P, c, best, pend, direct != nullptr ? ix.heads : heads, oc.off, oc.cnt);
        }
    }
    pend.flush(P, c, best);
}

// Is primer p among getPrimersForAlignmentTasksWithCommonKmer(read) (PM:108-116)?  Exact, by walking the read's k-mers.
template <int RW, bool ST>
__device__ __noinline__ bool kmer_member_exact(const DevPanel &P, const ScanCtx<RW, ST> &c, int p) {
    const DevKmerIndex &ix = P.kidx[0];
    const int n_bases = min(ix.window, c.len);
    const uint64_t kmask4 = ix.k >= 16 ? ~0ull : ((1ull << (4 * ix.k)) - 1ull);
    uint64_t key4 = 0;
    for (int i = 0; i < n_bases; ++i) {
        uint32_t word = 0;
#pragma unroll
        for (int w = 0; w < RW; ++w) if ((i >> 3) == w) word = c.win[w];
        key4 = ((key4 << 4) | ((word >> (4 * (i & 7))) & 15u)) & kmask4;
        OffCnt oc;
        if (i >= ix.k - 1 && kmer_lookup(ix, key4, oc))
            for (uint32_t j = 0; j < oc.cnt; ++j) if ((int)__ldg(&ix.postings[oc.off + j]) == p) return true;
    }
    return false;
}

// A run of k literally equal letters at the same offset in read and primer (inside both the search window and the primer)
// proves that the primer shares a k-mer with the read.  Nibble codes are letters: equal code <=> equal letter.
template <int RW, bool ST>
__device__ __forceinline__ bool diagonal_kmer(const DevPanel &P, const ScanCtx<RW, ST> &c, int p, int seq_len, int k, int n_bases) {
    const uint32_t *pm = c.pmask + (size_t)p * P.ww;
    uint64_t eq = 0;
#pragma unroll
    for (int w = 0; w < RW; ++w)
        if (w < P.ww && w < 8) eq |= (uint64_t)gather_nibble_bits(~nib_nonzero(c.win[w] ^ pm[w]) & 0x11111111u) << (8 * w);
    const int lim = min(min(n_bases, seq_len), 64);
    if (lim < k) return false;
    if (lim < 64) eq &= (1ull << lim) - 1ull;
    uint64_t run = eq;
    for (int t = 1; t < k; ++t) run &= eq >> t;
    return run != 0ull;
}

template <int... I, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, I...>, F &&f) { (f(std::integral_constant<int, I>{}), ...); }
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f) { static_for_impl(std::make_integer_sequence<int, N>{}, f); }

// The bit-sliced filter for a compile-time number of 32-primer groups, tables in shared memory at a 128-byte aligned address
// `tab`.  The per-position table offsets (nibble * 8, OR-ed into the aligned base) are formed once per read; every lookup is
// then one 64-bit LDS with an immediate offset that serves two groups, and the two bit planes of a group cost two LOP3.
// Reports the number of survivors and the first one (lowest primer index); the caller re-runs the general loop only when a
// read has more than one survivor.
template <int G, int RW>
__device__ __forceinline__ void bs_filter_fixed(uint32_t tab, uint32_t acc, const uint32_t *win, int &n_surv, uint32_t &first_word, int &first_group) {
    uint32_t o[kBsPositions];
#pragma unroll
    for (int i = 0; i < kBsPositions; ++i) {
        const uint32_t w = win[(i >> 3) < RW ? (i >> 3) : 0];
        const int s = 4 * (i & 7);
        o[i] = ((s < 3 ? (w << (3 - s)) : (w >> (s - 3))) & 0x78u) | tab;
    }
    static_for<G / 2>([&](auto gc) {
        constexpr int gp = decltype(gc)::value;
        // four positions per step: "seen twice" gains the majority of (two new vectors, seen once) for each half of the
        // step, which is five 3-input logic operations per four positions instead of eight
        static_assert(kBsPositions % 4 == 0, "positions are taken four at a time");
        uint32_t ones0 = 0, twos0 = 0, ones1 = 0, twos1 = 0;
        static_for<kBsPositions / 4>([&](auto ic) {
            constexpr int i = 4 * decltype(ic)::value;
            uint32_t a0, a1, b0, b1, c0, c1, d0, d1;
            asm("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(a0), "=r"(a1) : "r"(o[i]), "n"((gp * kBsStride + i) * 128));
            asm("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(b0), "=r"(b1) : "r"(o[i + 1]), "n"((gp * kBsStride + i + 1) * 128));
            asm("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(c0), "=r"(c1) : "r"(o[i + 2]), "n"((gp * kBsStride + i + 2) * 128));
            asm("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(d0), "=r"(d1) : "r"(o[i + 3]), "n"((gp * kBsStride + i + 3) * 128));
            const uint32_t m0 = (a0 & b0) | (ones0 & (a0 | b0)), h0 = ones0 | a0 | b0;
            const uint32_t n0 = (c0 & d0) | (h0 & (c0 | d0));
            ones0 = h0 | c0 | d0;
            twos0 = twos0 | m0 | n0;
            const uint32_t m1 = (a1 & b1) | (ones1 & (a1 | b1)), h1 = ones1 | a1 | b1;
            const uint32_t n1 = (c1 & d1) | (h1 & (c1 | d1));
            ones1 = h1 | c1 | d1;
            twos1 = twos1 | m1 | n1;
        });
        uint32_t ax0, ay0, ax1, ay1;
        asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(ax0), "=r"(ay0), "=r"(ax1), "=r"(ay1) : "r"(acc), "n"(gp * 16));
        // at most acc (0 or 1) mismatches in the positions looked at
        const uint32_t surv0 = ax0 & ~twos0 & (~ones0 | ay0), surv1 = ax1 & ~twos1 & (~ones1 | ay1);
        const bool first0 = surv0 != 0u && n_surv == 0;
        first_word = first0 ? surv0 : first_word;
        first_group = first0 ? 2 * gp : first_group;
        n_surv += __popc(surv0);
        const bool first1 = surv1 != 0u && n_surv == 0;
        first_word = first1 ? surv1 : first_word;
        first_group = first1 ? 2 * gp + 1 : first_group;
        n_surv += __popc(surv1);
    });
}

// Seed lookup (DevPanel::seed_ok): the read's two leading S-base seeds are hashed into the panel's seed tables; every
// listed primer is checked on the first 2*S bases (at most one mismatch) and the survivors are kept in ascending primer
// order (file order) in s0..s3, INT_MAX = empty.  Exact for every read whose first 2*S codes are non-zero: an accepted
// primer has at most one mismatch, so it matches one seed without a mismatch and is listed under the read's spelling of
// it.  Returns false when the read has to take the general loop instead (a zero code in the head: a read shorter than
// 2*S or the '=' code; or more than four survivors).
template <int RW, bool ST>
__device__ __forceinline__ bool seed_filter(const DevPanel &P, ScanCtx<RW, ST> &c, const uint32_t *tab, int &s0, int &s1, int &s2, int &s3) {
    const bool six = P.seed_S == 6;
    const uint32_t w0 = c.win[0];
    const uint32_t h1 = RW > 1 ? (six ? (c.win[RW > 1 ? 1 : 0] & 0xFFFFu) : c.win[RW > 1 ? 1 : 0]) : 0u;  // head = w0 + h1
    if (nib_nonzero(w0) != 0x11111111u || nib_nonzero(h1) != (six ? 0x1111u : 0x11111111u)) return false;
    const uint32_t key0 = six ? (w0 & 0xFFFFFFu) : w0;
    const uint32_t key1 = six ? ((w0 >> 24) | (h1 << 8)) : h1;
    s0 = s1 = s2 = s3 = INT_MAX;
    bool fits = true;
    unsigned verified = 0;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const uint32_t e = tab[(uint32_t)t * (uint32_t)P.seed_slots + (((t ? key1 : key0) * 0x9E3779B1u) >> P.seed_shift)];
        const uint32_t cnt = e & 0xFFu, val = e >> 8;
        for (uint32_t j = 0; j < cnt; ++j) {
            const int p = cnt == 1u ? (int)val : (int)__ldg(&P.seed_list[val + j]);
            if (p == s0 || p == s1 || p == s2 || p == s3) continue;
            const uint32_t *pm = c.pmask + (size_t)p * P.ww;
            ++verified;
            if (__popc(nib_nonzero(w0 & ~pm[0])) + __popc(nib_nonzero(h1 & ~pm[1])) > 1) continue;
            int x = p, y;  // sorted insertion; whatever falls off the end did not fit
            y = s0; s0 = min(x, y); x = max(x, y);
            y = s1; s1 = min(x, y); x = max(x, y);
            y = s2; s2 = min(x, y); x = max(x, y);
            y = s3; s3 = min(x, y); x = max(x, y);
            fits = fits && x == INT_MAX;
        }
    }
    c.count(verified * 2u * (unsigned)P.seed_S);
    return fits;
}

// ONCHIP: the read tiles are known at compile time to sit in shared memory, so the window loads are LDS rather than generic
// loads; the unrolled bit-sliced filter addresses its tables as shared memory as well (the general loop does not)
template <int RW, bool FAST, bool ONCHIP, bool ST>
__global__ void __launch_bounds__(kScanThreads, kScanMinBlocks) scan_kernel(const __grid_constant__ DevPanel P, const __grid_constant__ DevReads R, const __grid_constant__ ScanLaunch S, const ResultOut five,
                                                               uint8_t *__restrict__ rtype, Counters *__restrict__ ctr) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t tile_bar[kScanWarps][2];
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const DevKmerIndex &ix = P.kidx[0];
    // ---- panel-side tables into shared memory (once per persistent CTA) ----
    const uint32_t *direct = ix.direct;
    const uint2 *heads = ix.heads;
    const uint32_t *bs_mm = P.bs_mm;
    const uint2 *bs_acc = P.bs_acc;
    ScanCtx<RW, ST> c;
    c.pmask = P.pmask;
    c.pinfo = P.pinfo;
    if (!FAST && S.off_direct != 0xFFFFFFFFu) {
        uint32_t *d = reinterpret_cast<uint32_t *>(smem + S.off_direct);
        for (uint32_t i = tid; i < S.direct_entries; i += blockDim.x) d[i] = ix.direct[i];
        direct = d;
    }
    if (!FAST && S.off_heads != 0xFFFFFFFFu) {
        uint2 *d = reinterpret_cast<uint2 *>(smem + S.off_heads);
        for (uint32_t i = tid; i < ix.n_postings; i += blockDim.x) d[i] = ix.heads[i];
        heads = d;
    }
    if (S.off_pmask != 0xFFFFFFFFu) {
        uint32_t *d = reinterpret_cast<uint32_t *>(smem + S.off_pmask);
        for (int i = tid; i < P.n_primers * P.ww; i += blockDim.x) d[i] = P.pmask[i];
        c.pmask = d;
    }
    if (S.off_pinfo != 0xFFFFFFFFu) {
        int4 *d = reinterpret_cast<int4 *>(smem + S.off_pinfo);
        for (int i = tid; i < P.n_primers; i += blockDim.x) d[i] = P.pinfo[i];
        c.pinfo = d;
    }
    if (FAST && S.off_bsmm != 0xFFFFFFFFu) {
        uint32_t *d = reinterpret_cast<uint32_t *>(smem + S.off_bsmm);
        for (int i = tid; i < P.bs_groups * kBsStride * 16; i += blockDim.x) d[i] = P.bs_mm[i];
        bs_mm = d;
    }
    if (FAST && S.off_bsacc != 0xFFFFFFFFu) {
        uint2 *d = reinterpret_cast<uint2 *>(smem + S.off_bsacc);
        for (int i = tid; i < P.bs_groups; i += blockDim.x) d[i] = P.bs_acc[i];
        bs_acc = d;
    }
    const uint2 *diag = P.diag;
    if (FAST && P.diag != nullptr && S.off_diag != 0xFFFFFFFFu) {
        uint2 *d = reinterpret_cast<uint2 *>(smem + S.off_diag);
        for (int i = tid; i < P.n_primers; i += blockDim.x) d[i] = P.diag[i];
        diag = d;
    }
    const uint32_t *seed_tab = P.seed_tab;
    if (FAST && P.seed_ok && S.off_seed != 0xFFFFFFFFu) {
        uint32_t *d = reinterpret_cast<uint32_t *>(smem + S.off_seed);
        for (int i = tid; i < 2 * P.seed_slots; i += blockDim.x) d[i] = P.seed_tab[i];
        seed_tab = d;
    }
    const bool bs_onchip = ONCHIP && S.off_bsmm != 0xFFFFFFFFu && S.off_bsacc != 0xFFFFFFFFu;  // large panels keep them in global memory
    const uint32_t tab_addr = bs_onchip ? smem_u32(smem) + S.off_bsmm : 0u, acc_addr = bs_onchip ? smem_u32(smem) + S.off_bsacc : 0u;
    const bool fixed_groups = bs_onchip && (tab_addr & 127u) == 0u && (acc_addr & 15u) == 0u && P.bs_groups >= 2 && P.bs_groups <= 8 && (P.bs_groups & 1) == 0;
    if (S.stage_tile && lane == 0) {
        mbar_init(&tile_bar[wib][0], 1);
        mbar_init(&tile_bar[wib][1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // ---- every warp walks its own 32-read tiles: no CTA-wide barrier from here on ----
    auto wbuf = [&](int b) { return smem + (size_t)(2 * wib + b) * S.tile_smem_bytes; };  // two tile buffers per warp
    uint32_t phase = 0u;  // bit b: parity of the next completion of buffer b
    const int warps_total = gridDim.x * kScanWarps;
    unsigned long long compares_total = 0;
    auto issue_tile = [&](int wt, int b) {  // this warp's 32 reads: one TMA bulk copy + a tail shorter than 16 bytes
        const size_t g0 = (size_t)wt * 32 * S.stride;
        const uint32_t bytes = (uint32_t)min(32, R.n - wt * 32) * S.stride;
        const uint32_t bulk = bytes & ~15u;
        const uint32_t fbytes = S.stage_flags ? (uint32_t)min(32, R.n - wt * 32) * 2u : 0u, fbulk = fbytes & ~15u;
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&tile_bar[wib][b], bulk + fbulk);
            if (bulk) tma_load_1d(wbuf(b), R.seq4 + g0, bulk, &tile_bar[wib][b]);
            if (fbulk) tma_load_1d(wbuf(b) + S.flags_off, R.flags + (size_t)wt * 32, fbulk, &tile_bar[wib][b]);
        }
        if (bytes != bulk && lane < (int)(bytes - bulk)) wbuf(b)[bulk + lane] = R.seq4[g0 + bulk + lane];  // only the last tile has tails
        if (fbytes != fbulk && 2u * lane >= fbulk && 2u * lane < fbytes)
            reinterpret_cast<uint16_t *>(wbuf(b) + S.flags_off)[lane] = R.flags[(size_t)wt * 32 + lane];
    };
    int cur = 0;
    const int n_items = S.list != nullptr ? (int)ctr->n_parked : R.n;
    const int n_wtiles = (n_items + 31) / 32;
    const int wt0 = blockIdx.x * kScanWarps + wib;
    if (S.stage_tile && wt0 < n_wtiles) issue_tile(wt0, 0);
    const bool flags_staged = S.stage_tile && S.stage_flags;
    for (int wt = wt0; wt < n_wtiles; wt += warps_total) {
        if (S.stage_tile) {
            if (wt + warps_total < n_wtiles) issue_tile(wt + warps_total, cur ^ 1);  // prefetch; that buffer was released below
            mbar_wait(&tile_bar[wib][cur], (phase >> cur) & 1u);
            phase ^= 1u << cur;
            __syncwarp();  // tail bytes visible
        }
        {
            const int item = wt * 32 + lane;
            const bool active = item < n_items;
            int r = item;
            uint16_t fl = 0;
            bool search = false;   // this lane still needs the ungapped matcher
            bool defer = false;    // the record written below is provisional (ScanLaunch::defer_mask)
            idp_result res = none_result();
            c.compares = 0;
            if (active) {
                if (S.list != nullptr) r = (int)S.list[item];
                fl = flags_staged ? reinterpret_cast<const uint16_t *>(smem + (size_t)(2 * wib + cur) * S.tile_smem_bytes + S.flags_off)[lane]
                                  : R.flags[r];
                const uint8_t *s;
                read_geom(R, r, s, c.len);
                const bool rc = seq_order_is_rc(fl);
                if (ONCHIP && S.stage_tile) {  // index the shared array itself (16-byte aligned) so the loads stay LDS; the buffer has slack past what is read
                    const uint32_t bo = (uint32_t)(2 * wib + cur) * S.tile_smem_bytes + (uint32_t)lane * S.stride;
                    if (!rc || c.len >= 8 * RW) window_from_words<RW>(reinterpret_cast<const uint32_t *>(smem) + (bo >> 2), (int)(bo & 3u), c.len, rc, c.win);
                    else window_by_bases<RW>(smem + bo, c.len, rc, c.win);
                } else if (S.stage_tile) {  // the same through a generic pointer
                    const uintptr_t a = reinterpret_cast<uintptr_t>(wbuf(cur) + (size_t)lane * S.stride);
                    if (!rc || c.len >= 8 * RW) window_from_words<RW>(reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3)), (int)(a & 3), c.len, rc, c.win);
                    else window_by_bases<RW>(reinterpret_cast<const uint8_t *>(a), c.len, rc, c.win);
                } else {
                    window_from_global<RW>(R, r, s, c.len, rc, c.win);
                }
                // ---- LocationBasedPrimerMatcher.find (PM:194-213) ----
                if (!(fl & IDP_F_UNMAPPED) && R.ref_idx != nullptr) {
                    const int strand = (fl & IDP_F_REVERSE) ? 1 : 0;
                    const int nl = P.n_loc[strand];
                    const int pos = strand ? R.uend[r] : R.ustart[r];  // PM:180, 186
                    const int ref = R.ref_idx[r];
                    if (nl > 0 && ref >= 0 && pos + P.slop >= 1) {
                        const int64_t lo = ((int64_t)ref << 32) | (uint32_t)max(pos - P.slop, 0);
                        const int64_t hi = ((int64_t)ref << 32) | (uint32_t)(pos + P.slop);
                        const int64_t *keys = P.loc_sortkey[strand];
                        int a = 0, b = nl;
                        while (a < b) { const int m = (a + b) >> 1; if (__ldg(&keys[m]) < lo) a = m + 1; else b = m; }
                        // exactly one primer within slop, else "well I give up" (PM:198-206)
                        if (a < nl && __ldg(&keys[a]) <= hi && !(a + 1 < nl && __ldg(&keys[a + 1]) <= hi)) {
                            const int p = __ldg(&P.loc_primer[strand][a]);
                            int mm, plen;
                            if (mismatch_align<RW>(P, c, p, mm, plen)) { res.primer = p; res.type = IDP_LOCATION; res.mm = mm; }
                        }
                    }
                }
                // UngappedAlignmentBasedPrimerMatcher.find (PM:230-248); a read shorter than k has no candidates (PM:109)
                search = res.type == IDP_NONE && !(ix.k > 0 && c.len < ix.k);
            }
            if (search) {
                Best2 best;
                // the out-of-line walks get COPIES: an object whose address reaches a call lives in local memory for the
                // whole kernel, and the read window / running best are touched by every instruction of the hot path
                auto scan_exact = [&](const uint32_t *dir, const uint2 *hd) {
                    ScanCtx<RW, ST> cc = c;
                    Best2 bb = best;
                    kmer_scan_exact<RW>(P, cc, bb, dir, hd);
                    best = bb;
                    if constexpr (ST) c.compares = cc.compares;
                };
                auto member_exact = [&](int p) {
                    const ScanCtx<RW, ST> cc = c;
                    return kmer_member_exact<RW>(P, cc, p);
                };
                if (FAST) {
                    // survivors of the 12-base filter, ascending primer index (file order)
                    int n_acc = 0, a_p = -1, a_mm = 0, a_plen = 1, a_seq = 0;
                    uint32_t a_where = 0;
                    int q0 = -1, q1 = -1, q2 = -1, q3 = -1;
                    auto evaluate = [&](int p) {
                        int mm, plen;
                        uint32_t where;
                        if (mismatch_align<RW>(P, c, p, mm, plen, &where)) {
                            if (ix.k <= 0) best.offer(p, mm, __ddiv_rn((double)mm, (double)plen));  // PM:100: file order
                            else if (n_acc == 0) { a_p = p; a_mm = mm; a_plen = plen; a_seq = c.pinfo[p].w; a_where = where; }
                            ++n_acc;
                        }
                    };
                    auto flush = [&]() {
                        if (q0 >= 0) evaluate(q0);
                        if (q1 >= 0) evaluate(q1);
                        if (q2 >= 0) evaluate(q2);
                        if (q3 >= 0) evaluate(q3);
                        q0 = q1 = q2 = q3 = -1;
                    };
                    int n_surv = 2;  // "unknown": the general loop below decides
                    bool seeded = false;
                    if (P.seed_ok) {
                        int s0, s1, s2, s3;
                        if (seed_filter<RW>(P, c, seed_tab, s0, s1, s2, s3)) {
                            seeded = true;
                            n_surv = 0;
                            q0 = s0 == INT_MAX ? -1 : s0; q1 = s1 == INT_MAX ? -1 : s1;
                            q2 = s2 == INT_MAX ? -1 : s2; q3 = s3 == INT_MAX ? -1 : s3;
                        }
                    }
                    if (!seeded && ONCHIP && fixed_groups) {
                        uint32_t first_word = 0;
                        int first_group = 0;
                        n_surv = 0;
                        switch (P.bs_groups) {
                            case 2: bs_filter_fixed<2, RW>(tab_addr, acc_addr, c.win, n_surv, first_word, first_group); break;
                            case 4: bs_filter_fixed<4, RW>(tab_addr, acc_addr, c.win, n_surv, first_word, first_group); break;
                            case 6: bs_filter_fixed<6, RW>(tab_addr, acc_addr, c.win, n_surv, first_word, first_group); break;
                            default: bs_filter_fixed<8, RW>(tab_addr, acc_addr, c.win, n_surv, first_word, first_group); break;
                        }
                        if (n_surv == 1) q0 = first_group * 32 + __ffs(first_word) - 1;
                    }
                    for (int g = 0; n_surv >= 2 && g < P.bs_groups; ++g) {
                        const uint32_t *row = bs_mm + (size_t)(g >> 1) * (kBsStride * 32) + (g & 1);  // group pairs are interleaved
                        uint32_t ones = 0, twos = 0;
#pragma unroll
                        for (int i = 0; i < kBsPositions; ++i) {
                            const uint32_t nib = (c.win[(i >> 3) < RW ? (i >> 3) : 0] >> (4 * (i & 7))) & 15u;
                            const uint32_t v = row[(i * 16 + nib) * 2];
                            twos |= ones & v;
                            ones |= v;
                        }
                        const uint2 a = bs_acc[g];
                        uint32_t surv = a.x & ~twos & (~ones | a.y);  // at most acc (0 or 1) mismatches in the first 12 bases
                        while (surv) {
                            const int p = g * 32 + __ffs(surv) - 1;
                            surv &= surv - 1;
                            if (q3 >= 0) flush();
                            if (q0 < 0) q0 = p; else if (q1 < 0) q1 = p; else if (q2 < 0) q2 = p; else q3 = p;
                        }
                    }
                    if (!seeded) c.count((unsigned)(P.n_primers * kBsPositions));
                    flush();
                    if (ix.k > 0) {
                        if (n_acc == 1) {
                            const int n_bases = min(ix.window, c.len);
                            bool proved;
                            // a window without a zero code lies inside the read and spells every matched A/C/G/T
                            // primer letter literally; the table then knows whether a k-run avoids the one mismatch
                            bool full = diag != nullptr && (a_where & (a_where - 1u)) == 0u;
#pragma unroll
                            for (int w = 0; w < RW; ++w) full = full && (((c.win[w] - 0x11111111u) & ~c.win[w] & 0x88888888u) == 0u);
                            if (full) {
                                const uint2 d = diag[a_p];
                                proved = a_where ? ((d.x >> (__ffs((int)a_where) - 1)) & 1u) != 0u : d.y != 0u;
                            } else {
                                proved = diagonal_kmer<RW>(P, c, a_p, a_seq, ix.k, n_bases);
                            }
                            if (!proved && S.defer_mask != nullptr) { defer = true; proved = true; }
                            if (proved || member_exact(a_p))
                                best.offer(a_p, a_mm, 0.0);  // the only accepted candidate: its sort key is never compared
                        } else if (n_acc >= 2) {
                            scan_exact(ix.direct, ix.heads);  // ties: the reference's candidate order decides
                        }
                    }
                } else if (ix.k <= 0) {  // PM:100: every primer, file order
                    for (int p = 0; p < P.n_primers; ++p) {
                        int mm, plen;
                        if (mismatch_align<RW>(P, c, p, mm, plen)) best.offer(p, mm, __ddiv_rn((double)mm, (double)plen));
                    }
                } else {
                    scan_exact(direct, heads);
                }
                if (best.bp >= 0) {
                    res.primer = best.bp; res.type = IDP_UNGAPPED; res.mm = best.bmm;
                    res.next_mm = best.sp >= 0 ? best.smm : IDP_NEXT_NA;                   // PM:244-245
                    if (!(res.mm <= res.next_mm)) res.status = IDP_STATUS_REQ_MM_LE_NEXT;  // PMa:102
                }
            }
            bool need_gapped = false;
            if (active) {
                need_gapped = res.type == IDP_NONE;
                write_result(five, r, res);
                rtype[r] = res.type;  // what metrics_kernel reads: one byte instead of the whole record
            }
            // fall-through reads (IP:444-450): one bit each; compact_mask_kernel turns the mask into the list
            if (S.list == nullptr) {
                const unsigned ballot = __ballot_sync(0xffffffffu, need_gapped);
                if (lane == 0) S.fall_mask[wt] = ballot;
            } else {
                if (need_gapped) atomicOr(&S.fall_mask[r >> 5], 1u << (r & 31));  // result unused: a fire-and-forget reduction
                if (S.defer_mask != nullptr) {
                    const unsigned ballot = __ballot_sync(0xffffffffu, defer);
                    if (lane == 0) S.defer_mask[wt] = ballot;
                }
            }
            if constexpr (ST) compares_total += c.compares;
        }
        __syncwarp();  // all lanes are done with this tile buffer: the next iteration may overwrite it
        cur ^= 1;
    }
    // statistics: base comparisons performed, one atomic per warp
    if constexpr (ST) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) compares_total += __shfl_xor_sync(0xffffffffu, compares_total, d);
        if (lane == 0 && compares_total) atomicAdd(&ctr->base_compares, compares_total);
    }
}

// The same question without the index: primer p is a candidate iff one of the read's k-mers (search window, PM:110) equals one of
// the primer's, letter by letter (PM:113).  Every read k-mer is first tested against the primer's 1,024 hash bits
// (DevPanel::member_bits: about 2 % of them set, independent loads from a 128-byte row that stays in L1); only a hit is compared
// with the primer's k-mers themselves.  The index walk is a chain of ~3 dependent table loads per read k-mer.
template <int RW>
__device__ __forceinline__ bool kmer_member_bits(const DevPanel &P, const uint32_t (&win)[RW], int len, int p) {
    const DevKmerIndex &ix = P.kidx[0];
    const int k = ix.k, n_bases = min(min(ix.window, len), 8 * RW), seq_len = __ldg(&P.pinfo[p]).w;
    if (k <= 0 || n_bases < k || seq_len < k) return false;
    const uint64_t kmask4 = k >= 16 ? ~0ull : ((1ull << (4 * k)) - 1ull);
    const uint32_t *bits = P.member_bits + (size_t)p * kMemberWords;
    uint64_t key = 0;
    for (int i = 0; i < n_bases; ++i) {
        uint32_t word = 0;
#pragma unroll
        for (int w = 0; w < RW; ++w) if ((i >> 3) == w) word = win[w];
        key = ((key << 4) | ((word >> (4 * (i & 7))) & 15u)) & kmask4;
        if (i < k - 1) continue;
        const uint32_t h = member_bits_hash(key);
        if (!((__ldg(&bits[h >> 5]) >> (h & 31u)) & 1u)) continue;
        uint64_t pk = 0;  // a hit: the primer's k-mers themselves
        for (int b = 0; b < seq_len; ++b) {
            pk = ((pk << 4) | ((__ldg(&P.pmask[(size_t)p * P.ww + (b >> 3)]) >> (4 * (b & 7))) & 15u)) & kmask4;
            if (b >= k - 1 && pk == key) return true;
        }
    }
    return false;
}

// The reads scan_kernel deferred (ScanLaunch::defer_mask): is the provisionally recorded primer among the read's k-mer
// candidates (PM:108-116)?  One read per thread, every lane busy with the same walk.  No: the record goes back to "no match"
// and the read joins the fall-through reads.
template <int RW>
__global__ void __launch_bounds__(128) member_check_kernel(DevPanel P, DevReads R, const unsigned int *__restrict__ parked,
                                                           const unsigned int *__restrict__ items, const Counters *__restrict__ ctr, ResultOut five,
                                                           uint8_t *__restrict__ rtype, uint32_t *__restrict__ fall_mask) {
    const unsigned n = ctr->n_defer;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = (int)parked[items[i]];
        ScanCtx<RW, false> c;
        c.compares = 0; c.pmask = nullptr; c.pinfo = nullptr;
        const uint8_t *s;
        read_geom(R, r, s, c.len);
        window_from_global<RW>(R, r, s, c.len, seq_order_is_rc(R.flags[r]), c.win);
        int type, primer;
        read_type_primer(five, r, type, primer);
        // (inlined: an out-of-line call would take the address of the kernel's parameter block, which then costs every thread of
        // the launch a local-memory copy of it; the launcher defers only when DevPanel::member_bits exists)
        if (!kmer_member_bits<RW>(P, c.win, c.len, primer)) {
            write_result(five, r, none_result());
            rtype[r] = (uint8_t)IDP_NONE;
            atomicOr(&fall_mask[r >> 5], 1u << (r & 31));
        }
    }
}

}  // namespace idp
