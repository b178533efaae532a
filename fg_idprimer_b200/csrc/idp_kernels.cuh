// CUDA kernels of the IdentifyPrimers matching path (sm_100a).  Host launchers are in idp_ctx.cu.
//
// Data model on the device
//   reads   : BAM-native 4-bit bases, SoA (DevReads); never unpacked to ASCII.
//   panel   : DevPanel (idp_internal.h) -- primer IUPAC masks 8 bases / 32-bit word, k-mer hash tables whose posting
//             lists are in the reference's Set iteration order, per-strand sorted location arrays.
//   results : idp_result (32 B) or idp_result16 (16 B) per read (ResultOut, idp_device.cuh).
//
// Kernels (reference function each one covers)
//   scan_fx_kernel (idp_scanfx.cuh), scan_kernel / member_check_kernel (idp_scan.cuh)
//                      basesInSequencingOrder (PM:39-45) + LocationBasedPrimerMatcher.find (PM:194-213) +
//                      getPrimersForAlignmentTasks (PM:98-116) + mismatchAlign / numMismatches (PM:128-147) +
//                      getBestUngappedAlignment (PM:237-248); reads with no match go to the fall-through list
//   cand5_thread_kernel / cand5_kernel   GappedAlignmentBasedPrimerMatcher.toAlignmentTasks candidates under -K
//                      (PM:287-296, 108-116): a thread per read with a register-resident distinct list, a warp per read
//                      with a shared-memory bitmap for the reads that overflow it
//   cand3_kernel       matchThreePrime's primersToCheck (IP:490-499)
//   sw_pairs_kernel    fgbio Aligner.align of the (read, primer) pairs (IP:446-448, 509-513), score + targetStart / targetEnd in
//                      one pass: two pairs per thread in 16-bit half-words (sw_trace2_glocal16 / sw_trace2_local16) where the
//                      scores allow it, else one pair per thread on 32-bit values (sw_trace_reg) or the generic forms
//   sw_all_kernel      the same, score only, when every primer is a candidate: one block per read, best two kept on chip
//   finalize_*_kernel  getBestGappedAlignment (PM:305-328) + (all-primers path) re-alignment of the winner with start
//                      tracking + finalizedGappedAlignment (IP:459-471)
//   metrics_kernel     TemplateType + TemplateTypeMetric counter (TemplateType.scala:52-63, IP:413)
#pragma once
#include "idp_device.cuh"

namespace idp {

// ---------------------------------------------------------------------------------------------------------------
// Candidate emission for the 5' gapped stage under -K, common case: one THREAD per fall-through read keeps up to
// kCandLocal distinct candidates in registers (first-hit order, PM:112-114).  Reads with more go to the spill list and
// are handled by the warp kernel below.  Thousands of independent lookup chains in flight hide the table latency.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kCandLocal = 8;

// bm_words > 0: the presence bits of the direct table (DevKmerIndex::direct_bits, at most 128 KB) sit in shared memory and
// answer most lookups -- the fall-through reads are mostly off-target templates whose k-mers are not on the panel, and the
// 4^k-entry table itself lives in L2, where one 4-byte probe costs a whole sector (the kernel was bound by exactly that).
constexpr int kCand5Threads = 1024;
__device__ __forceinline__ uint32_t smem_ld32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__global__ void __launch_bounds__(kCand5Threads) cand5_thread_kernel(DevPanel P, DevReads R, const unsigned int *__restrict__ fall,
                                                           Counters *__restrict__ ctr, unsigned int *__restrict__ pair_work,
                                                           unsigned int *__restrict__ pair_primer, unsigned int pair_cap,
                                                           unsigned int *__restrict__ seg_off, int *__restrict__ seg_cnt,
                                                           unsigned int *__restrict__ spill, unsigned int bm_words) {
    extern __shared__ uint32_t cand5_bm[];
    const DevKmerIndex &ix = P.kidx[1];
    const unsigned n_work = ctr->n_fall;
    if (bm_words) {
        if (blockIdx.x * blockDim.x >= n_work) return;  // nothing for this block: skip the copy
        const uint4 *src = reinterpret_cast<const uint4 *>(ix.direct_bits);
        uint4 *dst = reinterpret_cast<uint4 *>(cand5_bm);
        for (unsigned i = threadIdx.x; i < bm_words / 4; i += blockDim.x) dst[i] = src[i];
        for (unsigned i = (bm_words & ~3u) + threadIdx.x; i < bm_words; i += blockDim.x) cand5_bm[i] = ix.direct_bits[i];
        for (unsigned i = threadIdx.x; i < 2048u; i += blockDim.x) cand5_bm[bm_words + i] = ix.iupac_bloom[i];
        __syncthreads();
    }
    const uint32_t *bloom = cand5_bm + bm_words;
    const int lane = threadIdx.x & 31;
    const uint64_t kmask = ix.k >= 16 ? ~0ull : ((1ull << (4 * ix.k)) - 1ull);
    const uint32_t kmask2 = ix.k >= 16 ? ~0u : ((1u << (2 * ix.k)) - 1u);
    for (unsigned base = blockIdx.x * blockDim.x; base < n_work; base += gridDim.x * blockDim.x) {
        const unsigned w = base + threadIdx.x;
        const bool active = w < n_work;
        int cand[kCandLocal];
        int nc = 0;
        bool spilled = false;
        if (active) {
            const int r = (int)fall[w];
            const uint8_t *s; int len;
            read_geom(R, r, s, len);
            const bool rc = seq_order_is_rc(R.flags[r]);
            TargetStream ts(Oriented{s, len, rc});
            const int n_bases = len >= ix.k ? min(ix.window, len) : 0;  // k-mers start at 0 .. n_bases - k (PM:110)
            uint64_t key = 0;
            uint32_t key2 = 0;
            int bad = 0;  // > 0 while the current k-mer holds a code that is not exactly one of A/C/G/T
            auto visit = [&](uint32_t nib, int i) {  // one more base of the window; false = the candidate list overflowed
                key = ((key << 4) | nib) & kmask;
                key2 = ((key2 << 2) | ((0x30210u >> (2 * nib)) & 3u)) & kmask2;  // A,C,G,T (1,2,4,8) -> 0..3
                bad = max(bad - 1, 0);
                if (__popc(nib) != 1) bad = ix.k;
                OffCnt oc;
                if (i < ix.k - 1) return true;
                if (ix.direct != nullptr && bad == 0) {  // one load, no probing: the lanes of a warp stay together
                    const uint32_t e = __ldg(&ix.direct[key2]);
                    if (!e) return true;
                    oc.off = e & 0xFFFFFu; oc.cnt = e >> 20;
                } else if (!kmer_lookup(ix, key, oc)) return true;
                for (uint32_t j = 0; j < oc.cnt; ++j) {
                    const int p = (int)__ldg(&ix.postings[oc.off + j]);
                    bool seen = false;
#pragma unroll
                    for (int t = 0; t < kCandLocal; ++t) seen |= (t < nc && cand[t] == p);
                    if (seen) continue;
                    if (nc == kCandLocal) return false;
#pragma unroll
                    for (int t = 0; t < kCandLocal; ++t) if (t == nc) cand[t] = p;
                    ++nc;
                }
                return true;
            };
            if (n_bases <= 32 && bm_words) {
                // Bit planes of the window (DevKmerIndex::direct_lohi): lo / hi = the two bits of every base's A/C/G/T code, badv = the
                // base is not exactly one of A/C/G/T (incl. the zero codes past the read's end).  A k-mer then costs two shifts, two
                // ANDs and a shift-add for its table index, one AND for "all A/C/G/T", and one shared-memory bit test that
                // answers most lookups (the fall-through reads are mostly off-target: their k-mers are not on the panel).
                uint32_t sr[4];
                window_from_global<4>(R, r, s, len, rc, sr);
                uint32_t lo = 0, hi = 0, badv = 0;
#pragma unroll
                for (int w4 = 0; w4 < 4; ++w4) {
                    const uint32_t x = sr[w4];
                    const uint32_t c2 = (x & 0x55555555u) + ((x >> 1) & 0x55555555u);
                    const uint32_t d4 = (c2 & 0x33333333u) + ((c2 >> 2) & 0x33333333u);  // bits set per nibble
                    lo |= gather_nibble_bits(((x >> 1) | (x >> 3)) & 0x11111111u) << (8 * w4);
                    hi |= gather_nibble_bits(((x >> 2) | (x >> 3)) & 0x11111111u) << (8 * w4);
                    badv |= gather_nibble_bits(nib_nonzero(d4 ^ 0x11111111u)) << (8 * w4);
                }
                const uint32_t km = (1u << ix.k) - 1u;
                // pass 1, registers and shared memory only: which k-mer positions can hit at all.  The k-mers holding a code other than
                // A/C/G/T are found for all positions at once (bk) and handled after the loop; the loop itself is branch-free: two
                // shifts and two ANDs for the planes, a shift-or for the table index, one shared-memory load, one bit test
                const int kk = ix.k, n_pos = max(n_bases - kk + 1, 0);
                uint32_t bk = badv;
                for (int t = 1; t < kk; ++t) bk |= badv >> t;
                const uint32_t bm_addr = (uint32_t)__cvta_generic_to_shared(cand5_bm), wmask = bm_words - 1u;
                uint32_t hits = 0u;
                for (int p0 = 0; p0 < n_pos; ++p0) {
                    const uint32_t idx = ((((hi >> p0) & km) << kk) | ((lo >> p0) & km));
                    const uint32_t word = smem_ld32(bm_addr + (((idx >> 5) & wmask) << 2));  // bm_words is a power of two (folded map)
                    hits |= ((word >> (idx & 31u)) & 1u) << p0;
                }
                const uint32_t pos_mask = n_pos >= 32 ? ~0u : ((1u << n_pos) - 1u);
                hits &= ~bk & pos_mask;
                if (ix.n_iupac_kmers != 0u) {  // a k-mer holding another code can only equal a primer k-mer that spells the same code (PM:113)
                    uint32_t todo = bk & pos_mask;
                    while (todo) {
                        const int p0 = __ffs((int)todo) - 1;
                        todo &= todo - 1u;
                        const uint32_t sig = kmer_plane_signature((lo >> p0) & km, (hi >> p0) & km, (badv >> p0) & km, kk);
                        hits |= ((bloom[sig >> 5] >> (sig & 31u)) & 1u) << p0;
                    }
                }
                // pass 2: the hits in read order (PM:110-113); the table entry of the next hit is requested while this one is walked
                auto entry_of = [&](int p0) -> uint32_t {  // all-A/C/G/T k-mers only; 0 for the others (they take the hashed table)
                    const uint32_t l = (lo >> p0) & km, h = (hi >> p0) & km, bad = (badv >> p0) & km;
                    return bad == 0u ? __ldg(&ix.direct_lohi[(h << ix.k) | l]) : 0u;
                };
                int last_p = -1;
                uint32_t e_next = hits ? entry_of(__ffs((int)hits) - 1) : 0u;
                while (hits) {
                    const int p0 = __ffs((int)hits) - 1;
                    hits &= hits - 1u;
                    const uint32_t e = e_next;
                    if (hits) e_next = entry_of(__ffs((int)hits) - 1);
                    OffCnt oc;
                    int single = -1;  // the entry's one primer, stored in the entry itself (DevKmerIndex::direct_lohi)
                    if (((badv >> p0) & km) == 0u) {
                        if (e & 0x80000000u) single = (int)(e & 0x7FFFFFFFu);
                        else { oc.off = e & 0xFFFFFu; oc.cnt = e >> 20; }
                    } else {
                        uint64_t key = 0;  // rare: spell the k-mer out, first base in the highest nibble
                        for (int j = 0; j < ix.k; ++j) {
                            const int b = p0 + j;
                            const uint32_t word = b < 8 ? sr[0] : b < 16 ? sr[1] : b < 24 ? sr[2] : sr[3];
                            key = (key << 4) | ((word >> (4 * (b & 7))) & 15u);
                        }
                        if (!kmer_lookup(ix, key, oc)) continue;  // the Bloom filter said "maybe"
                    }
                    bool full = false;
                    const uint32_t cnt = single >= 0 ? 1u : oc.cnt;
                    for (uint32_t j = 0; j < cnt; ++j) {
                        const int p = single >= 0 ? single : (int)__ldg(&ix.postings[oc.off + j]);
                        if (p == last_p) continue;  // consecutive k-mers of an on-target read list the same primer again and again
                        bool seen = false;
#pragma unroll
                        for (int t = 0; t < kCandLocal; ++t) seen |= (t < nc && cand[t] == p);
                        last_p = p;
                        if (seen) continue;
                        if (nc == kCandLocal) { full = true; break; }
#pragma unroll
                        for (int t = 0; t < kCandLocal; ++t) if (t == nc) cand[t] = p;
                        ++nc;
                    }
                    if (full) { spilled = true; break; }
                }
            } else if (n_bases <= 32) {
                // a window of at most 32 bases comes in as a 4-word shift register (base i in nibble i: the word loaders of the
                // scan kernel, forward or reverse complement), so a base costs a handful of register operations instead of a byte fetch
                uint32_t sr[4];
                window_from_global<4>(R, r, s, len, rc, sr);
                for (int i = 0; i < n_bases; ++i) {
                    const uint32_t nib = sr[0] & 15u;
                    sr[0] = __funnelshift_r(sr[0], sr[1], 4);
                    sr[1] = __funnelshift_r(sr[1], sr[2], 4);
                    sr[2] = __funnelshift_r(sr[2], sr[3], 4);
                    sr[3] >>= 4;
                    if (!visit(nib, i)) { spilled = true; break; }
                }
            } else {
                for (int i = 0; i < n_bases; ++i)
                    if (!visit(ts.nib(i), i)) { spilled = true; break; }
            }
        }
        const int count = (active && !spilled) ? nc : 0;
        int incl = count;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned rbase = 0;
        if (lane == 31 && total) rbase = atomicAdd(&ctr->n_pairs, (unsigned)total);
        rbase = __shfl_sync(0xffffffffu, rbase, 31);
        if (active) {
            if (spilled) { spill[atomicAdd(&ctr->n_spill, 1u)] = w; }
            else if ((unsigned long long)rbase + total > pair_cap) { if (count) atomicAdd(&ctr->overflow, 1u); seg_off[w] = 0; seg_cnt[w] = count ? -1 : 0; }
            else {
                const unsigned at = rbase + (unsigned)(incl - count);
                seg_off[w] = at; seg_cnt[w] = count;
#pragma unroll
                for (int t = 0; t < kCandLocal; ++t) if (t < count) { pair_work[at + t] = w; pair_primer[at + t] = (unsigned)cand[t]; }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Candidate emission for the 5' gapped stage under -K: one warp per fall-through read.  Walks the read's k-mers in
// order; a per-warp bitmap in shared memory implements `.distinct` (PM:114).  Pass 1 sets bits and counts, pass 2
// clears them again while emitting, so the bitmap is zero for the next read.
// ---------------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) cand5_kernel(DevPanel P, DevReads R, const unsigned int *__restrict__ fall,
                                                    const unsigned int *__restrict__ spill,
                                                    Counters *__restrict__ ctr, unsigned int *__restrict__ pair_work,
                                                    unsigned int *__restrict__ pair_primer, unsigned int pair_cap,
                                                    unsigned int *__restrict__ seg_off, int *__restrict__ seg_cnt) {
    extern __shared__ uint32_t bitmap_all[];
    const int warps_per_block = blockDim.x >> 5, wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int words = (P.n_primers + 31) >> 5;
    uint32_t *bitmap = bitmap_all + (size_t)wib * words;
    for (int i = lane; i < words; i += 32) bitmap[i] = 0;
    __syncwarp();
    const DevKmerIndex &ix = P.kidx[1];
    const unsigned n_work = ctr->n_spill;
    for (unsigned wi = blockIdx.x * warps_per_block + wib; wi < n_work; wi += gridDim.x * warps_per_block) {
        const unsigned w = spill[wi];
        const int r = (int)fall[w];
        const uint16_t fl = R.flags[r];
        const uint8_t *s; int len;
        read_geom(R, r, s, len);
        Oriented o{s, len, seq_order_is_rc(fl)};
        const int end = len >= ix.k ? min(ix.window, len) - ix.k + 1 : 0;  // k-mers start at 0 .. end-1 (PM:110)
        unsigned base = 0, count = 0;
        OffCnt cached{0u, 0u};  // with at most 32 k-mer positions (the usual case) the second pass reuses the first pass's lookups
        for (int pass = 0; pass < 2; ++pass) {
            unsigned emitted = 0;
            for (int p0 = 0; p0 < end; p0 += 32) {
                // every lane looks up the k-mer starting at its own position: one table latency per 32 positions
                const int pos = p0 + lane;
                OffCnt oc{0u, 0u};
                if (pass == 1 && end <= 32) oc = cached;
                else if (pos < end) {
                    uint64_t key = 0;
                    for (int t = 0; t < ix.k; ++t) key = (key << 4) | o.nib(pos + t);
                    if (!kmer_lookup(ix, key, oc)) oc.cnt = 0;
                }
                cached = oc;
                unsigned hits = __ballot_sync(0xffffffffu, oc.cnt > 0);
                while (hits) {  // positions in ascending order = the order the reference visits them in
                    const int src = __ffs(hits) - 1;
                    hits &= hits - 1;
                    const uint32_t off = __shfl_sync(0xffffffffu, oc.off, src), cnt = __shfl_sync(0xffffffffu, oc.cnt, src);
                    for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
                        const uint32_t j = j0 + lane;
                        bool fresh = false;
                        uint32_t p = 0;
                        if (j < cnt) {
                            p = __ldg(&ix.postings[off + j]);
                            const uint32_t bit = 1u << (p & 31);
                            if (pass == 0) fresh = !(atomicOr(&bitmap[p >> 5], bit) & bit);
                            else fresh = (atomicAnd(&bitmap[p >> 5], ~bit) & bit) != 0;
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, fresh);
                        if (pass == 1 && fresh && count != 0xffffffffu) {
                            const unsigned at = base + emitted + __popc(m & ((1u << lane) - 1u));
                            pair_work[at] = w;
                            pair_primer[at] = p;
                        }
                        emitted += __popc(m);
                    }
                }
            }
            if (pass == 0) {
                count = emitted;
                if (lane == 0 && count) base = atomicAdd(&ctr->n_pairs, count);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (count && (unsigned long long)base + count > pair_cap) {  // does not fit: report, emit nothing
                    if (lane == 0) { atomicAdd(&ctr->overflow, 1u); seg_off[w] = 0; seg_cnt[w] = -1; }
                    count = 0xffffffffu;
                } else if (lane == 0) { seg_off[w] = base; seg_cnt[w] = (int)count; }
                if (count == 0) break;  // no candidate set any bit: nothing to clear, nothing to emit
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Candidate selection for 3' matching (IP:490-499): one thread per read.  Reads with a short candidate list emit
// pairs; reads that must try every primer go to the `all` list (handled by sw_all_kernel).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int mate_of(const DevReads &R, int r) {
    const uint16_t fl = R.flags[r];
    if ((fl & IDP_F_MATE_NEXT) && r + 1 < R.n) return r + 1;
    if (r > 0 && (R.flags[r - 1] & IDP_F_MATE_NEXT)) return r - 1;
    return -1;
}

__global__ void __launch_bounds__(256) cand3_kernel(DevPanel P, DevReads R, const ResultOut five,
                                                    Counters *__restrict__ ctr, unsigned int *__restrict__ all_list,
                                                    unsigned int *__restrict__ pair_work, unsigned int *__restrict__ pair_primer,
                                                    unsigned int pair_cap, unsigned int *__restrict__ seg_off, int *__restrict__ seg_cnt) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int count = 0, first = -1, mode = 0;  // mode 0: nothing, 1: single primer, 2: mates of own primer, 3: all primers
    if (r < R.n) {
        const int mate = mate_of(R, r);
        int own_type, own_primer, other_type = IDP_NONE, other_primer = -1;
        read_type_primer(five, r, own_type, own_primer);
        if (mate >= 0) read_type_primer(five, mate, other_type, other_primer);
        if (other_type != IDP_NONE) { mode = 1; count = 1; first = other_primer; }                       // IP:495
        else if (own_type != IDP_NONE) { mode = 2; first = own_primer; count = P.mate_off[own_primer + 1] - P.mate_off[own_primer]; }  // IP:496
        else mode = 3;                                                                                   // IP:497
    }
    // all-primer reads
    const unsigned ball = __ballot_sync(0xffffffffu, mode == 3);
    if (ball) {
        unsigned b = 0;
        if (lane == __ffs(ball) - 1) b = atomicAdd(&ctr->n_all, __popc(ball));
        b = __shfl_sync(0xffffffffu, b, __ffs(ball) - 1);
        if (mode == 3) all_list[b + __popc(ball & ((1u << lane) - 1u))] = (unsigned)r;
    }
    // pair emission, warp-aggregated reservation
    int incl = count;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned base = 0;
    if (lane == 31 && total) base = atomicAdd(&ctr->n_pairs, (unsigned)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (r < R.n) {
        if (mode == 3) { seg_off[r] = 0; seg_cnt[r] = -2; }  // handled through the all list
        else if ((unsigned long long)base + total > pair_cap) { if (count) atomicAdd(&ctr->overflow, 1u); seg_off[r] = 0; seg_cnt[r] = count ? -1 : 0; }
        else {
            const unsigned at = base + (unsigned)(incl - count);
            seg_off[r] = at; seg_cnt[r] = count;
            if (mode == 1) { pair_work[at] = (unsigned)r; pair_primer[at] = (unsigned)first; }
            else if (mode == 2) {
                const int o0 = P.mate_off[first];
                for (int j = 0; j < count; ++j) { pair_work[at + j] = (unsigned)r; pair_primer[at + j] = (unsigned)P.mate_idx[o0 + j]; }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Smith-Waterman, score only.  fgbio Aligner's three matrices (Diagonal / Left / Up) with gap cost open + L*extend,
// free leading and trailing target gaps in Glocal, clamp-at-zero Diagonal in Local.  One thread owns one alignment:
// the query (primer) indexes NQ register-resident rows, the target (read) is streamed column by column.
//   D(i,j) = s(q_i,t_j) + max(D,L,U)(i-1,j-1)      U(i,j) = max(D(i-1,j)+o+e, U(i-1,j)+e)
//   L(i,j) = max(D(i,j-1)+o+e, L(i,j-1)+e)          M = max(D,L,U) is kept per row for the next column's diagonal
// ---------------------------------------------------------------------------------------------------------------
template <int NQ, bool LOCAL, bool SMAT>
__device__ __forceinline__ int sw_score(const DevPanel &P, const int32_t *__restrict__ smat, const uint32_t *__restrict__ qwords, int n,
                                        int n_warp_max, const Oriented &T, int m, bool &missing) {
    const int oe = P.gap_open + P.gap_extend, e = P.gap_extend;
    if (m == 0) return LOCAL ? 0 : P.gap_open + n * P.gap_extend;
    uint32_t q[(NQ + 7) / 8];
#pragma unroll
    for (int w = 0; w < (NQ + 7) / 8; ++w) q[w] = w < P.ww ? __ldg(&qwords[w]) : 0u;
    int D[NQ], L[NQ], M[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        if (LOCAL) { D[i] = 0; L[i] = 0; M[i] = 0; }
        else { D[i] = kNegInf; L[i] = kNegInf; M[i] = P.gap_open + (i + 1) * P.gap_extend; }  // Up(i,0)
    }
    int best = LOCAL ? 0 : kNegInf;
    TargetStream ts(T);
    for (int j = 0; j < m; ++j) {
        const uint32_t t = ts.nib(j);
        int mdiag = 0, d_up = 0, u_up = 0;  // row 0 is all zero in Glocal and Local
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            if (i >= n_warp_max) break;
            if (i < n) {
                const uint32_t qc = (q[i >> 3] >> (4 * (i & 7))) & 15u;
                int s;
                if (SMAT) { s = smat[qc * 16 + t]; if (s == INT32_MIN) { missing = true; s = 0; } }
                else s = qc == t ? P.match_score : P.mismatch_score;
                int dn = mdiag + s;
                if (LOCAL) dn = max(dn, 0);
                const int un = __viaddmax_s32(d_up, oe, u_up + e);
                const int ln = __viaddmax_s32(D[i], oe, L[i] + e);
                mdiag = M[i];
                const int mn = __vimax3_s32(dn, ln, un);
                D[i] = dn; L[i] = ln; M[i] = mn;
                d_up = dn; u_up = un;
                if (LOCAL) best = max(best, dn);
                else if (i == n - 1) best = max(best, mn);
            }
        }
    }
    return best;
}

// generic fallback for primers longer than the register-resident variants cover, and for every alignment when AlignPolicy is
// not the default: rows live in local memory; all four policy switches are honoured (ties do not change a score, the extra
// gap-to-gap transitions do)
template <bool LOCAL, bool SMAT>
__device__ int sw_score_generic(const DevPanel &P, const int32_t *__restrict__ smat, const uint32_t *__restrict__ qwords, int n,
                                const Oriented &T, int m, bool &missing) {
    const int oe = P.gap_open + P.gap_extend, e = P.gap_extend;
    if (m == 0) return LOCAL ? 0 : P.gap_open + n * P.gap_extend;
    int D[kMaxPrimerBases], L[kMaxPrimerBases], M[kMaxPrimerBases];
    int U[AlignPolicy::kLeftFromUp ? kMaxPrimerBases : 1];  // Up of the previous column, only when Left may open from it
    for (int i = 0; i < n; ++i) {
        if (LOCAL) { D[i] = 0; L[i] = 0; M[i] = 0; if (AlignPolicy::kLeftFromUp) U[i] = 0; }
        else { D[i] = kNegInf; L[i] = kNegInf; M[i] = P.gap_open + (i + 1) * P.gap_extend; if (AlignPolicy::kLeftFromUp) U[i] = M[i]; }  // Up(i,0)
    }
    int best = LOCAL ? 0 : kNegInf;
    for (int j = 0; j < m; ++j) {
        const uint32_t t = T.nib(j);
        int mdiag = 0, d_up = 0, u_up = 0, l_up = 0;  // row 0 is all zero in Glocal and Local
        for (int i = 0; i < n; ++i) {
            const uint32_t qc = (__ldg(&qwords[i >> 3]) >> (4 * (i & 7))) & 15u;
            int s;
            if (SMAT) { s = smat[qc * 16 + t]; if (s == INT32_MIN) { missing = true; s = 0; } }
            else s = qc == t ? P.match_score : P.mismatch_score;
            int dn = mdiag + s;
            if (LOCAL) dn = max(dn, 0);
            int un = max(d_up + oe, u_up + e);
            if (AlignPolicy::kUpFromLeft) un = max(un, l_up + oe);
            int ln = max(D[i] + oe, L[i] + e);
            if (AlignPolicy::kLeftFromUp) { ln = max(ln, U[i] + oe); U[i] = un; }
            mdiag = M[i];
            const int mn = max(dn, max(ln, un));
            D[i] = dn; L[i] = ln; M[i] = mn;
            d_up = dn; u_up = un; l_up = ln;
            if (LOCAL) best = max(best, dn);
            else if (i == n - 1) best = max(best, mn);
        }
    }
    return best;
}

// ---------------------------------------------------------------------------------------------------------------
// Fast register-resident variants (default byte-equality scorer).  The query occupies the LAST n of NQ register rows
// (DevPanel::qbot holds every primer bottom-aligned), so the last query row is always row NQ-1 and rows above the
// query simply are not executed: the unrolled rows are entered through a switch at the warp's longest query, shorter
// queries skip their leading rows by predicate.  Per column one SIMD compare produces the match mask of all rows.
// ---------------------------------------------------------------------------------------------------------------
#define IDP_ROWS8(b) IDP_ROW((b) + 0) IDP_ROW((b) + 1) IDP_ROW((b) + 2) IDP_ROW((b) + 3) IDP_ROW((b) + 4) IDP_ROW((b) + 5) IDP_ROW((b) + 6) IDP_ROW((b) + 7)
#define IDP_ROWS32 IDP_ROWS8(0) IDP_ROWS8(8) IDP_ROWS8(16) IDP_ROWS8(24)
#define IDP_ROWS64 IDP_ROWS32 IDP_ROWS8(32) IDP_ROWS8(40) IDP_ROWS8(48) IDP_ROWS8(56)
// Rows [lo, hi) are "ragged": some lanes' queries start below them.  They run branch-free (a lane whose query has not
// started keeps its three carried values by select), because a divergent branch inside the fall-through switch would
// keep the lanes apart until the end of the column.  Rows [hi, NQ) belong to every lane's query and run plain.
// lo = NQ - (longest query in the warp), hi = NQ - (shortest); both are warp-uniform.
#define IDP_SWITCH_ROWS(NQ_, first)                     \
    if constexpr ((NQ_) == 32) { switch (first) { IDP_ROWS32 default: break; } } \
    else { switch (first) { IDP_ROWS64 default: break; } }

// Glocal, score only, in "shifted" coordinates X'(i,j) = X(i,j) - extend*(i+j): a gap extension then costs nothing
// (U'(i,j) = max(D'(i-1,j) + open, U'(i-1,j)), same for Left) and the diagonal move adds s - 2*extend, so a cell is
// LOP3 + SEL + IADD + 2 x VIADDMNMX + VIMNMX3.  The end cell converts back: M(n,j) = M'(n,j) + extend*(n+j).
template <int NQ>
__device__ __forceinline__ int sw_score_glocal_shifted(const DevPanel &P, const uint32_t *__restrict__ qbot, int n, int lo, int hi,
                                                       const Oriented &T, int m) {
    const int o = P.gap_open, e = P.gap_extend;
    int ma = P.match_score - 2 * e, mi = P.mismatch_score - 2 * e;
    asm volatile("" : "+r"(ma), "+r"(mi));  // live in registers: do not reload and re-derive them from the parameters in every row
    const int off = NQ - n;
    uint32_t q[NQ / 8];
#pragma unroll
    for (int w = 0; w < NQ / 8; ++w) q[w] = __ldg(&qbot[w]);
    // LN[r] is Left'(r, next column) = max(Diagonal'(r, j) + open, Left'(r, j)), formed as soon as column j's cell is known:
    // two register arrays instead of three
    int LN[NQ], M[NQ];
#pragma unroll
    for (int r = 0; r < NQ; ++r) { LN[r] = kNegInf; M[r] = o; }  // Up'(i,0) = open + i*extend - extend*i
    int best = kNegInf;
    TargetStream ts(T);
    for (int j = 1; j <= m; ++j) {
        uint32_t t = ts.nib(j - 1);
        asm volatile("" : "+r"(t));
        const uint32_t trep = t * 0x11111111u;
        uint32_t ne[NQ / 8];
#pragma unroll
        for (int w = 0; w < NQ / 8; ++w) ne[w] = nib_nonzero(q[w] ^ trep);
        int mdiag = -e * (j - 1), d_up = -e * j, u_up = -e * j;  // row 0 is all zero: shifted by -extend*(0+j)
#define IDP_CELL(r)                                                                    \
            const int dn = mdiag + ((ne[(r) >> 3] & (1u << (4 * ((r) & 7)))) ? mi : ma); \
            const int un = __viaddmax_s32(d_up, o, u_up);                              \
            const int ln = LN[(r)];                                                    \
            const int mold = M[(r)];                                                   \
            M[(r)] = __vimax3_s32(dn, ln, un);                                         \
            LN[(r)] = __viaddmax_s32(dn, o, ln);
#define IDP_ROW(r)                                                                     \
    case (r): {                                                                        \
        if ((r) >= hi) break;                                                          \
        const bool act = (r) >= off;                                                   \
        IDP_CELL(r)                                                                    \
        mdiag = act ? mold : mdiag; d_up = act ? dn : d_up; u_up = act ? un : u_up;    \
    }
        IDP_SWITCH_ROWS(NQ, lo)
#undef IDP_ROW
#define IDP_ROW(r)                                                                     \
    case (r): {                                                                        \
        IDP_CELL(r)                                                                    \
        mdiag = mold; d_up = dn; u_up = un;                                            \
    }
        IDP_SWITCH_ROWS(NQ, hi)
#undef IDP_ROW
#undef IDP_CELL
        best = max(best, M[NQ - 1] + e * (n + j));  // last query row, this column, back in true coordinates
    }
    return best;
}

template <int NQ, bool LOCAL>
__device__ __forceinline__ int sw_score_fast(const DevPanel &P, const uint32_t *__restrict__ qbot, int n, int lo, int hi, const Oriented &T, int m) {
    static_assert(NQ == 32 || NQ == 64, "row count");
    if (m == 0) return LOCAL ? 0 : P.gap_open + n * P.gap_extend;
    if constexpr (!LOCAL) return sw_score_glocal_shifted<NQ>(P, qbot, n, lo, hi, T, m);
    int oe = P.gap_open + P.gap_extend, e = P.gap_extend, ma = P.match_score, mi = P.mismatch_score;
    asm volatile("" : "+r"(oe), "+r"(e), "+r"(ma), "+r"(mi));  // keep the constants in registers across the unrolled rows
    const int off = NQ - n;
    uint32_t q[NQ / 8];
#pragma unroll
    for (int w = 0; w < NQ / 8; ++w) q[w] = __ldg(&qbot[w]);
    int LN[NQ], M[NQ];  // LN[r] = Left(r, next column) = max(Diagonal(r, j) + open + extend, Left(r, j) + extend)
#pragma unroll
    for (int r = 0; r < NQ; ++r) {
        if (LOCAL) { LN[r] = max(oe, e); M[r] = 0; }  // column 0 is all zero
        else { LN[r] = kNegInf; M[r] = P.gap_open + (r - off + 1) * P.gap_extend; }  // Up(i,0), i = r-off+1
    }
    int best = LOCAL ? 0 : kNegInf;
    TargetStream ts(T);
    for (int j = 0; j < m; ++j) {
        uint32_t t = ts.nib(j);
        asm volatile("" : "+r"(t));  // keep the base in a register: do not re-derive it inside every row
        const uint32_t trep = t * 0x11111111u;
        uint32_t ne[NQ / 8];          // low bit of nibble r%8 of ne[r/8] is set when query row r differs from t
#pragma unroll
        for (int w = 0; w < NQ / 8; ++w) ne[w] = nib_nonzero(q[w] ^ trep);
        int mdiag = 0, d_up = 0, u_up = 0;  // row 0 is all zero in Glocal and Local
#define IDP_CELL(r)                                                                    \
            int dn = mdiag + ((ne[(r) >> 3] & (1u << (4 * ((r) & 7)))) ? mi : ma);      \
            if (LOCAL) dn = max(dn, 0);                                                \
            const int un = __viaddmax_s32(d_up, oe, u_up + e);                         \
            const int ln = LN[(r)];                                                    \
            const int mold = M[(r)];                                                   \
            M[(r)] = __vimax3_s32(dn, ln, un);                                         \
            LN[(r)] = __viaddmax_s32(dn, oe, ln + e);
#define IDP_ROW(r)                                                                     \
    case (r): {                                                                        \
        if ((r) >= hi) break;                                                          \
        const bool act = (r) >= off;                                                   \
        IDP_CELL(r)                                                                    \
        mdiag = act ? mold : mdiag; d_up = act ? dn : d_up; u_up = act ? un : u_up;    \
        if (LOCAL) best = max(best, act ? dn : 0);                                     \
    }
        IDP_SWITCH_ROWS(NQ, lo)
#undef IDP_ROW
#define IDP_ROW(r)                                                                     \
    case (r): {                                                                        \
        IDP_CELL(r)                                                                    \
        mdiag = mold; d_up = dn; u_up = un;                                            \
        if (LOCAL) best = max(best, dn);                                               \
    }
        IDP_SWITCH_ROWS(NQ, hi)
#undef IDP_ROW
#undef IDP_CELL
        if (!LOCAL) best = max(best, M[NQ - 1]);  // last query row, this column (Glocal end cell: best over the last row)
    }
    return best;
}

// Two alignments of the same target per thread, one in each 16-bit half of every register, on the packed DPX forms
// (VIADDMNMX.S16x2, VIMNMX3.S16x2, VIADD.16x2): twice the cells per issued instruction.  Queries a and b (bottom-aligned
// like above) may differ in length; lo / hi bound the ragged rows over both.  Scores only; the caller guarantees the
// value range (DevPanel::fits16_glocal / fits16_local).  Glocal runs in the shifted coordinates of
// sw_score_glocal_shifted, Local in plain coordinates with the zero clamp folded into the .RELU form.
__device__ __forceinline__ uint32_t pack16(int x) { return ((uint32_t)x & 0xFFFFu) * 0x00010001u; }
// PRMT with the selector's sign bits honoured (__byte_perm drops them): a selector nibble 8 | b replicates the TOP BIT of byte b
template <uint32_t SEL>
__device__ __forceinline__ uint32_t prmt_sign(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "n"(SEL));
    return d;
}
// the top bit of every nibble that is non-zero (the other bits are left as they fall)
__device__ __forceinline__ uint32_t nib_nonzero_top(uint32_t x) {
    x |= x << 1;
    x |= x << 2;
    return x;
}
// selector of the PRMT that turns the flags of row r into two half-word masks: byte (r % 8) / 2 of the first word for the low half,
// of the second word for the high half, each as its replicated top bit
#define IDP_SEL16(r) ((8u | (((r) & 7) >> 1)) * 0x0011u | (8u | (4u + (((r) & 7) >> 1))) * 0x1100u)

template <int NQ, bool LOCAL>
__device__ __forceinline__ void sw_score2_fast(const DevPanel &P, const uint32_t *__restrict__ qbot_a, const uint32_t *__restrict__ qbot_b, int na,
                                               int nb, int lo, int hi, const Oriented &T, int m_a, int m_b, int &score_a, int &score_b) {
    static_assert(NQ == 32 || NQ == 64, "row count");
    const int o = P.gap_open, e = P.gap_extend;
    const int m = max(m_a, m_b);  // the shorter target is a prefix of the longer one; its half stops taking end cells at its own m
    if (m == 0) { score_a = LOCAL ? 0 : o + na * e; score_b = LOCAL ? 0 : o + nb * e; return; }
    uint32_t ma2 = pack16(LOCAL ? P.match_score : P.match_score - 2 * e);
    uint32_t selx = ma2 ^ pack16(LOCAL ? P.mismatch_score : P.mismatch_score - 2 * e);  // score = ma2 ^ (mismatch mask & selx)
    uint32_t o2 = pack16(LOCAL ? o + e : o), e2 = pack16(e);
    asm volatile("" : "+r"(ma2), "+r"(selx), "+r"(o2), "+r"(e2));
    constexpr uint32_t NEG2 = 0xC000C000u;  // -16384 in both halves
    const int offa = NQ - na, offb = NQ - nb;
    uint32_t qa[NQ / 8], qb[NQ / 8];
#pragma unroll
    for (int w = 0; w < NQ / 8; ++w) { qa[w] = __ldg(&qbot_a[w]); qb[w] = __ldg(&qbot_b[w]); }
    uint32_t LN[NQ], M[NQ];  // LN[r] = Left(r, next column), formed as soon as this column's cell is known
#pragma unroll
    for (int r = 0; r < NQ; ++r) {
        if (LOCAL) { LN[r] = __vmaxs2(o2, e2); M[r] = 0u; }  // column 0 is all zero
        else { LN[r] = NEG2; M[r] = o2; }                  // Up'(i,0) = open + i*extend - extend*i
    }
    uint32_t best = LOCAL ? 0u : NEG2;
    uint32_t conv = (((uint32_t)(e * nb) & 0xFFFFu) << 16) | ((uint32_t)(e * na) & 0xFFFFu);  // Glocal: extend * (n + j), per half
    uint32_t edge = 0u;                                                                   // Glocal: -extend * j in both halves
    const uint32_t ne2 = pack16(-e);
    TargetStream ts(T);
    // TWO target columns per sweep, as in sw_trace_reg_glocal_shifted: stream A walks column j, stream B walks column j+1 one
    // row behind it, so every step holds two independent max-plus chains and one hides the latency of the other.  An odd last
    // column pairs with a dummy column (code 0: a mismatch in every row) whose end cell is not taken; in Local its Diagonal
    // values stay below the running best (they are an earlier cell's value plus a mismatch), so `best` is unaffected.
    for (int j = 1; j <= m; j += 2) {
        uint32_t tA = ts.nib(j - 1), tB = j < m ? ts.nib(j) : 0u;
        asm volatile("" : "+r"(tA), "+r"(tB));
        const uint32_t trepA = tA * 0x11111111u, trepB = tB * 0x11111111u;
        // mismatch flags in the TOP BIT of a byte, where the sign-replicating PRMT of a row picks them up: rows 8w+2h+1 in byte h
        // of ?o, rows 8w+2h in byte h of ?e (a, b = the two queries; A, B = the two columns)
        uint32_t aoA[NQ / 8], aeA[NQ / 8], boA[NQ / 8], beA[NQ / 8], aoB[NQ / 8], aeB[NQ / 8], boB[NQ / 8], beB[NQ / 8];
#pragma unroll
        for (int w = 0; w < NQ / 8; ++w) {
            aoA[w] = nib_nonzero_top(qa[w] ^ trepA); aeA[w] = aoA[w] << 4; boA[w] = nib_nonzero_top(qb[w] ^ trepA); beA[w] = boA[w] << 4;
            aoB[w] = nib_nonzero_top(qa[w] ^ trepB); aeB[w] = aoB[w] << 4; boB[w] = nib_nonzero_top(qb[w] ^ trepB); beB[w] = boB[w] << 4;
        }
        uint32_t mdiagA, d_upA, u_upA, mdiagB, d_upB, u_upB;
        if (LOCAL) { mdiagA = d_upA = u_upA = 0u; mdiagB = d_upB = u_upB = 0u; }
        else {  // row 0 is all zero: shifted by -extend*(0+j)
            mdiagA = edge; edge = __vadd2(edge, ne2); d_upA = edge; u_upA = edge;
            mdiagB = edge; edge = __vadd2(edge, ne2); d_upB = edge; u_upB = edge;
        }
#define IDP_CELL2(r, S)                                                                                    \
            const uint32_t mask##S = ((r) & 1) ? prmt_sign<IDP_SEL16(r)>(ao##S[(r) >> 3], bo##S[(r) >> 3])   \
                                               : prmt_sign<IDP_SEL16(r)>(ae##S[(r) >> 3], be##S[(r) >> 3]);  \
            const uint32_t s2##S = ma2 ^ (mask##S & selx);                                                 \
            const uint32_t dn##S = LOCAL ? __viaddmax_s16x2_relu(mdiag##S, s2##S, 0u) : __vadd2(mdiag##S, s2##S); \
            const uint32_t un##S = LOCAL ? __viaddmax_s16x2(d_up##S, o2, __vadd2(u_up##S, e2)) : __viaddmax_s16x2(d_up##S, o2, u_up##S); \
            const uint32_t ln##S = LN[(r)];                                                                \
            const uint32_t mold##S = M[(r)];                                                               \
            M[(r)] = __vimax3_s16x2(dn##S, ln##S, un##S);                                                  \
            LN[(r)] = LOCAL ? __viaddmax_s16x2(dn##S, o2, __vadd2(ln##S, e2)) : __viaddmax_s16x2(dn##S, o2, ln##S);
#define IDP_PREV(r) ((r) > 0 ? (r) - 1 : 0)
#define IDP_ACT(r) ((((r) >= offa) ? 0x0000FFFFu : 0u) | (((r) >= offb) ? 0xFFFF0000u : 0u))
#define IDP_ROW(r)                                                                                         \
    case (r): {                                                                                            \
        if ((r) > hi) break;                                                                               \
        {                                                                                                  \
            const uint32_t act = IDP_ACT(r);                                                               \
            IDP_CELL2(r, A)                                                                                \
            mdiagA = (moldA & act) | (mdiagA & ~act); d_upA = (dnA & act) | (d_upA & ~act); u_upA = (unA & act) | (u_upA & ~act); \
            if (LOCAL) best = __vmaxs2(best, dnA & act);                                                   \
        }                                                                                                  \
        if ((r) > 0) {                                                                                     \
            const uint32_t act = IDP_ACT(IDP_PREV(r));                                                     \
            IDP_CELL2(IDP_PREV(r), B)                                                                      \
            mdiagB = (moldB & act) | (mdiagB & ~act); d_upB = (dnB & act) | (d_upB & ~act); u_upB = (unB & act) | (u_upB & ~act); \
            if (LOCAL) best = __vmaxs2(best, dnB & act);                                                   \
        }                                                                                                  \
    }
        IDP_SWITCH_ROWS(NQ, lo)
#undef IDP_ROW
#define IDP_ROW(r)                                                                                         \
    case (r): {                                                                                            \
        { IDP_CELL2(r, A) mdiagA = moldA; d_upA = dnA; u_upA = unA; if (LOCAL) best = __vmaxs2(best, dnA); } \
        { IDP_CELL2(IDP_PREV(r), B) mdiagB = moldB; d_upB = dnB; u_upB = unB; if (LOCAL) best = __vmaxs2(best, dnB); } \
    }
        IDP_SWITCH_ROWS(NQ, hi + 1)
#undef IDP_ROW
        { IDP_CELL2(NQ - 1, B) d_upB = dnB; u_upB = unB; (void)moldB; if (LOCAL) best = __vmaxs2(best, dnB); }
#undef IDP_ACT
#undef IDP_PREV
#undef IDP_CELL2
        if (!LOCAL) {  // last query row, these columns, back in true coordinates: M(n,j) = M'(n,j) + extend*(n+j)
            // (d_up / u_up hold Diagonal / Up of the last row; Left there never exceeds an earlier column's Diagonal, and only the
            // score matters here, so max(Diagonal, Up) is taken)
            conv = __vadd2(conv, e2);
            const uint32_t updA = __viaddmax_s16x2(__vmaxs2(d_upA, u_upA), conv, best);
            const uint32_t inA = (j <= m_a ? 0x0000FFFFu : 0u) | (j <= m_b ? 0xFFFF0000u : 0u);
            best = (updA & inA) | (best & ~inA);
            conv = __vadd2(conv, e2);
            const uint32_t updB = __viaddmax_s16x2(__vmaxs2(d_upB, u_upB), conv, best);
            const uint32_t inB = (j + 1 <= m_a ? 0x0000FFFFu : 0u) | (j + 1 <= m_b ? 0xFFFF0000u : 0u);
            best = (updB & inB) | (best & ~inB);
        }
    }
    score_a = (int)(int16_t)(best & 0xFFFFu);
    score_b = (int)(int16_t)(best >> 16);
    if (!LOCAL && m_a == 0) score_a = o + na * e;
    if (!LOCAL && m_b == 0) score_b = o + nb * e;
}

// Same, carrying where each path started and resolving ties the way an explicit traceback would, in ONE integer per
// matrix value:      value = (score << 14) | (rank << 12) | start
// `start` (12 bits) is the target offset at which the path left the free first row / a zero cell.  `rank` (2 bits) is the
// matrix the value belongs to -- Diagonal 3, Left 2, Up 1 -- so that a plain integer max of two candidates with equal scores
// picks the one the reference prefers: Diagonal over Left over Up for the diagonal move, a gap opened from Diagonal over an
// extended gap.  After each max the rank is re-stamped with one LOP3.  Requires |scores| * (rows + columns) < 2^16
// (DevPanel::packed_trace_ok) and targets shorter than 4096.
//   a.score > b.score  <=>  a > (b | 0x3FFF)
struct TracedReg { int score, t_start, t_end; };

// Glocal in the shifted coordinates of sw_score_glocal_shifted: every candidate of a cell carries the same shift, so the
// comparisons (and the rank / start bits below the score) are untouched, a gap extension costs nothing and the two
// "+ extend" additions per cell disappear.  End cells of different columns are compared after shifting back.
// TWO target columns per sweep: stream A walks column j, stream B walks column j+1 one row behind it
// (B at row r-1 reads what A left in M / LN at row r-1 a step earlier), so every step holds two independent cell updates
// and the max-plus chain of one hides the latency of the other.  Rows lo..hi run in the ragged (select) form, which is
// also what keeps B's first step -- a row above the query -- without effect; the rows after hi run the plain form and B
// finishes with row NQ-1 after the sweep.  An odd last column pairs with a dummy column whose end cell is not taken.
template <int NQ>
__device__ __forceinline__ TracedReg sw_trace_reg_glocal_shifted(const DevPanel &P, const uint32_t *__restrict__ qbot, int n, int lo, int hi,
                                                                  const Oriented &T, int m) {
    constexpr int SH = 14, LOW = 0x3FFF, RANK = 0x3000, RD = 0x3000, RL = 0x2000, RU = 0x1000;
    TracedReg out{0, 1, 0};
    const int e1 = P.gap_extend;
    const int o = P.tr_open;
    const unsigned am = __activemask();
    const int self = (int)(threadIdx.x & 31u);
    const int ma = __shfl_sync(am, P.tr_match_shifted, self), mi = __shfl_sync(am, P.tr_mismatch_shifted, self), keep = __shfl_sync(am, ~RANK, self);
    constexpr int NEG = -(1 << 30);
    const int off = NQ - n;
    uint32_t q[NQ / 8];
#pragma unroll
    for (int w = 0; w < NQ / 8; ++w) q[w] = __ldg(&qbot[w]);
    int LN[NQ], M[NQ];
#pragma unroll
    for (int r = 0; r < NQ; ++r) { LN[r] = NEG | RL; M[r] = o | RU; }
    int best = NEG, bj = 0;
    const int ne1 = -P.tr_ext;
    int edge = 0;
    int back = (e1 * n) * (1 << SH);
    TargetStream ts(T);
    for (int j = 1; j <= m; j += 2) {
        uint32_t tA = ts.nib(j - 1), tB = j < m ? ts.nib(j) : 0u;
        asm volatile("" : "+r"(tA), "+r"(tB));
        const uint32_t trepA = tA * 0x11111111u, trepB = tB * 0x11111111u;
        uint32_t neA[NQ / 8], neB[NQ / 8];
#pragma unroll
        for (int w = 0; w < NQ / 8; ++w) { neA[w] = nib_nonzero(q[w] ^ trepA); neB[w] = nib_nonzero(q[w] ^ trepB); }
        int mdiagA = edge + (j - 1);  // row 0: score 0, path starts at this column
        edge += ne1;
        int d_upA = edge | RD | j, u_upA = edge | RU | j;
        int mdiagB = edge + j;
        edge += ne1;
        int d_upB = edge | RD | (j + 1), u_upB = edge | RU | (j + 1);
#define IDP_CELL2(r, S)                                                                                        \
            const int dn##S = ((mdiag##S + ((ne##S[(r) >> 3] & (1u << (4 * ((r) & 7)))) ? mi : ma)) & ~RANK) | RD; \
            const int un##S = (__viaddmax_s32(d_up##S, o, u_up##S) & keep) | RU;                                 \
            const int ln##S = LN[(r)];                                                                         \
            const int mold##S = M[(r)];                                                                        \
            M[(r)] = __vimax3_s32(dn##S, ln##S, un##S);                                                         \
            LN[(r)] = (__viaddmax_s32(dn##S, o, ln##S) & keep) | RL;
#define IDP_PREV(r) ((r) > 0 ? (r) - 1 : 0)
#define IDP_ROW(r)                                                                                 \
    case (r): {                                                                                    \
        if ((r) > hi) break;                                                                       \
        {                                                                                          \
            const bool act = (r) >= off;                                                           \
            IDP_CELL2(r, A)                                                                        \
            mdiagA = act ? moldA : mdiagA; d_upA = act ? dnA : d_upA; u_upA = act ? unA : u_upA;   \
        }                                                                                          \
        if ((r) > 0) {                                                                             \
            const bool act = (r) - 1 >= off;                                                       \
            IDP_CELL2(IDP_PREV(r), B)                                                              \
            mdiagB = act ? moldB : mdiagB; d_upB = act ? dnB : d_upB; u_upB = act ? unB : u_upB;   \
        }                                                                                          \
    }
        IDP_SWITCH_ROWS(NQ, lo)
#undef IDP_ROW
#define IDP_ROW(r)                                                                                 \
    case (r): {                                                                                    \
        { IDP_CELL2(r, A) mdiagA = moldA; d_upA = dnA; u_upA = unA; }                              \
        { IDP_CELL2(IDP_PREV(r), B) mdiagB = moldB; d_upB = dnB; u_upB = unB; }                    \
    }
        IDP_SWITCH_ROWS(NQ, hi + 1)
#undef IDP_ROW
        { IDP_CELL2(NQ - 1, B) d_upB = dnB; u_upB = unB; (void)moldB; }
#undef IDP_PREV
#undef IDP_CELL2
        // last row, columns ascending, Diagonal / Left / Up, strict '>'
        back += P.tr_ext;
        const int dtA = d_upA + back, utA = u_upA + back;
        if (dtA > (best | LOW)) { best = dtA; bj = j; }
        if (utA > (best | LOW)) { best = utA; bj = j; }
        back += P.tr_ext;
        if (j < m) {
            const int dtB = d_upB + back, utB = u_upB + back;
            if (dtB > (best | LOW)) { best = dtB; bj = j + 1; }
            if (utB > (best | LOW)) { best = utB; bj = j + 1; }
        }
    }
    out.score = best >> SH; out.t_start = (best & 0xFFF) + 1; out.t_end = bj;
    return out;
}

// Local (3' matching, IP:509-513), traceback-equivalent, in plain coordinates.  Differences from the generic form below, all
// to spend fewer instructions per cell (the kernel is bound by the ALU pipe):
//  * the start column is stored INVERTED (4095 - start) in the low 12 bits, so the zero clamp of a Diagonal cell is ONE max
//    against "score 0, path starts after this column": a negative value loses, a positive one wins, and a cell whose score
//    is exactly 0 keeps its own (earlier, hence larger when inverted) start -- fgbio clamps only scores below zero;
//  * the best cell (highest score, then lowest row, then lowest column; only Diagonal values can win) is tracked as one
//    key per column -- score bits | (NQ-1 - row), a LOP3 and a max per cell -- and compared with the running best once per
//    column; the start bits of a new best cell are then read back from M[row]: at that cell the Diagonal value is the cell's
//    maximum (Left / Up there are bounded by earlier Diagonal values minus a gap, and rank breaks a tie in its favour).
template <int NQ>
__device__ __forceinline__ TracedReg sw_trace_reg_local(const DevPanel &P, const uint32_t *__restrict__ qbot, int n, int lo, int hi, const Oriented &T, int m) {
    constexpr int SH = 14, LOW = 0x3FFF, RANK = 0x3000, RD = 0x3000, RL = 0x2000, RU = 0x1000, INV = 0xFFF;
    TracedReg out{0, 1, 0};
    const int oe = P.tr_open_ext, e = P.tr_ext;  // constant-bank operands
    const unsigned am = __activemask();
    const int self = (int)(threadIdx.x & 31u);
    const int ma = __shfl_sync(am, P.tr_match, self), mi = __shfl_sync(am, P.tr_mismatch, self), keep = __shfl_sync(am, ~RANK, self);  // registers
    const int keeplow = __shfl_sync(am, ~LOW, self);
    const int off = NQ - n;
    uint32_t q[NQ / 8];
#pragma unroll
    for (int w = 0; w < NQ / 8; ++w) q[w] = __ldg(&qbot[w]);
    int LN[NQ], M[NQ];  // LN[r] = Left(r, next column), stamped
#pragma unroll
    for (int r = 0; r < NQ; ++r) { LN[r] = (__viaddmax_s32(RD | INV, oe, (RL | INV) + e) & ~RANK) | RL; M[r] = RD | INV; }  // column 0: all zero, start 0
    int best = RD | INV, bestkey = INT_MIN, bj = 0;
    TargetStream ts(T);
    for (int j = 1; j <= m; ++j) {
        uint32_t t = ts.nib(j - 1);
        asm volatile("" : "+r"(t));
        const uint32_t trep = t * 0x11111111u;
        uint32_t ne[NQ / 8];
#pragma unroll
        for (int w = 0; w < NQ / 8; ++w) ne[w] = nib_nonzero(q[w] ^ trep);
        const int fresh = RD | (INV - j);  // score 0, Diagonal, the path starts after this column
        int mdiag = INV + 1 - j, d_up = fresh, u_up = RU | (INV - j);  // row 0: score 0, path starts at this column
        int colmax = INT_MIN;
#define IDP_CELL(r)                                                                                              \
            int dn = ((mdiag + ((ne[(r) >> 3] & (1u << (4 * ((r) & 7)))) ? mi : ma)) & ~RANK) | RD;               \
            dn = max(dn, fresh); /* the clamp at zero, see above */                                               \
            const int un = (__viaddmax_s32(d_up, oe, u_up + e) & keep) | RU; /* equal scores: opened from Diagonal */ \
            const int ln = LN[(r)];                                                                               \
            const int mold = M[(r)];                                                                              \
            M[(r)] = __vimax3_s32(dn, ln, un); /* equal scores: Diagonal, then Left, then Up */                   \
            LN[(r)] = (__viaddmax_s32(dn, oe, ln + e) & keep) | RL;
#define IDP_ROW(r)                                                                     \
    case (r): {                                                                        \
        if ((r) >= hi) break;                                                          \
        const bool act = (r) >= off;                                                   \
        IDP_CELL(r)                                                                    \
        mdiag = act ? mold : mdiag; d_up = act ? dn : d_up; u_up = act ? un : u_up;    \
        colmax = max(colmax, act ? ((dn & keeplow) | (NQ - 1 - (r))) : INT_MIN);       \
    }
        IDP_SWITCH_ROWS(NQ, lo)
#undef IDP_ROW
#define IDP_ROW(r)                                                                     \
    case (r): {                                                                        \
        IDP_CELL(r)                                                                    \
        mdiag = mold; d_up = dn; u_up = un;                                            \
        colmax = max(colmax, (dn & keeplow) | (NQ - 1 - (r)));                         \
    }
        IDP_SWITCH_ROWS(NQ, hi)
#undef IDP_ROW
#undef IDP_CELL
        if (colmax > bestkey) {  // a new best cell in this column (rare after the first few columns): fetch its start bits
            bestkey = colmax; bj = j;
            const int row = NQ - 1 - (colmax & 63);
#pragma unroll
            for (int r = 0; r < NQ; ++r) best = r == row ? M[r] : best;
        }
    }
    out.score = best >> SH; out.t_start = (INV - (best & INV)) + 1; out.t_end = bj;
    return out;
}

// The same, TWO alignments per thread (different queries AND different targets), one in each 16-bit half of every register:
//      value = (score << 10) | (rank << 8) | (255 - start)
// A Local Diagonal value is never negative, so Left / Up never fall below open + extend and six signed score bits are
// enough when  rows * match <= 31,  mismatch >= -32  and  open + 2 * extend >= -32  (DevPanel::local16_ok and the caller's check
// of the rows); start columns need targets of at most 254 bases.  Per step of two cells: one PRMT turns the two mismatch flags
// into the two half-word masks (sign-replicating byte select on the pre-shifted flag words), then the packed DPX forms.
// The zero clamp rides in the third operand of the diagonal VIADDMNMX: "score 0, rank 0, path starts after this column" loses
// against any positive value and, rank 0 being below every stamped rank, against the cell's own value at exactly zero.
template <int NQ>
__device__ __forceinline__ void sw_trace2_local16(const DevPanel &P, const uint32_t *__restrict__ qbot_a, const uint32_t *__restrict__ qbot_b, int na, int nb,
                                                  int lo, int hi, const Oriented &Ta, int m_a, const Oriented &Tb, int m_b, TracedReg &out_a, TracedReg &out_b) {
    static_assert(NQ == 32, "six score bits hold at most 31 rows of match score 1");
    constexpr uint32_t LOW = 0x03FF03FFu, RANK = 0x03000300u, RD = 0x03000300u, RL = 0x02000200u, RU = 0x01000100u, INV2 = 0x00FF00FFu, MIN2 = 0x80008000u;
    constexpr int SH = 10, INV = 0xFF;
    out_a = TracedReg{0, 1, 0}; out_b = TracedReg{0, 1, 0};
    const int m = max(m_a, m_b);
    if (m == 0) return;
    uint32_t ma2 = pack16(P.match_score << SH), selx = ma2 ^ pack16(P.mismatch_score << SH);
    uint32_t oe2 = pack16((P.gap_open + P.gap_extend) << SH), e2 = pack16(P.gap_extend << SH);
    asm volatile("" : "+r"(ma2), "+r"(selx), "+r"(oe2), "+r"(e2));
    const unsigned am = __activemask();
    const int self = (int)(threadIdx.x & 31u);
    const uint32_t keep = __shfl_sync(am, ~RANK, self), keeplow = __shfl_sync(am, ~LOW, self);  // in registers: "& keep | stamp" is then ONE LOP3
    const int offa = NQ - na, offb = NQ - nb;
    uint32_t qa[NQ / 8], qb[NQ / 8];
#pragma unroll
    for (int w = 0; w < NQ / 8; ++w) { qa[w] = __ldg(&qbot_a[w]); qb[w] = __ldg(&qbot_b[w]); }
    uint32_t LN[NQ], M[NQ];  // LN[r] = Left(r, next column), stamped
#pragma unroll
    for (int r = 0; r < NQ; ++r) { LN[r] = (__viaddmax_s16x2(RD | INV2, oe2, __vadd2(RL | INV2, e2)) & ~RANK) | RL; M[r] = RD | INV2; }  // column 0: all zero, start 0
    uint32_t bestkey = MIN2;
    int best_a = (int)((RD | INV2) & 0xFFFFu), best_b = best_a, bj_a = 0, bj_b = 0;
    TargetStream tsa(Ta), tsb(Tb);
    for (int j = 1; j <= m; ++j) {
        uint32_t ta = j <= m_a ? tsa.nib(j - 1) : 0u, tb = j <= m_b ? tsb.nib(j - 1) : 0u;
        asm volatile("" : "+r"(ta), "+r"(tb));
        const uint32_t trepa = ta * 0x11111111u, trepb = tb * 0x11111111u;
        // mismatch flags of rows 8w..8w+7, moved to the top bit of a byte: even rows 8w+2h in byte h of ?e, odd rows 8w+2h+1 in byte h of ?o
        uint32_t ae[NQ / 8], ao[NQ / 8], be[NQ / 8], bo[NQ / 8];
#pragma unroll
        for (int w = 0; w < NQ / 8; ++w) {
            ao[w] = nib_nonzero_top(qa[w] ^ trepa); ae[w] = ao[w] << 4; bo[w] = nib_nonzero_top(qb[w] ^ trepb); be[w] = bo[w] << 4;
        }
        const uint32_t fresh0 = pack16(INV - j);  // score 0, rank 0, the path starts after this column
        uint32_t mdiag = pack16(INV + 1 - j), d_up = RD | fresh0, u_up = RU | fresh0;  // row 0: score 0, path starts at this column
        uint32_t colmax = MIN2;
#define IDP_CELL(r)                                                                                                   \
            const uint32_t mask = ((r) & 1) ? prmt_sign<IDP_SEL16(r)>(ao[(r) >> 3], bo[(r) >> 3])                     \
                                            : prmt_sign<IDP_SEL16(r)>(ae[(r) >> 3], be[(r) >> 3]);                    \
            const uint32_t s2 = ma2 ^ (mask & selx);                                                                  \
            const uint32_t dn = (__viaddmax_s16x2(mdiag, s2, fresh0) & keep) | RD;                                    \
            const uint32_t un = (__viaddmax_s16x2(d_up, oe2, __vadd2(u_up, e2)) & keep) | RU;                         \
            const uint32_t ln = LN[(r)];                                                                              \
            const uint32_t mold = M[(r)];                                                                             \
            M[(r)] = __vimax3_s16x2(dn, ln, un);                                                                      \
            LN[(r)] = (__viaddmax_s16x2(dn, oe2, __vadd2(ln, e2)) & keep) | RL;
#define IDP_ROW(r)                                                                                                    \
    case (r): {                                                                                                       \
        if ((r) >= hi) break;                                                                                         \
        const uint32_t act = (((r) >= offa) ? 0x0000FFFFu : 0u) | (((r) >= offb) ? 0xFFFF0000u : 0u);                 \
        IDP_CELL(r)                                                                                                   \
        mdiag = (mold & act) | (mdiag & ~act); d_up = (dn & act) | (d_up & ~act); u_up = (un & act) | (u_up & ~act);  \
        colmax = __vmaxs2(colmax, ((((dn & keeplow) | (uint32_t)(NQ - 1 - (r)) * 0x00010001u) & act) | (MIN2 & ~act)));  \
    }
        IDP_SWITCH_ROWS(NQ, lo)
#undef IDP_ROW
#define IDP_ROW(r)                                                                                                    \
    case (r): {                                                                                                       \
        IDP_CELL(r)                                                                                                   \
        mdiag = mold; d_up = dn; u_up = un;                                                                           \
        colmax = __vmaxs2(colmax, (dn & keeplow) | (uint32_t)(NQ - 1 - (r)) * 0x00010001u);                           \
    }
        IDP_SWITCH_ROWS(NQ, hi)
#undef IDP_ROW
#undef IDP_CELL
        const uint32_t valid = (j <= m_a ? 0x0000FFFFu : 0u) | (j <= m_b ? 0xFFFF0000u : 0u);
        const uint32_t gt = __vcmpgts2(colmax, bestkey) & valid;
        if (gt) {  // a new best cell in this column (rare after the first few columns): fetch its start bits
            bestkey = (colmax & gt) | (bestkey & ~gt);
            if (gt & 0xFFFFu) {
                bj_a = j;
                const int row = NQ - 1 - (int)(colmax & 31u);
#pragma unroll
                for (int r = 0; r < NQ; ++r) best_a = r == row ? (int)(M[r] & 0xFFFFu) : best_a;
            }
            if (gt >> 16) {
                bj_b = j;
                const int row = NQ - 1 - (int)((colmax >> 16) & 31u);
#pragma unroll
                for (int r = 0; r < NQ; ++r) best_b = r == row ? (int)(M[r] >> 16) : best_b;
            }
        }
    }
    if (m_a > 0) { out_a.score = (int)(int16_t)best_a >> SH; out_a.t_start = (INV - (best_a & INV)) + 1; out_a.t_end = bj_a; }
    if (m_b > 0) { out_b.score = (int)(int16_t)best_b >> SH; out_b.t_start = (INV - (best_b & INV)) + 1; out_b.t_end = bj_b; }
}

// Glocal (5' matching) the same way: two alignments per thread in 16-bit half-words, shifted coordinates as in
// sw_trace_reg_glocal_shifted:      value = (score << 8) | (rank << 6) | start
// In shifted coordinates every value of row >= 1 lies between  min(open, 0) + min(mismatch - 2 extend, 0) + open  and
// rows * match - extend * (rows + columns): eight signed score bits hold that for the default scores up to 32 rows and 63
// columns (DevPanel::glocal16_ok and the caller's per-pair check); six start bits need targets of at most 63 bases.
template <int NQ>
__device__ __forceinline__ void sw_trace2_glocal16(const DevPanel &P, const uint32_t *__restrict__ qbot_a, const uint32_t *__restrict__ qbot_b, int na, int nb,
                                                   int lo, int hi, const Oriented &Ta, int m_a, const Oriented &Tb, int m_b, TracedReg &out_a, TracedReg &out_b) {
    static_assert(NQ == 32, "eight score bits: at most 32 rows");
    constexpr uint32_t LOW = 0x00FF00FFu, RANK = 0x00C000C0u, RD = 0x00C000C0u, RL = 0x00800080u, RU = 0x00400040u, NEG2 = 0x80008000u;
    constexpr int SH = 8;
    out_a = TracedReg{P.gap_open + na * P.gap_extend, 1, 0}; out_b = TracedReg{P.gap_open + nb * P.gap_extend, 1, 0};  // an empty target
    const int m = max(m_a, m_b);
    if (m == 0) return;
    const int e = P.gap_extend;
    uint32_t ma2 = pack16((P.match_score - 2 * e) << SH), selx = ma2 ^ pack16((P.mismatch_score - 2 * e) << SH);
    uint32_t o2 = pack16(P.gap_open << SH), ext2 = pack16(e << SH), ne2 = pack16((-e) << SH);
    asm volatile("" : "+r"(ma2), "+r"(selx), "+r"(o2), "+r"(ext2), "+r"(ne2));
    const unsigned am = __activemask();
    const int self = (int)(threadIdx.x & 31u);
    const uint32_t keep = __shfl_sync(am, ~RANK, self);  // in a register: "& keep | stamp" is then ONE LOP3
    const int offa = NQ - na, offb = NQ - nb;
    uint32_t qa[NQ / 8], qb[NQ / 8];
#pragma unroll
    for (int w = 0; w < NQ / 8; ++w) { qa[w] = __ldg(&qbot_a[w]); qb[w] = __ldg(&qbot_b[w]); }
    uint32_t LN[NQ], M[NQ];
#pragma unroll
    for (int r = 0; r < NQ; ++r) { LN[r] = NEG2 | RL; M[r] = o2 | RU; }  // column 0: Up'(i, 0) = open, start 0
    uint32_t best = NEG2, bj = 0u;
    uint32_t edge = 0u;                                                              // -extend * j in both halves
    uint32_t back = (pack16((e * na) << SH) & 0x0000FFFFu) | (pack16((e * nb) << SH) & 0xFFFF0000u);  // extend * (rows + j), per half
    TargetStream tsa(Ta), tsb(Tb);
    for (int j = 1; j <= m; ++j) {
        uint32_t ta = j <= m_a ? tsa.nib(j - 1) : 0u, tb = j <= m_b ? tsb.nib(j - 1) : 0u;
        asm volatile("" : "+r"(ta), "+r"(tb));
        const uint32_t trepa = ta * 0x11111111u, trepb = tb * 0x11111111u;
        uint32_t ae[NQ / 8], ao[NQ / 8], be[NQ / 8], bo[NQ / 8];  // mismatch flags in the top bit of a byte, see sw_trace2_local16
#pragma unroll
        for (int w = 0; w < NQ / 8; ++w) {
            ao[w] = nib_nonzero_top(qa[w] ^ trepa); ae[w] = ao[w] << 4; bo[w] = nib_nonzero_top(qb[w] ^ trepb); be[w] = bo[w] << 4;
        }
        const uint32_t jj = pack16(j);
        uint32_t mdiag = edge + (jj - 0x00010001u);  // row 0: score 0 (shifted: -extend * (j - 1)), rank 0, the path starts at this column
        edge = __vadd2(edge, ne2);
        uint32_t d_up = edge | RD | jj, u_up = edge | RU | jj;
#define IDP_CELL(r)                                                                                                   \
            const uint32_t mask = ((r) & 1) ? prmt_sign<IDP_SEL16(r)>(ao[(r) >> 3], bo[(r) >> 3])                     \
                                            : prmt_sign<IDP_SEL16(r)>(ae[(r) >> 3], be[(r) >> 3]);                    \
            const uint32_t s2 = ma2 ^ (mask & selx);                                                                  \
            const uint32_t dn = (__vadd2(mdiag, s2) & keep) | RD;                                                     \
            const uint32_t un = (__viaddmax_s16x2(d_up, o2, u_up) & keep) | RU; /* equal scores: opened from Diagonal */  \
            const uint32_t ln = LN[(r)];                                                                              \
            const uint32_t mold = M[(r)];                                                                             \
            M[(r)] = __vimax3_s16x2(dn, ln, un); /* equal scores: Diagonal, then Left, then Up */                     \
            LN[(r)] = (__viaddmax_s16x2(dn, o2, ln) & keep) | RL;
#define IDP_ROW(r)                                                                                                    \
    case (r): {                                                                                                       \
        if ((r) >= hi) break;                                                                                         \
        const uint32_t act = (((r) >= offa) ? 0x0000FFFFu : 0u) | (((r) >= offb) ? 0xFFFF0000u : 0u);                 \
        IDP_CELL(r)                                                                                                   \
        mdiag = (mold & act) | (mdiag & ~act); d_up = (dn & act) | (d_up & ~act); u_up = (un & act) | (u_up & ~act);  \
    }
        IDP_SWITCH_ROWS(NQ, lo)
#undef IDP_ROW
#define IDP_ROW(r)                                                                                                    \
    case (r): {                                                                                                       \
        IDP_CELL(r)                                                                                                   \
        mdiag = mold; d_up = dn; u_up = un;                                                                           \
    }
        IDP_SWITCH_ROWS(NQ, hi)
#undef IDP_ROW
#undef IDP_CELL
        // last row, columns ascending, Diagonal then Up, strict '>' on the score, back in true coordinates
        back = __vadd2(back, ext2);
        const uint32_t valid = (j <= m_a ? 0x0000FFFFu : 0u) | (j <= m_b ? 0xFFFF0000u : 0u);
        const uint32_t dt = __vadd2(d_up, back), ut = __vadd2(u_up, back);
        const uint32_t g1 = __vcmpgts2(dt, best | LOW) & valid;
        best = (dt & g1) | (best & ~g1); bj = (jj & g1) | (bj & ~g1);
        const uint32_t g2 = __vcmpgts2(ut, best | LOW) & valid;
        best = (ut & g2) | (best & ~g2); bj = (jj & g2) | (bj & ~g2);
    }
    if (m_a > 0) { const int b = (int)(int16_t)(best & 0xFFFFu); out_a.score = b >> SH; out_a.t_start = (b & 0x3F) + 1; out_a.t_end = (int)(bj & 0xFFFFu); }
    if (m_b > 0) { const int b = (int)(int16_t)(best >> 16); out_b.score = b >> SH; out_b.t_start = (b & 0x3F) + 1; out_b.t_end = (int)(bj >> 16); }
}

template <int NQ, bool LOCAL>
__device__ __forceinline__ TracedReg sw_trace_reg(const DevPanel &P, const uint32_t *__restrict__ qbot, int n, int lo, int hi, const Oriented &T, int m) {
    static_assert(NQ == 32 || NQ == 64, "row count");
    TracedReg out{0, 1, 0};
    if (m == 0) { out.score = LOCAL ? 0 : P.gap_open + n * P.gap_extend; return out; }
    if constexpr (!LOCAL) return sw_trace_reg_glocal_shifted<NQ>(P, qbot, n, lo, hi, T, m);
    else return sw_trace_reg_local<NQ>(P, qbot, n, lo, hi, T, m);
}

// the target of one (read, primer) task: 5' = read prefix of |primer|+slop bases in sequencing order (PM:292-293);
// 3' = reverse complement of the whole read in sequencing order (IP:502-506)
__device__ __forceinline__ Oriented make_target(const DevPanel &P, const DevReads &R, int r, int primer_seq_len, bool three_prime, int &m) {
    const uint8_t *s; int len;
    read_geom(R, r, s, len);
    const bool so_rc = seq_order_is_rc(R.flags[r]);
    if (three_prime) { m = len; return Oriented{s, len, !so_rc}; }
    m = min(primer_seq_len + P.slop, len);
    return Oriented{s, len, so_rc};
}

constexpr int32_t kScoreMissing = INT32_MIN;  // pair touched a base pair the scoring matrix has no entry for

template <int NQ, bool LOCAL, bool SMAT, bool TRACE, bool PACK2 = false>
__global__ void __launch_bounds__(128, 4) sw_pairs_kernel(DevPanel P, DevReads R, const unsigned int *__restrict__ work_read,
                                                       Counters *ctr, const unsigned int *__restrict__ pair_work,
                                                       const unsigned int *__restrict__ pair_primer, int *__restrict__ pair_score,
                                                       unsigned int *__restrict__ pair_se, unsigned int pair_cap) {
    __shared__ int32_t smat_s[256];
    if (SMAT) { for (int i = threadIdx.x; i < 256; i += blockDim.x) smat_s[i] = P.smat[i]; __syncthreads(); }
    // after an overflow the counter runs past the buffer and slots below pair_cap may hold stale entries: clamp and
    // range-check (the batch is redone with a larger buffer, nothing reads these scores)
    const unsigned n_pairs = min(ctr->n_pairs, pair_cap);
    const unsigned n_work = work_read ? ctr->n_fall : (unsigned)R.n;
    unsigned long long cells = 0;
    // A block takes `tile` x blockDim pairs at a time and deals them to its warps sorted by primer length (a counting sort
    // in shared memory), so that the 32 alignments of a warp have (nearly) the same number of rows: the unrolled rows are
    // entered at the longest query of the warp and the rows between the longest and the shortest run the slower ragged form.
    constexpr int kTileMax = 4, kBins = kMaxPrimerBases + 2;
    __shared__ unsigned short order_s[128 * kTileMax];
    __shared__ int hist_s[kBins];
    const unsigned per_round = gridDim.x * blockDim.x;
    // a span is tile x 128 pairs; half of the grid is resident (4 blocks per SM), so spans stay single (every pair gets its own
    // thread at once) until the batch exceeds one round of the whole grid
    const int tile = n_pairs > 2u * per_round ? 4 : (n_pairs > per_round || PACK2) ? 2 : 1;  // PACK2: a thread takes two pairs at a time
    const unsigned span = (unsigned)tile * blockDim.x;
    for (unsigned base = blockIdx.x * span; base < n_pairs; base += gridDim.x * span) {
        for (int i = threadIdx.x; i < kBins; i += blockDim.x) hist_s[i] = 0;
        __syncthreads();
        int bin[kTileMax], rank[kTileMax];
#pragma unroll
        for (int k = 0; k < kTileMax; ++k) {
            bin[k] = kBins - 1;  // not a pair: sorts to the end
            if (k < tile) {
                const unsigned t = base + (unsigned)k * blockDim.x + threadIdx.x;
                if (t < n_pairs) {
                    const unsigned pp = pair_primer[t];
                    if (pp < (unsigned)P.n_primers) bin[k] = min(__ldg(&P.pinfo[pp]).w, kBins - 2);
                }
                rank[k] = atomicAdd(&hist_s[bin[k]], 1);
            }
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
        for (int k = 0; k < kTileMax; ++k)
            if (k < tile) order_s[hist_s[bin[k]] + rank[k]] = (unsigned short)(k * blockDim.x + threadIdx.x);
        __syncthreads();
        if constexpr (PACK2) {
            // traceback-equivalent alignments, two pairs per thread (sw_trace2_local16 / sw_trace2_glocal16): NEIGHBOURS of the sorted span share a thread, so
            // the two queries have (nearly) the same number of rows
            static_assert(!PACK2 || (NQ == 32 && TRACE && !SMAT), "packed traced form: 32 rows, byte-equality scorer");
            for (int k2 = 0; k2 < tile / 2; ++k2) {
                const unsigned slot = (k2 & 1) ? (blockDim.x - 32u - (threadIdx.x & ~31u)) + (threadIdx.x & 31u) : threadIdx.x;
                const unsigned idx = 2u * ((unsigned)k2 * blockDim.x + slot);
                const unsigned t2[2] = {base + order_s[idx], base + order_s[idx + 1]};
                bool act2[2];
                int n2[2] = {0, 0}, m2[2] = {0, 0}, p2[2] = {0, 0};
                Oriented T2[2] = {Oriented{nullptr, 0, false}, Oriented{nullptr, 0, false}};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    act2[h] = t2[h] < n_pairs;
                    if (act2[h]) {
                        const unsigned w = pair_work[t2[h]];
                        p2[h] = (int)pair_primer[t2[h]];
                        act2[h] = w < n_work && (unsigned)p2[h] < (unsigned)P.n_primers;
                        if (act2[h]) {
                            const int r = work_read ? (int)work_read[w] : (int)w;
                            n2[h] = __ldg(&P.pinfo[p2[h]]).w;
                            T2[h] = make_target(P, R, r, n2[h], LOCAL, m2[h]);
                        }
                    }
                }
                if (!act2[1]) { n2[1] = n2[0]; m2[1] = m2[0]; p2[1] = p2[0]; T2[1] = T2[0]; }  // the sort puts non-pairs last: a lone pair is half 0
                const bool active = act2[0];
                const int n_warp_max = __reduce_max_sync(0xffffffffu, max(n2[0], n2[1]));
                const int n_warp_min = __reduce_min_sync(0xffffffffu, active ? min(n2[0], n2[1]) : n_warp_max);
                if (active) {
                    TracedReg ra, rb;
                    const int nmax = max(n2[0], n2[1]), mmax = max(m2[0], m2[1]);
                    const bool fits = LOCAL ? (nmax * P.match_score <= 31 && mmax <= 254)
                                            : (nmax * P.match_score - P.gap_extend * (nmax + mmax) <= 127 && mmax <= 63);
                    if (fits) {
                        if constexpr (LOCAL) sw_trace2_local16<32>(P, P.qbot + (size_t)p2[0] * 4, P.qbot + (size_t)p2[1] * 4, n2[0], n2[1], 32 - n_warp_max, 32 - n_warp_min,
                                                                   T2[0], m2[0], T2[1], m2[1], ra, rb);
                        else sw_trace2_glocal16<32>(P, P.qbot + (size_t)p2[0] * 4, P.qbot + (size_t)p2[1] * 4, n2[0], n2[1], 32 - n_warp_max, 32 - n_warp_min,
                                                    T2[0], m2[0], T2[1], m2[1], ra, rb);
                    } else {  // a long primer or a long target: one at a time, 32-bit values
                        ra = sw_trace_reg<32, LOCAL>(P, P.qbot + (size_t)p2[0] * 4, n2[0], 32 - n_warp_max, 32 - n_warp_min, T2[0], m2[0]);
                        rb = ra;
                        if (act2[1]) rb = sw_trace_reg<32, LOCAL>(P, P.qbot + (size_t)p2[1] * 4, n2[1], 32 - n_warp_max, 32 - n_warp_min, T2[1], m2[1]);
                    }
                    pair_score[t2[0]] = ra.score;
                    pair_se[t2[0]] = ((unsigned)ra.t_start & 0xFFFFu) | ((unsigned)ra.t_end << 16);
                    cells += (unsigned long long)n2[0] * (unsigned long long)m2[0];
                    if (act2[1]) {
                        pair_score[t2[1]] = rb.score;
                        pair_se[t2[1]] = ((unsigned)rb.t_start & 0xFFFFu) | ((unsigned)rb.t_end << 16);
                        cells += (unsigned long long)n2[1] * (unsigned long long)m2[1];
                    }
                }
            }
        } else
        for (int k = 0; k < tile; ++k) {
            // the sorted span is dealt to the warps in snake order (0 1 2 3 / 3 2 1 0 / ...), so every warp gets the same mix
            // of short and long primers and the warps reach the barrier below together
            const unsigned slot = (k & 1) ? (blockDim.x - 32u - (threadIdx.x & ~31u)) + (threadIdx.x & 31u) : threadIdx.x;
            const unsigned t = base + order_s[k * blockDim.x + slot];
            bool active = t < n_pairs;
            int n = 0, m = 0, p = 0;
            Oriented T{nullptr, 0, false};
            if (active) {
                const unsigned w = pair_work[t];
                p = (int)pair_primer[t];
                active = w < n_work && (unsigned)p < (unsigned)P.n_primers;
                if (active) {
                    const int r = work_read ? (int)work_read[w] : (int)w;
                    n = __ldg(&P.pinfo[p]).w;
                    T = make_target(P, R, r, n, LOCAL, m);
                }
            }
            const int n_warp_max = __reduce_max_sync(0xffffffffu, n);
            const int n_warp_min = __reduce_min_sync(0xffffffffu, active ? n : n_warp_max);
            constexpr int NQF = NQ == 64 ? 64 : 32;
            if (active) {
                bool missing = false;
                int sc;
                if (TRACE) {  // few candidates per read: carry the start column too, so no second alignment is needed
                    const TracedReg tr = sw_trace_reg<NQF, LOCAL>(P, P.qbot + (size_t)p * (NQF / 8), n, NQF - n_warp_max, NQF - n_warp_min, T, m);
                    sc = tr.score;
                    pair_se[t] = ((unsigned)tr.t_start & 0xFFFFu) | ((unsigned)tr.t_end << 16);
                } else if (NQ > 0 && !SMAT) sc = sw_score_fast<NQF, LOCAL>(P, P.qbot + (size_t)p * (NQF / 8), n, NQF - n_warp_max, NQF - n_warp_min, T, m);
                else if (NQ > 0) sc = sw_score<(NQ > 0 ? NQ : 8), LOCAL, SMAT>(P, smat_s, P.pmask + (size_t)p * P.ww, n, n_warp_max, T, m, missing);
                else sc = sw_score_generic<LOCAL, SMAT>(P, smat_s, P.pmask + (size_t)p * P.ww, n, T, m, missing);
                pair_score[t] = missing ? kScoreMissing : sc;
                cells += (unsigned long long)n * (unsigned long long)m;
            }
        }
        __syncthreads();  // order_s / hist_s are rewritten by the next tile
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, d);
    if ((threadIdx.x & 31) == 0 && cells) atomicAdd(&ctr->sw_cells, cells);
}

// running best two by the stable sortBy(-score / query.length) of PM:308-310; ties keep the lower candidate ordinal
struct GBest2 {
    int b_ord = -1, b_primer = 0, b_score = 0, b_qlen = 1;
    int s_ord = -1, s_primer = 0, s_score = 0, s_qlen = 1;
    __device__ __forceinline__ static bool before(int sa, int qa, int oa, int sb, int qb, int ob) {
        const double ka = -__ddiv_rn((double)sa, (double)qa), kb = -__ddiv_rn((double)sb, (double)qb);
        return ka < kb || (ka == kb && oa < ob);
    }
    __device__ __forceinline__ void offer(int ord, int primer, int score, int qlen) {
        if (ord < 0) return;
        if (b_ord < 0 || before(score, qlen, ord, b_score, b_qlen, b_ord)) {
            s_ord = b_ord; s_primer = b_primer; s_score = b_score; s_qlen = b_qlen;
            b_ord = ord; b_primer = primer; b_score = score; b_qlen = qlen;
        } else if (s_ord < 0 || before(score, qlen, ord, s_score, s_qlen, s_ord)) {
            s_ord = ord; s_primer = primer; s_score = score; s_qlen = qlen;
        }
    }
    __device__ __forceinline__ void merge_from_lane(int src_lane_xor) {
        const int o1 = __shfl_xor_sync(0xffffffffu, b_ord, src_lane_xor), p1 = __shfl_xor_sync(0xffffffffu, b_primer, src_lane_xor);
        const int c1 = __shfl_xor_sync(0xffffffffu, b_score, src_lane_xor), q1 = __shfl_xor_sync(0xffffffffu, b_qlen, src_lane_xor);
        const int o2 = __shfl_xor_sync(0xffffffffu, s_ord, src_lane_xor), p2 = __shfl_xor_sync(0xffffffffu, s_primer, src_lane_xor);
        const int c2 = __shfl_xor_sync(0xffffffffu, s_score, src_lane_xor), q2 = __shfl_xor_sync(0xffffffffu, s_qlen, src_lane_xor);
        offer(o1, p1, c1, q1);
        offer(o2, p2, c2, q2);
    }
};

// Every primer is a candidate (no -K, or 3' with no 5' match anywhere): one block per read, threads stride the panel
// in file order, the best two are reduced on chip.  Nothing per-pair ever touches HBM.
template <int NQ, bool LOCAL, bool SMAT, bool PACKED>
__global__ void __launch_bounds__(128, 4) sw_all_kernel(DevPanel P, DevReads R, const unsigned int *__restrict__ work_read,
                                                     const unsigned int *__restrict__ n_work_ptr, Fin *__restrict__ fin,
                                                     Counters *__restrict__ ctr_out, int count_as_5p) {
    __shared__ int32_t smat_s[256];
    __shared__ int red[4][8];  // per-warp GBest2
    if (SMAT) { for (int i = threadIdx.x; i < 256; i += blockDim.x) smat_s[i] = P.smat[i]; __syncthreads(); }
    const unsigned n_work = *n_work_ptr;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    unsigned long long cells = 0;
    for (unsigned w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int r = (int)work_read[w];
        GBest2 g;
        bool any_missing = false;
        constexpr int NQF = NQ == 64 ? 64 : 32;
        constexpr bool packed = PACKED && NQ > 0 && !SMAT;  // the launcher checked DevPanel::fits16_*
        for (int p0 = 0; packed && p0 < P.n_primers; p0 += 2 * blockDim.x) {
            // two primers of (nearly) equal length per thread, one per 16-bit half
            const int ia = p0 + 2 * (int)threadIdx.x, ib = ia + 1;
            const bool active = ia < P.n_primers, both = ib < P.n_primers;
            const int pa = active ? (int)__ldg(&P.by_len[ia]) : 0, pb = both ? (int)__ldg(&P.by_len[ib]) : pa;
            int na = 0, nb = 0, m_a = 0, m_b = 0;
            Oriented T{nullptr, 0, false};
            if (active) {
                na = __ldg(&P.pinfo[pa]).w; nb = __ldg(&P.pinfo[pb]).w;
                T = make_target(P, R, r, na, LOCAL, m_a);
                make_target(P, R, r, nb, LOCAL, m_b);
            }
            const int n_warp_max = __reduce_max_sync(0xffffffffu, max(na, nb));
            const int n_warp_min = __reduce_min_sync(0xffffffffu, active ? min(na, nb) : n_warp_max);
            if (active) {
                int sa, sb;
                sw_score2_fast<NQF, LOCAL>(P, P.qbot + (size_t)pa * (NQF / 8), P.qbot + (size_t)pb * (NQF / 8), na, nb, NQF - n_warp_max,
                                           NQF - n_warp_min, T, m_a, m_b, sa, sb);
                g.offer(pa, pa, sa, na);
                cells += (unsigned long long)na * (unsigned long long)m_a;
                if (both) { g.offer(pb, pb, sb, nb); cells += (unsigned long long)nb * (unsigned long long)m_b; }
            }
        }
        for (int p0 = 0; !packed && p0 < P.n_primers; p0 += blockDim.x) {
            // primers are visited sorted by length so that a warp's alignments have (nearly) the same number of rows;
            // ties are still broken by the primer's position in the file (the ordinal offered below)
            const bool active = p0 + (int)threadIdx.x < P.n_primers;
            const int p = active ? (int)__ldg(&P.by_len[p0 + threadIdx.x]) : 0;
            int n = 0, m = 0;
            Oriented T{nullptr, 0, false};
            if (active) { n = __ldg(&P.pinfo[p]).w; T = make_target(P, R, r, n, LOCAL, m); }
            const int n_warp_max = __reduce_max_sync(0xffffffffu, n);
            const int n_warp_min = __reduce_min_sync(0xffffffffu, active ? n : n_warp_max);
            if (active) {
                bool missing = false;
                int sc;
                if (NQ > 0 && !SMAT) sc = sw_score_fast<NQF, LOCAL>(P, P.qbot + (size_t)p * (NQF / 8), n, NQF - n_warp_max, NQF - n_warp_min, T, m);
                else if (NQ > 0) sc = sw_score<(NQ > 0 ? NQ : 8), LOCAL, SMAT>(P, smat_s, P.pmask + (size_t)p * P.ww, n, n_warp_max, T, m, missing);
                else sc = sw_score_generic<LOCAL, SMAT>(P, smat_s, P.pmask + (size_t)p * P.ww, n, T, m, missing);
                any_missing |= missing;
                g.offer(p, p, sc, n);
                cells += (unsigned long long)n * (unsigned long long)m;
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) g.merge_from_lane(d);
        const bool miss_block = __syncthreads_or(any_missing);
        if (lane == 0) { red[wib][0] = g.b_ord; red[wib][1] = g.b_primer; red[wib][2] = g.b_score; red[wib][3] = g.b_qlen;
                         red[wib][4] = g.s_ord; red[wib][5] = g.s_primer; red[wib][6] = g.s_score; red[wib][7] = g.s_qlen; }
        __syncthreads();
        if (threadIdx.x == 0) {
            GBest2 t;
            for (int k = 0; k < nwarps; ++k) { t.offer(red[k][0], red[k][1], red[k][2], red[k][3]); t.offer(red[k][4], red[k][5], red[k][6], red[k][7]); }
            Fin f;
            f.count = P.n_primers; f.primer = t.b_primer; f.raw = miss_block ? kScoreMissing : t.b_score; f.qlen = (int16_t)t.b_qlen;
            f.raw_next = t.s_ord >= 0 ? t.s_score : 0; f.qlen_next = (int16_t)(t.s_ord >= 0 ? t.s_qlen : 0);
            fin[w] = f;
            if (count_as_5p) atomicAdd(&ctr_out->n_align5, (unsigned long long)P.n_primers);
            else atomicAdd(&ctr_out->n_align3, (unsigned long long)P.n_primers);
        }
        __syncthreads();
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, d);
    if (lane == 0 && cells) atomicAdd(&ctr_out->sw_cells, cells);
}

// ---------------------------------------------------------------------------------------------------------------
// Traceback-equivalent: re-align the winner carrying, next to every score, the target offset at which its path left
// the free first row (Glocal) / a zero cell (Local).  The tie rules are the ones an explicit traceback would follow:
// Diagonal over Left over Up for the diagonal move, Diagonal over gap extension, strict '>' for the end cell scanning
// target positions (and query rows in Local) in ascending order.  Returns score, targetStart (1-based), targetEnd.
// ---------------------------------------------------------------------------------------------------------------
struct Traced { int score, t_start, t_end, q_start, q_end; bool missing; };

// QS: also carry the query row at which the path started (Local only: in Glocal / Global the whole query is aligned).
// Every choice of AlignPolicy is honoured here, in the order the oracle applies them (oracle/idp_oracle.c, align_impl).
template <int TN, bool QS = false>
__device__ Traced sw_trace(const DevPanel &P, const uint32_t *__restrict__ qwords, int n, const Oriented &T, int m, int mode) {
    using Pol = AlignPolicy;
    const int oe = P.gap_open + P.gap_extend, e = P.gap_extend;
    Traced out{0, 1, 0, 1, n, false};
    const bool local = mode == IDP_MODE_LOCAL, global = mode == IDP_MODE_GLOBAL;
    if (m == 0) { out.score = local ? 0 : P.gap_open + n * P.gap_extend; if (local) out.q_start = n + 1; return out; }  // Local: the empty alignment after the last query base
    int D[TN], L[TN], M[TN];
    short sD[TN], sL[TN], sM[TN];
    short qD[QS ? TN : 1], qL[QS ? TN : 1], qM[QS ? TN : 1];  // 1-based row of the boundary / clamped cell the path left
    int U[Pol::kLeftFromUp ? TN : 1];                          // Up of the previous column (value, target start, query start)
    short sU[Pol::kLeftFromUp ? TN : 1], qU[Pol::kLeftFromUp && QS ? TN : 1];
    for (int i = 0; i < n; ++i) {
        sD[i] = 0; sL[i] = 0; sM[i] = 0;
        if (QS) { qD[i] = (short)(i + 1); qL[i] = (short)(i + 1); qM[i] = (short)(i + 1); }
        if (local) { D[i] = 0; L[i] = 0; M[i] = 0; }
        else { D[i] = kNegInf; L[i] = kNegInf; M[i] = P.gap_open + (i + 1) * P.gap_extend; }
        if (Pol::kLeftFromUp) { U[i] = M[i]; sU[i] = 0; if (QS) qU[i] = local ? (short)(i + 1) : (short)0; }  // column 0: Up(i,0) (Local: zero, Done)
    }
    int best = kNegInf, bi = 0, bj = 0, bstart = 0, bq = 0;
    for (int j = 1; j <= m; ++j) {
        const uint32_t t = T.nib(j - 1);
        // row 0: all matrices zero / Done in Glocal and Local; Global has only Left = open + j*extend
        int mdiag, smdiag, d_up, sd_up, u_up, su_up, l_up, sl_up;
        int qmdiag = 0, qd_up = 0, qu_up = 0, ql_up = 0;
        if (!global) { mdiag = 0; smdiag = j - 1; d_up = 0; sd_up = j; u_up = 0; su_up = j; l_up = 0; sl_up = j; }
        else {
            mdiag = j == 1 ? 0 : P.gap_open + (j - 1) * P.gap_extend; smdiag = 0; d_up = kNegInf; sd_up = 0; u_up = kNegInf; su_up = 0;
            l_up = P.gap_open + j * P.gap_extend; sl_up = 0;  // Left(0, j)
        }
        for (int i = 0; i < n; ++i) {
            const uint32_t qc = (__ldg(&qwords[i >> 3]) >> (4 * (i & 7))) & 15u;
            int s;
            if (P.smat) { s = __ldg(&P.smat[qc * 16 + t]); if (s == INT32_MIN) { out.missing = true; s = 0; } }
            else s = qc == t ? P.match_score : P.mismatch_score;
            int dn = mdiag + s, sdn = smdiag, qdn = qmdiag;
            if (local && dn < 0) { dn = 0; sdn = j; qdn = i + 1; }
            int un, sun, qun, ln, sln, qln;
            {   // Up: opened from Diagonal, extended, or (policy) opened from Left -- in that order, a later one needs a strictly higher score
                const int a = d_up + oe, b = u_up + e;
                if (Pol::kExtendOverOpen ? a > b : a >= b) { un = a; sun = sd_up; qun = qd_up; } else { un = b; sun = su_up; qun = qu_up; }
                if (Pol::kUpFromLeft) { const int x = l_up + oe; if (x > un) { un = x; sun = sl_up; qun = ql_up; } }
            }
            {   // Left: opened from Diagonal, extended, or (policy) opened from Up
                const int a = D[i] + oe, b = L[i] + e;
                if (Pol::kExtendOverOpen ? a > b : a >= b) { ln = a; sln = sD[i]; qln = QS ? qD[i] : 0; } else { ln = b; sln = sL[i]; qln = QS ? qL[i] : 0; }
                if (Pol::kLeftFromUp) { const int x = U[i] + oe; if (x > ln) { ln = x; sln = sU[i]; qln = QS ? qU[i] : 0; } }
            }
            int mn = dn, smn = sdn, qmn = qdn;  // the diagonal move of the next column takes the best of the three: Diagonal first on ties
            if (Pol::kTieUpBeforeLeft) {
                if (un > mn) { mn = un; smn = sun; qmn = qun; }
                if (ln > mn) { mn = ln; smn = sln; qmn = qln; }
            } else {
                if (ln > mn) { mn = ln; smn = sln; qmn = qln; }
                if (un > mn) { mn = un; smn = sun; qmn = qun; }
            }
            mdiag = M[i]; smdiag = sM[i];
            if (QS) qmdiag = qM[i];
            D[i] = dn; sD[i] = (short)sdn; L[i] = ln; sL[i] = (short)sln; M[i] = mn; sM[i] = (short)smn;
            if (QS) { qD[i] = (short)qdn; qL[i] = (short)qln; qM[i] = (short)qmn; }
            if (Pol::kLeftFromUp) { U[i] = un; sU[i] = (short)sun; if (QS) qU[i] = (short)qun; }
            d_up = dn; sd_up = sdn; u_up = un; su_up = sun; qd_up = qdn; qu_up = qun; l_up = ln; sl_up = sln; ql_up = qln;
            // end cell: strict '>' scanning rows then columns ascending, matrices in the policy's order
            const int c2 = Pol::kTieUpBeforeLeft ? un : ln, sc2 = Pol::kTieUpBeforeLeft ? sun : sln, qc2 = Pol::kTieUpBeforeLeft ? qun : qln;
            const int c3 = Pol::kTieUpBeforeLeft ? ln : un, sc3 = Pol::kTieUpBeforeLeft ? sln : sun, qc3 = Pol::kTieUpBeforeLeft ? qln : qun;
            if (local) {  // best over all cells: lowest row, then lowest column
                if (dn > best || (dn == best && i + 1 < bi)) { best = dn; bi = i + 1; bj = j; bstart = sdn; bq = qdn; }
                if (c2 > best || (c2 == best && i + 1 < bi)) { best = c2; bi = i + 1; bj = j; bstart = sc2; bq = qc2; }
                if (c3 > best || (c3 == best && i + 1 < bi)) { best = c3; bi = i + 1; bj = j; bstart = sc3; bq = qc3; }
            } else if (i == n - 1 && (!global || j == m)) {
                if (dn > best) { best = dn; bj = j; bstart = sdn; }
                if (c2 > best) { best = c2; bj = j; bstart = sc2; }
                if (c3 > best) { best = c3; bj = j; bstart = sc3; }
            }
        }
    }
    out.score = best; out.t_start = bstart + 1; out.t_end = bj;
    if (local) { out.q_start = bq + 1; out.q_end = bi; }
    return out;
}

// getBestGappedAlignment (PM:305-328) + GappedAlignmentPrimerMatch requires (PMa:113-114) + the score-rate filter (IP:464-469).
// `pre` (score, start, end) is the winner's traceback-equivalent when a kernel already produced it, else it is computed here:
// NQ > 0 selects the register-resident re-alignment, otherwise the generic one with TN rows in local memory.
template <int TN, int NQ>
__device__ void finish_gapped(const DevPanel &P, const DevReads &R, int r, const Fin &f, bool three_prime, const TracedReg *pre,
                              const ResultOut &out, uint8_t *__restrict__ rtype) {
    idp_result res = none_result();
    if (f.count > 0) {
        if (f.raw == kScoreMissing) res.status = IDP_STATUS_SCORE_MISSING;
        else if (!(f.count == 1 && f.raw < 0)) {  // PM:314
            TracedReg tr;
            if (pre) tr = *pre;
            else {
                int m;
                const int n = __ldg(&P.pinfo[f.primer]).w;
                Oriented T = make_target(P, R, r, n, three_prime, m);
                const uint32_t *qw = P.pmask + (size_t)f.primer * P.ww;
                if (NQ > 0 && P.packed_trace_ok) {
                    constexpr int NQF = NQ == 64 ? 64 : 32;
                    const uint32_t *qb = P.qbot + (size_t)f.primer * (NQF / 8);
                    tr = three_prime ? sw_trace_reg<NQF, true>(P, qb, n, 0, NQF, T, m) : sw_trace_reg<NQF, false>(P, qb, n, 0, NQF, T, m);
                } else {
                    const Traced g = sw_trace<TN>(P, qw, n, T, m, three_prime ? IDP_MODE_LOCAL : IDP_MODE_GLOCAL);
                    tr.score = g.score; tr.t_start = g.t_start; tr.t_end = g.t_end;
                }
            }
            res.primer = f.primer; res.type = IDP_GAPPED; res.multi = f.count > 1;
            res.raw_score = tr.score; res.q_len = f.qlen;
            res.raw_next = f.count > 1 ? f.raw_next : 0; res.q_len_next = f.count > 1 ? f.qlen_next : 0;
            res.t_start = (int16_t)tr.t_start; res.t_end = (int16_t)tr.t_end;
            const double score = res.multi ? __ddiv_rn((double)res.raw_score, (double)res.q_len) : (double)res.raw_score;
            const double second = res.multi ? __ddiv_rn((double)res.raw_next, (double)res.q_len_next) : 0.0;
            if (!(score >= second)) res.status = IDP_STATUS_REQ_SCORE_GE_SECOND;
            else if (!(tr.t_start <= tr.t_end)) res.status = IDP_STATUS_REQ_START_LE_END;
            const double rate = __ddiv_rn(score, (double)__ldg(&P.pinfo[f.primer]).x);  // IP:467
            if (!(rate >= P.min_alignment_score_rate)) { const uint8_t st = res.status; res = none_result(); res.status = st; }
        }
    }
    write_result(out, r, res);
    if (rtype != nullptr && res.type != IDP_NONE) rtype[r] = res.type;  // the scan kernel left IDP_NONE there
}

// pair mode: one thread per work item picks the best two of its segment, then finishes
template <int TN, int NQ>
__global__ void __launch_bounds__(128) finalize_pairs_kernel(DevPanel P, DevReads R, const unsigned int *__restrict__ work_read,
                                                             const unsigned int *__restrict__ n_work_ptr, unsigned int n_work_fixed,
                                                             const unsigned int *__restrict__ seg_off, const int *__restrict__ seg_cnt,
                                                             const unsigned int *__restrict__ pair_primer, const int *__restrict__ pair_score,
                                                             const unsigned int *__restrict__ pair_se, const ResultOut out,
                                                             uint8_t *__restrict__ rtype, Counters *__restrict__ ctr_out, int three_prime) {
    const unsigned n_work = n_work_ptr ? *n_work_ptr : n_work_fixed;
    unsigned long long aligned = 0;
    for (unsigned w = blockIdx.x * blockDim.x + threadIdx.x; w < n_work; w += gridDim.x * blockDim.x) {
        const int cnt = seg_cnt[w];
        if (cnt < 0) continue;  // -1: did not fit (reported through Counters::overflow); -2: handled by the all-primers path
        if (cnt == 0 && !three_prime) continue;  // no candidate: the scan already wrote this read's "no match" record (most fall-through reads are off-target)
        const int r = work_read ? (int)work_read[w] : (int)w;
        Fin f;
        f.count = cnt; f.primer = -1; f.raw = 0; f.raw_next = 0; f.qlen = 0; f.qlen_next = 0;
        TracedReg pre{0, 1, 0};
        if (cnt > 0) {
            GBest2 g;
            bool missing = false;
            const unsigned o = seg_off[w];
            for (int j = 0; j < cnt; ++j) {
                const int p = (int)pair_primer[o + j], sc = pair_score[o + j];
                if (sc == kScoreMissing) missing = true;
                g.offer(j, p, sc, __ldg(&P.pinfo[p]).w);
            }
            f.primer = g.b_primer; f.raw = missing ? kScoreMissing : g.b_score; f.qlen = (int16_t)g.b_qlen;
            f.raw_next = g.s_ord >= 0 ? g.s_score : 0; f.qlen_next = (int16_t)(g.s_ord >= 0 ? g.s_qlen : 0);
            aligned += (unsigned)cnt;
            if (pair_se) { const unsigned se = pair_se[o + g.b_ord]; pre.score = g.b_score; pre.t_start = (int)(se & 0xFFFFu); pre.t_end = (int)(se >> 16); }
        }
        finish_gapped<TN, NQ>(P, R, r, f, three_prime != 0, pair_se ? &pre : nullptr, out, rtype);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) aligned += __shfl_xor_sync(0xffffffffu, aligned, d);
    if ((threadIdx.x & 31) == 0 && aligned) atomicAdd(three_prime ? &ctr_out->n_align3 : &ctr_out->n_align5, aligned);
}

// all-primers mode: sw_all_kernel left one Fin per work item
template <int TN, int NQ>
__global__ void __launch_bounds__(128) finalize_all_kernel(DevPanel P, DevReads R, const unsigned int *__restrict__ work_read,
                                                           const unsigned int *__restrict__ n_work_ptr, const Fin *__restrict__ fin,
                                                           const ResultOut out, uint8_t *__restrict__ rtype, int three_prime) {
    const unsigned n_work = *n_work_ptr;
    for (unsigned w = blockIdx.x * blockDim.x + threadIdx.x; w < n_work; w += gridDim.x * blockDim.x)
        finish_gapped<TN, NQ>(P, R, (int)work_read[w], fin[w], three_prime != 0, nullptr, out, rtype);
}

// ---------------------------------------------------------------------------------------------------------------
// TemplateTypeMetric counter (IP:413; TemplateType.scala:52-63): one thread per read, R1 of each template counts.
// Reads the one-byte match type per read that the scan and finalize kernels keep next to the 32-byte results.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) metrics_kernel(DevReads R, const uint8_t *__restrict__ rtype, Counters *__restrict__ ctr) {
    __shared__ unsigned int hist[160];
    for (int i = threadIdx.x; i < 160; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    // a thread takes 8 consecutive reads: one 16-byte load of flags, one 8-byte load of match types, plus the neighbours on
    // either side (the read before decides whether the first one is an R2, the one after is the last one's mate)
    const bool vec = (reinterpret_cast<uintptr_t>(R.flags) & 15) == 0 && (reinterpret_cast<uintptr_t>(rtype) & 7) == 0;
    const long long n8 = ((long long)R.n + 7) / 8;
    // A batch puts nearly all of its templates into a handful of bins.  Each warp picks up to four of them from the first
    // templates it sees and counts those in registers for the whole launch (same-address shared-memory atomics serialise:
    // they were this kernel's whole cost); only the templates of other bins go to the shared histogram one by one.
    int B[4] = {-1, -1, -1, -1};
    unsigned c[4] = {0u, 0u, 0u, 0u};
    bool chosen = false;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < n8; base += (long long)gridDim.x * blockDim.x) {
        const long long g = base + threadIdx.x;
        int bin[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) bin[u] = -1;
        if (g < n8) {
            const int r0 = (int)(g * 8);
            uint16_t f[10];  // f[0] = flags[r0 - 1], f[1..8] = flags[r0 .. r0 + 7], f[9] = flags[r0 + 8]; zero outside the batch
            uint8_t t[9];    // t[0..7] = rtype[r0 .. r0 + 7], t[8] = rtype[r0 + 8]
            bool done = false;
            if (vec && r0 + 8 <= R.n) {
                const uint4 fv = *reinterpret_cast<const uint4 *>(R.flags + r0);
                const uint2 tv = *reinterpret_cast<const uint2 *>(rtype + r0);
                const uint32_t fw[4] = {fv.x, fv.y, fv.z, fv.w};
                const uint32_t tw[2] = {tv.x, tv.y};
                // the usual batch: R1 at the even positions, each followed by its R2 -- four templates, a word of flags each
                static_assert(IDP_F_UNMAPPED == (1 << 2) && IDP_F_FR_PAIR == (1 << 12), "bit positions used below");
                const uint32_t all_and = fw[0] & fw[1] & fw[2] & fw[3], all_or = fw[0] | fw[1] | fw[2] | fw[3];
                const uint16_t before = r0 > 0 ? R.flags[r0 - 1] : (uint16_t)0;
                if ((all_and & IDP_F_MATE_NEXT) && !(all_or & ((uint32_t)IDP_F_MATE_NEXT << 16)) && !(before & IDP_F_MATE_NEXT)) {
                    done = true;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t w = fw[k];
                        // TemplateType of a pair: both mapped 0, one mapped 1, none 2 = the number of unmapped reads
                        const uint32_t tt = ((w >> 2) & 1u) + ((w >> 18) & 1u), fr = (w >> 12) & 1u;
                        const uint32_t ty = tw[k >> 1] >> (16 * (k & 1));  // type of R1 in the low byte, of R2 in the next
                        bin[2 * k] = (int)(tt * 32u + fr * 16u + (ty & 0xFFu) * 4u + ((ty >> 8) & 0xFFu));
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) { f[u + 1] = (uint16_t)(fw[u >> 1] >> (16 * (u & 1))); t[u] = (uint8_t)(tw[u >> 2] >> (8 * (u & 3))); }
            } else {
#pragma unroll
                for (int u = 0; u < 8; ++u) { const bool in = r0 + u < R.n; f[u + 1] = in ? R.flags[r0 + u] : (uint16_t)0; t[u] = in ? rtype[r0 + u] : (uint8_t)IDP_NONE; }
            }
            if (!done) {
            f[0] = r0 > 0 ? R.flags[r0 - 1] : (uint16_t)0;
            const bool after = r0 + 8 < R.n;
            f[9] = after ? R.flags[r0 + 8] : (uint16_t)0;
            t[8] = after ? rtype[r0 + 8] : (uint8_t)IDP_NONE;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + u;
                if (r >= R.n || (f[u] & IDP_F_MATE_NEXT)) continue;  // outside the batch, or an R2
                const uint16_t fl = f[u + 1];
                const bool has_r2 = (fl & IDP_F_MATE_NEXT) && r + 1 < R.n;
                const bool m1 = !(fl & IDP_F_UNMAPPED), m2 = !(f[u + 2] & IDP_F_UNMAPPED);
                const int tt = !has_r2 ? (m1 ? 3 : 4) : ((m1 && m2) ? 0 : (m1 != m2 ? 1 : 2));
                bin[u] = ((tt * 2 + ((fl & IDP_F_FR_PAIR) ? 1 : 0)) * 4 + t[u]) * 4 + (has_r2 ? (int)t[u + 1] : (int)IDP_NONE);
            }
            }
        }
        if (!chosen) {  // warp-uniform: the loop bound is the same for the whole block
            chosen = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int want = -1;  // this lane's first bin that is not picked yet
#pragma unroll
                for (int u = 7; u >= 0; --u) {
                    bool taken = bin[u] < 0;
#pragma unroll
                    for (int j = 0; j < k; ++j) taken |= bin[u] == B[j];
                    if (!taken) want = bin[u];
                }
                const unsigned m = __ballot_sync(0xffffffffu, want >= 0);
                if (m) B[k] = __shfl_sync(0xffffffffu, want, __ffs(m) - 1);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (bin[u] < 0) continue;
            bool counted = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) { const bool h = bin[u] == B[k]; c[k] += h ? 1u : 0u; counted |= h; }
            if (!counted) atomicAdd(&hist[bin[u]], 1u);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned sum = __reduce_add_sync(0xffffffffu, c[k]);
        if ((threadIdx.x & 31) == 0 && sum) atomicAdd(&hist[B[k]], sum);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 160; i += blockDim.x) if (hist[i]) atomicAdd(&ctr->metrics[i], (unsigned long long)hist[i]);
}

// ---------------------------------------------------------------------------------------------------------------
// Downstream classification of templates (scripts/post_process_bam.py:56-78) on the forward / reverse 5' slots of
// addTagsToPair (IP:530-540), as an epilogue on the device-resident 5' records: one thread per read, the thread of a
// template's first read does the work.  counts: one per template whose first record is read 1 (post_process_bam.py:159).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int classify_primer_pair(const DevPanel &P, int f5, int r5, int min_insert, int max_insert) {
    if (f5 < 0 && r5 < 0) return IDP_CLASS_NO_MATCH;
    if (f5 < 0 || r5 < 0) return IDP_CLASS_SINGLE;
    const int4 a = __ldg(&P.cls_info[f5]), b = __ldg(&P.cls_info[r5]);  // {pair, primer_id key, ref_name key, positive}
    if (a.x != b.x) {  // from different primer pairs
        if (a.z != b.z) return IDP_CLASS_CROSS_DIMER;
        const int2 sa = __ldg(&P.cls_span[f5]), sb = __ldg(&P.cls_span[r5]);
        const int insert_length = max(sa.y, sb.y) - min(sa.x, sb.x) + 1;  // :47-53
        return (insert_length < min_insert || max_insert < insert_length) ? IDP_CLASS_CROSS_DIMER : IDP_CLASS_NON_CANONICAL;
    }
    if (a.y == b.y) return IDP_CLASS_SELF_DIMER;
    if (a.w == b.w) return IDP_CLASS_NON_CANONICAL;
    return IDP_CLASS_CANONICAL;
}

__global__ void __launch_bounds__(256) classify_kernel(DevPanel P, const uint16_t *__restrict__ flags, const ResultOut five, int n, int min_insert,
                                                       int max_insert, uint8_t *__restrict__ cls_per_read, Counters *__restrict__ ctr) {
    __shared__ unsigned int hist[6];
    if (threadIdx.x < 6) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        if (r > 0 && (flags[r - 1] & IDP_F_MATE_NEXT)) continue;  // an R2: its R1 speaks for the template
        const uint16_t fl = flags[r];
        const bool has_r2 = (fl & IDP_F_MATE_NEXT) && r + 1 < n;
        int t1, p1, t2 = IDP_NONE, p2 = -1;
        read_type_primer(five, r, t1, p1);
        if (has_r2) read_type_primer(five, r + 1, t2, p2);
        const int m1 = t1 != IDP_NONE ? p1 : -1, m2 = t2 != IDP_NONE ? p2 : -1;
        int f5 = m1, r5 = m2;
        if (has_r2) {  // forward / reverse slots (IP:530-540)
            bool fwd_is_r1 = true;
            if (m1 >= 0) fwd_is_r1 = __ldg(&P.cls_info[m1]).w != 0;
            else if (m2 >= 0) fwd_is_r1 = __ldg(&P.cls_info[m2]).w == 0;
            if (!fwd_is_r1) { f5 = m2; r5 = m1; }
        }
        const int cls = classify_primer_pair(P, f5, r5, min_insert, max_insert);
        if (cls_per_read) { cls_per_read[r] = (uint8_t)cls; if (has_r2) cls_per_read[r + 1] = (uint8_t)cls; }
        if (fl & IDP_F_FIRST) atomicAdd(&hist[cls], 1u);  // "do not double-count": only record.is_read1
    }
    __syncthreads();
    if (threadIdx.x < 6 && hist[threadIdx.x]) atomicAdd(&ctr->classes[threadIdx.x], (unsigned long long)hist[threadIdx.x]);
}

}  // namespace idp

