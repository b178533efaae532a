// Context, device upload of the panel, and the batch pipeline behind idp_match_batch / idp_match_batch_device.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "idp_launch.h"

namespace idp {

int base_to_nibble(unsigned char c);

#define CUDA_TRY(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t err__ = (expr);                                                                         \
        if (err__ != cudaSuccess) {                                                                         \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__);        \
            return IDP_ERR_CUDA;                                                                            \
        }                                                                                                   \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return IDP_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); return IDP_ERR_NOMEM; }
        cap = want;
        return IDP_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Pageable host data -> device.  With a stream the copy is ordered on it (kernels launched there afterwards see the data);
// without one it goes through the legacy default stream, which the contexts' non-blocking streams do NOT order against:
// the caller then synchronises the device before anything is launched (upload_panel does).
template <class T>
static int upload(const std::vector<T> &v, DevBuf &b, size_t min_elems = 1, cudaStream_t st = nullptr, bool on_stream = false) {
    size_t n = std::max(v.size(), min_elems);
    int rc = b.ensure(n * sizeof(T));
    if (rc) return rc;
    if (v.empty()) return IDP_OK;
    if (on_stream) CUDA_TRY(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    else CUDA_TRY(cudaMemcpy(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return IDP_OK;
}

enum { EV_START = 0, EV_SCAN, EV_GAPPED, EV_THREE, EV_END, EV_SW5_A, EV_SW5_B, EV_SW3_A, EV_SW3_B, EV_COUNT };

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[EV_COUNT] = {};
    // staging for the host-pointer path
    DevBuf seq4, seq_off, len, ref_idx, ustart, uend, flags, five, three;
    // scratch of the chain
    DevBuf fall, parked, masks, rtype, cls, spill, all_list, seg_off, seg_cnt, pair_work, pair_primer, pair_score, pair_se, fin;
    Counters *d_ctr = nullptr, *h_ctr = nullptr;
    size_t pair_cap = 0;
    bool busy = false;
    // chunk bookkeeping for the host-pointer path
    int64_t chunk_begin = 0, chunk_n = 0;
    DevReads R{};
    long launches = 0;
};

constexpr int kSlots = 3;

}  // namespace idp

using namespace idp;

struct idp_ctx {
    const idp_panel *panel = nullptr;
    int device = 0;
    LaunchCfg cfg{};
    DevBuf b_fx, b_pm4, b_dfx, b_cinfo, b_cspan, b_qbot, b_mbits, b_bylen, b_pmask, b_pinfo, b_positive, b_loc_key[2], b_loc_primer[2], b_mate_off, b_mate_idx, b_smat;
    DevBuf b_keys[2], b_offcnt[2], b_postings[2], b_heads[2], b_direct[2], b_dbits[2], b_dlohi, b_bloom, b_bsmm, b_bsacc, b_seedtab, b_seedlist, b_diag;
    Slot slots[kSlots];
    cudaStream_t user_stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_finish = nullptr;  // first copy / last harvest of an idp_match_batch call
    idp_timings last{};
    bool compact_ok = false;  // every value of this panel's results fits idp_result16
    // asynchronous device-resident mode (idp_ctx_set_async): batches are enqueued back to back, counters accumulate on the device
    bool async_on = false;
    bool async_lean = false;   // asynchronous mode without the per-stage events and the per-batch counter copy (idp_ctx_set_async(ctx, 2))
    bool stage_events_recorded = false;
    int async_pending = 0;
    // classification epilogue (idp_ctx_set_classify)
    bool cls_on = false;
    int32_t cls_min = 0, cls_max = 0;
    uint8_t *cls_user = nullptr;    // host (idp_match_batch*) or device (*_device) array of the caller
    int64_t *cls_counts = nullptr;  // host, added to
    // idp_align_batch staging (device + pinned host), grown on demand and kept
    DevBuf al_q, al_ql, al_t, al_to, al_tl, al_tp, al_out;  // ASCII queries, query offsets, ASCII targets, query words, target offsets, packed targets, results
};

namespace idp {

static int upload_panel(idp_ctx *c) {
    const idp_panel &p = *c->panel;
    DevPanel &d = c->cfg.dp;
    int rc;
    d.n_primers = p.n; d.ww = p.ww; d.window_bases = p.window_bases; d.max_seq_len = p.max_seq_len;
    if ((rc = upload(p.pmask, c->b_pmask))) return rc;
    d.pmask = c->b_pmask.as<uint32_t>();
    std::vector<int4> info(p.n);
    for (int i = 0; i < p.n; ++i) info[i] = make_int4(p.plen[i], p.cap_full[i], p.acc_full[i], p.seq_len[i]);
    if ((rc = upload(info, c->b_pinfo))) return rc;
    d.pinfo = c->b_pinfo.as<int4>();
    if ((rc = upload(p.positive, c->b_positive))) return rc;
    d.positive = c->b_positive.as<uint8_t>();
    for (int s = 0; s < 2; ++s) {
        if ((rc = upload(p.loc_sortkey[s], c->b_loc_key[s]))) return rc;
        if ((rc = upload(p.loc_primer[s], c->b_loc_primer[s]))) return rc;
        d.loc_sortkey[s] = c->b_loc_key[s].as<int64_t>();
        d.loc_primer[s] = c->b_loc_primer[s].as<int32_t>();
        d.n_loc[s] = (int32_t)p.loc_primer[s].size();
    }
    for (int w = 0; w < 2; ++w) {
        const KmerIndexHost &h = p.kidx[w];
        DevKmerIndex &k = d.kidx[w];
        k.k = h.k > 0 ? h.k : -1;
        k.window = p.max_primer_length;
        k.slot_mask = 0; k.shift = 63; k.keys = nullptr; k.off_cnt = nullptr; k.postings = nullptr;
        k.heads = nullptr; k.direct = nullptr; k.direct_lohi = nullptr; k.direct_bits = nullptr; k.direct_bits_words = 0; k.iupac_bloom = nullptr; k.n_iupac_kmers = 0; k.n_postings = 0;
        if (h.k > 0) {
            std::vector<OffCnt> oc(h.keys.size());
            for (size_t i = 0; i < oc.size(); ++i) oc[i] = OffCnt{h.off[i], h.cnt[i]};
            if ((rc = upload(h.keys, c->b_keys[w]))) return rc;
            if ((rc = upload(oc, c->b_offcnt[w]))) return rc;
            if ((rc = upload(h.postings, c->b_postings[w]))) return rc;
            int bits = 0;
            while ((size_t(1) << bits) < h.keys.size()) ++bits;
            k.slot_mask = (uint32_t)h.keys.size() - 1; k.shift = 64 - bits;
            k.keys = c->b_keys[w].as<uint64_t>(); k.off_cnt = c->b_offcnt[w].as<OffCnt>(); k.postings = c->b_postings[w].as<uint32_t>();
            if ((rc = upload(h.heads, c->b_heads[w]))) return rc;
            k.heads = c->b_heads[w].as<uint2>(); k.n_postings = (uint32_t)h.postings.size();
            if (!h.direct.empty()) {
                if ((rc = upload(h.direct, c->b_direct[w]))) return rc;
                k.direct = c->b_direct[w].as<uint32_t>();
                if (h.k <= 10 && w == 1) {  // cand5_thread_kernel's copies: the table indexed by bit planes, its presence bits, the IUPAC Bloom filter
                    const int kk = h.k;
                    std::vector<uint32_t> lohi(h.direct.size(), 0u), bits((h.direct.size() + 31) / 32, 0u), bloom(2048, 0u);
                    bool inline_ok = true;  // no ordinary entry has bit 31 set (counts below 2048)
                    for (size_t i = 0; i < h.direct.size(); ++i) {
                        if (!h.direct[i]) continue;
                        uint32_t lo = 0, hi = 0;  // i holds the 2-bit codes, first base in the highest bits
                        for (int j = 0; j < kk; ++j) { const uint32_t code = (uint32_t)(i >> (2 * (kk - 1 - j))) & 3u; lo |= (code & 1u) << j; hi |= (code >> 1) << j; }
                        const uint32_t idx = (hi << kk) | lo;
                        lohi[idx] = h.direct[i];
                        bits[idx >> 5] |= 1u << (idx & 31);
                        inline_ok = inline_ok && !(h.direct[i] & 0x80000000u);
                    }
                    if (inline_ok)  // a k-mer with one primer carries it in the entry: one dependent load less per hit
                        for (uint32_t &e : lohi)
                            if ((e >> 20) == 1u) { const uint32_t pr = h.postings[e & 0xFFFFFu]; if (!(pr & 0x80000000u)) e = 0x80000000u | pr; }
                    uint32_t n_iupac = 0;
                    for (uint64_t key : h.keys) {  // 4-bit codes, first base in the highest nibble; 0 = empty slot
                        if (!key) continue;
                        uint32_t lo = 0, hi = 0, bad = 0;
                        for (int j = 0; j < kk; ++j) {
                            const uint32_t nib = (uint32_t)(key >> (4 * (kk - 1 - j))) & 15u;
                            lo |= plane_lo(nib) << j; hi |= plane_hi(nib) << j; bad |= (uint32_t)(__builtin_popcount(nib) != 1) << j;
                        }
                        if (!bad) continue;
                        ++n_iupac;
                        const uint32_t sig = kmer_plane_signature(lo, hi, bad, kk);
                        bloom[sig >> 5] |= 1u << (sig & 31);
                    }
                    // a sparse map is folded (OR of its halves, up to three times) while at most 3 % of its bits stay set: a small panel
                    // then takes 16 KB instead of 128 KB of shared memory and two blocks of the candidate kernel fit an SM.  A bit set
                    // by another k-mer only sends the lookup on to the table, which says "no primer".
                    {
                        size_t set = 0;
                        for (uint32_t v : bits) set += (size_t)__builtin_popcount(v);
                        const size_t full_words = bits.size();
                        while (bits.size() > full_words / 8 && bits.size() >= 2048 && set * 100 <= 3 * (bits.size() * 32 / 2)) {
                            const size_t half = bits.size() / 2;
                            for (size_t i = 0; i < half; ++i) bits[i] |= bits[i + half];
                            bits.resize(half);
                        }
                    }
                    k.direct_bits_words = (uint32_t)bits.size();
                    if ((rc = upload(lohi, c->b_dlohi)) || (rc = upload(bits, c->b_dbits[w])) || (rc = upload(bloom, c->b_bloom))) return rc;
                    k.direct_lohi = c->b_dlohi.as<uint32_t>(); k.direct_bits = c->b_dbits[w].as<uint32_t>();
                    k.iupac_bloom = c->b_bloom.as<uint32_t>(); k.n_iupac_kmers = n_iupac;
                }
            }
        }
    }
    if ((rc = upload(p.bs_mm, c->b_bsmm)) || (rc = upload(p.bs_acc, c->b_bsacc))) return rc;
    d.bs_mm = c->b_bsmm.as<uint32_t>(); d.bs_acc = c->b_bsacc.as<uint2>(); d.bs_groups = (int)(p.bs_acc.size() / 2);
    d.bs_ok = p.acc_max <= 1 && p.window_bases <= 64;
    d.diag = nullptr;
    d.member_bits = nullptr;
    if (p.kidx[0].k > 0 && p.kidx[0].k <= 16) {  // hash bits of every primer's k-mers (member_check_kernel)
        const int k = p.kidx[0].k;
        const uint64_t kmask4 = k >= 16 ? ~0ull : ((1ull << (4 * k)) - 1ull);
        std::vector<uint32_t> mb((size_t)p.n * kMemberWords, 0u);
        for (int i = 0; i < p.n; ++i) {
            uint64_t key = 0;
            for (int b = 0; b < p.seq_len[i]; ++b) {
                key = ((key << 4) | ((p.pmask[(size_t)i * p.ww + b / 8] >> (4 * (b % 8))) & 15u)) & kmask4;
                if (b >= k - 1) { const uint32_t h = member_bits_hash(key); mb[(size_t)i * kMemberWords + (h >> 5)] |= 1u << (h & 31); }
            }
        }
        if ((rc = upload(mb, c->b_mbits))) return rc;
        d.member_bits = c->b_mbits.as<uint32_t>();
    }
    if (d.bs_ok && p.kidx[0].k > 0 && p.ww <= 4) {
        // a position holding one of A/C/G/T that the read matches carries the same letter in the read (PM:143 on one-bit
        // masks), so the literally equal positions are the non-degenerate ones minus the mismatch position
        const int k = p.kidx[0].k;
        std::vector<uint2> diag(p.n);
        for (int i = 0; i < p.n; ++i) {
            const int L = std::min(p.max_primer_length, p.seq_len[i]);  // the k-mer window (PM:110) inside the primer
            uint32_t plain = 0;
            for (int b = 0; b < L && b < 32; ++b) {
                const uint32_t code = (p.pmask[(size_t)i * p.ww + b / 8] >> (4 * (b % 8))) & 15u;
                if (__builtin_popcount(code) == 1) plain |= 1u << b;
            }
            auto has_run = [&](uint32_t e) { uint32_t r = e; for (int t = 1; t < k; ++t) r &= e >> t; return k <= 32 && r != 0u; };
            uint32_t ok = 0;
            for (int b = 0; b < 32; ++b) if (has_run(plain & ~(1u << b))) ok |= 1u << (4 * (b % 8) + b / 8);
            diag[i] = make_uint2(ok, has_run(plain) ? 1u : 0u);
        }
        if ((rc = upload(diag, c->b_diag))) return rc;
        d.diag = c->b_diag.as<uint2>();
    }
    d.seed_ok = 0; d.seed_tab = nullptr; d.seed_list = nullptr; d.seed_S = 0; d.seed_shift = 0; d.seed_slots = 0;
    // panels of up to 8 x 32 primers are scanned faster by the unrolled bit-sliced filter (measured on c2: 0.36 vs 0.43 ms per 8 M reads)
    const bool want_seed = getenv("IDP_SCAN_SEED") ? atoi(getenv("IDP_SCAN_SEED")) != 0 : d.bs_groups > 8;
    if (!p.seed_tab.empty() && want_seed) {
        if ((rc = upload(p.seed_tab, c->b_seedtab)) || (rc = upload(p.seed_list, c->b_seedlist))) return rc;
        d.seed_tab = c->b_seedtab.as<uint32_t>(); d.seed_list = c->b_seedlist.as<uint32_t>();
        d.seed_ok = 1; d.seed_S = p.seed_S; d.seed_shift = p.seed_shift; d.seed_slots = p.seed_slots;
    }
    d.fx_tab = nullptr; d.fx_shift = 0; d.fx_slots = 0; d.fx_mul = p.fx_mul;
    if (!p.fx_tab.empty() && d.bs_ok) {
        if ((rc = upload(p.fx_tab, c->b_fx))) return rc;
        d.fx_tab = c->b_fx.as<uint16_t>(); d.fx_shift = p.fx_shift; d.fx_slots = p.fx_slots;
    }
    c->cfg.pm4 = nullptr; c->cfg.rule_fx = nullptr;
    if (d.fx_tab != nullptr && p.ww <= 4) {
        // scan_fx_kernel's copies, in the nibble order of the stored rows (base i of a word in nibble i ^ 1): the masks padded to
        // four words per primer (one 16-byte load), and the k-mer rule table with its bits where the kernel's mismatch flags are
        auto stored = [](uint32_t x) { return ((x & 0x0F0F0F0Fu) << 4) | ((x >> 4) & 0x0F0F0F0Fu); };
        std::vector<uint4> pm4(p.n);
        std::vector<uint2> dfx(p.n, make_uint2(0u, 0u));  // FxLaunch::rule
        const int k = p.kidx[0].k;
        for (int i = 0; i < p.n; ++i) {
            uint32_t w[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
            for (int t = 0; t < p.ww; ++t) w[t] = stored(p.pmask[(size_t)i * p.ww + t]);
            pm4[i] = make_uint4(w[0], w[1], w[2], w[3]);
            if (k > 0) {  // as DevPanel::diag above: a run of k non-degenerate primer letters that avoids position b
                const int L = std::min(p.max_primer_length, p.seq_len[i]);
                uint32_t plain = 0;
                for (int b = 0; b < L && b < 32; ++b) {
                    const uint32_t code = (p.pmask[(size_t)i * p.ww + b / 8] >> (4 * (b % 8))) & 15u;
                    if (__builtin_popcount(code) == 1) plain |= 1u << b;
                }
                auto has_run = [&](uint32_t e) { uint32_t r = e; for (int t = 1; t < k; ++t) r &= e >> t; return k <= 32 && r != 0u; };
                uint32_t ok = 0;
                for (int b = 0; b < 32; ++b) if (has_run(plain & ~(1u << b))) ok |= 1u << fx_where_bit(b / 8, (b % 8) ^ 1);
                dfx[i] = make_uint2(has_run(plain) ? 1u << 30 : 0u, ok);
            }
            dfx[i].x |= (uint32_t)p.plen[i] | ((uint32_t)std::min(p.cap_full[i], 1023) << 10) | ((uint32_t)std::max(p.acc_full[i], 0) << 20);
        }
        if ((rc = upload(pm4, c->b_pm4)) || (rc = upload(dfx, c->b_dfx))) return rc;
        c->cfg.pm4 = c->b_pm4.as<uint4>();
        c->cfg.rule_fx = c->b_dfx.as<uint2>();
    }
    {   // downstream classification tables
        std::vector<int4> ci(p.n);
        std::vector<int2> cs(p.n);
        for (int i = 0; i < p.n; ++i) { ci[i] = make_int4(p.pair_of[i], p.primer_id_key[i], p.ref_name_key[i], p.positive[i]); cs[i] = make_int2(p.start[i], p.end[i]); }
        if ((rc = upload(ci, c->b_cinfo)) || (rc = upload(cs, c->b_cspan))) return rc;
        d.cls_info = c->b_cinfo.as<int4>(); d.cls_span = c->b_cspan.as<int2>();
    }
    if ((rc = upload(p.mate_off, c->b_mate_off))) return rc;
    if ((rc = upload(p.mate_idx, c->b_mate_idx))) return rc;
    d.mate_off = c->b_mate_off.as<int32_t>(); d.mate_idx = c->b_mate_idx.as<int32_t>();
    d.match_score = p.params.match_score; d.mismatch_score = p.params.mismatch_score;
    d.gap_open = p.params.gap_open; d.gap_extend = p.params.gap_extend;
    d.smat = nullptr;
    if (!p.smat16.empty()) { if ((rc = upload(p.smat16, c->b_smat))) return rc; d.smat = c->b_smat.as<int32_t>(); }
    {   // (score << 14 | rank << 12 | start) register traceback: |score| * (rows + columns) must stay below 2^16
        const long maxabs = std::max<long>({std::labs(d.match_score), std::labs(d.mismatch_score), std::labs((long)d.gap_open + d.gap_extend), std::labs(d.gap_extend)});
        d.packed_trace_ok = d.smat == nullptr && maxabs * (kMaxPrimerBases + kMaxReadLen + 1) < (1L << 16);
        // 16-bit halves (sw_score2_fast): Glocal runs in coordinates shifted by extend * (row + column) over at most
        // max_seq_len + slop columns, Local values are bounded by the query length alone (gaps open from non-negative cells)
        const long m16 = std::max<long>({maxabs, std::labs((long)d.match_score - 2L * d.gap_extend), std::labs((long)d.mismatch_score - 2L * d.gap_extend)});
        const long span = 64 + (long)p.max_seq_len + std::max(p.params.slop, 0) + 2;
        const long ma = d.match_score, mi = d.mismatch_score, go = d.gap_open, ge = d.gap_extend;
        d.tr_open = (int32_t)(go * 16384); d.tr_open_ext = (int32_t)((go + ge) * 16384); d.tr_ext = (int32_t)(ge * 16384);
        d.tr_match = (int32_t)(ma * 16384); d.tr_mismatch = (int32_t)(mi * 16384);
        d.tr_match_shifted = (int32_t)((ma - 2 * ge) * 16384); d.tr_mismatch_shifted = (int32_t)((mi - 2 * ge) * 16384);
        const bool no16 = getenv("IDP_SW_NO_16X2") != nullptr;
        d.fits16_glocal = !no16 && d.smat == nullptr && 2 * m16 * span < 16000;
        d.fits16_local = !no16 && d.smat == nullptr && m16 * (64 + 2) < 16000;
        // sw_trace2_local16: six signed score bits per half-word (the rows * match <= 31 part is checked per pair)
        // sw_trace2_glocal16: eight signed score bits in shifted coordinates (the upper end is checked per pair: rows, columns)
        d.glocal16_ok = d.packed_trace_ok && !getenv("IDP_SW_NO_GLOCAL16") && ma >= 0 && mi <= 0 && go <= 0 && ge <= 0 &&
                        std::min(go, 0L) + std::min(mi - 2 * ge, 0L) + go + ge * (32 + 63) >= -127 && (ma - 2 * ge) <= 127 && -ge * 64 <= 127;
        d.local16_ok = d.packed_trace_ok && !getenv("IDP_SW_NO_LOCAL16") && ma >= 1 && ma <= 31 && mi <= 0 && mi >= -32 && go <= 0 && ge <= 0 && go + 2 * ge >= -32;
    }
    d.slop = p.params.slop; d.three_prime = p.params.three_prime;
    d.max_mismatch_rate = p.params.max_mismatch_rate; d.min_alignment_score_rate = p.params.min_alignment_score_rate;
    const int wb = p.window_bases;
    c->cfg.rw = wb <= 32 ? 4 : wb <= 40 ? 5 : wb <= 64 ? 8 : wb <= 128 ? 16 : 32;
    c->cfg.nq = p.max_seq_len <= 32 ? 32 : p.max_seq_len <= 64 ? 64 : 0;
    if (!AlignPolicy::kIsDefault) {  // the register-resident aligners implement the default policy only: everything takes the generic kernels
        c->cfg.nq = 0; d.packed_trace_ok = 0; d.fits16_glocal = 0; d.fits16_local = 0; d.local16_ok = 0; d.glocal16_ok = 0;
    }
    const int nq = c->cfg.nq;
    d.qbot = nullptr;
    if (nq > 0) {  // primers bottom-aligned in nq rows for the register-resident aligners
        const int words = nq / 8;
        std::vector<uint32_t> qbot((size_t)p.n * words, 0u);
        for (int i = 0; i < p.n; ++i) {
            const int off = nq - p.seq_len[i];
            for (int b = 0; b < p.seq_len[i]; ++b) {
                const uint32_t nib = (p.pmask[(size_t)i * p.ww + b / 8] >> (4 * (b % 8))) & 15u;
                qbot[(size_t)i * words + (off + b) / 8] |= nib << (4 * ((off + b) % 8));
            }
        }
        if ((rc = upload(qbot, c->b_qbot))) return rc;
        d.qbot = c->b_qbot.as<uint32_t>();
    }
    {
        std::vector<uint32_t> order(p.n);
        for (int i = 0; i < p.n; ++i) order[i] = (uint32_t)i;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return p.seq_len[a] < p.seq_len[b]; });
        if ((rc = upload(order, c->b_bylen))) return rc;
        d.by_len = c->b_bylen.as<uint32_t>();
    }
    {   // idp_result16: scores in 16 bits, primer + 1 in 24 bits.  |Alignment.score| <= largest |score| * (rows + columns); the
        // 5' target has at most max_seq_len + slop columns and a Local (3') score lies in [0, match * rows]
        long maxabs = std::max<long>({std::labs(d.match_score), std::labs(d.mismatch_score), std::labs((long)d.gap_open + d.gap_extend), std::labs(d.gap_extend)});
        for (int32_t v : p.smat16) if (v != INT32_MIN) maxabs = std::max<long>(maxabs, std::labs((long)v));
        c->compact_ok = p.n < (1 << 24) - 1 && maxabs * (2L * p.max_seq_len + std::max(p.params.slop, 0) + 2L) < 32767L;
    }
    // the copies above went through the legacy default stream; the contexts' streams are non-blocking and do not order
    // against it, so nothing may be launched before they have landed
    CUDA_TRY(cudaDeviceSynchronize());
    return IDP_OK;
}

// Enqueue the whole matcher chain for reads already on the device.  Nothing here waits on the host.
// five / three: where the per-read records go, in either format (ResultOut).
// keep_totals: only the per-batch scratch counters are cleared, the metric / alignment / class totals keep accumulating
static int enqueue_chain(idp_ctx *c, Slot &s, cudaStream_t st, const DevReads &R, ResultOut five, ResultOut three, uint8_t *cls_dev,
                         bool keep_totals = false) {
    const idp_panel &p = *c->panel;
    const LaunchCfg &cfg = c->cfg;
    const size_t n = (size_t)R.n;
    int rc;
    const bool k_gapped = p.kidx[1].k > 0, do3 = p.params.three_prime >= 0;
    // the stage events each cost the stream a small bubble, the counter copy another: the lean asynchronous mode drops both
    // (idp_ctx_sync then reports ms_total of the last batch only and copies the counters itself)
    const bool stage = !(keep_totals && c->async_lean);
    c->stage_events_recorded = stage;
    if ((rc = s.fall.ensure(n * 4)) || (rc = s.parked.ensure(n * 4)) || (rc = s.masks.ensure(((n + 31) / 32) * 8 + 8))) return rc;
    if (k_gapped && (rc = s.spill.ensure(n * 4))) return rc;
    if (k_gapped || do3) {
        if (s.pair_cap < 2 * n + 4096) s.pair_cap = 2 * n + 4096;
        if ((rc = s.seg_off.ensure(n * 4)) || (rc = s.seg_cnt.ensure(n * 4)) || (rc = s.pair_work.ensure(s.pair_cap * 4)) ||
            (rc = s.pair_primer.ensure(s.pair_cap * 4)) || (rc = s.pair_score.ensure(s.pair_cap * 4)) || (rc = s.pair_se.ensure(s.pair_cap * 4))) return rc;
    }
    if (!k_gapped || do3) { if ((rc = s.fin.ensure(n * sizeof(Fin)))) return rc; }
    if (do3) { if ((rc = s.all_list.ensure(n * 4))) return rc; }
    if ((rc = s.rtype.ensure(n))) return rc;
    Counters *ctr = s.d_ctr;
    unsigned *fall = s.fall.as<unsigned>();
    uint8_t *rtype = s.rtype.as<uint8_t>();
    s.launches = 0;
    CUDA_TRY(cudaMemsetAsync(ctr, 0, keep_totals ? offsetof(Counters, overflow) : sizeof(Counters), st));  // the per-batch list lengths (overflow and what follows are totals)
    CUDA_TRY(cudaEventRecord(s.ev[EV_START], st));
    if (n > 0) {
        if ((rc = launch_scan(cfg, st, R, five, rtype, fall, s.parked.as<unsigned>(), s.masks.as<uint32_t>(), ctr, &s.launches))) return rc;
    }
    if (stage) CUDA_TRY(cudaEventRecord(s.ev[EV_SCAN], st));
    // ---- 5' gapped stage over the fall-through list ----
    if (n > 0) {
        if (k_gapped) {
            launch_cand5(cfg, st, R, fall, ctr, s.pair_work.as<unsigned>(), s.pair_primer.as<unsigned>(), (unsigned)s.pair_cap,
                         s.seg_off.as<unsigned>(), s.seg_cnt.as<int>(), s.spill.as<unsigned>());
            if (stage) CUDA_TRY(cudaEventRecord(s.ev[EV_SW5_A], st));
            launch_sw_pairs(cfg, st, false, R, fall, ctr, s.pair_work.as<unsigned>(), s.pair_primer.as<unsigned>(), s.pair_score.as<int>(), s.pair_se.as<unsigned>(), (unsigned)s.pair_cap);
            if (stage) CUDA_TRY(cudaEventRecord(s.ev[EV_SW5_B], st));
            launch_finalize_pairs(cfg, st, R, fall, &ctr->n_fall, 0u, s.seg_off.as<unsigned>(), s.seg_cnt.as<int>(),
                                  s.pair_primer.as<unsigned>(), s.pair_score.as<int>(), s.pair_se.as<unsigned>(), five, rtype, ctr, 0);
            s.launches += 4;
        } else {
            if (stage) CUDA_TRY(cudaEventRecord(s.ev[EV_SW5_A], st));
            launch_sw_all(cfg, st, false, R, fall, &ctr->n_fall, s.fin.as<Fin>(), ctr, 1);
            if (stage) CUDA_TRY(cudaEventRecord(s.ev[EV_SW5_B], st));
            launch_finalize_all(cfg, st, R, fall, &ctr->n_fall, s.fin.as<Fin>(), five, rtype, 0);
            s.launches += 2;
        }
    } else { if (stage) CUDA_TRY(cudaEventRecord(s.ev[EV_SW5_A], st)); if (stage) CUDA_TRY(cudaEventRecord(s.ev[EV_SW5_B], st)); }
    if (c->cls_on && n > 0) {  // classification epilogue on the finished 5' records
        launch_classify(cfg, st, R.flags, five, R.n, c->cls_min, c->cls_max, cls_dev, ctr);
        ++s.launches;
    }
    if (stage) CUDA_TRY(cudaEventRecord(s.ev[EV_GAPPED], st));
    // ---- 3' stage (IP:483-516) ----
    if (do3 && n > 0 && three.p) {
        CUDA_TRY(cudaMemsetAsync(&ctr->n_pairs, 0, sizeof(unsigned), st));
        launch_cand3(cfg, st, R, five, ctr, s.all_list.as<unsigned>(), s.pair_work.as<unsigned>(), s.pair_primer.as<unsigned>(),
                     (unsigned)s.pair_cap, s.seg_off.as<unsigned>(), s.seg_cnt.as<int>());
        if (stage) CUDA_TRY(cudaEventRecord(s.ev[EV_SW3_A], st));
        launch_sw_pairs(cfg, st, true, R, nullptr, ctr, s.pair_work.as<unsigned>(), s.pair_primer.as<unsigned>(), s.pair_score.as<int>(), s.pair_se.as<unsigned>(), (unsigned)s.pair_cap);
        launch_sw_all(cfg, st, true, R, s.all_list.as<unsigned>(), &ctr->n_all, s.fin.as<Fin>(), ctr, 0);
        if (stage) CUDA_TRY(cudaEventRecord(s.ev[EV_SW3_B], st));
        launch_finalize_pairs(cfg, st, R, nullptr, nullptr, (unsigned)R.n, s.seg_off.as<unsigned>(), s.seg_cnt.as<int>(),
                              s.pair_primer.as<unsigned>(), s.pair_score.as<int>(), s.pair_se.as<unsigned>(), three, nullptr, ctr, 1);
        launch_finalize_all(cfg, st, R, s.all_list.as<unsigned>(), &ctr->n_all, s.fin.as<Fin>(), three, nullptr, 1);
        s.launches += 5;
    } else { if (stage) CUDA_TRY(cudaEventRecord(s.ev[EV_SW3_A], st)); if (stage) CUDA_TRY(cudaEventRecord(s.ev[EV_SW3_B], st)); }
    if (stage) CUDA_TRY(cudaEventRecord(s.ev[EV_THREE], st));
    if (n > 0) { launch_metrics(cfg, st, R, rtype, ctr); ++s.launches; }
    if (stage) CUDA_TRY(cudaMemcpyAsync(s.h_ctr, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(s.ev[EV_END], st));
    CUDA_TRY(cudaGetLastError());
    return IDP_OK;
}

static void add_timings(idp_ctx *c, Slot &s, bool first) {
    idp_timings &t = c->last;
    float ms = 0;
    auto el = [&](int a, int b) { ms = 0; cudaEventElapsedTime(&ms, s.ev[a], s.ev[b]); return ms; };
    if (first) t.ms_total = 0;
    if (c->stage_events_recorded) {
        t.ms_scan += el(EV_START, EV_SCAN);
        t.ms_gapped += el(EV_SCAN, EV_GAPPED);
        t.ms_three_prime += el(EV_GAPPED, EV_THREE);
        t.ms_sw_score += el(EV_SW5_A, EV_SW5_B) + el(EV_SW3_A, EV_SW3_B);
    }
    t.n_fallthrough += s.h_ctr->n_fall;
    t.n_alignments_5p += (int64_t)s.h_ctr->n_align5;
    t.n_alignments_3p += (int64_t)s.h_ctr->n_align3;
    t.sw_cells += (int64_t)s.h_ctr->sw_cells;
    t.scan_base_compares += (int64_t)s.h_ctr->base_compares;
    t.kernel_launches += s.launches;
    t.n_scan_parked += (int64_t)s.h_ctr->n_parked;
    t.n_scan_deferred += (int64_t)s.h_ctr->n_defer;
}

// a failed call must not leave copies in flight into the caller's arrays, nor slots that a later call would harvest
static void drain_slots(idp_ctx *c) {
    for (Slot &s : c->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        s.busy = false;
    }
    cudaGetLastError();
}

static int check_reads(const idp_ctx *c, const idp_reads *reads, int32_t n_reads, const void *five, const void *three, bool compact) {
    if (!c) { set_error("NULL context"); return IDP_ERR_ARG; }
    if (!reads || n_reads < 0 || !five) { set_error("NULL reads / results or negative count"); return IDP_ERR_ARG; }
    if (n_reads > 0 && (!reads->seq4 || !reads->flags)) { set_error("reads.seq4 and reads.flags are required"); return IDP_ERR_ARG; }
    if (reads->fixed_len <= 0 && n_reads > 0 && (!reads->seq_off || !reads->len)) { set_error("reads.seq_off and reads.len are required unless fixed_len > 0"); return IDP_ERR_ARG; }
    if (reads->fixed_len > kMaxReadLen) { set_error("reads longer than %d bases are not supported", kMaxReadLen); return IDP_ERR_UNSUPPORTED; }
    if (c->panel->params.three_prime >= 0 && !three) { set_error("three_prime >= 0 needs a `three` result array"); return IDP_ERR_ARG; }
    if (compact && !c->compact_ok) {
        set_error("this panel's alignment scores or size do not fit the 16-byte record; use the 32-byte idp_match_batch calls");
        return IDP_ERR_UNSUPPORTED;
    }
    return IDP_OK;
}

static int match_batch_device(idp_ctx *c, const idp_reads *reads, int32_t n_reads, void *five, void *three, bool compact,
                              int64_t *metrics160, int64_t *n_gapped_alignments) {
    int rc = check_reads(c, reads, n_reads, five, three, compact);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(c->device));
    Slot &s = c->slots[0];
    cudaStream_t st = c->user_stream ? c->user_stream : s.stream;
    DevReads R{reads->seq4, reads->seq_off, reads->len, reads->ref_idx, reads->unclipped_start, reads->unclipped_end, reads->flags, reads->fixed_len, n_reads};
    c->last = idp_timings{};
    if (c->async_on) {  // enqueue and return: idp_ctx_sync() collects the counters of everything enqueued since the last one
        if (c->async_pending == 0) CUDA_TRY(cudaMemsetAsync(s.d_ctr, 0, sizeof(Counters), st));
        if ((rc = enqueue_chain(c, s, st, R, ResultOut{five, compact}, ResultOut{three, compact}, c->cls_user, true))) { cudaStreamSynchronize(st); cudaGetLastError(); c->async_pending = 0; return rc; }
        ++c->async_pending;
        c->last.n_reads = n_reads;
        return IDP_OK;
    }
    for (int attempt = 0;; ++attempt) {
        if ((rc = enqueue_chain(c, s, st, R, ResultOut{five, compact}, ResultOut{three, compact}, c->cls_user))) { cudaStreamSynchronize(st); cudaGetLastError(); return rc; }
        CUDA_TRY(cudaStreamSynchronize(st));
        if (s.h_ctr->overflow == 0) break;
        if (attempt >= 6) { set_error("candidate pair buffer overflow persists after %d retries", attempt); return IDP_ERR_NOMEM; }
        s.pair_cap *= 4;  // rare: a read listed far more -K candidates than budgeted; redo the batch with more room
    }
    c->last = idp_timings{};
    add_timings(c, s, true);
    float ms = 0; cudaEventElapsedTime(&ms, s.ev[EV_START], s.ev[EV_END]);
    c->last.ms_total = ms; c->last.n_reads = n_reads;
    if (metrics160) for (int i = 0; i < 160; ++i) metrics160[i] += (int64_t)s.h_ctr->metrics[i];
    if (n_gapped_alignments) *n_gapped_alignments += (int64_t)s.h_ctr->n_align5;
    if (c->cls_on && c->cls_counts) for (int i = 0; i < 6; ++i) c->cls_counts[i] += (int64_t)s.h_ctr->classes[i];
    return IDP_OK;
}

// host-pointer path: chunks of reads flow through kSlots streams so that H2D, kernels and D2H overlap
static int match_batch_host(idp_ctx *c, const idp_reads *reads, int32_t n_reads, void *five, void *three, bool compact,
                            int64_t *metrics160, int64_t *n_gapped_alignments) {
    int rc = check_reads(c, reads, n_reads, five, three, compact);
    if (rc) return rc;
    const bool do3 = c->panel->params.three_prime >= 0;
    const size_t rec = compact ? sizeof(idp_result16) : sizeof(idp_result);
    CUDA_TRY(cudaSetDevice(c->device));
    c->last = idp_timings{};
    c->last.n_reads = n_reads;
    const char *env = getenv("IDP_CHUNK_READS");
    int64_t chunk = env ? atoll(env) : (1 << 20);
    if (chunk < 2) chunk = 2;
    const bool have_loc = reads->ref_idx && reads->unclipped_start && reads->unclipped_end;
    bool first = true;
    int64_t total_metrics[160] = {0}, total_align = 0, total_classes[6] = {0};
    const bool cls_bytes = c->cls_on && c->cls_user != nullptr;
    uint8_t *five_b = static_cast<uint8_t *>(five), *three_b = static_cast<uint8_t *>(three);

    auto harvest = [&](Slot &s) -> int {
        if (!s.busy) return IDP_OK;
        CUDA_TRY(cudaStreamSynchronize(s.stream));
        s.busy = false;
        for (int attempt = 0; s.h_ctr->overflow; ++attempt) {  // rare: redo this chunk (still resident in the slot) with more room
            if (attempt >= 6) { set_error("candidate pair buffer overflow persists after %d retries", attempt); return IDP_ERR_NOMEM; }
            s.pair_cap *= 4;
            int rc2 = enqueue_chain(c, s, s.stream, s.R, ResultOut{s.five.p, compact}, ResultOut{do3 ? s.three.p : nullptr, compact}, cls_bytes ? s.cls.as<uint8_t>() : nullptr);
            if (rc2) return rc2;
            if (cls_bytes) CUDA_TRY(cudaMemcpyAsync(c->cls_user + s.chunk_begin, s.cls.p, s.chunk_n, cudaMemcpyDeviceToHost, s.stream));
            CUDA_TRY(cudaMemcpyAsync(five_b + s.chunk_begin * rec, s.five.p, s.chunk_n * rec, cudaMemcpyDeviceToHost, s.stream));
            if (do3) CUDA_TRY(cudaMemcpyAsync(three_b + s.chunk_begin * rec, s.three.p, s.chunk_n * rec, cudaMemcpyDeviceToHost, s.stream));
            CUDA_TRY(cudaStreamSynchronize(s.stream));
        }
        add_timings(c, s, first);
        first = false;
        for (int i = 0; i < 160; ++i) total_metrics[i] += (int64_t)s.h_ctr->metrics[i];
        total_align += (int64_t)s.h_ctr->n_align5;
        for (int i = 0; i < 6; ++i) total_classes[i] += (int64_t)s.h_ctr->classes[i];
        return IDP_OK;
    };

    // one chunk: stage, run the chain, bring the records back -- all asynchronous on the slot's stream
    auto submit = [&](Slot &s, int64_t b, int64_t e) -> int {
        const int64_t cn = e - b;
        cudaStream_t st = s.stream;
        int rc2;
        uint64_t byte0, byte1;  // byte range of this chunk's bases
        if (reads->fixed_len > 0) { const uint64_t stride = (uint64_t)((reads->fixed_len + 1) >> 1); byte0 = (uint64_t)b * stride; byte1 = (uint64_t)e * stride; }
        else {
            byte0 = reads->seq_off[b]; byte1 = 0;
            for (int64_t r = b; r < e; ++r) {
                const int32_t len = reads->len[r];
                if (len < 0 || len > kMaxReadLen) { set_error("read %lld: length %d outside [0, %d]", (long long)r, len, kMaxReadLen); return IDP_ERR_UNSUPPORTED; }
                byte0 = std::min<uint64_t>(byte0, reads->seq_off[r]);
                byte1 = std::max<uint64_t>(byte1, reads->seq_off[r] + (uint64_t)((len + 1) >> 1));
            }
            if (byte1 < byte0) byte1 = byte0;
        }
        if ((rc2 = s.seq4.ensure(byte1 - byte0 + 16)) || (rc2 = s.flags.ensure(cn * 2)) || (rc2 = s.five.ensure(cn * rec))) return rc2;
        if (do3 && (rc2 = s.three.ensure(cn * rec))) return rc2;
        if (cls_bytes && (rc2 = s.cls.ensure(cn))) return rc2;
        CUDA_TRY(cudaMemcpyAsync(s.seq4.p, reads->seq4 + byte0, byte1 - byte0, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(s.flags.p, reads->flags + b, cn * 2, cudaMemcpyHostToDevice, st));
        c->last.h2d_bytes += (int64_t)(byte1 - byte0) + cn * 2;
        DevReads R{};
        R.n = (int32_t)cn; R.fixed_len = reads->fixed_len; R.flags = s.flags.as<uint16_t>();
        R.seq4 = s.seq4.as<uint8_t>() - (reads->fixed_len > 0 ? 0 : byte0);
        if (reads->fixed_len <= 0) {
            if ((rc2 = s.seq_off.ensure(cn * 8)) || (rc2 = s.len.ensure(cn * 4))) return rc2;
            CUDA_TRY(cudaMemcpyAsync(s.seq_off.p, reads->seq_off + b, cn * 8, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(s.len.p, reads->len + b, cn * 4, cudaMemcpyHostToDevice, st));
            R.seq_off = s.seq_off.as<uint64_t>(); R.len = s.len.as<int32_t>();
            c->last.h2d_bytes += cn * 12;
        }
        if (have_loc) {
            if ((rc2 = s.ref_idx.ensure(cn * 4)) || (rc2 = s.ustart.ensure(cn * 4)) || (rc2 = s.uend.ensure(cn * 4))) return rc2;
            CUDA_TRY(cudaMemcpyAsync(s.ref_idx.p, reads->ref_idx + b, cn * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(s.ustart.p, reads->unclipped_start + b, cn * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(s.uend.p, reads->unclipped_end + b, cn * 4, cudaMemcpyHostToDevice, st));
            R.ref_idx = s.ref_idx.as<int32_t>(); R.ustart = s.ustart.as<int32_t>(); R.uend = s.uend.as<int32_t>();
            c->last.h2d_bytes += cn * 12;
        }
        s.R = R; s.chunk_begin = b; s.chunk_n = cn;
        s.busy = true;  // from here on the slot has work in flight, whatever happens below
        if ((rc2 = enqueue_chain(c, s, st, R, ResultOut{s.five.p, compact}, ResultOut{do3 ? s.three.p : nullptr, compact}, cls_bytes ? s.cls.as<uint8_t>() : nullptr))) return rc2;
        CUDA_TRY(cudaMemcpyAsync(five_b + b * rec, s.five.p, cn * rec, cudaMemcpyDeviceToHost, st));
        if (cls_bytes) { CUDA_TRY(cudaMemcpyAsync(c->cls_user + b, s.cls.p, cn, cudaMemcpyDeviceToHost, st)); c->last.d2h_bytes += cn; }
        c->last.d2h_bytes += cn * (int64_t)rec;
        if (do3) { CUDA_TRY(cudaMemcpyAsync(three_b + b * rec, s.three.p, cn * rec, cudaMemcpyDeviceToHost, st)); c->last.d2h_bytes += cn * (int64_t)rec; }
        return IDP_OK;
    };

    int slot_i = 0;
    bool began = false;
    for (int64_t b = 0; b < n_reads;) {
        int64_t e = std::min<int64_t>(b + chunk, n_reads);
        if (e < n_reads && (reads->flags[e - 1] & IDP_F_MATE_NEXT)) ++e;  // never split a template
        Slot &s = c->slots[slot_i];
        slot_i = (slot_i + 1) % kSlots;
        if ((rc = harvest(s))) { drain_slots(c); return rc; }
        if (!began) { if (cudaEventRecord(c->ev_begin, s.stream) != cudaSuccess) { set_error("cudaEventRecord failed"); drain_slots(c); return IDP_ERR_CUDA; } began = true; }
        if ((rc = submit(s, b, e))) { drain_slots(c); return rc; }
        b = e;
    }
    for (int k = 0; k < kSlots; ++k) if ((rc = harvest(c->slots[k]))) { drain_slots(c); return rc; }
    if (began) {
        CUDA_TRY(cudaEventRecord(c->ev_finish, c->slots[0].stream));
        CUDA_TRY(cudaEventSynchronize(c->ev_finish));
        float ms = 0; cudaEventElapsedTime(&ms, c->ev_begin, c->ev_finish);
        c->last.ms_total = ms;
    }
    if (metrics160) for (int i = 0; i < 160; ++i) metrics160[i] += total_metrics[i];
    if (n_gapped_alignments) *n_gapped_alignments += total_align;
    if (c->cls_on && c->cls_counts) for (int i = 0; i < 6; ++i) c->cls_counts[i] += total_classes[i];
    return IDP_OK;
}

}  // namespace idp

extern "C" {

int idp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void *idp_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); set_error("cudaHostAlloc(%zu) failed", bytes); return nullptr; }
    return p;
}
void idp_host_free(void *p) { if (p) cudaFreeHost(p); }

int idp_ctx_create(const idp_panel *panel, int device, idp_ctx **out) {
    if (!panel || !out) { set_error("idp_ctx_create: NULL argument"); return IDP_ERR_ARG; }
    int ndev = idp_device_count();
    if (ndev <= 0) { set_error("no CUDA device is usable; the IdentifyPrimers GPU path has no CPU fallback"); return IDP_ERR_CUDA; }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (0..%d)", device, ndev - 1); return IDP_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    // the library holds sm_100a code only: architecture-specific cubins do not run on any other major version (nor on sm_12x)
    if (prop.major != 10) { set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return IDP_ERR_CUDA; }
    auto *c = new idp_ctx();
    c->panel = panel; c->device = device; c->cfg.sm_count = prop.multiProcessorCount;
    c->cfg.scan_stats = getenv("IDP_SCAN_STATS") && atoi(getenv("IDP_SCAN_STATS")) != 0;
    int rc = upload_panel(c);
    if (rc) { idp_ctx_destroy(c); return rc; }
    if (cudaEventCreate(&c->ev_begin) != cudaSuccess || cudaEventCreate(&c->ev_finish) != cudaSuccess) { set_error("cudaEventCreate failed"); idp_ctx_destroy(c); return IDP_ERR_CUDA; }
    for (Slot &s : c->slots) {
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); idp_ctx_destroy(c); return IDP_ERR_CUDA; }
        for (auto &e : s.ev) if (cudaEventCreate(&e) != cudaSuccess) { set_error("cudaEventCreate failed"); idp_ctx_destroy(c); return IDP_ERR_CUDA; }
        if (cudaMalloc(&s.d_ctr, sizeof(Counters)) != cudaSuccess || cudaHostAlloc(&s.h_ctr, sizeof(Counters), cudaHostAllocDefault) != cudaSuccess) {
            set_error("counter allocation failed"); idp_ctx_destroy(c); return IDP_ERR_NOMEM;
        }
    }
    *out = c;
    return IDP_OK;
}

void idp_ctx_destroy(idp_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (Slot &s : c->slots) {
        for (DevBuf *b : {&s.seq4, &s.seq_off, &s.len, &s.ref_idx, &s.ustart, &s.uend, &s.flags, &s.five, &s.three, &s.fall, &s.parked, &s.masks, &s.cls, &s.spill, &s.all_list,
                          &s.seg_off, &s.seg_cnt, &s.pair_work, &s.pair_primer, &s.pair_score, &s.pair_se, &s.fin, &s.rtype}) b->release();
        if (s.d_ctr) cudaFree(s.d_ctr);
        if (s.h_ctr) cudaFreeHost(s.h_ctr);
        for (auto &e : s.ev) if (e) cudaEventDestroy(e);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    if (c->ev_begin) cudaEventDestroy(c->ev_begin);
    if (c->ev_finish) cudaEventDestroy(c->ev_finish);
    for (DevBuf *b : {&c->b_fx, &c->b_pm4, &c->b_dfx, &c->b_cinfo, &c->b_cspan, &c->b_qbot, &c->b_mbits, &c->b_bylen, &c->b_pmask, &c->b_pinfo, &c->b_positive, &c->b_loc_key[0], &c->b_loc_key[1], &c->b_loc_primer[0], &c->b_loc_primer[1],
                      &c->b_mate_off, &c->b_mate_idx, &c->b_smat, &c->b_keys[0], &c->b_keys[1], &c->b_offcnt[0], &c->b_offcnt[1],
                      &c->b_postings[0], &c->b_postings[1], &c->b_heads[0], &c->b_heads[1], &c->b_direct[0], &c->b_direct[1], &c->b_dbits[0], &c->b_dbits[1], &c->b_dlohi, &c->b_bloom, &c->b_bsmm, &c->b_bsacc, &c->b_seedtab, &c->b_seedlist, &c->b_diag,
                      &c->al_q, &c->al_ql, &c->al_t, &c->al_to, &c->al_tl, &c->al_tp, &c->al_out}) b->release();
    delete c;
}

int idp_ctx_set_stream(idp_ctx *ctx, void *cuda_stream) {
    if (!ctx) { set_error("idp_ctx_set_stream: NULL context"); return IDP_ERR_ARG; }
    ctx->user_stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return IDP_OK;
}

int idp_ctx_timings(const idp_ctx *ctx, idp_timings *out) {
    if (!ctx || !out) { set_error("idp_ctx_timings: NULL argument"); return IDP_ERR_ARG; }
    *out = ctx->last;
    return IDP_OK;
}

int idp_match_batch_device(idp_ctx *c, const idp_reads *reads, int32_t n_reads, idp_result *five, idp_result *three,
                           int64_t *metrics160, int64_t *n_gapped_alignments) {
    return match_batch_device(c, reads, n_reads, five, three, false, metrics160, n_gapped_alignments);
}
int idp_match_batch16_device(idp_ctx *c, const idp_reads *reads, int32_t n_reads, idp_result16 *five, idp_result16 *three,
                             int64_t *metrics160, int64_t *n_gapped_alignments) {
    return match_batch_device(c, reads, n_reads, five, three, true, metrics160, n_gapped_alignments);
}
int idp_match_batch(idp_ctx *c, const idp_reads *reads, int32_t n_reads, idp_result *five, idp_result *three,
                    int64_t *metrics160, int64_t *n_gapped_alignments) {
    return match_batch_host(c, reads, n_reads, five, three, false, metrics160, n_gapped_alignments);
}
int idp_match_batch16(idp_ctx *c, const idp_reads *reads, int32_t n_reads, idp_result16 *five, idp_result16 *three,
                      int64_t *metrics160, int64_t *n_gapped_alignments) {
    return match_batch_host(c, reads, n_reads, five, three, true, metrics160, n_gapped_alignments);
}

int idp_align_policy(void) { return AlignPolicy::kBits; }

int idp_ctx_set_async(idp_ctx *c, int enable) {
    if (!c) { set_error("idp_ctx_set_async: NULL context"); return IDP_ERR_ARG; }
    if (c->async_pending) { set_error("idp_ctx_set_async: batches are pending; call idp_ctx_sync first"); return IDP_ERR_ARG; }
    c->async_on = enable != 0;
    c->async_lean = enable == 2;
    return IDP_OK;
}

int idp_ctx_sync(idp_ctx *c, int64_t *metrics160, int64_t *n_gapped_alignments) {
    if (!c) { set_error("idp_ctx_sync: NULL context"); return IDP_ERR_ARG; }
    if (c->async_pending == 0) return IDP_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    Slot &s = c->slots[0];
    cudaStream_t st = c->user_stream ? c->user_stream : s.stream;
    c->async_pending = 0;
    if (c->async_lean) CUDA_TRY(cudaMemcpyAsync(s.h_ctr, s.d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    if (s.h_ctr->overflow) {
        s.pair_cap *= 4;  // the next batches get more room
        set_error("a batch listed more -K candidates than the pair buffer holds (asynchronous mode cannot redo it): rerun the batches since the last idp_ctx_sync");
        return IDP_ERR_NOMEM;
    }
    c->last = idp_timings{};
    add_timings(c, s, true);  // stage times of the last batch; the work counters cover every batch since the last sync
    float ms = 0; cudaEventElapsedTime(&ms, s.ev[EV_START], s.ev[EV_END]);
    c->last.ms_total = ms;
    if (metrics160) for (int i = 0; i < 160; ++i) metrics160[i] += (int64_t)s.h_ctr->metrics[i];
    if (n_gapped_alignments) *n_gapped_alignments += (int64_t)s.h_ctr->n_align5;
    if (c->cls_on && c->cls_counts) for (int i = 0; i < 6; ++i) c->cls_counts[i] += (int64_t)s.h_ctr->classes[i];
    return IDP_OK;
}

int idp_ctx_set_classify(idp_ctx *c, int32_t min_insert_length, int32_t max_insert_length, uint8_t *cls_per_read, int64_t *counts6) {
    if (!c) { set_error("idp_ctx_set_classify: NULL context"); return IDP_ERR_ARG; }
    c->cls_on = min_insert_length >= 0;
    c->cls_min = min_insert_length; c->cls_max = max_insert_length;
    c->cls_user = c->cls_on ? cls_per_read : nullptr;
    c->cls_counts = c->cls_on ? counts6 : nullptr;
    return IDP_OK;
}

int idp_classify_batch_device(idp_ctx *c, const uint16_t *flags, const void *five, int compact, int32_t n_reads, int32_t min_insert_length,
                              int32_t max_insert_length, uint8_t *cls_per_read, int64_t *counts6) {
    if (!c || n_reads < 0 || (n_reads > 0 && (!flags || !five))) { set_error("idp_classify_batch_device: NULL argument"); return IDP_ERR_ARG; }
    if (n_reads == 0) return IDP_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    Slot &s = c->slots[0];
    cudaStream_t st = c->user_stream ? c->user_stream : s.stream;
    CUDA_TRY(cudaMemsetAsync(s.d_ctr->classes, 0, sizeof(s.d_ctr->classes), st));
    launch_classify(c->cfg, st, flags, ResultOut{const_cast<void *>(five), compact != 0}, n_reads, min_insert_length, max_insert_length, cls_per_read, s.d_ctr);
    CUDA_TRY(cudaMemcpyAsync(s.h_ctr->classes, s.d_ctr->classes, sizeof(s.d_ctr->classes), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    if (counts6) for (int i = 0; i < 6; ++i) counts6[i] += (int64_t)s.h_ctr->classes[i];
    return IDP_OK;
}

void idp_result16_unpack(const idp_result16 *in, int64_t n, idp_result *out) {
    for (int64_t i = 0; i < n; ++i) {
        const idp_result16 r = in[i];
        idp_result o;
        memset(&o, 0, sizeof o);
        o.primer = (int32_t)(r.head & 0xFFFFFFu) - 1;
        o.type = (uint8_t)((r.head >> 24) & 3u);
        o.multi = (uint8_t)((r.head >> 26) & 1u);
        o.status = (uint8_t)((r.head >> 27) & 7u);
        if (o.type == IDP_GAPPED) { o.raw_score = r.a; o.raw_next = r.b; }
        else { o.mm = r.a; o.next_mm = r.b == IDP_NEXT16_NA ? IDP_NEXT_NA : r.b; }
        o.q_len = (int16_t)r.q_len; o.q_len_next = (int16_t)r.q_len_next;
        o.t_start = (int16_t)r.t_start; o.t_end = (int16_t)r.t_end;
        out[i] = o;
    }
}

// ---- batched AsyncAligner.align (AA:37-54) ----
// Persistent per-context device buffers, every copy and kernel on the context's stream; the ASCII is packed and validated on
// the GPU (align_pack_kernel), queries of up to 64 bases run on the register-resident aligners.
int idp_align_batch(idp_ctx *c, int mode, const char *queries, const int64_t *q_off, const char *targets, const int64_t *t_off,
                    int32_t n, int32_t *score, int32_t *target_start, int32_t *target_end, int32_t *query_start, int32_t *query_end) {
    if (!c || !queries || !q_off || !targets || !t_off || n < 0 || !score || !target_start || !target_end) { set_error("idp_align_batch: NULL argument"); return IDP_ERR_ARG; }
    if ((query_start == nullptr) != (query_end == nullptr)) { set_error("idp_align_batch: query_start and query_end go together"); return IDP_ERR_ARG; }
    if (mode != IDP_MODE_LOCAL && mode != IDP_MODE_GLOCAL && mode != IDP_MODE_GLOBAL) { set_error("unknown alignment mode %d", mode); return IDP_ERR_ARG; }
    if (n == 0) return IDP_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    int maxq = 1;
    if (q_off[0] != 0 || t_off[0] != 0) { set_error("idp_align_batch: offsets must start at 0"); return IDP_ERR_ARG; }
    for (int i = 0; i < n; ++i) {
        const int64_t ql = q_off[i + 1] - q_off[i], tl = t_off[i + 1] - t_off[i];
        if (ql < 1 || ql > kMaxPrimerBases) { set_error("query %d: length %lld outside [1, %d]", i, (long long)ql, kMaxPrimerBases); return IDP_ERR_UNSUPPORTED; }
        if (tl < 0 || tl > kMaxReadLen) { set_error("target %d: length %lld outside [0, %d]", i, (long long)tl, kMaxReadLen); return IDP_ERR_UNSUPPORTED; }
        maxq = std::max<int>(maxq, (int)ql);
    }
    const int64_t q_bytes = q_off[n], t_bytes = t_off[n];
    const size_t off_bytes = (size_t)(n + 1) * 8;
    const size_t words = (size_t)n * std::max(8, (maxq + 7) / 8);
    int rc;
    if ((rc = c->al_q.ensure((size_t)q_bytes + 16)) || (rc = c->al_t.ensure((size_t)t_bytes + 16)) || (rc = c->al_ql.ensure(off_bytes)) ||
        (rc = c->al_tl.ensure(off_bytes)) || (rc = c->al_to.ensure(words * 4)) || (rc = c->al_tp.ensure(align_target_bytes(t_bytes, n))) ||
        (rc = c->al_out.ensure((size_t)n * 20 + 16))) return rc;
    cudaStream_t st = c->user_stream ? c->user_stream : c->slots[0].stream;
    int32_t *o = c->al_out.as<int32_t>();
    unsigned *bad = reinterpret_cast<unsigned *>(o + 5 * (size_t)n);
    cudaError_t e = cudaMemsetAsync(bad, 0, 4, st);
    // uploads ordered on the stream the kernels run on (a pageable source is staged before each call returns)
    if (e == cudaSuccess && q_bytes) e = cudaMemcpyAsync(c->al_q.p, queries, (size_t)q_bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && t_bytes) e = cudaMemcpyAsync(c->al_t.p, targets, (size_t)t_bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->al_ql.p, q_off, off_bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->al_tl.p, t_off, off_bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        launch_align(c->cfg, st, mode, c->al_q.as<char>(), c->al_ql.as<int64_t>(), c->al_t.as<char>(), c->al_tl.as<int64_t>(), n, maxq,
                     c->al_to.as<uint32_t>(), c->al_tp.as<uint8_t>(), bad, o, o + n, o + 2 * (size_t)n,
                     query_start ? o + 3 * (size_t)n : nullptr, query_start ? o + 4 * (size_t)n : nullptr);
        e = cudaGetLastError();
    }
    unsigned h_bad = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(score, o, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(target_start, o + n, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(target_end, o + 2 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && query_start) e = cudaMemcpyAsync(query_start, o + 3 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && query_end) e = cudaMemcpyAsync(query_end, o + 4 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, st);
    const cudaError_t es = cudaStreamSynchronize(st);  // always: nothing may stay in flight into the caller's arrays
    if (e != cudaSuccess || es != cudaSuccess) {
        set_error("idp_align_batch: CUDA failure: %s", cudaGetErrorString(e != cudaSuccess ? e : es));
        cudaGetLastError();
        return IDP_ERR_CUDA;
    }
    if (h_bad) { set_error("idp_align_batch: only upper-case IUPAC bases and '=' are supported"); return IDP_ERR_UNSUPPORTED; }
    return IDP_OK;
}

}  // extern "C"
