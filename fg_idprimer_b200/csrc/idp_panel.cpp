// Host-side construction of the immutable panel state: what the reference builds once in the constructors of
// LocationBasedPrimerMatcher / UngappedAlignmentBasedPrimerMatcher / GappedAlignmentBasedPrimerMatcher
// (IdentifyPrimers.scala:223-234; PrimerMatcher.scala:62-117, 156-175), laid out for the GPU kernels.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <numeric>

#include "idp_internal.h"

namespace idp {

static thread_local char g_error[1024] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}
const char *last_error() { return g_error; }

// BAM nibble <-> base letter ("=ACMGRSVTWYHKDBN"); the nibble IS the IUPAC bit mask (A=1, C=2, G=4, T=8)
static const char kNibbleToChar[17] = "=ACMGRSVTWYHKDBN";
int base_to_nibble(unsigned char c) {
    switch (c) {
        case '=': return 0;
        case 'A': case 'a': return 1;  case 'C': case 'c': return 2;  case 'M': case 'm': return 3;
        case 'G': case 'g': return 4;  case 'R': case 'r': return 5;  case 'S': case 's': return 6;
        case 'V': case 'v': return 7;  case 'T': case 't': return 8;  case 'W': case 'w': return 9;
        case 'Y': case 'y': return 10; case 'H': case 'h': return 11; case 'K': case 'k': return 12;
        case 'D': case 'd': return 13; case 'B': case 'b': return 14;
        default: return 15;  // htsjdk bytesToCompressedBases maps everything else to N
    }
}
char nibble_to_base(int n) { return kNibbleToChar[n & 15]; }

// ------------------------------------------------------------------------------------------------------------
// Iteration order of the reference's `Map[String, Set[Primer]]` values (PrimerMatcher.scala:78-91).  The order in
// which a k-mer's primers are visited decides ties of the stable sorts at PM:241 and PM:309, and it is a property
// of Scala 2.12's collections: primers are added to a mutable.HashSet (open-addressing FlatHashTable keyed on the
// case-class hashCode) and the set is then rebuilt as an immutable Set (insertion order up to 4 elements, a
// 32-way hash trie from 5 on).  Nothing in the reference pins this; it is reproduced from the library's algorithm.
// ------------------------------------------------------------------------------------------------------------
namespace scala212 {

inline uint32_t rotl(uint32_t v, int r) { return (v << r) | (v >> (32 - r)); }
inline uint32_t rotr(uint32_t v, int r) { return r == 0 ? v : (v >> r) | (v << (32 - r)); }

uint32_t string_hash(const std::string &s) {  // java.lang.String.hashCode for ASCII
    uint32_t h = 0;
    for (unsigned char c : s) h = h * 31u + c;
    return h;
}

struct Murmur {  // scala.runtime.Statics.mix / finalizeHash
    uint32_t acc = 0xcafebabeu;
    int fields = 0;
    void add(uint32_t data) {
        uint32_t k = rotl(data * 0xcc9e2d51u, 15) * 0x1b873593u;
        acc = rotl(acc ^ k, 13) * 5u + 0xe6546b64u;
        ++fields;
    }
    uint32_t done() const {
        uint32_t h = acc ^ (uint32_t)fields;
        h ^= h >> 16; h *= 0x85ebca6bu;
        h ^= h >> 13; h *= 0xc2b2ae35u;
        return h ^ (h >> 16);
    }
};

// hashCode of case class Primer(pair_id, primer_id, sequence, positive_strand, ref_name, start, end) (Primer.scala:182-188)
uint32_t primer_hash(const std::string &pair_id, const std::string &primer_id, const std::string &sequence, bool positive,
                     const std::string &ref_name, int32_t start, int32_t end) {
    Murmur m;
    m.add(string_hash(pair_id));
    m.add(string_hash(primer_id));
    m.add(string_hash(sequence));
    m.add(positive ? 1231u : 1237u);  // java.lang.Boolean.hashCode
    m.add(string_hash(ref_name));
    m.add((uint32_t)start);
    m.add((uint32_t)end);
    return m.done();
}

// mutable.HashSet: scala.collection.mutable.FlatHashTable with the default load factor (450/1000) and 32 initial slots
class FlatHashTable {
  public:
    explicit FlatHashTable(const std::vector<uint32_t> &hashes) : hashes_(hashes), slots_(32, -1) {}
    void add(int id) {
        place(slots_, id);
        if (++size_ >= threshold()) grow();
    }
    std::vector<int> in_iteration_order() const {
        std::vector<int> out;
        for (int v : slots_) if (v >= 0) out.push_back(v);
        return out;
    }
  private:
    size_t threshold() const { return (size_t)((uint64_t)slots_.size() * 450 / 1000); }
    static uint32_t byteswap32(uint32_t v) { return __builtin_bswap32(v * 0x9e3775cdu) * 0x9e3775cdu; }
    uint32_t index_of(uint32_t hcode, size_t len) const {
        uint32_t ones = (uint32_t)len - 1;
        int bits = __builtin_popcount(ones);            // seedvalue = tableSizeSeed = bitCount(length - 1)
        uint32_t improved = rotr(byteswap32(hcode), bits % 32);
        return (improved >> (32 - bits)) & ones;
    }
    void place(std::vector<int> &table, int id) const {
        size_t h = index_of(hashes_[id], table.size());
        while (table[h] >= 0) h = (h + 1) % table.size();
        table[h] = id;
    }
    void grow() {
        std::vector<int> bigger(slots_.size() * 2, -1);
        for (int v : slots_) if (v >= 0) place(bigger, v);  // growTable re-adds in old slot order
        slots_.swap(bigger);
    }
    const std::vector<uint32_t> &hashes_;
    std::vector<int> slots_;
    size_t size_ = 0;
};

// immutable.HashSet (hash trie): depth-first iteration = ascending 5-bit digits of improve(hash), low digits first
bool trie_less(uint32_t ha, uint32_t hb) {
    auto improve = [](uint32_t hcode) {
        uint32_t h = hcode + ~(hcode << 9);
        h ^= h >> 14;
        h += h << 4;
        return h ^ (h >> 10);
    };
    uint32_t a = improve(ha), b = improve(hb);
    for (int lvl = 0; lvl < 7; ++lvl) {
        uint32_t da = (a >> (5 * lvl)) & 31u, db = (b >> (5 * lvl)) & 31u;
        if (da != db) return da < db;
    }
    return false;
}

// primers (ids, in the order they were added) -> order in which `set.toSet` iterates
std::vector<int> set_iteration_order(const std::vector<int> &added, const std::vector<uint32_t> &hashes) {
    if (added.size() <= 1) return added;
    FlatHashTable t(hashes);
    for (int id : added) t.add(id);
    std::vector<int> order = t.in_iteration_order();
    if (order.size() >= 5)  // Set1..Set4 keep builder order; the fifth element switches to the trie
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return trie_less(hashes[x], hashes[y]); });
    return order;
}

}  // namespace scala212

static bool valid_primer_base(char c) { return strchr("ACGTMRWSYKVHDBN", c) != nullptr && c != 0; }

// PM:78-91 for one matcher
static bool build_kmer_index(const idp_panel &p, int k, KmerIndexHost &out) {
    out = KmerIndexHost();
    out.k = k;
    if (k <= 0) return true;
    std::map<std::string, std::vector<int>> kmers;  // the reference uses a TreeMap too; only per-key order matters
    for (int i = 0; i < p.n; ++i) {
        const std::string &s = p.sequence[i];
        if ((int)s.size() < k) continue;  // sliding(k) of a shorter string yields that string, which no read k-mer equals
        for (size_t w = 0; w + k <= s.size(); ++w) {
            auto &v = kmers[s.substr(w, k)];
            if (std::find(v.begin(), v.end(), i) == v.end()) v.push_back(i);
        }
    }
    size_t slots = 16;
    while (slots < 8 * kmers.size() + 2 && slots < (size_t(1) << 22)) slots *= 2;  // sparse: a warp probes until its slowest lane is done
    while (slots < 2 * kmers.size() + 2) slots *= 2;
    out.keys.assign(slots, 0);
    out.off.assign(slots, 0);
    out.cnt.assign(slots, 0);
    int bits = 0;
    while ((size_t(1) << bits) < slots) ++bits;
    for (auto &kv : kmers) {
        std::vector<int> order = scala212::set_iteration_order(kv.second, p.hash);
        uint64_t key = 0;
        for (char c : kv.first) key = (key << 4) | (uint64_t)base_to_nibble((unsigned char)c);
        uint32_t off = (uint32_t)out.postings.size();
        for (int id : order) out.postings.push_back((uint32_t)id);
        out.by_string[kv.first] = {off, (uint32_t)order.size()};
        size_t h = (size_t)((key * 0x9E3779B97F4A7C15ull) >> (64 - bits));
        while (out.keys[h] != 0) h = (h + 1) & (slots - 1);
        out.keys[h] = key;
        out.off[h] = off;
        out.cnt[h] = (uint32_t)order.size();
    }
    if (out.postings.empty()) out.postings.push_back(0);
    // head-filtered postings for the scan kernel: {primer id | (accept threshold + 1) << 24, first 8 bases of the mask}
    out.heads.resize(out.postings.size() * 2);
    for (size_t j = 0; j < out.postings.size(); ++j) {
        const uint32_t id = out.postings[j];
        const int acc1 = std::min(p.acc_full[id] + 1, 255);
        out.heads[2 * j] = id | ((uint32_t)acc1 << 24);
        out.heads[2 * j + 1] = p.pmask[(size_t)id * p.ww];
    }
    // direct-address table over the 4^k all-ACGT k-mers (2 bits per base): entry = posting offset (20 bits) | count (12 bits)
    if (k <= 11 && out.postings.size() < (1u << 20) && p.n < (1 << 20)) {  // up to 16 MB; the scan kernel stages it in shared memory when k <= 6
        bool fits = true;
        std::vector<uint32_t> direct((size_t)1 << (2 * k), 0);
        for (auto &kv : out.by_string) {
            uint32_t code = 0;
            bool acgt = true;
            for (char c : kv.first) {
                const char *at = strchr("ACGT", c);
                if (!at || !c) { acgt = false; break; }
                code = (code << 2) | (uint32_t)(at - "ACGT");
            }
            if (!acgt) continue;
            if (kv.second.second >= (1u << 12)) { fits = false; break; }
            direct[code] = kv.second.first | (kv.second.second << 20);
        }
        if (fits) out.direct.swap(direct);
    }
    return true;
}

}  // namespace idp

using namespace idp;

extern "C" {

int idp_abi_version(void) { return IDP_ABI_VERSION; }
const char *idp_last_error(void) { return idp::last_error(); }

int idp_panel_create(const idp_primer *primers, int32_t n, const idp_params *params, idp_panel **out) {
    if (!primers || !params || !out || n <= 0) { set_error("idp_panel_create: NULL argument or empty panel"); return IDP_ERR_ARG; }
    // the tool's own validation (IdentifyPrimers.scala:216-221)
    if (!(params->max_mismatch_rate >= 0)) { set_error("--max-mismatches must be >= 0"); return IDP_ERR_ARG; }
    if (params->match_score < 0) { set_error("--match-score must be >= 0"); return IDP_ERR_ARG; }
    if (params->mismatch_score > 0) { set_error("--mismatch-score must be <= 0"); return IDP_ERR_ARG; }
    if (params->gap_open > 0) { set_error("--gap-open must be <= 0"); return IDP_ERR_ARG; }
    if (params->gap_extend > 0) { set_error("--gap-extend must be <= 0"); return IDP_ERR_ARG; }
    if (!(params->min_alignment_score_rate >= 0)) { set_error("--min-alignment-score-rate must be >= 0"); return IDP_ERR_ARG; }
    if (params->slop < 0 || params->slop > 1024) { set_error("slop %d outside [0, 1024]", params->slop); return IDP_ERR_UNSUPPORTED; }
    const int32_t kScoreLimit = 1 << 15;
    if (params->match_score > kScoreLimit || params->mismatch_score < -kScoreLimit || params->gap_open < -kScoreLimit || params->gap_extend < -kScoreLimit) {
        set_error("alignment scores beyond +-%d are not supported", kScoreLimit); return IDP_ERR_UNSUPPORTED;
    }
    if (params->k_ungapped > kMaxK || params->k_gapped > kMaxK) { set_error("k-mer lengths above %d are not supported", kMaxK); return IDP_ERR_UNSUPPORTED; }

    auto *p = new idp_panel();
    p->n = n;
    p->params = *params;
    p->params.score_matrix = nullptr;
    std::map<std::string, int> pair_index;
    for (int i = 0; i < n; ++i) {
        const idp_primer &in = primers[i];
        if (!in.pair_id || !in.primer_id || !in.sequence) { set_error("primer %d: NULL string", i); delete p; return IDP_ERR_ARG; }
        std::string seq = in.sequence, ref = in.ref_name ? in.ref_name : "";
        if (strchr(in.pair_id, ',') || strchr(in.primer_id, ',')) { set_error("primer %d: id contains a comma (Primer.scala:190-191)", i); delete p; return IDP_ERR_ARG; }
        if (seq.empty() || (int)seq.size() > kMaxPrimerBases) { set_error("primer %s: sequence length %zu outside [1, %d]", in.primer_id, seq.size(), kMaxPrimerBases); delete p; return IDP_ERR_UNSUPPORTED; }
        for (char c : seq) {
            if (c == '.') { set_error("primer %s: '.' bases are not supported by the GPU path", in.primer_id); delete p; return IDP_ERR_UNSUPPORTED; }
            if (!valid_primer_base(c)) { set_error("Found an invalid base for primer %s/%s: [%c] (sequence must be upper-case IUPAC)", in.pair_id, in.primer_id, c); delete p; return IDP_ERR_ARG; }
        }
        int32_t plen = ref.empty() ? (int32_t)seq.size() : in.end - in.start + 1;  // Primer.scala:206
        if (plen < 1 || plen > kMaxPrimerBases) { set_error("primer %s: Primer.length %d outside [1, %d] is not supported", in.primer_id, plen, kMaxPrimerBases); delete p; return IDP_ERR_UNSUPPORTED; }
        if (!ref.empty() && in.start < 1) { set_error("primer %s: mapped primer with start < 1 is not supported", in.primer_id); delete p; return IDP_ERR_UNSUPPORTED; }
        p->pair_id.push_back(in.pair_id); p->primer_id.push_back(in.primer_id); p->sequence.push_back(seq); p->ref_name.push_back(ref);
        p->ref_idx.push_back(ref.empty() ? -1 : (in.ref_idx >= 0 ? in.ref_idx : -2));
        p->start.push_back(in.start); p->end.push_back(in.end);
        p->seq_len.push_back((int32_t)seq.size()); p->plen.push_back(plen);
        p->positive.push_back(in.positive_strand ? 1 : 0);
        p->hash.push_back(scala212::primer_hash(in.pair_id, in.primer_id, seq, in.positive_strand != 0, ref, in.start, in.end));
        auto it = pair_index.find(in.pair_id);
        if (it == pair_index.end()) it = pair_index.emplace(in.pair_id, (int)pair_index.size()).first;
        p->pair_of.push_back(it->second);
        p->max_seq_len = std::max(p->max_seq_len, (int32_t)seq.size());
        p->max_primer_length = std::max(p->max_primer_length, plen);
    }
    {   // integer keys for the strings the downstream classification compares (post_process_bam.py:60-75)
        std::map<std::string, int> ids, refs;
        for (int i = 0; i < n; ++i) {
            p->primer_id_key.push_back(ids.emplace(p->primer_id[i], (int)ids.size()).first->second);
            p->ref_name_key.push_back(refs.emplace(p->ref_name[i], (int)refs.size()).first->second);
        }
    }
    {   // rows that are equal as case classes would collapse in the reference's Sets; refuse them
        std::map<std::array<std::string, 4>, std::vector<int>> seen;
        for (int i = 0; i < n; ++i) {
            auto &v = seen[{p->pair_id[i], p->primer_id[i], p->sequence[i], p->ref_name[i]}];
            for (int j : v) if (p->positive[j] == p->positive[i] && p->start[j] == p->start[i] && p->end[j] == p->end[i]) {
                set_error("primers %d and %d are identical rows", j, i); delete p; return IDP_ERR_UNSUPPORTED;
            }
            v.push_back(i);
        }
    }
    // scan window: the read prefix that mismatch counting (PM:139) and the k-mer search (PM:110) can touch
    p->window_bases = std::max(p->max_seq_len, p->max_primer_length);
    p->ww = (p->max_seq_len + 7) / 8;
    p->pmask.assign((size_t)n * p->ww, 0xFFFFFFFFu);  // nibble 0xF beyond the primer: any read base "matches"
    for (int i = 0; i < n; ++i)
        for (int b = 0; b < p->seq_len[i]; ++b) {
            uint32_t &w = p->pmask[(size_t)i * p->ww + b / 8];
            int sh = 4 * (b % 8);
            w = (w & ~(0xFu << sh)) | ((uint32_t)base_to_nibble((unsigned char)p->sequence[i][b]) << sh);
        }
    // mismatchAlign thresholds for a read at least as long as the primer (PM:130-134), in the JVM's double arithmetic
    const double rate = params->max_mismatch_rate;
    for (int i = 0; i < n; ++i) {
        int length = p->plen[i];
        double max_mismatches = length * rate;
        // `while (.. && count <= maxMismatches)`: the count stops at the first integer above maxMismatches
        double capd = std::floor(max_mismatches) + 1.0;
        p->cap_full.push_back(capd > 1e9 ? 1000000000 : (int32_t)capd);
        int acc = -1;
        for (int mm = 0; mm <= p->seq_len[i]; ++mm) { if (mm / (double)length <= rate) acc = mm; else break; }
        p->acc_full.push_back(acc);
    }
    // bit-sliced filter over the first 16 bases, 32 primers per word
    {
        const int groups = (((n + 31) / 32) + 1) & ~1;  // an even count: the kernel unrolls 2 / 4 / 6 / 8 groups; a padding group accepts nothing
        p->bs_mm.assign((size_t)groups * 16 * 16, 0);
        p->bs_acc.assign((size_t)groups * 2, 0);
        for (int i = 0; i < n; ++i) {
            const int g = i / 32, b = i % 32;
            p->acc_max = std::max(p->acc_max, p->acc_full[i]);
            if (p->acc_full[i] >= 0) p->bs_acc[2 * g] |= 1u << b;
            if (p->acc_full[i] >= 1) p->bs_acc[2 * g + 1] |= 1u << b;
            for (int pos = 0; pos < 16; ++pos) {
                const uint32_t mask = pos < p->seq_len[i] ? (uint32_t)base_to_nibble((unsigned char)p->sequence[i][pos]) : 0xFu;
                for (uint32_t nib = 0; nib < 16; ++nib)
                    if (nib & ~mask) p->bs_mm[((((size_t)(g / 2) * 16 + pos) * 16 + nib) << 1) + (g & 1)] |= 1u << b;  // group pairs interleaved
            }
        }
    }
    // seed index for the ungapped scan: when every primer tolerates at most one mismatch (PM:130-134), a primer accepted for
    // a read matches one of the read's two leading S-base seeds without a mismatch.  Each primer seed is entered under every
    // read spelling it matches (PM:143: the read code must be a non-empty submask of the primer code).
    {
        int min_len = INT32_MAX;
        for (int i = 0; i < n; ++i) min_len = std::min(min_len, p->seq_len[i]);
        const int S = (min_len >= 16 && n > 512) ? 8 : 6;
        bool ok = n > 0 && p->acc_max <= 1 && min_len >= 2 * S && n < (1 << 24);
        std::vector<std::pair<uint32_t, uint32_t>> ent[2];  // (seed key, primer)
        for (int i = 0; ok && i < n; ++i)
            for (int t = 0; ok && t < 2; ++t) {
                uint32_t codes[8];
                size_t variants = 1;
                for (int j = 0; j < S; ++j) {
                    codes[j] = (uint32_t)base_to_nibble((unsigned char)p->sequence[i][t * S + j]) & 15u;
                    variants *= (size_t(1) << __builtin_popcount(codes[j])) - 1;
                }
                if (variants == 0 || variants > 1024) { ok = false; break; }
                uint32_t sub[8];
                for (int j = 0; j < S; ++j) sub[j] = codes[j];  // largest submask first
                for (;;) {
                    uint32_t key = 0;
                    for (int j = 0; j < S; ++j) key |= sub[j] << (4 * j);
                    ent[t].emplace_back(key, (uint32_t)i);
                    int j = 0;  // odometer over the non-empty submasks of every position
                    for (; j < S; ++j) {
                        sub[j] = (sub[j] - 1) & codes[j];
                        if (sub[j] != 0) break;
                        sub[j] = codes[j];
                    }
                    if (j == S) break;
                }
            }
        const size_t most = std::max(ent[0].size(), ent[1].size());
        if (ok && most > (size_t(1) << 22)) ok = false;
        if (ok) {
            int bits = 10;
            while ((size_t(1) << bits) < 4 * most) ++bits;
            const uint32_t slots = 1u << bits;
            std::vector<uint32_t> tab(2 * (size_t)slots, 0u), list;
            for (int t = 0; ok && t < 2; ++t) {
                std::vector<std::vector<uint32_t>> per(slots);
                for (auto &e : ent[t]) per[(e.first * 0x9E3779B1u) >> (32 - bits)].push_back(e.second);
                for (uint32_t s = 0; ok && s < slots; ++s) {
                    auto &v = per[s];
                    if (v.empty()) continue;
                    std::sort(v.begin(), v.end());
                    v.erase(std::unique(v.begin(), v.end()), v.end());
                    if (v.size() > 255 || list.size() + v.size() >= (size_t(1) << 24)) { ok = false; break; }
                    if (v.size() == 1) { tab[(size_t)t * slots + s] = 1u | (v[0] << 8); continue; }
                    tab[(size_t)t * slots + s] = (uint32_t)v.size() | ((uint32_t)list.size() << 8);
                    list.insert(list.end(), v.begin(), v.end());
                }
            }
            if (ok) {
                p->seed_tab.swap(tab);
                p->seed_list.swap(list);
                p->seed_S = S; p->seed_shift = 32 - bits; p->seed_slots = (int32_t)slots;
            }
        }
    }
    // exact-prefix table for the scan kernel's common case (DevPanel::fx_tab): every spelling of the first 12 bases of every
    // ISOLATED primer.  Isolated = every other primer q has at least two positions among the first 12 whose codes share no
    // base with this primer's: then no read that spells this prefix without a mismatch can be within one mismatch of q
    // (PM:143: a read code must be a subset of the primer code to match), so the primer is the only one mismatchAlign can
    // accept whenever every primer tolerates at most one mismatch (acc_max <= 1).
    {
        int min_len = INT32_MAX;
        for (int i = 0; i < n; ++i) min_len = std::min(min_len, p->seq_len[i]);
        const bool ok = p->acc_max <= 1 && min_len >= kFxBases && n <= 16384 && n < 65535;
        if (ok) {
            std::vector<uint32_t> a0(n), a1(n);
            for (int i = 0; i < n; ++i) { a0[i] = p->pmask[(size_t)i * p->ww]; a1[i] = p->pmask[(size_t)i * p->ww + 1] | 0xFFFF0000u; }
            auto zero_nibbles = [](uint32_t x) { x |= x >> 1; x |= x >> 2; return __builtin_popcount(~x & 0x11111111u); };
            std::vector<uint8_t> isolated(n, 1);
            for (int i = 0; i < n; ++i)
                for (int j = i + 1; j < n; ++j)
                    if (zero_nibbles(a0[i] & a0[j]) + zero_nibbles(a1[i] & a1[j]) < 2) isolated[i] = isolated[j] = 0;
            // spellings in A/C/G/T only: a read with an ambiguity code in its first 12 bases takes the filter instead
            std::vector<std::array<uint32_t, 3>> ent;  // (w0, w1 low half, primer)
            for (int i = 0; i < n; ++i) {
                if (!isolated[i]) continue;
                uint32_t codes[kFxBases], sub[kFxBases];
                size_t variants = 1;
                for (int j = 0; j < kFxBases; ++j) {
                    codes[j] = ((j < 8 ? a0[i] : a1[i]) >> (4 * (j & 7))) & 15u;
                    variants *= (size_t)__builtin_popcount(codes[j]);
                    if (variants > 256) break;
                }
                if (variants == 0 || variants > 256) continue;  // such a read falls back to the filter, which is exact for every read
                for (int j = 0; j < kFxBases; ++j) sub[j] = codes[j] & (~codes[j] + 1u);  // lowest base of every position
                for (;;) {
                    uint32_t w0 = 0, w1 = 0;
                    for (int j = 0; j < 8; ++j) w0 |= sub[j] << (4 * j);
                    for (int j = 8; j < kFxBases; ++j) w1 |= sub[j] << (4 * (j - 8));
                    // the table is keyed on the row words as STORED (first base of a byte in its high nibble): see idp_scanfx.cuh
                    auto stored = [](uint32_t x) { return ((x & 0x0F0F0F0Fu) << 4) | ((x >> 4) & 0x0F0F0F0Fu); };
                    ent.push_back({stored(w0), stored(w1), (uint32_t)i});
                    int j = 0;  // odometer over the single bases of every position
                    for (; j < kFxBases; ++j) {
                        const uint32_t rest = codes[j] & ~((sub[j] << 1) - 1u);  // bases above the current one
                        if (rest != 0) { sub[j] = rest & (~rest + 1u); break; }
                        sub[j] = codes[j] & (~codes[j] + 1u);
                    }
                    if (j == kFxBases) break;
                }
            }
            if (!ent.empty() && ent.size() <= (size_t(1) << 16)) {
                int bits = 10;
                while ((size_t(1) << bits) < 8 * ent.size()) ++bits;
                p->fx_slots = 1 << bits; p->fx_shift = 32 - bits;
                // a slot holds one spelling; a spelling that loses its slot sends its reads through the filter (still exact), so
                // the multiplier with the fewest collisions among a few candidates is kept
                size_t best_lost = ent.size() + 1;
                std::vector<uint8_t> used;
                for (uint32_t t = 0; t < 32 && best_lost > 0; ++t) {
                    const uint32_t mul = 0x9E3779B1u + 2u * t * 0x632BE5ABu;
                    used.assign((size_t)p->fx_slots, 0);
                    size_t lost = 0;
                    for (auto &e : ent) { uint8_t &u = used[fx_hash(e[0], e[1], mul, p->fx_shift)]; lost += u; u = 1; }
                    if (lost < best_lost) { best_lost = lost; p->fx_mul = mul; }
                }
                p->fx_tab.assign((size_t)p->fx_slots, 0);
                for (auto &e : ent) {
                    uint16_t &slot = p->fx_tab[fx_hash(e[0], e[1], p->fx_mul, p->fx_shift)];
                    if (slot == 0) slot = (uint16_t)(e[2] + 1);
                }
            }
        }
    }
    // location matcher (PM:162-172): one sorted array per strand
    for (int strand = 0; strand < 2; ++strand) {  // 0 = positive, 1 = negative
        std::vector<int> ids;
        for (int i = 0; i < n; ++i)
            if (!p->ref_name[i].empty() && (p->positive[i] != 0) == (strand == 0) && p->start[i] <= p->end[i] /* OverlapDetector drops empty intervals */
                && p->ref_idx[i] >= 0)
                ids.push_back(i);
        auto key_of = [&](int i) { return ((int64_t)p->ref_idx[i] << 32) | (uint32_t)(strand == 0 ? p->start[i] : p->end[i]); };
        std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) { return key_of(a) < key_of(b); });
        for (int i : ids) { p->loc_sortkey[strand].push_back(key_of(i)); p->loc_primer[strand].push_back(i); }
    }
    // pairIdToPrimers(pair_id).filter(_.positive_strand == primer.negativeStrand) (IP:496), file order
    p->mate_off.assign(n + 1, 0);
    {
        std::vector<std::vector<int>> members(pair_index.size());
        for (int i = 0; i < n; ++i) members[p->pair_of[i]].push_back(i);
        for (int i = 0; i < n; ++i) {
            for (int j : members[p->pair_of[i]]) if (p->positive[j] != p->positive[i]) p->mate_idx.push_back(j);
            p->mate_off[i + 1] = (int32_t)p->mate_idx.size();
        }
        if (p->mate_idx.empty()) p->mate_idx.push_back(0);
    }
    if (params->score_matrix) {  // IP:252-278
        p->score_matrix.assign(params->score_matrix, params->score_matrix + 128 * 128);
        for (char a : std::string("ACGT")) for (char b : std::string("ACGT"))
            if (p->score_matrix[(int)a * 128 + (int)b] == INT32_MIN) { set_error("Missing base combination 'primer_pair' '%c' and 'read_base' '%c'", a, b); delete p; return IDP_ERR_ARG; }
        p->smat16.assign(256, INT32_MIN);
        for (int q = 1; q < 16; ++q) for (int t = 0; t < 16; ++t) {
            int32_t s = p->score_matrix[(int)nibble_to_base(q) * 128 + (int)nibble_to_base(t)];
            if (s != INT32_MIN && (s > kScoreLimit || s < -kScoreLimit)) { set_error("scoring matrix entries beyond +-%d are not supported", kScoreLimit); delete p; return IDP_ERR_UNSUPPORTED; }
            p->smat16[q * 16 + t] = s;
        }
        p->params.score_matrix = p->score_matrix.data();
    }
    build_kmer_index(*p, params->k_ungapped, p->kidx[0]);
    build_kmer_index(*p, params->k_gapped, p->kidx[1]);
    *out = p;
    return IDP_OK;
}

void idp_panel_destroy(idp_panel *panel) { delete panel; }
int32_t idp_panel_size(const idp_panel *panel) { return panel ? panel->n : 0; }
// The leading bases (sequencing order) that can influence the 5' chain: the k-mer search window min(maxPrimerLength, len)
// (PM:110), mismatchAlign's min(len, Primer.length) and numMismatches' min(len, sequence.length) (PM:130, 139), the gapped
// target min(sequence.length + slop, len) (PM:293), and the `len < k` tests (PM:109).  A read cut to this many bases gives the
// same results as the whole read.  -1 with --three-prime: matchThreePrime aligns against the whole read (IP:501-505).
int32_t idp_panel_window(const idp_panel *panel) {
    if (!panel || panel->params.three_prime >= 0) return -1;
    int32_t w = std::max(panel->max_primer_length, panel->max_seq_len + std::max(panel->params.slop, 0));
    w = std::max(w, std::max(panel->params.k_ungapped, panel->params.k_gapped));
    return std::max(w, 1);
}
int32_t idp_panel_primer_length(const idp_panel *panel, int32_t i) { return (panel && i >= 0 && i < panel->n) ? panel->plen[i] : -1; }
uint32_t idp_panel_primer_hash(const idp_panel *panel, int32_t i) { return (panel && i >= 0 && i < panel->n) ? panel->hash[i] : 0; }

int32_t idp_panel_kmer_postings(const idp_panel *panel, int which, const char *kmer, int32_t *out, int32_t cap) {
    if (!panel || !kmer || which < 0 || which > 1) return -1;
    const KmerIndexHost &ix = panel->kidx[which];
    auto it = ix.by_string.find(kmer);
    if (it == ix.by_string.end()) return 0;
    for (uint32_t j = 0; j < it->second.second && (int32_t)j < cap; ++j) out[j] = (int32_t)ix.postings[it->second.first + j];
    return (int32_t)it->second.second;
}

// ---- pure host helpers ----
double idp_result_score(const idp_result *r) {
    if (!r || r->type != IDP_GAPPED) return 0.0;
    return r->multi ? r->raw_score / (double)r->q_len : (double)r->raw_score;  // PM:321 vs PM:315
}
double idp_result_second(const idp_result *r) {
    if (!r || r->type != IDP_GAPPED || !r->multi) return 0.0;
    return r->raw_next / (double)r->q_len_next;                                // PM:322
}
int idp_metric_bin(int template_type, int is_fr, int r1_type, int r2_type) {
    return ((template_type * 2 + (is_fr ? 1 : 0)) * 4 + r1_type) * 4 + r2_type;
}
void idp_summary_metrics(const int64_t *m, int64_t *o) {  // IdentifyPrimersMetric.scala:34-89
    int64_t by_template[5] = {0}, by_type[4] = {0}, matched_ends[3] = {0}, fr = 0;
    for (int bin = 0; bin < 160; ++bin) {
        int64_t c = m[bin];
        int r2 = bin & 3, r1 = (bin >> 2) & 3, is_fr = (bin >> 4) & 1, tt = bin >> 5;
        by_template[tt] += c;
        by_type[r1] += c;
        by_type[r2] += c;
        matched_ends[(r1 != IDP_NONE) + (r2 != IDP_NONE)] += c;
        if (is_fr) fr += c;
    }
    int64_t pairs = by_template[0] + by_template[1] + by_template[2], frags = by_template[3] + by_template[4];
    int64_t vals[17] = {pairs + frags, pairs, frags, by_template[0], by_template[1], by_template[2], fr, by_template[3], by_template[4],
                        matched_ends[2], matched_ends[1], matched_ends[0], by_type[0] + by_type[1] + by_type[2] + by_type[3],
                        by_type[IDP_LOCATION], by_type[IDP_UNGAPPED], by_type[IDP_GAPPED], by_type[IDP_NONE]};
    memcpy(o, vals, sizeof vals);
}

// scripts/post_process_bam.py:56-78
int idp_classify_primer_pair(const idp_panel *p, int32_t f5, int32_t r5, int32_t min_insert_length, int32_t max_insert_length) {
    if (!p || f5 >= p->n || r5 >= p->n) return -1;
    if (f5 < 0 && r5 < 0) return IDP_CLASS_NO_MATCH;
    if (f5 < 0 || r5 < 0) return IDP_CLASS_SINGLE;
    if (p->pair_id[f5] != p->pair_id[r5]) {  // from different primer pairs
        if (p->ref_name[f5] != p->ref_name[r5]) return IDP_CLASS_CROSS_DIMER;
        const int32_t insert_length = std::max(p->end[f5], p->end[r5]) - std::min(p->start[f5], p->start[r5]) + 1;  // :47-53
        return (insert_length < min_insert_length || max_insert_length < insert_length) ? IDP_CLASS_CROSS_DIMER : IDP_CLASS_NON_CANONICAL;
    }
    if (p->primer_id[f5] == p->primer_id[r5]) return IDP_CLASS_SELF_DIMER;
    if (p->positive[f5] == p->positive[r5]) return IDP_CLASS_NON_CANONICAL;
    return IDP_CLASS_CANONICAL;
}

int idp_classify_templates(const idp_panel *p, const uint16_t *flags, const idp_result *five, int32_t n_reads, int32_t min_insert_length,
                           int32_t max_insert_length, uint8_t *cls_per_read, int64_t *counts6) {
    if (!p || !flags || !five || n_reads < 0) { set_error("idp_classify_templates: NULL argument"); return IDP_ERR_ARG; }
    for (int32_t r = 0; r < n_reads;) {
        const bool has_r2 = (flags[r] & IDP_F_MATE_NEXT) && r + 1 < n_reads;
        const int32_t m1 = five[r].type != IDP_NONE ? five[r].primer : -1;
        const int32_t m2 = has_r2 && five[r + 1].type != IDP_NONE ? five[r + 1].primer : -1;
        int32_t f5 = m1, r5 = m2;
        if (has_r2) {  // forward / reverse slots (IP:530-540)
            bool fwd_is_r1 = true;
            if (m1 >= 0) fwd_is_r1 = p->positive[m1] != 0;
            else if (m2 >= 0) fwd_is_r1 = p->positive[m2] == 0;
            if (!fwd_is_r1) { f5 = m2; r5 = m1; }
        }
        const int cls = idp_classify_primer_pair(p, f5, r5, min_insert_length, max_insert_length);
        if (cls < 0) { set_error("idp_classify_templates: primer index out of range at read %d", r); return IDP_ERR_ARG; }
        if (cls_per_read) { cls_per_read[r] = (uint8_t)cls; if (has_r2) cls_per_read[r + 1] = (uint8_t)cls; }
        if (counts6 && (flags[r] & IDP_F_FIRST)) counts6[cls] += 1;  // "do not double-count": only record.is_read1
        r += has_r2 ? 2 : 1;
    }
    return IDP_OK;
}

void idp_pack_bases(const char *ascii, int32_t n, uint8_t *seq4) {
    for (int32_t i = 0; i < n; i += 2) {
        int hi = base_to_nibble((unsigned char)ascii[i]);
        int lo = i + 1 < n ? base_to_nibble((unsigned char)ascii[i + 1]) : 0;
        seq4[i / 2] = (uint8_t)((hi << 4) | lo);
    }
}
void idp_unpack_bases(const uint8_t *seq4, int32_t n, char *ascii) {
    for (int32_t i = 0; i < n; ++i) ascii[i] = nibble_to_base((seq4[i / 2] >> ((i & 1) ? 0 : 4)) & 15);
}

// Java's f"$x%.2f": shortest round-tripping decimal digits (FloatingDecimal) rounded HALF_UP (PMa:115)
static std::string java_format_2f(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "Infinity" : "-Infinity";
    char buf[64];
    double a = std::fabs(v);
    int prec = 1;
    for (; prec <= 17; ++prec) { snprintf(buf, sizeof buf, "%.*e", prec - 1, a); if (strtod(buf, nullptr) == a) break; }
    std::string digits; int exp10 = 0;
    for (const char *c = buf; *c; ++c) { if (*c == 'e') { exp10 = atoi(c + 1); break; } if (*c >= '0' && *c <= '9') digits.push_back(*c); }
    // a = 0.d1d2d3... * 10^(exp10+1); scale to hundredths as a decimal string and round half up on the next digit
    int point = exp10 + 1;
    std::string ip, fp;
    for (int i = 0; i < point; ++i) ip.push_back(i < (int)digits.size() ? digits[i] : '0');
    for (int i = point; i < std::max((int)digits.size(), point + 3); ++i) fp.push_back(i < 0 ? '0' : (i < (int)digits.size() ? digits[i] : '0'));
    if (ip.empty()) ip = "0";
    while (fp.size() < 3) fp.push_back('0');
    bool carry = fp[2] >= '5';
    std::string num = ip + fp.substr(0, 2);
    for (int i = (int)num.size() - 1; i >= 0 && carry; --i) { if (num[i] == '9') num[i] = '0'; else { ++num[i]; carry = false; } }
    if (carry) num.insert(num.begin(), '1');
    std::string outp = num.substr(0, num.size() - 2) + "." + num.substr(num.size() - 2);
    return (std::signbit(v) ? "-" : "") + outp;
}

int idp_format_info(const idp_panel *panel, const idp_result *r, uint16_t fl, char *buf, int32_t cap) {
    if (!panel || !r || !buf || cap <= 0) return -1;
    std::string s;
    if (r->type == IDP_NONE || r->primer < 0 || r->primer >= panel->n) s = "none";  // IP:242
    else {
        int i = r->primer;
        int read_num = (!(fl & IDP_F_PAIRED) || (fl & IDP_F_FIRST)) ? 1 : 2;  // PMa:84
        s = panel->pair_id[i] + "," + panel->primer_id[i] + "," + panel->ref_name[i] + ":" + std::to_string(panel->start[i]) + "-" +
            std::to_string(panel->end[i]) + "," + (panel->positive[i] ? "+" : "-") + "," + std::to_string(read_num) + ",";
        if (r->type == IDP_LOCATION) s += "Location," + std::to_string(r->mm);
        else if (r->type == IDP_UNGAPPED) s += "Ungapped," + std::to_string(r->mm) + "," + (r->next_mm == IDP_NEXT_NA ? std::string("na") : std::to_string(r->next_mm));
        else s += "Gapped," + java_format_2f(idp_result_score(r)) + "," + java_format_2f(idp_result_second(r)) + "," + std::to_string(r->t_start) + "," + std::to_string(r->t_end);
    }
    snprintf(buf, (size_t)cap, "%s", s.c_str());
    return (int)s.size();
}

}  // extern "C"
