// idprimer.hpp -- header-only C++17 host mirror of the reference's interface for the matching path, over the C-ABI
// of idprimer.h.  The reference is compiled (Scala/JVM) code and no JVM exists in the build image, so this is the
// compiled-language host side: same names, argument meaning and error behaviour as the Scala it stands in for.
//
//   idprimer::Primer                       Primer.scala:182-221 (validation :189-201, length :206)
//   idprimer::LocationBasedPrimerMatch ... PrimerMatch.scala:96-117 (requires :102, :113-114 throw std::invalid_argument)
//   idprimer::IdentifyPrimers              IdentifyPrimers.scala:165-186 (options), processBatch :381-436,
//                                          addTagsToPair :522-554, addTagsToFragment :557-569
//   idprimer::AsyncCudaAligner             AsyncAligner.scala:37-54 (batched)
// Failures of the library surface as std::runtime_error(idp_last_error()), as the JNI shim would throw RuntimeException.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

#include "idprimer.h"

namespace idprimer {

inline void check(int rc) {
    if (rc != IDP_OK) throw std::runtime_error(std::string("idprimer error ") + std::to_string(rc) + ": " + idp_last_error());
}

struct Primer {  // case class Primer(pair_id, primer_id, sequence, positive_strand, ref_name = "", start = 0, end = 0)
    std::string pair_id, primer_id, sequence;
    bool positive_strand = true;
    std::string ref_name;
    int32_t start = 0, end = 0;

    Primer(std::string pair, std::string id, std::string seq, bool positive, std::string ref = "", int32_t s = 0, int32_t e = 0)
        : pair_id(std::move(pair)), primer_id(std::move(id)), sequence(std::move(seq)), positive_strand(positive), ref_name(std::move(ref)), start(s), end(e) {
        if (pair_id.find(',') != std::string::npos) throw std::invalid_argument("requirement failed: pair_id contained a comma: '" + pair_id + "'");
        if (primer_id.find(',') != std::string::npos) throw std::invalid_argument("requirement failed: primer_id contained a comma: '" + pair_id + "'");
        static const std::string ok = "ACGTNacgtn.aAbBcCdDgGhHkKmMnNrRsStTvVwWyY";  // htsjdk isValidBase || isIUPAC
        for (size_t i = 0; i < sequence.size(); ++i)
            if (ok.find(sequence[i]) == std::string::npos)
                throw std::invalid_argument("Found an invalid base for primer " + pair_id + "/" + primer_id + ": " + sequence.substr(0, i) + "[" + sequence[i] + "]" + sequence.substr(i + 1));
    }
    int32_t length() const { return ref_name.empty() ? (int32_t)sequence.size() : end - start + 1; }  // Primer.scala:206
    bool mapped() const { return !ref_name.empty(); }
    bool negativeStrand() const { return !positive_strand; }
};

struct LocationBasedPrimerMatch { const Primer *primer; int32_t numMismatches; };
struct UngappedAlignmentPrimerMatch {
    const Primer *primer; int32_t numMismatches, nextNumMismatches;
    UngappedAlignmentPrimerMatch(const Primer *p, int32_t mm, int32_t next) : primer(p), numMismatches(mm), nextNumMismatches(next) {
        if (!(numMismatches <= nextNumMismatches)) throw std::invalid_argument("requirement failed");  // PrimerMatch.scala:102
    }
};
struct GappedAlignmentPrimerMatch {
    const Primer *primer; double score, secondBestScore; int32_t start, end;
    GappedAlignmentPrimerMatch(const Primer *p, double s, double second, int32_t st, int32_t en) : primer(p), score(s), secondBestScore(second), start(st), end(en) {
        if (!(score >= secondBestScore)) throw std::invalid_argument("requirement failed: second best score must be less than or equal to the score");  // :113
        if (!(start <= end)) throw std::invalid_argument("requirement failed");                                                                       // :114
    }
};
using PrimerMatch = std::variant<LocationBasedPrimerMatch, UngappedAlignmentPrimerMatch, GappedAlignmentPrimerMatch>;

struct ReadBatch {  // structure of arrays in the layout of idp_reads; bases stay BAM 4-bit
    std::vector<uint8_t> seq4;
    std::vector<uint64_t> seq_off;
    std::vector<int32_t> len, ref_idx, unclipped_start, unclipped_end;
    std::vector<uint16_t> flags;

    size_t size() const { return flags.size(); }
    // one record: ASCII bases as stored in the BAM record, SAM-style flags (IDP_F_*), mapping fields when mapped
    void add(const std::string &bases, uint16_t f, int32_t ref = -1, int32_t ustart = 0, int32_t uend = 0) {
        seq_off.push_back(seq4.size());
        seq4.resize(seq4.size() + (bases.size() + 1) / 2);
        idp_pack_bases(bases.data(), (int32_t)bases.size(), seq4.data() + seq_off.back());
        len.push_back((int32_t)bases.size()); ref_idx.push_back(ref); unclipped_start.push_back(ustart); unclipped_end.push_back(uend); flags.push_back(f);
    }
    static ReadBatch fromBam(const uint8_t *bam, size_t n_bytes, std::vector<int32_t> *rec_index = nullptr) {
        ReadBatch b; int32_t n = 0; uint64_t bytes = 0;
        check(idp_bam_scan(bam, n_bytes, &n, &bytes));
        b.seq4.resize(bytes); b.seq_off.resize(n); b.len.resize(n); b.ref_idx.resize(n); b.unclipped_start.resize(n); b.unclipped_end.resize(n); b.flags.resize(n);
        if (rec_index) rec_index->resize(n);
        check(idp_bam_to_batch(bam, n_bytes, b.seq4.data(), b.seq_off.data(), b.len.data(), b.ref_idx.data(), b.unclipped_start.data(), b.unclipped_end.data(),
                               b.flags.data(), rec_index ? rec_index->data() : nullptr));
        return b;
    }
    // the same for a run without --three-prime: only the bases the 5' chain can reach are kept (idp_bam_to_batch_window)
    static ReadBatch fromBamWindow(const idp_panel *panel, const uint8_t *bam, size_t n_bytes, std::vector<int32_t> *rec_index = nullptr) {
        ReadBatch b; int32_t n = 0; uint64_t bytes = 0;
        check(idp_bam_scan(bam, n_bytes, &n, &bytes));
        b.seq4.resize(bytes); b.seq_off.resize(n); b.len.resize(n); b.ref_idx.resize(n); b.unclipped_start.resize(n); b.unclipped_end.resize(n); b.flags.resize(n);
        if (rec_index) rec_index->resize(n);
        check(idp_bam_to_batch_window(panel, bam, n_bytes, b.seq4.data(), b.seq_off.data(), b.len.data(), b.ref_idx.data(), b.unclipped_start.data(),
                                      b.unclipped_end.data(), b.flags.data(), rec_index ? rec_index->data() : nullptr));
        if (n > 0) b.seq4.resize(b.seq_off.back() + (size_t)((b.len.back() + 1) / 2));
        return b;
    }
    idp_reads view() const { return idp_reads{seq4.data(), seq_off.data(), len.data(), ref_idx.data(), unclipped_start.data(), unclipped_end.data(), flags.data(), 0}; }
};

class IdentifyPrimers {
  public:
    struct Options {  // IdentifyPrimers.scala:170-185
        int32_t slop = 5, threePrime = -1;
        double maxMismatchRate = 0.05, minAlignmentScoreRate = 0.01;
        int32_t matchScore = 1, mismatchScore = -4, gapOpen = -6, gapExtend = -1;
        std::optional<int32_t> ungappedKmerLength, gappedKmerLength;
        std::optional<std::vector<int32_t>> scoringMatrix;  // 128*128, [primer_base*128 + read_base], INT32_MIN = no entry
    };

    // device < 0: host-only (panel queries, formatting), no matching possible
    IdentifyPrimers(std::vector<Primer> primers, const Options &o, const std::vector<std::string> &ref_names = {}, int device = 0)
        : primers_(std::move(primers)), threePrime_(o.threePrime) {
        std::vector<idp_primer> raw(primers_.size());
        for (size_t i = 0; i < primers_.size(); ++i) {
            const Primer &p = primers_[i];
            int32_t ref = -1;
            for (size_t r = 0; r < ref_names.size(); ++r) if (ref_names[r] == p.ref_name) ref = (int32_t)r;
            raw[i] = idp_primer{p.pair_id.c_str(), p.primer_id.c_str(), p.sequence.c_str(), p.ref_name.c_str(), ref, p.start, p.end, (uint8_t)(p.positive_strand ? 1 : 0)};
        }
        if (o.scoringMatrix) matrix_ = *o.scoringMatrix;
        idp_params prm{o.slop, o.threePrime, o.matchScore, o.mismatchScore, o.gapOpen, o.gapExtend, o.ungappedKmerLength.value_or(-1),
                       o.gappedKmerLength.value_or(-1), o.maxMismatchRate, o.minAlignmentScoreRate, matrix_.empty() ? nullptr : matrix_.data()};
        check(idp_panel_create(raw.data(), (int32_t)raw.size(), &prm, &panel_));
        if (device >= 0) {
            const int rc = idp_ctx_create(panel_, device, &ctx_);
            if (rc != IDP_OK) { idp_panel_destroy(panel_); panel_ = nullptr; check(rc); }
        }
        metricCounter.fill(0);
    }
    ~IdentifyPrimers() { if (ctx_) idp_ctx_destroy(ctx_); if (panel_) idp_panel_destroy(panel_); }
    IdentifyPrimers(const IdentifyPrimers &) = delete;
    IdentifyPrimers &operator=(const IdentifyPrimers &) = delete;

    struct BatchResult { std::vector<idp_result> fivePrime, threePrime; };

    // processBatch (IdentifyPrimers.scala:381-436): matches every read, adds to metricCounter and numAlignments
    BatchResult processBatch(const ReadBatch &reads) {
        if (!ctx_) throw std::runtime_error("IdentifyPrimers was created host-only; matching needs a GPU context");
        BatchResult out;
        out.fivePrime.resize(reads.size());
        if (threePrime_ >= 0) out.threePrime.resize(reads.size());
        const idp_reads view = reads.view();
        int64_t aligned = 0;
        check(idp_match_batch(ctx_, &view, (int32_t)reads.size(), out.fivePrime.data(), threePrime_ >= 0 ? out.threePrime.data() : nullptr, metricCounter.data(), &aligned));
        numAlignments += aligned;
        return out;
    }

    // idp_result -> the PrimerMatch the reference would hold; a non-zero status is the require() that would have thrown
    std::optional<PrimerMatch> toMatch(const idp_result &r) const {
        if (r.status == IDP_STATUS_REQ_MM_LE_NEXT || r.status == IDP_STATUS_REQ_SCORE_GE_SECOND || r.status == IDP_STATUS_REQ_START_LE_END)
            throw std::invalid_argument("requirement failed");
        if (r.status == IDP_STATUS_SCORE_MISSING) throw std::out_of_range("key not found");  // NoSuchElementException, IdentifyPrimers.scala:277
        switch (r.type) {
            case IDP_LOCATION: return PrimerMatch{LocationBasedPrimerMatch{&primers_[r.primer], r.mm}};
            case IDP_UNGAPPED: return PrimerMatch{UngappedAlignmentPrimerMatch{&primers_[r.primer], r.mm, r.next_mm}};
            case IDP_GAPPED: return PrimerMatch{GappedAlignmentPrimerMatch{&primers_[r.primer], idp_result_score(&r), idp_result_second(&r), r.t_start, r.t_end}};
            default: return std::nullopt;
        }
    }
    std::string info(const idp_result &r, uint16_t read_flags) const {  // PrimerMatch.info / "none"
        char buf[1024];
        idp_format_info(panel_, &r, read_flags, buf, sizeof buf);
        return buf;
    }
    // addTagsToPair (IdentifyPrimers.scala:522-554): the tags BOTH reads of the pair receive
    std::map<std::string, std::string> pairTags(uint16_t f1, uint16_t f2, const idp_result &r1Five, const idp_result &r2Five,
                                                const idp_result *r1Three = nullptr, const idp_result *r2Three = nullptr) const {
        bool fwdIsR1 = true;
        if (r1Five.type != IDP_NONE) fwdIsR1 = primers_[r1Five.primer].positive_strand;
        else if (r2Five.type != IDP_NONE) fwdIsR1 = primers_[r2Five.primer].negativeStrand();
        std::map<std::string, std::string> tags;
        tags["f5"] = fwdIsR1 ? info(r1Five, f1) : info(r2Five, f2);
        tags["r5"] = fwdIsR1 ? info(r2Five, f2) : info(r1Five, f1);
        if (threePrime_ >= 0 && r1Three && r2Three) {
            tags["f3"] = fwdIsR1 ? info(*r1Three, f1) : info(*r2Three, f2);
            tags["r3"] = fwdIsR1 ? info(*r2Three, f2) : info(*r1Three, f1);
        }
        return tags;
    }
    // addTagsToFragment (IdentifyPrimers.scala:557-569), including the overwrite of f5 by the 3' match
    std::map<std::string, std::string> fragmentTags(uint16_t f, const idp_result &five, const idp_result *three = nullptr) const {
        std::map<std::string, std::string> tags{{"f5", info(five, f)}, {"r5", "none"}};
        if (threePrime_ >= 0 && three) tags["f5"] = info(*three, f);
        return tags;
    }
    // the records ReadBatch::fromBam was made from, with the tags written back (idp_bam_annotate)
    std::vector<uint8_t> annotateBam(const uint8_t *bam, size_t n_bytes, const std::vector<idp_result> &five,
                                     const std::vector<idp_result> *three = nullptr) const {
        size_t need = 0;
        const idp_result *t = three ? three->data() : nullptr;
        check(idp_bam_annotate(panel_, bam, n_bytes, five.data(), t, (int32_t)five.size(), nullptr, 0, &need));
        std::vector<uint8_t> out(need);
        check(idp_bam_annotate(panel_, bam, n_bytes, five.data(), t, (int32_t)five.size(), out.data(), out.size(), &need));
        return out;
    }
    static std::string headerComments() {  // the @CO lines of IdentifyPrimers.scala:243-250
        char buf[2048];
        idp_header_comments(buf, sizeof buf);
        return buf;
    }
    std::array<int64_t, 17> summary() const {  // IdentifyPrimersMetric, field order of the case class
        std::array<int64_t, 17> out{};
        idp_summary_metrics(metricCounter.data(), out.data());
        return out;
    }
    std::vector<int32_t> kmerPostings(int which, const std::string &kmer) const {
        std::vector<int32_t> out(primers_.size() + 1);
        const int32_t n = idp_panel_kmer_postings(panel_, which, kmer.c_str(), out.data(), (int32_t)out.size());
        out.resize(n > 0 ? n : 0);
        return out;
    }
    const std::vector<Primer> &primers() const { return primers_; }
    idp_ctx *context() const { return ctx_; }
    const idp_panel *panel() const { return panel_; }

    std::array<int64_t, 160> metricCounter{};  // SimpleCounter[TemplateTypeMetric] (IdentifyPrimers.scala:285)
    int64_t numAlignments = 0;                  // IdentifyPrimers.scala:236

  private:
    std::vector<Primer> primers_;
    std::vector<int32_t> matrix_;
    int32_t threePrime_;
    idp_panel *panel_ = nullptr;
    idp_ctx *ctx_ = nullptr;
};

// AsyncAligner.align(query, target) for many pairs at once (AsyncAligner.scala:37-54); mode is IDP_MODE_*
class AsyncCudaAligner {
  public:
    AsyncCudaAligner(int matchScore, int mismatchScore, int gapOpen, int gapExtend, int mode = IDP_MODE_GLOCAL, int device = 0)
        : tool_({Primer("p", "p+", "A", true), Primer("p", "p-", "T", false)}, options(matchScore, mismatchScore, gapOpen, gapExtend), {}, device), mode_(mode) {}
    struct Alignment { int32_t score, targetStart, targetEnd, queryStart, queryEnd; };
    std::vector<Alignment> align(const std::vector<std::string> &queries, const std::vector<std::string> &targets) {
        const int32_t n = (int32_t)queries.size();
        std::string q, t;
        std::vector<int64_t> qo(n + 1, 0), to(n + 1, 0);
        for (int32_t i = 0; i < n; ++i) { q += queries[i]; t += targets[i]; qo[i + 1] = (int64_t)q.size(); to[i + 1] = (int64_t)t.size(); }
        std::vector<int32_t> s(n), ts(n), te(n), qs(n), qe(n);
        check(idp_align_batch(tool_.context(), mode_, q.data(), qo.data(), t.data(), to.data(), n, s.data(), ts.data(), te.data(), qs.data(), qe.data()));
        numAligned += n;
        std::vector<Alignment> out(n);
        for (int32_t i = 0; i < n; ++i) out[i] = Alignment{s[i], ts[i], te[i], qs[i], qe[i]};
        return out;
    }
    int64_t numAligned = 0;

  private:
    static IdentifyPrimers::Options options(int ma, int mi, int go, int ge) {
        IdentifyPrimers::Options o; o.matchScore = ma; o.mismatchScore = mi; o.gapOpen = go; o.gapExtend = ge; return o;
    }
    IdentifyPrimers tool_;
    int mode_;
};

}  // namespace idprimer
