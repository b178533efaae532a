// C++ host-mirror test, written the way the reference's own ScalaTest suites read (PrimerTest, PrimerMatcherTest,
// PrimerMatchTest, IdentifyPrimersTest).  "host" mode needs no GPU; "gpu" mode runs the end-to-end known answers.
#include <cstring>
#include <algorithm>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "idprimer.hpp"

using namespace idprimer;

#define EXPECT(cond) do { if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } } while (0)
template <class F> static bool throws(F f) { try { f(); } catch (const std::exception &) { return true; } return false; }

// PrimerMatcherTest.scala:34-48 / IdentifyPrimersTest.scala:73-87
static std::vector<Primer> panel(bool mapped = true) {
    auto P = [&](const char *pair, const char *id, const char *seq, bool pos, const char *ref, int s, int e) { return Primer(pair, id, seq, pos, mapped ? ref : "", s, e); };
    return {P("1", "1+", "AAAAAAAAAA", true, "chr1", 1, 10),    P("1", "1-", "TTTTTTTTTT", false, "chr1", 101, 110),
            P("2", "2+", "TTTTTTTTTT", true, "chr2", 1, 10),    P("2", "2-", "AAAAAAAAAA", false, "chr2", 101, 110),
            P("3", "3+", "TTTTTATTTT", true, "chr2", 1001, 1010), P("3", "3-", "AAAAATAAAA", false, "chr2", 1001, 1010),
            P("4", "4+", "GGGGGGGGGG", true, "chr3", 2001, 2010), P("4", "4-", "CCCCCCCCCC", false, "chr3", 2001, 2010),
            P("5", "5+", "GGGGAAAGGG", true, "chr3", 3001, 3010), P("5", "5-", "CCCCTTTCCC", false, "chr3", 3001, 3010),
            P("6", "6+", "GATTACAGAT", true, "chr4", 1001, 1010), P("6", "6-", "GATTACAGAT", false, "chr4", 1001, 1010)};
}
static std::vector<std::string> dict() {
    std::vector<std::string> d;
    for (int i = 1; i <= 22; ++i) d.push_back("chr" + std::to_string(i));
    d.push_back("chrX"); d.push_back("chrY"); d.push_back("chrM");
    return d;
}
static std::string revcomp(const std::string &s) {  // htsjdk 2.16: only ACGT are complemented
    std::string o(s.rbegin(), s.rend());
    for (char &c : o) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
    return o;
}

static void host_tests() {
    // PrimerTest: validation
    EXPECT(throws([] { Primer("1,2", "x", "ACGT", true); }));
    EXPECT(throws([] { Primer("1", "x,y", "ACGT", true); }));
    EXPECT(throws([] { Primer("1", "x", "ACGU", true); }));
    EXPECT(Primer("1", "1+", "GATTACA", true).length() == 7);
    EXPECT(Primer("1", "1+", "GATTACA", true, "chr1", 1, 10).length() == 10);   // Primer.scala:206
    // PrimerMatch requires
    Primer p("pair1", "primer1+", "GATTACA", true, "chr1", 1, 7);
    EXPECT(throws([&] { UngappedAlignmentPrimerMatch(&p, 3, 2); }));
    EXPECT(throws([&] { GappedAlignmentPrimerMatch(&p, 1.0, 2.0, 0, 0); }));
    EXPECT(throws([&] { GappedAlignmentPrimerMatch(&p, 2.0, 1.0, 3, 2); }));
    EXPECT(!throws([&] { GappedAlignmentPrimerMatch(&p, 3, 2, 0, 0); }));
    // PrimerMatcherWithKmerFilterTest (kmerLength 5): "AAAAA" is in 1+, 2-, 3-; "GATTA" in 6+, 6-
    IdentifyPrimers::Options o;
    o.maxMismatchRate = 0.1; o.ungappedKmerLength = 5;
    IdentifyPrimers tool(panel(), o, dict(), /*device=*/-1);
    EXPECT(tool.kmerPostings(0, "AAAAA").size() == 3);
    EXPECT(tool.kmerPostings(0, "GATTA").size() == 2);
    EXPECT(tool.kmerPostings(0, "ACGTA").empty());
    // PrimerMatchTest: info()
    idp_result r{};
    r.primer = 0; r.type = IDP_UNGAPPED; r.mm = 2; r.next_mm = IDP_NEXT_NA;
    EXPECT(tool.info(r, 0) == "1,1+,chr1:1-10,+,1,Ungapped,2,na");
    r.type = IDP_GAPPED; r.multi = 1; r.raw_score = 1; r.q_len = 8; r.raw_next = -3; r.q_len_next = 24; r.t_start = 1; r.t_end = 9;
    EXPECT(tool.info(r, IDP_F_PAIRED | IDP_F_SECOND) == "1,1+,chr1:1-10,+,2,Gapped,0.13,-0.13,1,9");
    r.type = IDP_NONE;
    EXPECT(tool.info(r, 0) == "none");
    EXPECT(throws([&] { tool.processBatch(ReadBatch()); }));  // host-only: matching must fail loudly
    // windowed ingest: 10-base primers (Primer.length 10), slop 5 -> the 5' chain reaches 15 bases of a read
    EXPECT(idp_panel_window(tool.panel()) == 15);
    {
        // one unmapped 24-base record, built by hand (SAMv1 4.2): block_size, then the fixed fields, name, no cigar, seq, qual
        const std::string bases = "ACGTACGTACGTACGTTTTTGGGG", name = "q";
        std::vector<uint8_t> rec(4 + 32 + name.size() + 1 + (bases.size() + 1) / 2 + bases.size(), 0);
        auto put32 = [&](size_t at, int32_t v) { std::memcpy(rec.data() + at, &v, 4); };
        put32(0, (int32_t)rec.size() - 4); put32(4, -1); put32(8, -1);
        rec[12] = (uint8_t)(name.size() + 1);
        const uint16_t flag = 0x4; std::memcpy(rec.data() + 18, &flag, 2);
        put32(20, (int32_t)bases.size()); put32(24, -1); put32(28, -1);
        std::memcpy(rec.data() + 36, name.c_str(), name.size() + 1);
        idp_pack_bases(bases.data(), (int32_t)bases.size(), rec.data() + 36 + name.size() + 1);
        const ReadBatch whole = ReadBatch::fromBam(rec.data(), rec.size()), cut = ReadBatch::fromBamWindow(tool.panel(), rec.data(), rec.size());
        EXPECT(whole.len[0] == 24 && cut.len[0] == 15 && cut.seq4.size() == 8 && cut.flags[0] == whole.flags[0]);
        char back[16] = {0};
        idp_unpack_bases(cut.seq4.data(), 15, back);
        EXPECT(std::string(back, 15) == bases.substr(0, 15));
    }
    std::puts("OK host");
}

// IdentifyPrimersTest.scala:100-235 read factory
struct Frag { std::string bases; bool unmapped, negative; int ref, start; int end() const { return start + (int)bases.size() - 1; } };
static Frag fromPrimer(const Primer &p, char edit = 0, int n = 0) {
    Primer q = p;
    if (edit) q.ref_name = "chrM";
    const std::string &s = p.sequence;
    const int len = (int)s.size();
    if (edit == 'm') {
        const int rem = len - n;
        if (rem <= 0) q.sequence = s;
        else if (rem == 1) q.sequence = p.positive_strand ? "N" + s.substr(1) : s.substr(1) + "N";
        else q.sequence = s.substr(0, (rem + 1) / 2) + std::string(n, 'N') + s.substr(len - rem / 2);
    } else if (edit == 'i') { q.sequence = s.substr(0, (len + 1) / 2) + std::string(n, 'N') + s.substr(len - len / 2); q.end += n; }
    else if (edit == 'd') { const int rem = len - n; q.sequence = s.substr(0, (rem + 1) / 2) + s.substr(len - rem / 2); q.start += n; }
    const auto d = dict();
    const int ref = (int)(std::find(d.begin(), d.end(), q.ref_name) - d.begin());
    return Frag{q.positive_strand ? q.sequence : revcomp(q.sequence), false, !q.positive_strand, ref, q.start};
}

static void gpu_tests() {
    const auto primers = panel();
    std::vector<Frag> frags;
    for (auto &p : primers) frags.push_back(fromPrimer(p));
    for (int n = 0; n <= 4; ++n) for (auto &p : primers) frags.push_back(fromPrimer(p, 'm', n));
    for (int L = 1; L <= 2; ++L) { for (auto &p : primers) frags.push_back(fromPrimer(p, 'd', L)); for (auto &p : primers) frags.push_back(fromPrimer(p, 'i', L)); }
    frags.push_back(Frag{std::string(10, 'G'), false, false, 0, 100000});
    frags.push_back(Frag{std::string(10, 'C'), false, true, 0, 100000});
    frags.push_back(Frag{std::string(10, 'G'), true, false, 0, 100000});
    frags.push_back(Frag{std::string(10, 'C'), true, true, 0, 100000});
    frags.push_back(fromPrimer(primers[0]));
    frags.push_back(fromPrimer(primers[3]));
    EXPECT(frags.size() == 126);
    ReadBatch batch;
    for (size_t i = 0; i < frags.size(); i += 2) {  // pairRecords (IdentifyPrimersTest.scala:137-151) + SamPairUtil.setMateInfo
        const Frag &a = frags[i], &b = frags[i + 1];
        bool fr = false;
        if (!a.unmapped && !b.unmapped && a.ref == b.ref && a.negative != b.negative) {
            const int p1 = a.negative ? a.end() : a.start, p2 = b.negative ? b.end() : b.start;
            const int isize = p2 - p1 + (p2 >= p1 ? 1 : -1);
            const long pos5 = a.negative ? b.start : a.start, neg5 = a.negative ? a.end() : (long)a.start + isize;
            fr = pos5 < neg5;
        }
        auto flags = [&](const Frag &f, bool first) {
            return (uint16_t)(IDP_F_PAIRED | (f.unmapped ? IDP_F_UNMAPPED : 0) | (f.negative ? IDP_F_REVERSE : 0) | (first ? IDP_F_FIRST | IDP_F_MATE_NEXT : IDP_F_SECOND) | (fr ? IDP_F_FR_PAIR : 0));
        };
        batch.add(a.bases, flags(a, true), a.unmapped ? -1 : a.ref, a.start, a.end());
        batch.add(b.bases, flags(b, false), b.unmapped ? -1 : b.ref, b.start, b.end());
    }
    IdentifyPrimers::Options o;  // IdentifyPrimersTest.scala:331-343
    o.slop = 1; o.maxMismatchRate = 3.0 / 11; o.minAlignmentScoreRate = 0; o.matchScore = 1; o.mismatchScore = -3; o.gapOpen = -6; o.gapExtend = -1;
    IdentifyPrimers tool(primers, o, dict(), 0);
    const auto res = tool.processBatch(batch);
    const auto s = tool.summary();
    // templates pairs fragments mapped_pairs unpaired unmapped_pairs fr_pairs mapped_fragments unmapped_fragments paired_matches
    // unpaired_matches no_paired_matches match_attempts location ungapped gapped no_match   (IdentifyPrimersTest.scala:361-397)
    const std::array<int64_t, 17> expected{63, 63, 0, 62, 0, 1, 61, 0, 0, 48, 0, 15, 126, 14, 78, 4, 30};
    for (int i = 0; i < 17; ++i) if (s[i] != expected[i]) { std::fprintf(stderr, "metric %d: got %lld expected %lld\n", i, (long long)s[i], (long long)expected[i]); std::exit(1); }
    // every read decodes into the reference's case classes without tripping a require
    int gapped = 0;
    for (auto &r : res.fivePrime) { auto m = tool.toMatch(r); if (m && std::holds_alternative<GappedAlignmentPrimerMatch>(*m)) ++gapped; }
    EXPECT(gapped == 4);
    const auto tags = tool.pairTags(batch.flags[0], batch.flags[1], res.fivePrime[0], res.fivePrime[1]);
    EXPECT(tags.at("f5") == "1,1+,chr1:1-10,+,1,Location,0" && tags.at("r5") == "1,1-,chr1:101-110,-,2,Location,0");
    // AsyncAlignerTest-style: the batched aligner agrees with the matcher's own numbers on an exact hit
    AsyncCudaAligner al(1, -3, -6, -1, IDP_MODE_GLOCAL, 0);
    const auto a = al.align({"GATTACA"}, {"TTTGATTACATTT"});
    EXPECT(a[0].score == 7 && a[0].targetStart == 4 && a[0].targetEnd == 10);
    std::puts("OK gpu");
}

int main(int argc, char **argv) {
    const std::string mode = argc > 1 ? argv[1] : "host";
    if (mode == "host") host_tests(); else gpu_tests();
    return 0;
}
