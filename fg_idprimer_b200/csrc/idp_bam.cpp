// Host ingest / egress: uncompressed BAM alignment records -> the SoA batch of include/idprimer.h, and the same records
// with the primer-match tags added.
//
// The reference reaches its matchers through fgbio's Bams.templateIterator (IdentifyPrimers.scala:288-289), which yields the
// primary R1 / R2 of each queryname group; the matchers then look at rec.bases, rec.unmapped, rec.positiveStrand, rec.refIndex,
// rec.unclippedStart / unclippedEnd (PrimerMatcher.scala:41, 179-186), rec.paired / firstOfPair (PrimerMatch.scala:84) and
// rec.isFrPair (IdentifyPrimers.scala:413).  All of that is derivable from the record's own bytes (SAMv1 section 4.2), and the
// 4-bit `seq` field is already the kernel's operand, so it is copied, never decoded to ASCII.
#include <cstring>
#include <string>
#include <vector>

#include "idp_internal.h"

namespace {

struct BamRec {
    const uint8_t *p;      // start of the record body (after block_size)
    int32_t block_size;
    int32_t ref_id, pos, l_read_name, n_cigar, flag, l_seq, next_ref_id, next_pos, tlen;
    const char *name;
    const uint8_t *cigar, *seq;
};

inline int32_t rd32(const uint8_t *p) { int32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

bool parse(const uint8_t *bam, size_t n_bytes, size_t &off, BamRec &r) {
    if (off + 4 > n_bytes) return false;
    r.block_size = rd32(bam + off);
    if (r.block_size < 32 || off + 4 + (size_t)r.block_size > n_bytes) return false;
    const uint8_t *p = bam + off + 4;
    r.p = p;
    r.ref_id = rd32(p); r.pos = rd32(p + 4);
    r.l_read_name = p[8];
    r.n_cigar = rd16(p + 12); r.flag = rd16(p + 14);
    r.l_seq = rd32(p + 16); r.next_ref_id = rd32(p + 20); r.next_pos = rd32(p + 24); r.tlen = rd32(p + 28);
    if (r.l_seq < 0) return false;
    // name, CIGAR, packed bases AND the l_seq quality bytes must all lie inside the record (idp_bam_annotate copies up to the aux fields)
    const size_t fixed = 32 + (size_t)r.l_read_name + 4 * (size_t)r.n_cigar;
    if (fixed > (size_t)r.block_size || fixed + (size_t)((r.l_seq + 1) / 2) + (size_t)r.l_seq > (size_t)r.block_size) return false;
    r.name = reinterpret_cast<const char *>(p + 32);
    r.cigar = p + 32 + r.l_read_name;
    r.seq = r.cigar + 4 * (size_t)r.n_cigar;
    off += 4 + (size_t)r.block_size;
    return true;
}

// htsjdk SAMRecord.getAlignmentEnd / getUnclippedStart / getUnclippedEnd from the CIGAR (1-based, inclusive)
void span(const BamRec &r, int32_t &start, int32_t &end, int32_t &ustart, int32_t &uend) {
    start = r.pos + 1;
    int32_t ref_len = 0, lead = 0, trail = 0;
    bool seen_aligned = false;
    for (int i = 0; i < r.n_cigar; ++i) {
        const uint32_t c = (uint32_t)rd32(r.cigar + 4 * i);
        const int32_t len = (int32_t)(c >> 4), op = (int32_t)(c & 15);
        const bool clip = op == 4 || op == 5;                                      // S, H
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_len += len;   // M D N = X consume the reference
        if (clip) { if (!seen_aligned) lead += len; else trail += len; }
        else { seen_aligned = true; trail = 0; }
    }
    end = start + ref_len - 1;
    ustart = start - lead;
    uend = end + trail;
}

}  // namespace

extern "C" {

// Counts what idp_bam_to_batch will emit for these bytes: primary records (not secondary 0x100, not supplementary 0x800) and
// the bytes their packed bases need.  Returns IDP_ERR_ARG on a truncated or malformed record.
int idp_bam_scan(const uint8_t *bam, size_t n_bytes, int32_t *n_reads, uint64_t *seq_bytes) {
    if (!bam || !n_reads || !seq_bytes) { idp::set_error("idp_bam_scan: NULL argument"); return IDP_ERR_ARG; }
    size_t off = 0;
    int32_t n = 0;
    uint64_t bytes = 0;
    BamRec r;
    while (off < n_bytes) {
        if (!parse(bam, n_bytes, off, r)) { idp::set_error("idp_bam_scan: malformed BAM record at byte %zu", off); return IDP_ERR_ARG; }
        if (r.flag & 0x900) continue;
        if (r.l_seq > idp::kMaxReadLen) { idp::set_error("read longer than %d bases", idp::kMaxReadLen); return IDP_ERR_UNSUPPORTED; }
        ++n;
        bytes += (uint64_t)((r.l_seq + 1) / 2);
    }
    *n_reads = n;
    *seq_bytes = bytes;
    return IDP_OK;
}

// Fills the SoA arrays (sized by idp_bam_scan) from queryname-grouped records.  Reads come out template by template, R1 before
// R2 (Template.r1 / r2 are chosen by the first/second-of-pair flags, not by file order); rec_index[i] is the ordinal of the BAM
// record read i came from, so that the host can write the tags back.  An unmapped record reports ref_idx -1.
// window > 0 keeps only the `window` bases of every read that the 5' matchers can look at (idp_panel_window): the first ones
// of a read whose sequencing order is its stored order, the LAST ones of a mapped reverse-strand read (PM:39-45: its
// sequencing order is the reverse complement, so its 5' end is the end of the stored bases).  The stored orientation and the
// flags are kept, unclipped ends still come from the whole CIGAR.
static int bam_to_batch(const uint8_t *bam, size_t n_bytes, uint8_t *seq4, uint64_t *seq_off, int32_t *len, int32_t *ref_idx,
                        int32_t *unclipped_start, int32_t *unclipped_end, uint16_t *flags, int32_t *rec_index, int32_t window) {
    if (!bam || !seq4 || !seq_off || !len || !ref_idx || !unclipped_start || !unclipped_end || !flags) { idp::set_error("idp_bam_to_batch: NULL argument"); return IDP_ERR_ARG; }
    size_t off = 0;
    int32_t ordinal = 0, out = 0;
    uint64_t byte_off = 0;
    std::vector<std::pair<BamRec, int32_t>> group;  // primary records of the current template
    std::string group_name;
    auto emit = [&](const BamRec &r, int32_t ord, bool mate_next) {
        int32_t start, end, us, ue;
        span(r, start, end, us, ue);
        const bool unmapped = (r.flag & 0x4) != 0;
        uint16_t f = (uint16_t)(r.flag & (IDP_F_PAIRED | IDP_F_UNMAPPED | IDP_F_REVERSE | IDP_F_FIRST | IDP_F_SECOND));
        // fgbio SamRecord.isFrPair: paired, both ends mapped, same reference, htsjdk SamPairUtil.getPairOrientation == FR
        if ((r.flag & 0x1) && !unmapped && !(r.flag & 0x8) && r.ref_id == r.next_ref_id) {
            const bool rev = (r.flag & 0x10) != 0, mate_rev = (r.flag & 0x20) != 0;
            if (rev != mate_rev) {
                const int64_t pos5 = rev ? (int64_t)r.next_pos + 1 : start;
                const int64_t neg5 = rev ? end : (int64_t)start + r.tlen;
                if (pos5 < neg5) f |= IDP_F_FR_PAIR;
            }
        }
        if (mate_next) f |= IDP_F_MATE_NEXT;
        int32_t keep = r.l_seq;
        size_t nb = (size_t)((r.l_seq + 1) / 2);
        if (window > 0 && r.l_seq > window) {
            keep = window;
            nb = (size_t)((window + 1) / 2);
            const bool from_end = !unmapped && (r.flag & 0x10);
            const int32_t first = from_end ? r.l_seq - window : 0;  // first stored base kept
            uint8_t *dst = seq4 + byte_off;
            if ((first & 1) == 0) memcpy(dst, r.seq + first / 2, nb);
            else {  // odd start: every output byte straddles two stored bytes
                const uint8_t *src = r.seq + first / 2;
                const size_t src_bytes = (size_t)((r.l_seq + 1) / 2) - (size_t)(first / 2);
                for (size_t i = 0; i < nb; ++i) {
                    const uint8_t hi = (uint8_t)(src[i] << 4);
                    const uint8_t lo = i + 1 < src_bytes ? (uint8_t)(src[i + 1] >> 4) : (uint8_t)0;
                    dst[i] = (uint8_t)(hi | lo);
                }
            }
            if (window & 1) dst[nb - 1] &= 0xF0;  // the unused low nibble of the last byte is zero, as in a BAM record
        } else {
            memcpy(seq4 + byte_off, r.seq, nb);
        }
        seq_off[out] = byte_off; len[out] = keep;
        ref_idx[out] = unmapped ? -1 : r.ref_id;
        unclipped_start[out] = unmapped ? 0 : us; unclipped_end[out] = unmapped ? 0 : ue;
        flags[out] = f;
        if (rec_index) rec_index[out] = ord;
        byte_off += nb; ++out;
    };
    auto flush = [&]() -> int {
        if (group.empty()) return IDP_OK;
        const BamRec *r1 = nullptr, *r2 = nullptr; int32_t o1 = 0, o2 = 0;
        for (auto &g : group) {
            const bool paired = g.first.flag & 0x1;
            if (!paired || (g.first.flag & 0x40)) { if (r1) { idp::set_error("template %s has more than one primary R1", group_name.c_str()); return IDP_ERR_ARG; } r1 = &g.first; o1 = g.second; }
            else { if (r2) { idp::set_error("template %s has more than one primary R2", group_name.c_str()); return IDP_ERR_ARG; } r2 = &g.first; o2 = g.second; }
        }
        if (!r1) { idp::set_error("Template did not have an R1: %s", group_name.c_str()); return IDP_ERR_ARG; }  // IdentifyPrimers.scala:411
        emit(*r1, o1, r2 != nullptr);
        if (r2) emit(*r2, o2, false);
        group.clear();
        return IDP_OK;
    };
    BamRec r;
    while (off < n_bytes) {
        if (!parse(bam, n_bytes, off, r)) { idp::set_error("idp_bam_to_batch: malformed BAM record at byte %zu", off); return IDP_ERR_ARG; }
        const int32_t ord = ordinal++;
        if (r.flag & 0x900) continue;
        const std::string name(r.name, r.l_read_name > 0 ? strnlen(r.name, (size_t)r.l_read_name) : 0);
        if (!group.empty() && name != group_name) { int rc = flush(); if (rc) return rc; }
        group_name = name;
        group.push_back({r, ord});
    }
    return flush();
}

int idp_bam_to_batch(const uint8_t *bam, size_t n_bytes, uint8_t *seq4, uint64_t *seq_off, int32_t *len, int32_t *ref_idx,
                     int32_t *unclipped_start, int32_t *unclipped_end, uint16_t *flags, int32_t *rec_index) {
    return bam_to_batch(bam, n_bytes, seq4, seq_off, len, ref_idx, unclipped_start, unclipped_end, flags, rec_index, 0);
}

// The same with every read cut down to the bases the 5' matchers can reach (see idp_panel_window): results of
// idp_match_batch are identical, the batch is a fraction of the size (150 bp reads against 30 bp primers: 18 of 75 bytes).
int idp_bam_to_batch_window(const idp_panel *panel, const uint8_t *bam, size_t n_bytes, uint8_t *seq4, uint64_t *seq_off, int32_t *len,
                            int32_t *ref_idx, int32_t *unclipped_start, int32_t *unclipped_end, uint16_t *flags, int32_t *rec_index) {
    const int32_t w = idp_panel_window(panel);
    if (w <= 0) { idp::set_error("idp_bam_to_batch_window: the panel has --three-prime set (3' matching reads the whole read) or is NULL"); return IDP_ERR_ARG; }
    return bam_to_batch(bam, n_bytes, seq4, seq_off, len, ref_idx, unclipped_start, unclipped_end, flags, rec_index, w);
}

// The three @CO header lines the tool adds (IdentifyPrimers.scala:243-250), '\n'-separated; returns the length.
int idp_header_comments(char *buf, int32_t cap) {
    static const char *text =
        "The f5/r5/f3/r3 tags store the primer match metadata for the forward and reverse strand respectively.\n"
        "The f5/r5/f3/r3 tags are formatted as follow: <pair_id>,<primer_id>,<ref_name>:<start>-<end>,<strand>,<read-num>,<match-offset>,"
        "<match-length>,<match-type>,<match-type-info>.\n"
        "The match-type is 'location', 'gapped', or 'ungapped' based on if the match was found using the location, mismatch-based (ungapped) "
        "alignment, or gapped-alignment.";
    if (buf && cap > 0) snprintf(buf, (size_t)cap, "%s", text);
    return (int)strlen(text);
}

// Host egress: the input records with the primer-match tags set the way addTagsToPair / addTagsToFragment do
// (IdentifyPrimers.scala:522-569): Z-type aux fields f5 / r5 (and f3 / r3 when --three-prime >= 0) on the primary R1 / R2 of
// every template, including the fragment quirk (with --three-prime the 3' match overwrites f5 and no f3 / r3 is written).
// five / three are the results of idp_match_batch for the batch idp_bam_to_batch made from the same bytes (n_reads rows, R1
// then R2 per template).  Records come out template by template in Template.allReads order: R1, R2, then supplementary and
// secondary records (R1's before R2's), untouched.  A record that already carries one of the tags has it replaced.
// out == NULL sizes the output (*out_bytes); IDP_ERR_ARG when out_cap is too small or the batch does not match the bytes.
int idp_bam_annotate(const idp_panel *panel, const uint8_t *bam, size_t n_bytes, const idp_result *five, const idp_result *three,
                     int32_t n_reads, uint8_t *out, size_t out_cap, size_t *out_bytes) {
    if (!panel || !bam || !five || !out_bytes) { idp::set_error("idp_bam_annotate: NULL argument"); return IDP_ERR_ARG; }
    const bool want3 = panel->params.three_prime >= 0;
    if (want3 && !three) { idp::set_error("idp_bam_annotate: the panel has --three-prime set but no 3' results were passed"); return IDP_ERR_ARG; }
    size_t off = 0, w = 0;
    int32_t read = 0;
    bool fits = true;
    auto put = [&](const void *src, size_t n) {
        if (out) { if (w + n <= out_cap) memcpy(out + w, src, n); else fits = false; }
        w += n;
    };
    auto info = [&](const idp_result *r, uint16_t fl) {
        char buf[1024];
        idp_format_info(panel, r, fl, buf, (int32_t)sizeof buf);
        return std::string(buf);
    };
    auto is_ours = [](const uint8_t *t) { return (t[0] == 'f' || t[0] == 'r') && (t[1] == '5' || t[1] == '3'); };
    // copies a record, dropping f5/r5/f3/r3 aux fields, and appends the given tags
    auto write_rec = [&](const BamRec &r, const std::vector<std::pair<const char *, std::string>> &tags) -> int {
        const uint8_t *aux = r.seq + (size_t)((r.l_seq + 1) / 2) + (size_t)r.l_seq, *end = r.p + r.block_size;
        std::vector<std::pair<const uint8_t *, size_t>> keep;  // aux byte ranges that stay
        const uint8_t *a = aux;
        while (a < end) {
            if (end - a < 3) { idp::set_error("idp_bam_annotate: truncated aux field"); return IDP_ERR_ARG; }
            const uint8_t *f = a;
            const char ty = (char)a[2];
            a += 3;
            auto fixed = [](char t) { return t == 'A' || t == 'c' || t == 'C' ? 1 : t == 's' || t == 'S' ? 2 : t == 'i' || t == 'I' || t == 'f' ? 4 : 0; };
            if (fixed(ty)) a += fixed(ty);
            else if (ty == 'Z' || ty == 'H') { while (a < end && *a) ++a; ++a; }
            else if (ty == 'B') {
                if (end - a < 5) { idp::set_error("idp_bam_annotate: truncated B aux field"); return IDP_ERR_ARG; }
                const int es = fixed((char)a[0]);
                const uint32_t cnt = (uint32_t)rd32(a + 1);
                if (!es) { idp::set_error("idp_bam_annotate: unknown B subtype"); return IDP_ERR_ARG; }
                a += 5 + (size_t)es * cnt;
            } else { idp::set_error("idp_bam_annotate: unknown aux type '%c'", ty); return IDP_ERR_ARG; }
            if (a > end) { idp::set_error("idp_bam_annotate: aux field runs past the record"); return IDP_ERR_ARG; }
            if (!(tags.size() && is_ours(f))) keep.push_back({f, (size_t)(a - f)});
        }
        size_t body = (size_t)(aux - r.p);
        for (auto &k : keep) body += k.second;
        for (auto &t : tags) body += 3 + t.second.size() + 1;
        const int32_t bs = (int32_t)body;
        put(&bs, 4);
        put(r.p, (size_t)(aux - r.p));
        for (auto &k : keep) put(k.first, k.second);
        for (auto &t : tags) { put(t.first, 2); put("Z", 1); put(t.second.c_str(), t.second.size() + 1); }
        return IDP_OK;
    };
    std::vector<BamRec> group;
    std::string group_name;
    auto flush = [&]() -> int {
        if (group.empty()) return IDP_OK;
        const BamRec *r1 = nullptr, *r2 = nullptr;
        for (auto &g : group) {
            if (g.flag & 0x900) continue;
            const bool paired = g.flag & 0x1;
            if (!paired || (g.flag & 0x40)) r1 = r1 ? r1 : &g; else r2 = r2 ? r2 : &g;
        }
        if (!r1) { idp::set_error("Template did not have an R1: %s", group_name.c_str()); return IDP_ERR_ARG; }
        const int32_t i1 = read, i2 = r2 ? read + 1 : -1;
        read += r2 ? 2 : 1;
        if (read > n_reads) { idp::set_error("idp_bam_annotate: more reads in the BAM bytes than results"); return IDP_ERR_ARG; }
        std::vector<std::pair<const char *, std::string>> tags;
        const uint16_t fl1 = (uint16_t)r1->flag, fl2 = r2 ? (uint16_t)r2->flag : 0;
        if (r2) {  // addTagsToPair: which read is "forward" follows R1's 5' primer, then R2's, else R1 (IP:530-540)
            bool fwd_is_r1 = true;
            if (five[i1].type != IDP_NONE && five[i1].primer >= 0 && five[i1].primer < panel->n) fwd_is_r1 = panel->positive[five[i1].primer] != 0;
            else if (five[i2].type != IDP_NONE && five[i2].primer >= 0 && five[i2].primer < panel->n) fwd_is_r1 = panel->positive[five[i2].primer] == 0;
            const int32_t fi = fwd_is_r1 ? i1 : i2, ri = fwd_is_r1 ? i2 : i1;
            const uint16_t ffl = fwd_is_r1 ? fl1 : fl2, rfl = fwd_is_r1 ? fl2 : fl1;
            tags.push_back({"f5", info(&five[fi], ffl)});
            tags.push_back({"r5", info(&five[ri], rfl)});
            if (want3) { tags.push_back({"f3", info(&three[fi], ffl)}); tags.push_back({"r3", info(&three[ri], rfl)}); }
        } else {   // addTagsToFragment (IP:557-569)
            tags.push_back({"f5", want3 ? info(&three[i1], fl1) : info(&five[i1], fl1)});
            tags.push_back({"r5", "none"});
        }
        int rc = write_rec(*r1, tags);
        if (!rc && r2) rc = write_rec(*r2, tags);
        // Template.allReads: r1, r2, r1 supplementals, r2 supplementals, r1 secondaries, r2 secondaries
        for (int pass = 0; pass < 4 && !rc; ++pass)
            for (auto &g : group) {
                if (rc || !(g.flag & 0x900)) continue;
                const bool supp = (g.flag & 0x800) != 0, second_end = (g.flag & 0x1) && !(g.flag & 0x40);
                if ((pass == 0 && supp && !second_end) || (pass == 1 && supp && second_end) || (pass == 2 && !supp && !second_end) ||
                    (pass == 3 && !supp && second_end))
                    rc = write_rec(g, {});
            }
        group.clear();
        return rc;
    };
    BamRec r;
    while (off < n_bytes) {
        if (!parse(bam, n_bytes, off, r)) { idp::set_error("idp_bam_annotate: malformed BAM record at byte %zu", off); return IDP_ERR_ARG; }
        const std::string name(r.name, r.l_read_name > 0 ? strnlen(r.name, (size_t)r.l_read_name) : 0);
        if (!group.empty() && name != group_name) { int rc = flush(); if (rc) return rc; }
        group_name = name;
        group.push_back(r);
    }
    int rc = flush();
    if (rc) return rc;
    if (read != n_reads) { idp::set_error("idp_bam_annotate: %d results for %d reads in the BAM bytes", n_reads, read); return IDP_ERR_ARG; }
    *out_bytes = w;
    if (out && !fits) { idp::set_error("idp_bam_annotate: output needs %zu bytes, %zu given", w, out_cap); return IDP_ERR_ARG; }
    return IDP_OK;
}

}  // extern "C"
