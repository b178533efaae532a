// Host ingest: uncompressed BAM alignment records -> the SoA batch of include/idprimer.h.
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
    r.name = reinterpret_cast<const char *>(p + 32);
    r.cigar = p + 32 + r.l_read_name;
    r.seq = r.cigar + 4 * (size_t)r.n_cigar;
    if ((size_t)(r.seq - p) + (size_t)((r.l_seq + 1) / 2) > (size_t)r.block_size || r.l_seq < 0) return false;
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
int idp_bam_to_batch(const uint8_t *bam, size_t n_bytes, uint8_t *seq4, uint64_t *seq_off, int32_t *len, int32_t *ref_idx,
                     int32_t *unclipped_start, int32_t *unclipped_end, uint16_t *flags, int32_t *rec_index) {
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
        const size_t nb = (size_t)((r.l_seq + 1) / 2);
        memcpy(seq4 + byte_off, r.seq, nb);
        seq_off[out] = byte_off; len[out] = r.l_seq;
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

}  // extern "C"
