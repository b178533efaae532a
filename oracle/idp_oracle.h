/*
 * idp_oracle.h -- CPU ORACLE for the IdentifyPrimers matching hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference
 * algorithm (fulcrumgenomics/fg-idprimer, Scala) used as the checker by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs.  Nothing in fg_idprimer_b200/ (the product) may call it.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   PINNED   by the reference's own known-answer tests (tests/golden/):
 *            IUPAC mismatch count + early exit (PrimerMatcherTest.scala:161-185),
 *            mismatchAlign accept rule (:187-218), k-mer candidate lookup
 *            (:89-145), location matching (:236-300), best/next-best ungapped
 *            (:322-374) and gapped (:396-422) selection, and the end-to-end
 *            metrics / per-read match types (IdentifyPrimersTest.scala:238-403).
 *   UNPINNED gapped alignment score / targetStart / targetEnd values and all
 *            alignment tie-breaks: the arithmetic lives in the un-vendored
 *            dependency com.fulcrumgenomics:fgbio_2.12:0.8.0-8d0176b-SNAPSHOT
 *            (build.sbt:127) which is absent from /root/reference and no JVM
 *            exists on this box.  orc_align() restates its published three-
 *            matrix affine-gap algorithm; the reference's tests only pin
 *            accept/reject outcomes at that boundary.
 *   UNPINNED candidate order under -k/-K (Scala 2.12 HashSet iteration order).
 *
 * Citations: PM = src/main/scala/com/fulcrumgenomics/identifyprimers/PrimerMatcher.scala
 *            PMa = .../PrimerMatch.scala  PR = .../Primer.scala
 *            IP = .../IdentifyPrimers.scala  AA = src/main/scala/com/fulcrumgenomics/alignment/AsyncAligner.scala
 */
#ifndef IDP_ORACLE_H
#define IDP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Read flag bits: the SAM flag bits the path looks at, plus two host-derived bits. */
#define ORC_F_PAIRED     0x0001  /* rec.paired                                    */
#define ORC_F_UNMAPPED   0x0004  /* rec.unmapped                                  */
#define ORC_F_REVERSE    0x0010  /* rec.negativeStrand                            */
#define ORC_F_FIRST      0x0040  /* rec.firstOfPair                               */
#define ORC_F_SECOND     0x0080
#define ORC_F_FR_PAIR    0x1000  /* fgbio SamRecord.isFrPair, host-computed (IP:413) */
#define ORC_F_MATE_NEXT  0x2000  /* this read is R1 and the next read is its R2   */

/* match types (PMa:33-37) */
enum { ORC_NONE = 0, ORC_LOCATION = 1, ORC_UNGAPPED = 2, ORC_GAPPED = 3 };
/* status: which require() of the reference would have thrown (the JVM tool aborts) */
enum { ORC_OK = 0, ORC_REQ_MM_LE_NEXT = 1 /* PMa:102 */, ORC_REQ_SCORE_GE_SECOND = 2 /* PMa:113 */,
       ORC_REQ_START_LE_END = 3 /* PMa:114 */ };
/* alignment modes (fgbio Mode; ksw mapping AA:124-128) */
enum { ORC_MODE_LOCAL = 0, ORC_MODE_GLOCAL = 1, ORC_MODE_GLOBAL = 3 };

#define ORC_NEXT_NA 2147483647 /* Int.MaxValue, rendered "na" (PMa:104) */

typedef struct {
    const char *pair_id, *primer_id, *sequence; /* sequence already upper-cased (PR:85) */
    int positive_strand;
    const char *ref_name;                        /* "" = unmapped primer (PR:209)      */
    int start, end;
    int ref_idx;                                 /* index of ref_name in the read dictionary; <0 if none */
} orc_primer;

typedef struct {
    int slop, three_prime;                       /* IP:170-171 */
    int match_score, mismatch_score, gap_open, gap_extend; /* IP:175-178 */
    int k_ungapped, k_gapped;                    /* IP:182-183; <=0 means None */
    double max_mismatch_rate, min_alignment_score_rate; /* IP:172-174 */
    const int *score_matrix;                     /* NULL or int[128*128] indexed [primer_base*128+read_base]; INT32_MIN = missing (IP:252-278) */
} orc_params;

typedef struct {
    const char *bases;          /* ASCII bases of all reads, concatenated, as stored in the BAM record */
    const int64_t *off;         /* n+1 offsets into bases */
    const int32_t *ref_idx;     /* -1 when unmapped */
    const int32_t *unclipped_start, *unclipped_end; /* PM:180,186 */
    const uint16_t *flags;
} orc_reads;

typedef struct {
    int32_t primer;             /* index into the panel, -1 = None */
    uint8_t type, status, multi /* gapped: >=2 candidates were aligned (PM:317) */, pad;
    int32_t mm, next_mm;        /* location / ungapped */
    int32_t raw_score, raw_next;/* gapped: Alignment.score of best and runner-up */
    int16_t q_len, q_len_next;  /* gapped: alignment.query.length of best / runner-up */
    int16_t t_start, t_end;     /* gapped: Alignment.targetStart / targetEnd (1-based) */
} orc_result;

typedef struct orc_panel orc_panel;

/* ---- panel / matcher construction (IP:223-234) ---- */
orc_panel *orc_panel_create(const orc_primer *primers, int n, const orc_params *params);
void       orc_panel_destroy(orc_panel *);
int        orc_primer_length(const orc_panel *, int i);           /* PR:206 */
uint32_t   orc_primer_scala_hash(const orc_panel *, int i);       /* case-class hashCode (Scala 2.12) */

/* ---- the individual reference functions, exposed for known-answer tests ---- */
/* PM:138-147 */
int orc_num_mismatches(const char *bases, int bases_len, const char *primer, int primer_len, int start, double max_mismatches);
/* PM:108-116; which=0 ungapped matcher (-k), 1 gapped matcher (-K).  Returns count, fills out[] (cap entries). */
int orc_candidates_with_common_kmer(const orc_panel *, int which, const char *bases, int len, int kmer_len, int *out, int cap);
/* PM:128-135; returns 1 and *mm if accepted */
int orc_mismatch_align(const orc_panel *, const char *seq_order_bases, int len, int primer, int *mm);
/* PM:177-191 on a mapped read; returns number of overlapping primers, fills out[] */
int orc_location_overlaps(const orc_panel *, int ref_idx, int negative, int ustart, int uend, int *out, int cap);
/* PM:237-248 over (primer, mm) pairs in order */
void orc_best_ungapped(const orc_panel *, const int *primers, const int *mms, int n, orc_result *out);
/* PM:305-328 over (primer, score, qlen, tstart, tend) in order */
void orc_best_gapped(const int *primers, const int *scores, const int *qlens, const int *tstarts, const int *tends, int n, orc_result *out);
/* fgbio Aligner.align restatement: returns score, 1-based starts/ends, CIGAR lengths */
typedef struct { int32_t score, query_start, target_start, query_end, target_end; } orc_alignment;
void orc_align(const orc_params *, int mode, const char *query, int qlen, const char *target, int tlen, orc_alignment *out);

/* ---- processBatch (IP:381-436): 5' chain, optional 3' chain, metrics ---- */
/* five[n], three[n] (may be NULL when three_prime<0), metrics160[160] is ADDED to, n_align is ADDED to.
 * n_threads>1 splits templates across pthreads the way IP:311-329 splits batches. */
int orc_process_batch(const orc_panel *, const orc_reads *, int32_t n_reads, orc_result *five, orc_result *three,
                      int64_t *metrics160, int64_t *n_align, int n_threads);

/* doubles the JVM would hold in GappedAlignmentPrimerMatch (PM:315, 321-322) */
double orc_result_score(const orc_result *);
double orc_result_second(const orc_result *);

/* metric bin = ((template_type*2 + is_fr)*4 + r1_type)*4 + r2_type, template_type order of TemplateType.scala:38-46 */
int orc_metric_bin(int template_type, int is_fr, int r1_type, int r2_type);
/* IdentifyPrimersMetric.scala:34-89: out17 in the field order of the case class (:112-130) */
void orc_summary_metrics(const int64_t *metrics160, int64_t *out17);

/* tag strings: PrimerMatch.info (PMa:78-88) and the f5/r5/f3/r3 slot logic (IP:522-569).  buf sizes >= 512. */
void orc_format_info(const orc_panel *, const orc_result *, uint16_t read_flags, char *buf, int cap);
void orc_java_format_2f(double v, char *buf, int cap); /* Java f"%.2f" (PMa:115) */

#ifdef __cplusplus
}
#endif
/* the unpinned alignment choices this build was compiled with (bit 0 left_from_up, 1 up_from_left, 2 extend_over_open, 3 tie_up_before_left) */
int orc_align_policy_bits(void);

#endif
