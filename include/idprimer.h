/*
 * idprimer.h -- C-ABI of the B200-native IdentifyPrimers matching path (libidprimer_b200.so).
 *
 * This is the drop-in boundary for fg-idprimer's Location -> Ungapped -> Gapped matcher chain.  One
 * idp_match_batch() call replaces one IdentifyPrimers.processBatch() (IdentifyPrimers.scala:381-436):
 * it runs toPrimeMatchFuture (IP:441-452) for every read, matchThreePrime (IP:483-516) when enabled,
 * and accumulates the TemplateTypeMetric counter (IP:413, 425-426).  Plain pointers and sizes only; no
 * exceptions cross the boundary; errors are a non-zero return plus idp_last_error().
 *
 * Citations: IP = src/main/scala/com/fulcrumgenomics/identifyprimers/IdentifyPrimers.scala,
 * PM = .../PrimerMatcher.scala, PMa = .../PrimerMatch.scala, PR = .../Primer.scala,
 * AA = src/main/scala/com/fulcrumgenomics/alignment/AsyncAligner.scala (all under /root/reference).
 *
 * There is no CPU fallback: every matching entry point fails with IDP_ERR_CUDA when no sm_100 device
 * is usable.  idp_panel_create() and the idp_panel_* / idp_format_* helpers are pure host code.
 */
#ifndef IDPRIMER_H
#define IDPRIMER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IDP_ABI_VERSION 2

/* ---- return codes ---- */
#define IDP_OK               0
#define IDP_ERR_ARG          1   /* NULL / out-of-range argument                                  */
#define IDP_ERR_UNSUPPORTED  2   /* input the GPU path does not cover (see idp_last_error())      */
#define IDP_ERR_CUDA         3   /* CUDA failure or no usable device                              */
#define IDP_ERR_NOMEM        4

/* ---- read flags: the SAM flag bits the path consults + two host-derived bits ---- */
#define IDP_F_PAIRED     0x0001  /* rec.paired                       (PMa:84)                      */
#define IDP_F_UNMAPPED   0x0004  /* rec.unmapped                     (PM:41, PM:194)               */
#define IDP_F_REVERSE    0x0010  /* rec.negativeStrand               (PM:41, PM:179)               */
#define IDP_F_FIRST      0x0040  /* rec.firstOfPair                  (PMa:84)                      */
#define IDP_F_SECOND     0x0080
#define IDP_F_FR_PAIR    0x1000  /* fgbio SamRecord.isFrPair of R1, computed by the host (IP:413)  */
#define IDP_F_MATE_NEXT  0x2000  /* this read is R1 of a template and the next read is its R2.
                                    Must not be set on an R2.  Reads of one template are adjacent. */

/* ---- match types (PMa:33-37) and per-read status ---- */
#define IDP_NONE      0
#define IDP_LOCATION  1   /* LocationBasedPrimerMatch      (PMa:96)  */
#define IDP_UNGAPPED  2   /* UngappedAlignmentPrimerMatch  (PMa:101) */
#define IDP_GAPPED    3   /* GappedAlignmentPrimerMatch    (PMa:111) */

#define IDP_STATUS_OK                  0
#define IDP_STATUS_REQ_MM_LE_NEXT      1  /* require(numMismatches <= nextNumMismatches) would throw (PMa:102) */
#define IDP_STATUS_REQ_SCORE_GE_SECOND 2  /* require(score >= secondBestScore) would throw           (PMa:113) */
#define IDP_STATUS_REQ_START_LE_END    3  /* require(start <= end) would throw                       (PMa:114) */
#define IDP_STATUS_SCORE_MISSING       4  /* scoring matrix has no entry for a base pair (IP:277 NoSuchElementException) */

#define IDP_NEXT_NA 2147483647            /* Int.MaxValue, rendered "na" (PMa:104) */

/* alignment modes of fgbio's Aligner, numbered like the ksw mapping at AA:124-128 */
#define IDP_MODE_LOCAL  0
#define IDP_MODE_GLOCAL 1
#define IDP_MODE_GLOBAL 3

/* ---- option surface of the tool that reaches the hot path (IP:170-185) ---- */
typedef struct {
    int32_t slop;                 /* -S, default 5                                               */
    int32_t three_prime;          /* -3, default -1 (3' matching off)                            */
    int32_t match_score;          /* --match-score 1                                             */
    int32_t mismatch_score;       /* --mismatch-score -4                                         */
    int32_t gap_open;             /* --gap-open -6                                               */
    int32_t gap_extend;           /* --gap-extend -1                                             */
    int32_t k_ungapped;           /* -k; <= 0 means None                                         */
    int32_t k_gapped;             /* -K; <= 0 means None                                         */
    double  max_mismatch_rate;    /* --max-mismatch-rate 0.05                                    */
    double  min_alignment_score_rate; /* --min-alignment-score-rate 0.01                         */
    const int32_t *score_matrix;  /* --scoring-matrix: NULL, or int32[128*128] indexed
                                     [primer_base*128 + read_base], INT32_MIN = no entry (IP:252-278) */
} idp_params;

/* one row of the primer TSV, already parsed and upper-cased by the host (PR:54-112, PR:182-188) */
typedef struct {
    const char *pair_id;
    const char *primer_id;
    const char *sequence;         /* sequencing order, upper case, A/C/G/T/IUPAC/'.'              */
    const char *ref_name;         /* "" (or NULL) = unmapped primer                                */
    int32_t ref_idx;              /* index of ref_name in the reads' sequence dictionary; <0 if absent */
    int32_t start, end;           /* 1-based inclusive; Primer.length = end-start+1 when mapped (PR:206) */
    uint8_t positive_strand;
} idp_primer;

/* a batch of reads, structure of arrays, template by template (R1 then R2) */
typedef struct {
    const uint8_t  *seq4;             /* BAM-native 4-bit bases (=ACMGRSVTWYHKDBN -> 0..15), first base in the
                                         high nibble, each read starting on a byte boundary                  */
    const uint64_t *seq_off;          /* [n] byte offset of each read in seq4; NULL when fixed_len > 0        */
    const int32_t  *len;              /* [n] read length in bases;             NULL when fixed_len > 0        */
    const int32_t  *ref_idx;          /* [n] reference index (-1 unmapped);    NULL if no read is mapped      */
    const int32_t  *unclipped_start;  /* [n] rec.unclippedStart (PM:180);      NULL if no read is mapped      */
    const int32_t  *unclipped_end;    /* [n] rec.unclippedEnd   (PM:186);      NULL if no read is mapped      */
    const uint16_t *flags;            /* [n] IDP_F_* bits                                                     */
    int32_t fixed_len;                /* >0: every read has this length and occupies (fixed_len+1)/2 bytes    */
} idp_reads;

/* one PrimerMatch (PMa:96-117) as integers; the JVM-side doubles are derived with idp_result_score() */
typedef struct {
    int32_t primer;       /* index into the panel; -1 = no match                                              */
    uint8_t type;         /* IDP_NONE / LOCATION / UNGAPPED / GAPPED                                          */
    uint8_t status;       /* IDP_STATUS_*: non-zero where the reference would have thrown                     */
    uint8_t multi;        /* gapped: 1 if two or more candidates were aligned (PM:317 vs PM:313)              */
    uint8_t reserved;
    int32_t mm;           /* location / ungapped: numMismatches                                               */
    int32_t next_mm;      /* ungapped: nextNumMismatches or IDP_NEXT_NA                                       */
    int32_t raw_score;    /* gapped: Alignment.score of the best candidate                                    */
    int32_t raw_next;     /* gapped: Alignment.score of the runner-up (0 when !multi)                         */
    int16_t q_len;        /* gapped: alignment.query.length of the best candidate (PM:301)                    */
    int16_t q_len_next;   /* gapped: same for the runner-up (0 when !multi)                                   */
    int16_t t_start;      /* gapped: Alignment.targetStart, 1-based (PM:315, 323)                             */
    int16_t t_end;        /* gapped: Alignment.targetEnd                                                      */
} idp_result;             /* 32 bytes */

/* The same PrimerMatch in 16 bytes: the record the windowed / compact entry points (idp_match_batch16*) write, half the
 * device->host bytes of idp_result.  A type-tagged union: `a` / `b` are numMismatches / nextNumMismatches for Location and
 * Ungapped matches (PMa:96-104; b = IDP_NEXT16_NA renders "na") and Alignment.score of the best / runner-up candidate for
 * Gapped ones (PM:315-322).  idp_result16_unpack() gives back the 32-byte record bit for bit. */
#define IDP_NEXT16_NA 32767
typedef struct {
    uint32_t head;        /* bits 0..23 primer index + 1 (0 = no match), 24..25 type, 26 multi, 27..29 status             */
    int16_t  a;           /* location / ungapped: mm;       gapped: raw_score                                          */
    int16_t  b;           /* ungapped: next_mm or IDP_NEXT16_NA; gapped: raw_next; else 0                              */
    uint16_t q_len;       /* gapped: as in idp_result                                                                  */
    uint16_t q_len_next;
    uint16_t t_start;
    uint16_t t_end;
} idp_result16;           /* 16 bytes */

/* device-side timing and work counters of the last idp_match_batch*() call on a context */
typedef struct {
    float    ms_total;            /* first H2D/kernel to last D2H/kernel, CUDA events                          */
    float    ms_scan;             /* kernel A+B: k-mer prefilter + location + ungapped scan                    */
    float    ms_gapped;           /* candidate emission + Smith-Waterman score + finalize + traceback (5')     */
    float    ms_three_prime;      /* the same for the 3' pass                                                  */
    float    ms_sw_score;         /* the score-only Smith-Waterman kernel alone (5' + 3')                      */
    int64_t  n_reads;
    int64_t  n_fallthrough;       /* reads that reached the gapped stage                                       */
    int64_t  n_alignments_5p;     /* gapped alignments, 5' (== the reference's numAlignments, IP:424)          */
    int64_t  n_alignments_3p;
    int64_t  sw_cells;            /* sum over alignments of |query| * |target|                                 */
    int64_t  scan_base_compares;  /* sum over (read, candidate) of min(readLen, primer bytes); counted only when
                                   * IDP_SCAN_STATS=1 is in the environment at idp_ctx_create (the counting variant
                                   * of the scan kernel is 3 % slower), else 0                                    */
    int64_t  kernel_launches;     /* kernels of this library launched by the call                              */
    int64_t  h2d_bytes, d2h_bytes;
    int64_t  n_scan_parked;       /* reads the exact-prefix table of the scan kernel did not decide (they took the filter pass) */
    int64_t  n_scan_deferred;     /* of those, reads whose one accepted primer needed the exact k-mer membership test    */
} idp_timings;

typedef struct idp_panel idp_panel;  /* immutable after create; shareable across threads and devices           */
typedef struct idp_ctx   idp_ctx;    /* one per calling thread and device: streams, staging, scratch           */

/* ---- library ---- */
int         idp_abi_version(void);
const char *idp_last_error(void);                       /* thread-local message of the last failing call     */
int         idp_device_count(void);                     /* number of usable CUDA devices (0 if none)         */
/* The alignment choices this build was compiled with that nothing in the reference pins (fgbio's Aligner is an un-vendored
 * dependency): bit 0 Left may open from Up, 1 Up may open from Left, 2 ties prefer extending a gap over opening one, 3 ties are
 * tried Diagonal, Up, Left.  0 = the published algorithm (the default build); see idp_internal.h, AlignPolicy. */
int         idp_align_policy(void);

/* ---- panel: the three matchers' immutable state (IP:223-234; PM:62-117, 156-175) ---- */
int  idp_panel_create(const idp_primer *primers, int32_t n, const idp_params *params, idp_panel **out);
void idp_panel_destroy(idp_panel *panel);
int32_t  idp_panel_size(const idp_panel *panel);
int32_t  idp_panel_primer_length(const idp_panel *panel, int32_t i);          /* Primer.length (PR:206)       */
uint32_t idp_panel_primer_hash(const idp_panel *panel, int32_t i);            /* Scala case-class hashCode    */
/* kmerToPrimers(kmer) in the iteration order the reference's Set has (PM:78-91, 113); which: 0 = -k, 1 = -K  */
int32_t  idp_panel_kmer_postings(const idp_panel *panel, int which, const char *kmer, int32_t *out, int32_t cap);

/* ---- context ---- */
int  idp_ctx_create(const idp_panel *panel, int device, idp_ctx **out);
void idp_ctx_destroy(idp_ctx *ctx);
/* run on the caller's CUDA stream (a cudaStream_t passed as void*); NULL = the context's own stream */
int  idp_ctx_set_stream(idp_ctx *ctx, void *cuda_stream);
int  idp_ctx_timings(const idp_ctx *ctx, idp_timings *out);

/* pinned host memory for batches that are copied asynchronously (plain malloc'd buffers also work, slower) */
void *idp_host_alloc(size_t bytes);
void  idp_host_free(void *p);

/* ---- the hot path ----
 * processBatch (IP:381-436).  five[n] receives each read's 5' match, three[n] its 3' match (may be NULL when
 * params.three_prime < 0).  metrics160 (may be NULL) is ADDED to: bin = ((template_type*2 + is_fr)*4 + r1_type)*4
 * + r2_type with template_type in the order of TemplateType.scala:38-46.  *n_gapped_alignments (may be NULL) is
 * ADDED to and equals the reference's numAlignments increment (5' alignments only, IP:424).
 * All read arrays and result arrays are HOST pointers; copies are pipelined with compute. */
int idp_match_batch(idp_ctx *ctx, const idp_reads *reads, int32_t n_reads, idp_result *five, idp_result *three,
                    int64_t *metrics160, int64_t *n_gapped_alignments);
/* Same, but every pointer inside `reads`, and `five`/`three`, are DEVICE pointers on the context's device
 * (metrics160 / n_gapped_alignments stay host pointers).  No host<->device copy of reads or results happens. */
int idp_match_batch_device(idp_ctx *ctx, const idp_reads *reads, int32_t n_reads, idp_result *five, idp_result *three,
                           int64_t *metrics160, int64_t *n_gapped_alignments);

/* Asynchronous device-resident mode.  After idp_ctx_set_async(ctx, 1) the two *_device calls only ENQUEUE the batch on the
 * context's stream (idp_ctx_set_stream) and return; their metrics160 / n_gapped_alignments arguments are ignored.  Batches run
 * back to back without a host round trip in between and their counters accumulate on the device; idp_ctx_sync() waits for
 * everything enqueued since the last sync and ADDS the accumulated counters to its arguments (either may be NULL).  Input and
 * output arrays of a batch must stay valid until then.  If a batch lists more -K candidates than the pair buffer holds (the
 * synchronous calls redo such a batch themselves), idp_ctx_sync returns IDP_ERR_NOMEM and enlarges the buffer: rerun the batches
 * enqueued since the last sync.  idp_ctx_timings then reports the stage times of the last batch and the work counters of all.
 * enable == 2: the same without the per-stage CUDA events and the per-batch copy of the counters (each is a small bubble on the
 * stream between two kernels): idp_ctx_timings then reports ms_total of the last batch and zero stage times. */
int idp_ctx_set_async(idp_ctx *ctx, int enable);
int idp_ctx_sync(idp_ctx *ctx, int64_t *metrics160, int64_t *n_gapped_alignments);

/* The same two calls writing 16-byte records (idp_result16).  This is what a caller that ships windowed batches
 * (idp_bam_to_batch_window) should use: 18 B of bases + 2 B of flags in and 16 B out per 150 bp read instead of 77 B in and
 * 32 B out.  IDP_ERR_UNSUPPORTED when the panel's scores or size do not fit the record (|alignment score| >= 32767 possible,
 * or 2^24 - 1 primers or more): use the 32-byte calls then. */
int idp_match_batch16(idp_ctx *ctx, const idp_reads *reads, int32_t n_reads, idp_result16 *five, idp_result16 *three,
                      int64_t *metrics160, int64_t *n_gapped_alignments);
int idp_match_batch16_device(idp_ctx *ctx, const idp_reads *reads, int32_t n_reads, idp_result16 *five, idp_result16 *three,
                             int64_t *metrics160, int64_t *n_gapped_alignments);
/* host: n 16-byte records -> the 32-byte records idp_match_batch would have written (pure function) */
void idp_result16_unpack(const idp_result16 *in, int64_t n, idp_result *out);

/* ---- the aligner seam: a batched AsyncAligner.align (AA:37-54) ----
 * queries/targets are ASCII (upper-case IUPAC or '='), concatenated, with n+1 offsets starting at 0; outputs are
 * Alignment.score, targetStart and targetEnd (1-based) and, when query_start / query_end are not NULL (both or neither),
 * queryStart / queryEnd.  mode is IDP_MODE_*; scoring comes from the context's panel parameters.  The letters are packed and
 * checked on the GPU; queries of up to 64 bases run on the register-resident aligners (Local with query coordinates, Global
 * and longer queries on the generic one).  Host pointers; the call returns when the results are in place. */
int idp_align_batch(idp_ctx *ctx, int mode, const char *queries, const int64_t *q_off, const char *targets,
                    const int64_t *t_off, int32_t n, int32_t *score, int32_t *target_start, int32_t *target_end,
                    int32_t *query_start, int32_t *query_end);

/* ---- host-side result helpers (pure functions, same IEEE double operations as the JVM) ---- */
double idp_result_score(const idp_result *r);        /* GappedAlignmentPrimerMatch.score        (PM:315, 321) */
double idp_result_second(const idp_result *r);       /* GappedAlignmentPrimerMatch.secondBestScore (PM:315, 322) */
int    idp_metric_bin(int template_type, int is_fr, int r1_type, int r2_type);
void   idp_summary_metrics(const int64_t *metrics160, int64_t *out17);  /* IdentifyPrimersMetric.scala:34-89 */
/* PrimerMatch.info(rec) (PMa:78-88), "none" for no match (IP:242); returns the string length */
int    idp_format_info(const idp_panel *panel, const idp_result *r, uint16_t read_flags, char *buf, int32_t cap);
/* Downstream pair classification (scripts/post_process_bam.py:56-78) on the forward / reverse 5' slots of addTagsToPair
 * (IP:530-540).  idp_classify_primer_pair takes the primers in the f5 / r5 tags (-1 = "none").  idp_classify_templates walks a
 * batch: cls_per_read[n] (may be NULL) receives the class of each read's template (both reads of a pair get the same value, as
 * the script tags both records); counts6 (may be NULL) is ADDED to once per template whose first record is read 1
 * (post_process_bam.py:159).  Pure host functions. */
#define IDP_CLASS_CANONICAL     0
#define IDP_CLASS_SELF_DIMER    1
#define IDP_CLASS_CROSS_DIMER   2
#define IDP_CLASS_NON_CANONICAL 3
#define IDP_CLASS_SINGLE        4
#define IDP_CLASS_NO_MATCH      5
int    idp_classify_primer_pair(const idp_panel *panel, int32_t f5_primer, int32_t r5_primer, int32_t min_insert_length, int32_t max_insert_length);
int    idp_classify_templates(const idp_panel *panel, const uint16_t *flags, const idp_result *five, int32_t n_reads,
                              int32_t min_insert_length, int32_t max_insert_length, uint8_t *cls_per_read, int64_t *counts6);
/* The same classification as a DEVICE epilogue on the 5' records of a batch (no record ever crosses PCIe for it).
 * idp_classify_batch_device: flags / five / cls_per_read are device pointers on the context's device; five holds n_reads
 * records, idp_result when compact = 0 and idp_result16 when compact = 1, as the idp_match_batch*_device calls leave them;
 * cls_per_read (may be NULL) receives one IDP_CLASS_* byte per read, counts6 (host, may be NULL) is ADDED to.
 * idp_ctx_set_classify makes every later idp_match_batch* call of the context run it right after the 5' matching (before the 3'
 * pass): cls_per_read (may be NULL) is then a HOST array for idp_match_batch / idp_match_batch16 and a DEVICE array for the
 * *_device calls, counts6 (host, may be NULL) is ADDED to by every call; min_insert_length < 0 switches the epilogue off. */
int    idp_classify_batch_device(idp_ctx *ctx, const uint16_t *flags, const void *five, int compact, int32_t n_reads,
                                 int32_t min_insert_length, int32_t max_insert_length, uint8_t *cls_per_read, int64_t *counts6);
int    idp_ctx_set_classify(idp_ctx *ctx, int32_t min_insert_length, int32_t max_insert_length, uint8_t *cls_per_read, int64_t *counts6);
/* Host ingest: uncompressed BAM alignment records (the bytes inside the BGZF blocks: block_size-prefixed records, SAMv1 4.2),
 * queryname-grouped as Bams.templateIterator consumes them (IP:288-289), into the arrays of idp_reads.  The record's own 4-bit
 * `seq` bytes are copied unchanged; unclipped ends come from the CIGAR (PM:180, 186), IDP_F_FR_PAIR from the mate fields (fgbio
 * SamRecord.isFrPair), IDP_F_MATE_NEXT from the grouping; secondary / supplementary records are skipped; reads come out R1 then
 * R2 per template and rec_index[i] (may be NULL) is the ordinal of the BAM record behind read i.  idp_bam_scan sizes the arrays. */
int    idp_bam_scan(const uint8_t *bam, size_t n_bytes, int32_t *n_reads, uint64_t *seq_bytes);
int    idp_bam_to_batch(const uint8_t *bam, size_t n_bytes, uint8_t *seq4, uint64_t *seq_off, int32_t *len, int32_t *ref_idx,
                        int32_t *unclipped_start, int32_t *unclipped_end, uint16_t *flags, int32_t *rec_index);
/* Windowed ingest for 5'-only runs (--three-prime off).  idp_panel_window is the number of leading bases (sequencing order)
 * that can influence the Location / Ungapped / Gapped chain: max(maxPrimerLength, longest primer sequence + slop, k, K)
 * (PM:109-110, 130, 139, 293); it is -1 when the panel has --three-prime set, because matchThreePrime aligns against the
 * whole read (IP:501-505).  idp_bam_to_batch_window fills the same arrays as idp_bam_to_batch but keeps only that many bases
 * of every read -- the first ones, or the last ones of a mapped reverse-strand read, whose 5' end is the end of the stored
 * bases (PM:39-45) -- so that idp_match_batch returns the same results from a batch a fraction of the size (150 bp reads
 * against 30 bp primers: 18 instead of 75 bytes per read over PCIe).  Arrays are sized by idp_bam_scan as before. */
int32_t idp_panel_window(const idp_panel *panel);
int    idp_bam_to_batch_window(const idp_panel *panel, const uint8_t *bam, size_t n_bytes, uint8_t *seq4, uint64_t *seq_off, int32_t *len,
                               int32_t *ref_idx, int32_t *unclipped_start, int32_t *unclipped_end, uint16_t *flags, int32_t *rec_index);
/* Host egress: the records of the same bytes with the primer-match tags set as addTagsToPair / addTagsToFragment do
 * (IP:522-569): Z-type aux fields f5 / r5 (+ f3 / r3 when --three-prime >= 0) on the primary R1 / R2 of every template, the tag
 * strings being PrimerMatch.info / "none" (idp_format_info).  five / three are the idp_match_batch results of the batch
 * idp_bam_to_batch made from these bytes.  Records come out in Template.allReads order; out == NULL only sizes (*out_bytes).
 * idp_header_comments returns the three @CO lines of IP:243-250 ('\n'-separated). */
int    idp_bam_annotate(const idp_panel *panel, const uint8_t *bam, size_t n_bytes, const idp_result *five, const idp_result *three,
                        int32_t n_reads, uint8_t *out, size_t out_cap, size_t *out_bytes);
int    idp_header_comments(char *buf, int32_t cap);
/* pack ASCII bases to BAM nibbles / back (htsjdk SAMUtils.bytesToCompressedBases semantics) */
void   idp_pack_bases(const char *ascii, int32_t n, uint8_t *seq4);
void   idp_unpack_bases(const uint8_t *seq4, int32_t n, char *ascii);

#ifdef __cplusplus
}
#endif
#endif /* IDPRIMER_H */
