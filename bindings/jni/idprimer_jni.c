/*
 * JNI shim over include/idprimer.h for com.fulcrumgenomics.identifyprimers.GpuPrimerMatcher (bindings/scala/GpuPrimerMatcher.scala)
 * and com.fulcrumgenomics.alignment.AsyncCudaAligner (bindings/scala/AsyncCudaAligner.scala).
 * NOT compiled in this image (no JDK, no jni.h).  Build on a machine with a JDK:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude bindings/jni/idprimer_jni.c \
 *       -Lfg_idprimer_b200 -lidprimer_b200 -o libidprimer_jni.so
 *
 * Every export of idprimer.h has an entry below (the table at the end of INTEGRATION.md lists them side by side); the two
 * aligner* entries are conveniences built from idp_panel_create + idp_ctx_create.  Bulk data crosses as direct ByteBuffers
 * (no copies; a null buffer is a NULL pointer), handles as jlong; a failing call throws RuntimeException(idp_last_error()).
 * Device pointers (the *_device entries, idp_ctx_set_stream) are passed as jlong addresses.
 */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "idprimer.h"

#define GPM(name) Java_com_fulcrumgenomics_identifyprimers_GpuPrimerMatcher_00024Native_##name
#define ACA(name) Java_com_fulcrumgenomics_alignment_AsyncCudaAligner_00024Native_##name
#define ADDR(buf) ((buf) ? (*env)->GetDirectBufferAddress(env, (buf)) : NULL)
#define PTR(T, h) ((T *)(intptr_t)(h))

static void throw_idp(JNIEnv *env) {
    jclass cls = (*env)->FindClass(env, "java/lang/RuntimeException");
    if (cls) (*env)->ThrowNew(env, cls, idp_last_error());
}
static void check(JNIEnv *env, int rc) { if (rc != IDP_OK) throw_idp(env); }
static idp_reads reads_of(JNIEnv *env, jobject seq4, jobject seqOff, jobject len, jobject refIdx, jobject uStart, jobject uEnd, jobject flags, jint fixedLen) {
    idp_reads r;
    r.seq4 = (const uint8_t *)ADDR(seq4); r.seq_off = (const uint64_t *)ADDR(seqOff); r.len = (const int32_t *)ADDR(len);
    r.ref_idx = (const int32_t *)ADDR(refIdx); r.unclipped_start = (const int32_t *)ADDR(uStart); r.unclipped_end = (const int32_t *)ADDR(uEnd);
    r.flags = (const uint16_t *)ADDR(flags); r.fixed_len = fixedLen;
    return r;
}
static idp_reads reads_of_device(jlong seq4, jlong seqOff, jlong len, jlong refIdx, jlong uStart, jlong uEnd, jlong flags, jint fixedLen) {
    idp_reads r;
    r.seq4 = PTR(const uint8_t, seq4); r.seq_off = PTR(const uint64_t, seqOff); r.len = PTR(const int32_t, len);
    r.ref_idx = PTR(const int32_t, refIdx); r.unclipped_start = PTR(const int32_t, uStart); r.unclipped_end = PTR(const int32_t, uEnd);
    r.flags = PTR(const uint16_t, flags); r.fixed_len = fixedLen;
    return r;
}

/* ---- library ---- */
JNIEXPORT jint JNICALL GPM(abiVersion)(JNIEnv *env, jclass c) { return idp_abi_version(); }
JNIEXPORT jstring JNICALL GPM(lastError)(JNIEnv *env, jclass c) { return (*env)->NewStringUTF(env, idp_last_error()); }
JNIEXPORT jint JNICALL GPM(deviceCount)(JNIEnv *env, jclass c) { return idp_device_count(); }
JNIEXPORT jint JNICALL GPM(alignPolicy)(JNIEnv *env, jclass c) { return idp_align_policy(); }

/* ---- panel ---- */
/* primers: idp_primer[n] laid out by the caller in a direct buffer whose char* fields point into other direct buffers it keeps alive */
JNIEXPORT jlong JNICALL GPM(panelCreate)(JNIEnv *env, jclass c, jobject primers, jint n, jobject params) {
    idp_panel *p = NULL;
    if (idp_panel_create((const idp_primer *)ADDR(primers), n, (const idp_params *)ADDR(params), &p) != IDP_OK) { throw_idp(env); return 0; }
    return (jlong)(intptr_t)p;
}
JNIEXPORT void JNICALL GPM(panelDestroy)(JNIEnv *env, jclass c, jlong h) { idp_panel_destroy(PTR(idp_panel, h)); }
JNIEXPORT jint JNICALL GPM(panelSize)(JNIEnv *env, jclass c, jlong h) { return idp_panel_size(PTR(const idp_panel, h)); }
JNIEXPORT jint JNICALL GPM(panelPrimerLength)(JNIEnv *env, jclass c, jlong h, jint i) { return idp_panel_primer_length(PTR(const idp_panel, h), i); }
JNIEXPORT jint JNICALL GPM(panelPrimerHash)(JNIEnv *env, jclass c, jlong h, jint i) { return (jint)idp_panel_primer_hash(PTR(const idp_panel, h), i); }
JNIEXPORT jint JNICALL GPM(panelKmerPostings)(JNIEnv *env, jclass c, jlong h, jint which, jstring kmer, jobject out, jint cap) {
    const char *k = (*env)->GetStringUTFChars(env, kmer, NULL);
    const jint n = idp_panel_kmer_postings(PTR(const idp_panel, h), which, k, (int32_t *)ADDR(out), cap);
    (*env)->ReleaseStringUTFChars(env, kmer, k);
    return n;
}
JNIEXPORT jint JNICALL GPM(panelWindow)(JNIEnv *env, jclass c, jlong h) { return idp_panel_window(PTR(const idp_panel, h)); }

/* ---- context ---- */
JNIEXPORT jlong JNICALL GPM(ctxCreate)(JNIEnv *env, jclass c, jlong panel, jint device) {
    idp_ctx *x = NULL;
    if (idp_ctx_create(PTR(const idp_panel, panel), device, &x) != IDP_OK) { throw_idp(env); return 0; }
    return (jlong)(intptr_t)x;
}
JNIEXPORT void JNICALL GPM(ctxDestroy)(JNIEnv *env, jclass c, jlong h) { idp_ctx_destroy(PTR(idp_ctx, h)); }
JNIEXPORT void JNICALL GPM(ctxSetStream)(JNIEnv *env, jclass c, jlong h, jlong cudaStream) { check(env, idp_ctx_set_stream(PTR(idp_ctx, h), PTR(void, cudaStream))); }
JNIEXPORT void JNICALL GPM(ctxSetAsync)(JNIEnv *env, jclass c, jlong h, jint enable) { check(env, idp_ctx_set_async(PTR(idp_ctx, h), enable)); }
JNIEXPORT void JNICALL GPM(ctxSync)(JNIEnv *env, jclass c, jlong h, jobject metrics, jobject nAligned) { check(env, idp_ctx_sync(PTR(idp_ctx, h), (int64_t *)ADDR(metrics), (int64_t *)ADDR(nAligned))); }
JNIEXPORT void JNICALL GPM(ctxTimings)(JNIEnv *env, jclass c, jlong h, jobject out /* idp_timings */) {
    /* the struct grows at its end from one ABI revision to the next: refuse a buffer cut for an older one instead of writing past it */
    if (out && (*env)->GetDirectBufferCapacity(env, out) < (jlong)sizeof(idp_timings)) {
        jclass cls = (*env)->FindClass(env, "java/lang/IllegalArgumentException");
        if (cls) (*env)->ThrowNew(env, cls, "ctxTimings: the buffer is smaller than idp_timings");
        return;
    }
    check(env, idp_ctx_timings(PTR(const idp_ctx, h), (idp_timings *)ADDR(out)));
}
/* clsPerRead: a direct buffer for matchBatch / matchBatch16 (or null); use ctxSetClassifyDevice for the *_device calls */
JNIEXPORT void JNICALL GPM(ctxSetClassify)(JNIEnv *env, jclass c, jlong h, jint minInsert, jint maxInsert, jobject clsPerRead, jobject counts6) {
    check(env, idp_ctx_set_classify(PTR(idp_ctx, h), minInsert, maxInsert, (uint8_t *)ADDR(clsPerRead), (int64_t *)ADDR(counts6)));
}
JNIEXPORT void JNICALL GPM(ctxSetClassifyDevice)(JNIEnv *env, jclass c, jlong h, jint minInsert, jint maxInsert, jlong clsPerRead, jobject counts6) {
    check(env, idp_ctx_set_classify(PTR(idp_ctx, h), minInsert, maxInsert, PTR(uint8_t, clsPerRead), (int64_t *)ADDR(counts6)));
}
/* pinned host memory as a direct ByteBuffer (idp_host_alloc / idp_host_free) */
JNIEXPORT jobject JNICALL GPM(hostAlloc)(JNIEnv *env, jclass c, jlong bytes) {
    void *p = idp_host_alloc((size_t)bytes);
    if (!p) { throw_idp(env); return NULL; }
    return (*env)->NewDirectByteBuffer(env, p, bytes);
}
JNIEXPORT void JNICALL GPM(hostFree)(JNIEnv *env, jclass c, jobject buf) { idp_host_free(ADDR(buf)); }

/* ---- the hot path: processBatch (IP:381-436) ---- */
JNIEXPORT void JNICALL GPM(matchBatch)(JNIEnv *env, jclass c, jlong ctx, jobject seq4, jobject seqOff, jobject len, jobject refIdx, jobject uStart,
                                       jobject uEnd, jobject flags, jint fixedLen, jint n, jobject five, jobject three, jobject metrics, jobject nAligned) {
    const idp_reads r = reads_of(env, seq4, seqOff, len, refIdx, uStart, uEnd, flags, fixedLen);
    check(env, idp_match_batch(PTR(idp_ctx, ctx), &r, n, (idp_result *)ADDR(five), (idp_result *)ADDR(three), (int64_t *)ADDR(metrics), (int64_t *)ADDR(nAligned)));
}
JNIEXPORT void JNICALL GPM(matchBatch16)(JNIEnv *env, jclass c, jlong ctx, jobject seq4, jobject seqOff, jobject len, jobject refIdx, jobject uStart,
                                         jobject uEnd, jobject flags, jint fixedLen, jint n, jobject five, jobject three, jobject metrics, jobject nAligned) {
    const idp_reads r = reads_of(env, seq4, seqOff, len, refIdx, uStart, uEnd, flags, fixedLen);
    check(env, idp_match_batch16(PTR(idp_ctx, ctx), &r, n, (idp_result16 *)ADDR(five), (idp_result16 *)ADDR(three), (int64_t *)ADDR(metrics), (int64_t *)ADDR(nAligned)));
}
JNIEXPORT void JNICALL GPM(matchBatchDevice)(JNIEnv *env, jclass c, jlong ctx, jlong seq4, jlong seqOff, jlong len, jlong refIdx, jlong uStart, jlong uEnd,
                                             jlong flags, jint fixedLen, jint n, jlong five, jlong three, jobject metrics, jobject nAligned) {
    const idp_reads r = reads_of_device(seq4, seqOff, len, refIdx, uStart, uEnd, flags, fixedLen);
    check(env, idp_match_batch_device(PTR(idp_ctx, ctx), &r, n, PTR(idp_result, five), PTR(idp_result, three), (int64_t *)ADDR(metrics), (int64_t *)ADDR(nAligned)));
}
JNIEXPORT void JNICALL GPM(matchBatch16Device)(JNIEnv *env, jclass c, jlong ctx, jlong seq4, jlong seqOff, jlong len, jlong refIdx, jlong uStart, jlong uEnd,
                                               jlong flags, jint fixedLen, jint n, jlong five, jlong three, jobject metrics, jobject nAligned) {
    const idp_reads r = reads_of_device(seq4, seqOff, len, refIdx, uStart, uEnd, flags, fixedLen);
    check(env, idp_match_batch16_device(PTR(idp_ctx, ctx), &r, n, PTR(idp_result16, five), PTR(idp_result16, three), (int64_t *)ADDR(metrics), (int64_t *)ADDR(nAligned)));
}
JNIEXPORT void JNICALL GPM(result16Unpack)(JNIEnv *env, jclass c, jobject in, jlong n, jobject out) {
    idp_result16_unpack((const idp_result16 *)ADDR(in), n, (idp_result *)ADDR(out));
}

/* ---- result helpers ---- */
JNIEXPORT jdouble JNICALL GPM(resultScore)(JNIEnv *env, jclass c, jobject results, jint i) { return idp_result_score((const idp_result *)ADDR(results) + i); }
JNIEXPORT jdouble JNICALL GPM(resultSecond)(JNIEnv *env, jclass c, jobject results, jint i) { return idp_result_second((const idp_result *)ADDR(results) + i); }
JNIEXPORT jint JNICALL GPM(metricBin)(JNIEnv *env, jclass c, jint templateType, jint isFr, jint r1, jint r2) { return idp_metric_bin(templateType, isFr, r1, r2); }
JNIEXPORT void JNICALL GPM(summaryMetrics)(JNIEnv *env, jclass c, jobject metrics160, jobject out17) { idp_summary_metrics((const int64_t *)ADDR(metrics160), (int64_t *)ADDR(out17)); }
/* PrimerMatch.info(rec) (PMa:78-88) of results[i] */
JNIEXPORT jstring JNICALL GPM(formatInfo)(JNIEnv *env, jclass c, jlong panel, jobject results, jint i, jint readFlags) {
    char buf[1024];
    if (idp_format_info(PTR(const idp_panel, panel), (const idp_result *)ADDR(results) + i, (uint16_t)readFlags, buf, (int32_t)sizeof buf) < 0) { throw_idp(env); return NULL; }
    return (*env)->NewStringUTF(env, buf);
}
JNIEXPORT jstring JNICALL GPM(headerComments)(JNIEnv *env, jclass c) {
    char buf[2048];
    idp_header_comments(buf, (int32_t)sizeof buf);
    return (*env)->NewStringUTF(env, buf);
}

/* ---- downstream classification (scripts/post_process_bam.py:56-78) ---- */
JNIEXPORT jint JNICALL GPM(classifyPrimerPair)(JNIEnv *env, jclass c, jlong panel, jint f5, jint r5, jint minInsert, jint maxInsert) {
    return idp_classify_primer_pair(PTR(const idp_panel, panel), f5, r5, minInsert, maxInsert);
}
JNIEXPORT void JNICALL GPM(classifyTemplates)(JNIEnv *env, jclass c, jlong panel, jobject flags, jobject five, jint n, jint minInsert, jint maxInsert,
                                              jobject clsPerRead, jobject counts6) {
    check(env, idp_classify_templates(PTR(const idp_panel, panel), (const uint16_t *)ADDR(flags), (const idp_result *)ADDR(five), n, minInsert, maxInsert,
                                      (uint8_t *)ADDR(clsPerRead), (int64_t *)ADDR(counts6)));
}
JNIEXPORT void JNICALL GPM(classifyBatchDevice)(JNIEnv *env, jclass c, jlong ctx, jlong flags, jlong five, jint compact, jint n, jint minInsert, jint maxInsert,
                                                jlong clsPerRead, jobject counts6) {
    check(env, idp_classify_batch_device(PTR(idp_ctx, ctx), PTR(const uint16_t, flags), PTR(const void, five), compact, n, minInsert, maxInsert,
                                         PTR(uint8_t, clsPerRead), (int64_t *)ADDR(counts6)));
}

/* ---- host ingest / egress: BAM records <-> batches (IP:288-289, 522-569) ---- */
/* out2: int32 n_reads, then uint64 seq_bytes (16 bytes, 8-byte aligned) */
JNIEXPORT void JNICALL GPM(bamScan)(JNIEnv *env, jclass c, jobject bam, jlong nBytes, jobject out2) {
    uint8_t *o = (uint8_t *)ADDR(out2);
    check(env, idp_bam_scan((const uint8_t *)ADDR(bam), (size_t)nBytes, (int32_t *)o, (uint64_t *)(o + 8)));
}
JNIEXPORT void JNICALL GPM(bamToBatch)(JNIEnv *env, jclass c, jobject bam, jlong nBytes, jobject seq4, jobject seqOff, jobject len, jobject refIdx,
                                       jobject uStart, jobject uEnd, jobject flags, jobject recIndex) {
    check(env, idp_bam_to_batch((const uint8_t *)ADDR(bam), (size_t)nBytes, (uint8_t *)ADDR(seq4), (uint64_t *)ADDR(seqOff), (int32_t *)ADDR(len),
                                (int32_t *)ADDR(refIdx), (int32_t *)ADDR(uStart), (int32_t *)ADDR(uEnd), (uint16_t *)ADDR(flags), (int32_t *)ADDR(recIndex)));
}
JNIEXPORT void JNICALL GPM(bamToBatchWindow)(JNIEnv *env, jclass c, jlong panel, jobject bam, jlong nBytes, jobject seq4, jobject seqOff, jobject len,
                                             jobject refIdx, jobject uStart, jobject uEnd, jobject flags, jobject recIndex) {
    check(env, idp_bam_to_batch_window(PTR(const idp_panel, panel), (const uint8_t *)ADDR(bam), (size_t)nBytes, (uint8_t *)ADDR(seq4), (uint64_t *)ADDR(seqOff),
                                       (int32_t *)ADDR(len), (int32_t *)ADDR(refIdx), (int32_t *)ADDR(uStart), (int32_t *)ADDR(uEnd), (uint16_t *)ADDR(flags),
                                       (int32_t *)ADDR(recIndex)));
}
/* returns the number of bytes the annotated records take; out == null only sizes */
JNIEXPORT jlong JNICALL GPM(bamAnnotate)(JNIEnv *env, jclass c, jlong panel, jobject bam, jlong nBytes, jobject five, jobject three, jint n, jobject out, jlong outCap) {
    size_t need = 0;
    check(env, idp_bam_annotate(PTR(const idp_panel, panel), (const uint8_t *)ADDR(bam), (size_t)nBytes, (const idp_result *)ADDR(five),
                                (const idp_result *)ADDR(three), n, (uint8_t *)ADDR(out), (size_t)outCap, &need));
    return (jlong)need;
}
JNIEXPORT void JNICALL GPM(packBases)(JNIEnv *env, jclass c, jobject ascii, jint n, jobject seq4) { idp_pack_bases((const char *)ADDR(ascii), n, (uint8_t *)ADDR(seq4)); }
JNIEXPORT void JNICALL GPM(unpackBases)(JNIEnv *env, jclass c, jobject seq4, jint n, jobject ascii) { idp_unpack_bases((const uint8_t *)ADDR(seq4), n, (char *)ADDR(ascii)); }

/* ---- the AsyncAligner seam (AA:37-54) ---- */
typedef struct { idp_panel *panel; idp_ctx *ctx; } aligner_handle;
/* idp_align_batch only reads the scoring parameters of the context's panel: a two-primer placeholder panel carries them */
JNIEXPORT jlong JNICALL ACA(alignerCreate)(JNIEnv *env, jclass c, jint matchScore, jint mismatchScore, jint gapOpen, jint gapExtend, jint device) {
    idp_params prm;
    memset(&prm, 0, sizeof prm);
    prm.slop = 5; prm.three_prime = -1; prm.match_score = matchScore; prm.mismatch_score = mismatchScore; prm.gap_open = gapOpen; prm.gap_extend = gapExtend;
    prm.k_ungapped = -1; prm.k_gapped = -1; prm.max_mismatch_rate = 0.05; prm.min_alignment_score_rate = 0.01;
    idp_primer pr[2];
    memset(pr, 0, sizeof pr);
    pr[0].pair_id = "p"; pr[0].primer_id = "p+"; pr[0].sequence = "A"; pr[0].ref_name = ""; pr[0].ref_idx = -1; pr[0].positive_strand = 1;
    pr[1].pair_id = "p"; pr[1].primer_id = "p-"; pr[1].sequence = "T"; pr[1].ref_name = ""; pr[1].ref_idx = -1; pr[1].positive_strand = 0;
    aligner_handle *h = (aligner_handle *)calloc(1, sizeof *h);
    if (!h || idp_panel_create(pr, 2, &prm, &h->panel) != IDP_OK || idp_ctx_create(h->panel, device, &h->ctx) != IDP_OK) {
        throw_idp(env);
        if (h) { idp_panel_destroy(h->panel); free(h); }
        return 0;
    }
    return (jlong)(intptr_t)h;
}
JNIEXPORT void JNICALL ACA(alignerDestroy)(JNIEnv *env, jclass c, jlong handle) {
    aligner_handle *h = PTR(aligner_handle, handle);
    if (!h) return;
    idp_ctx_destroy(h->ctx); idp_panel_destroy(h->panel); free(h);
}
JNIEXPORT void JNICALL ACA(alignBatch)(JNIEnv *env, jclass c, jlong handle, jint mode, jobject q, jobject qOff, jobject t, jobject tOff, jint n, jobject score,
                                       jobject ts, jobject te, jobject qs, jobject qe) {
    check(env, idp_align_batch(PTR(aligner_handle, handle)->ctx, mode, (const char *)ADDR(q), (const int64_t *)ADDR(qOff), (const char *)ADDR(t),
                               (const int64_t *)ADDR(tOff), n, (int32_t *)ADDR(score), (int32_t *)ADDR(ts), (int32_t *)ADDR(te), (int32_t *)ADDR(qs), (int32_t *)ADDR(qe)));
}
/* the same on a context the caller already has (GpuPrimerMatcher) */
JNIEXPORT void JNICALL GPM(alignBatch)(JNIEnv *env, jclass c, jlong ctx, jint mode, jobject q, jobject qOff, jobject t, jobject tOff, jint n, jobject score,
                                       jobject ts, jobject te, jobject qs, jobject qe) {
    check(env, idp_align_batch(PTR(idp_ctx, ctx), mode, (const char *)ADDR(q), (const int64_t *)ADDR(qOff), (const char *)ADDR(t), (const int64_t *)ADDR(tOff), n,
                               (int32_t *)ADDR(score), (int32_t *)ADDR(ts), (int32_t *)ADDR(te), (int32_t *)ADDR(qs), (int32_t *)ADDR(qe)));
}
