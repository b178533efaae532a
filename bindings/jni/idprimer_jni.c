/*
 * JNI shim: 1:1 wrapper of include/idprimer.h for com.fulcrumgenomics.identifyprimers.GpuPrimerMatcher.
 * NOT compiled in this image (no JDK, no jni.h).  Build on a machine with a JDK:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude bindings/jni/idprimer_jni.c \
 *       -Lfg_idprimer_b200 -lidprimer_b200 -o libidprimer_jni.so
 * All bulk data crosses as direct ByteBuffers (no copies); failures become RuntimeException(idp_last_error()).
 */
#include <jni.h>
#include <stdint.h>
#include "idprimer.h"

static void throw_idp(JNIEnv *env) {
    jclass cls = (*env)->FindClass(env, "java/lang/RuntimeException");
    if (cls) (*env)->ThrowNew(env, cls, idp_last_error());
}
#define ADDR(buf) ((buf) ? (*env)->GetDirectBufferAddress(env, (buf)) : NULL)

/* long panelCreate(ByteBuffer primers /* idp_primer[n], strings pinned by the caller *\/, int n, ByteBuffer params) */
JNIEXPORT jlong JNICALL Java_com_fulcrumgenomics_identifyprimers_GpuPrimerMatcher_00024Native_panelCreate(
        JNIEnv *env, jclass cls, jobject primers, jint n, jobject params) {
    idp_panel *p = NULL;
    if (idp_panel_create((const idp_primer *)ADDR(primers), n, (const idp_params *)ADDR(params), &p) != IDP_OK) { throw_idp(env); return 0; }
    return (jlong)(intptr_t)p;
}
JNIEXPORT void JNICALL Java_com_fulcrumgenomics_identifyprimers_GpuPrimerMatcher_00024Native_panelDestroy(JNIEnv *env, jclass cls, jlong h) {
    idp_panel_destroy((idp_panel *)(intptr_t)h);
}
JNIEXPORT jint JNICALL Java_com_fulcrumgenomics_identifyprimers_GpuPrimerMatcher_00024Native_panelWindow(JNIEnv *env, jclass cls, jlong h) {
    return idp_panel_window((const idp_panel *)(intptr_t)h);
}
JNIEXPORT jlong JNICALL Java_com_fulcrumgenomics_identifyprimers_GpuPrimerMatcher_00024Native_ctxCreate(JNIEnv *env, jclass cls, jlong panel, jint device) {
    idp_ctx *c = NULL;
    if (idp_ctx_create((const idp_panel *)(intptr_t)panel, device, &c) != IDP_OK) { throw_idp(env); return 0; }
    return (jlong)(intptr_t)c;
}
JNIEXPORT void JNICALL Java_com_fulcrumgenomics_identifyprimers_GpuPrimerMatcher_00024Native_ctxDestroy(JNIEnv *env, jclass cls, jlong h) {
    idp_ctx_destroy((idp_ctx *)(intptr_t)h);
}
/* void matchBatch(long ctx, ByteBuffer seq4, ByteBuffer seqOff, ByteBuffer len, ByteBuffer refIdx, ByteBuffer uStart, ByteBuffer uEnd,
 *                 ByteBuffer flags, int fixedLen, int n, ByteBuffer five, ByteBuffer three, ByteBuffer metrics160, ByteBuffer nAligned) */
JNIEXPORT void JNICALL Java_com_fulcrumgenomics_identifyprimers_GpuPrimerMatcher_00024Native_matchBatch(
        JNIEnv *env, jclass cls, jlong ctx, jobject seq4, jobject seqOff, jobject len, jobject refIdx, jobject uStart, jobject uEnd,
        jobject flags, jint fixedLen, jint n, jobject five, jobject three, jobject metrics, jobject nAligned) {
    idp_reads r;
    r.seq4 = (const uint8_t *)ADDR(seq4); r.seq_off = (const uint64_t *)ADDR(seqOff); r.len = (const int32_t *)ADDR(len);
    r.ref_idx = (const int32_t *)ADDR(refIdx); r.unclipped_start = (const int32_t *)ADDR(uStart); r.unclipped_end = (const int32_t *)ADDR(uEnd);
    r.flags = (const uint16_t *)ADDR(flags); r.fixed_len = fixedLen;
    if (idp_match_batch((idp_ctx *)(intptr_t)ctx, &r, n, (idp_result *)ADDR(five), (idp_result *)ADDR(three),
                        (int64_t *)ADDR(metrics), (int64_t *)ADDR(nAligned)) != IDP_OK) throw_idp(env);
}
/* void alignBatch(long ctx, int mode, ByteBuffer queries, ByteBuffer qOff, ByteBuffer targets, ByteBuffer tOff, int n,
 *                 ByteBuffer score, ByteBuffer targetStart, ByteBuffer targetEnd)  -- the AsyncAligner seam */
JNIEXPORT void JNICALL Java_com_fulcrumgenomics_alignment_AsyncCudaAligner_00024Native_alignBatch(
        JNIEnv *env, jclass cls, jlong ctx, jint mode, jobject q, jobject qOff, jobject t, jobject tOff, jint n, jobject score, jobject ts, jobject te) {
    if (idp_align_batch((idp_ctx *)(intptr_t)ctx, mode, (const char *)ADDR(q), (const int64_t *)ADDR(qOff), (const char *)ADDR(t),
                        (const int64_t *)ADDR(tOff), n, (int32_t *)ADDR(score), (int32_t *)ADDR(ts), (int32_t *)ADDR(te)) != IDP_OK) throw_idp(env);
}
