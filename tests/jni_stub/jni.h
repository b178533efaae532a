/* A stand-in for the JDK's jni.h with just the names bindings/jni/idprimer_jni.c uses, so that the shim can be type-checked
 * against include/idprimer.h on a machine without a JDK (tests/test_bindings.py).  Test infrastructure only. */
#ifndef IDP_TEST_JNI_STUB_H
#define IDP_TEST_JNI_STUB_H
#include <stdint.h>
typedef int32_t jint;
typedef int64_t jlong;
typedef double jdouble;
typedef void *jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef unsigned char jboolean;
struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
    jclass (*FindClass)(JNIEnv *, const char *);
    jint (*ThrowNew)(JNIEnv *, jclass, const char *);
    void *(*GetDirectBufferAddress)(JNIEnv *, jobject);
    jlong (*GetDirectBufferCapacity)(JNIEnv *, jobject);
    jobject (*NewDirectByteBuffer)(JNIEnv *, void *, jlong);
    jstring (*NewStringUTF)(JNIEnv *, const char *);
    const char *(*GetStringUTFChars)(JNIEnv *, jstring, jboolean *);
    void (*ReleaseStringUTFChars)(JNIEnv *, jstring, const char *);
};
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#endif
