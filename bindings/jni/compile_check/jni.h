/*
 * NOT the JDK's <jni.h>.  A declaration-only stand-in with just the names bindings/jni/chaos_jni.c uses, so that the shim
 * can be compile-checked (gcc -fsyntax-only) in an image without a JDK.  Member ORDER and therefore the ABI are not those
 * of the real header: never link or run anything built against this file.
 */
#ifndef CHAOS_JNI_COMPILE_CHECK_H
#define CHAOS_JNI_COMPILE_CHECK_H
#include <stdint.h>
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
typedef int32_t jint;
typedef int64_t jlong;
typedef uint8_t jboolean;
typedef jint jsize;
typedef struct chaos_jobject_ *jobject;
typedef jobject jclass, jstring, jarray, jintArray, jobjectArray;
struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
    jclass (*FindClass)(JNIEnv *, const char *);
    jint (*ThrowNew)(JNIEnv *, jclass, const char *);
    const char *(*GetStringUTFChars)(JNIEnv *, jstring, jboolean *);
    void (*ReleaseStringUTFChars)(JNIEnv *, jstring, const char *);
    jstring (*NewStringUTF)(JNIEnv *, const char *);
    jsize (*GetArrayLength)(JNIEnv *, jarray);
    jint *(*GetIntArrayElements)(JNIEnv *, jintArray, jboolean *);
    void (*ReleaseIntArrayElements)(JNIEnv *, jintArray, jint *, jint);
    jobjectArray (*NewObjectArray)(JNIEnv *, jsize, jclass, jobject);
    void (*SetObjectArrayElement)(JNIEnv *, jobjectArray, jsize, jobject);
    void *(*GetDirectBufferAddress)(JNIEnv *, jobject);
    jlong (*GetDirectBufferCapacity)(JNIEnv *, jobject);
    jobject (*NewDirectByteBuffer)(JNIEnv *, void *, jlong);
};
#endif
