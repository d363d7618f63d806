/* COMPILE-CHECK STUB ONLY — not the JDK's jni.h.  The build image has no JDK, so `make -C jni check` compiles
 * dpf_jni.c against this minimal declaration of the handful of JNI entries the shim uses, to catch syntax and
 * prototype errors.  A real build uses $JAVA_HOME/include/jni.h (the function table order below is NOT the JVM's). */
#ifndef DPF_JNI_STUB_H
#define DPF_JNI_STUB_H
#include <stdint.h>
typedef int32_t jint;
typedef int64_t jlong;
typedef double jdouble;
typedef uint8_t jboolean;
typedef int8_t jbyte;
typedef jint jsize;
typedef void* jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jdoubleArray;
typedef jarray jbyteArray;
typedef jobject jstring;
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
    jsize (*GetArrayLength)(JNIEnv*, jarray);
    void* (*GetPrimitiveArrayCritical)(JNIEnv*, jarray, jboolean*);
    void (*ReleasePrimitiveArrayCritical)(JNIEnv*, jarray, void*, jint);
    jintArray (*NewIntArray)(JNIEnv*, jsize);
    jlongArray (*NewLongArray)(JNIEnv*, jsize);
    void (*SetIntArrayRegion)(JNIEnv*, jintArray, jsize, jsize, const jint*);
    void (*SetLongArrayRegion)(JNIEnv*, jlongArray, jsize, jsize, const jlong*);
    jclass (*FindClass)(JNIEnv*, const char*);
    jint (*ThrowNew)(JNIEnv*, jclass, const char*);
    void* (*GetDirectBufferAddress)(JNIEnv*, jobject);
    jbyteArray (*NewByteArray)(JNIEnv*, jsize);
    void (*SetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, const jbyte*);
    void (*GetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, jbyte*);
};
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#endif
