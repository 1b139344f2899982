/* Minimal stand-in for the JDK's <jni.h>: just the types and the JNIEnv function-table slots that
 * isolation-forest_b200/jvm/ifb200_jni.c uses, with the JDK's signatures, so that the glue can be compiled
 * (-Wall -Werror) and linked against libifb200.so on a box without a JDK (tests/test_jni_glue.py).
 * TEST INFRASTRUCTURE ONLY -- a real build uses the JDK header; slot ORDER here is irrelevant because nothing is run. */
#ifndef IFB_TEST_JNI_STUB_H
#define IFB_TEST_JNI_STUB_H
#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_OK 0

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef float jfloat;
typedef double jdouble;
typedef jint jsize;

struct _jobject;
typedef struct _jobject *jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jobjectArray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jfloatArray;
typedef jarray jdoubleArray;
typedef jarray jbyteArray;
typedef jobject jthrowable;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;

struct JNINativeInterface_ {
    jclass (JNICALL *FindClass)(JNIEnv *env, const char *name);
    jint (JNICALL *ThrowNew)(JNIEnv *env, jclass clazz, const char *msg);
    jboolean (JNICALL *ExceptionCheck)(JNIEnv *env);
    jsize (JNICALL *GetArrayLength)(JNIEnv *env, jarray array);
    jobjectArray (JNICALL *NewObjectArray)(JNIEnv *env, jsize len, jclass clazz, jobject init);
    void (JNICALL *SetObjectArrayElement)(JNIEnv *env, jobjectArray array, jsize index, jobject val);
    jintArray (JNICALL *NewIntArray)(JNIEnv *env, jsize len);
    jlongArray (JNICALL *NewLongArray)(JNIEnv *env, jsize len);
    jfloatArray (JNICALL *NewFloatArray)(JNIEnv *env, jsize len);
    jdoubleArray (JNICALL *NewDoubleArray)(JNIEnv *env, jsize len);
    jbyteArray (JNICALL *NewByteArray)(JNIEnv *env, jsize len);
    void (JNICALL *GetIntArrayRegion)(JNIEnv *env, jintArray array, jsize start, jsize len, jint *buf);
    void (JNICALL *GetLongArrayRegion)(JNIEnv *env, jlongArray array, jsize start, jsize len, jlong *buf);
    void (JNICALL *GetFloatArrayRegion)(JNIEnv *env, jfloatArray array, jsize start, jsize len, jfloat *buf);
    void (JNICALL *GetDoubleArrayRegion)(JNIEnv *env, jdoubleArray array, jsize start, jsize len, jdouble *buf);
    void (JNICALL *GetByteArrayRegion)(JNIEnv *env, jbyteArray array, jsize start, jsize len, jbyte *buf);
    void (JNICALL *SetIntArrayRegion)(JNIEnv *env, jintArray array, jsize start, jsize len, const jint *buf);
    void (JNICALL *SetLongArrayRegion)(JNIEnv *env, jlongArray array, jsize start, jsize len, const jlong *buf);
    void (JNICALL *SetFloatArrayRegion)(JNIEnv *env, jfloatArray array, jsize start, jsize len, const jfloat *buf);
    void (JNICALL *SetDoubleArrayRegion)(JNIEnv *env, jdoubleArray array, jsize start, jsize len, const jdouble *buf);
    void (JNICALL *SetByteArrayRegion)(JNIEnv *env, jbyteArray array, jsize start, jsize len, const jbyte *buf);
    jobject (JNICALL *NewDirectByteBuffer)(JNIEnv *env, void *address, jlong capacity);
    void *(JNICALL *GetDirectBufferAddress)(JNIEnv *env, jobject buf);
    jlong (JNICALL *GetDirectBufferCapacity)(JNIEnv *env, jobject buf);
};
#endif
