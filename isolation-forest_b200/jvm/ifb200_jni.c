/* JNI glue over include/ifb200.h for the Scala object
 * com.linkedin.relevance.isolationforest.gpu.NativeForest (see NativeForest.scala).
 * NOT compiled in this repository's image (no jni.h); kept as the drop-in a maintainer builds against a JDK. */
#include <jni.h>
#include <stdint.h>

#include "ifb200.h"

static void throw_for(JNIEnv *env, int rc) { /* error contract of include/ifb200.h */
    const char *cls = rc == IFB_EINVAL   ? "java/lang/IllegalArgumentException"
                      : rc == IFB_ESTATE ? "java/lang/IllegalStateException"
                                         : "java/lang/RuntimeException";
    (*env)->ThrowNew(env, (*env)->FindClass(env, cls), ifb_last_error());
}

#define JFN(name) Java_com_linkedin_relevance_isolationforest_gpu_NativeForest_00024_##name

JNIEXPORT jobject JNICALL JFN(hostAlloc)(JNIEnv *env, jobject self, jlong bytes) {
    void *p = NULL;
    int rc = ifb_host_alloc((size_t)bytes, &p);
    if (rc != IFB_OK) { throw_for(env, rc); return NULL; }
    return (*env)->NewDirectByteBuffer(env, p, bytes);
}

JNIEXPORT void JNICALL JFN(hostFree)(JNIEnv *env, jobject self, jobject buf) {
    ifb_host_free((*env)->GetDirectBufferAddress(env, buf));
}

JNIEXPORT jlong JNICALL JFN(createStandard)(JNIEnv *env, jobject self, jint device, jintArray nodeOff, jintArray left,
                                            jintArray right, jintArray feature, jdoubleArray threshold,
                                            jlongArray numInstances, jint numSamples, jint totalNumFeatures) {
    const jsize T = (*env)->GetArrayLength(env, nodeOff) - 1;
    jint *no = (*env)->GetPrimitiveArrayCritical(env, nodeOff, NULL);
    jint *l = (*env)->GetPrimitiveArrayCritical(env, left, NULL);
    jint *r = (*env)->GetPrimitiveArrayCritical(env, right, NULL);
    jint *f = (*env)->GetPrimitiveArrayCritical(env, feature, NULL);
    jdouble *t = (*env)->GetPrimitiveArrayCritical(env, threshold, NULL);
    jlong *n = (*env)->GetPrimitiveArrayCritical(env, numInstances, NULL);
    ifb_forest *forest = NULL;
    int rc = ifb_forest_create_standard(device, (int32_t)T, (const int32_t *)no, (const int32_t *)l, (const int32_t *)r,
                                        (const int32_t *)f, t, (const int64_t *)n, numSamples, totalNumFeatures, &forest);
    (*env)->ReleasePrimitiveArrayCritical(env, numInstances, n, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, threshold, t, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, feature, f, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, right, r, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, left, l, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, nodeOff, no, JNI_ABORT);
    if (rc != IFB_OK) { throw_for(env, rc); return 0; }
    return (jlong)(intptr_t)forest;
}

JNIEXPORT jlong JNICALL JFN(createExtended)(JNIEnv *env, jobject self, jint device, jintArray nodeOff, jintArray left,
                                            jintArray right, jlongArray numInstances, jdoubleArray offset,
                                            jlongArray hpOff, jintArray hpIdx, jfloatArray hpW, jint numSamples,
                                            jint totalNumFeatures) {
    const jsize T = (*env)->GetArrayLength(env, nodeOff) - 1;
    jint *no = (*env)->GetPrimitiveArrayCritical(env, nodeOff, NULL);
    jint *l = (*env)->GetPrimitiveArrayCritical(env, left, NULL);
    jint *r = (*env)->GetPrimitiveArrayCritical(env, right, NULL);
    jlong *n = (*env)->GetPrimitiveArrayCritical(env, numInstances, NULL);
    jdouble *o = (*env)->GetPrimitiveArrayCritical(env, offset, NULL);
    jlong *ho = (*env)->GetPrimitiveArrayCritical(env, hpOff, NULL);
    jint *hi = (*env)->GetPrimitiveArrayCritical(env, hpIdx, NULL);
    jfloat *hw = (*env)->GetPrimitiveArrayCritical(env, hpW, NULL);
    ifb_forest *forest = NULL;
    int rc = ifb_forest_create_extended(device, (int32_t)T, (const int32_t *)no, (const int32_t *)l, (const int32_t *)r,
                                        (const int64_t *)n, o, (const int64_t *)ho, (const int32_t *)hi, hw, numSamples,
                                        totalNumFeatures, &forest);
    (*env)->ReleasePrimitiveArrayCritical(env, hpW, hw, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, hpIdx, hi, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, hpOff, ho, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, offset, o, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, numInstances, n, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, right, r, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, left, l, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, nodeOff, no, JNI_ABORT);
    if (rc != IFB_OK) { throw_for(env, rc); return 0; }
    return (jlong)(intptr_t)forest;
}

JNIEXPORT void JNICALL JFN(destroy)(JNIEnv *env, jobject self, jlong handle) {
    ifb_forest_destroy((ifb_forest *)(intptr_t)handle);
}

JNIEXPORT void JNICALL JFN(scoreHost)(JNIEnv *env, jobject self, jlong handle, jobject x, jlong nRows, jint d, jlong ld,
                                      jint layout, jobject scores) {
    const float *px = (const float *)(*env)->GetDirectBufferAddress(env, x);
    double *ps = (double *)(*env)->GetDirectBufferAddress(env, scores);
    int rc = ifb_score_host((const ifb_forest *)(intptr_t)handle, px, nRows, d, ld, layout, ps, NULL, NULL);
    if (rc != IFB_OK) throw_for(env, rc);
}

JNIEXPORT jlong JNICALL JFN(fitHost)(JNIEnv *env, jobject self, jint device, jobject x, jlong nRows, jint d, jlong ld,
                                     jint layout, jint numEstimators, jint numSamples, jint numFeatures,
                                     jboolean bootstrap, jlong randomSeed, jint numPartitions, jint extensionLevel,
                                     jint treeBegin, jint treeEnd) {
    ifb_fit_params p;
    p.num_estimators = numEstimators; p.num_samples = numSamples; p.num_features = numFeatures;
    p.bootstrap = bootstrap ? 1 : 0; p.random_seed = randomSeed; p.num_partitions = numPartitions;
    p.extension_level = extensionLevel; p.tree_begin = treeBegin; p.tree_end = treeEnd;
    ifb_forest *forest = NULL;
    int rc = ifb_fit_host(device, (const float *)(*env)->GetDirectBufferAddress(env, x), nRows, d, ld, layout, &p, &forest);
    if (rc != IFB_OK) { throw_for(env, rc); return 0; }
    return (jlong)(intptr_t)forest;
}

/* exportTables: ifb_forest_get_info for the sizes, NewIntArray/NewDoubleArray/..., GetPrimitiveArrayCritical on each,
 * one ifb_forest_export call, then construct com.linkedin.relevance.isolationforest.gpu.ForestTables. */
