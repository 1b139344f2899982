/* JNI glue over include/ifb200.h for the Scala object
 * com.linkedin.relevance.isolationforest.gpu.NativeForest (see NativeForest.scala).
 *
 * No JDK exists in this repository's image, so the glue is compile- and link-checked against a stand-in <jni.h>
 * (tests/jni_stub/jni.h, tests/test_jni_glue.py: gcc -Wall -Werror, then linked against libifb200.so); a maintainer
 * builds it against the JDK header unchanged:
 *     gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude ifb200_jni.c -L. -lifb200 -o libifb200_jni.so
 *
 * Rules followed here (JNI specification, "critical regions"): Java arrays are COPIED with Get<Type>ArrayRegion into
 * native buffers before any libifb200 call -- forest creation allocates device memory, copies and synchronises, none of
 * which may happen inside GetPrimitiveArrayCritical; every allocation and every JNI return value is checked; a pending
 * Java exception makes the native method return immediately. */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ifb200.h"

static void throw_for(JNIEnv *env, int rc) { /* error contract of include/ifb200.h */
    const char *cls = rc == IFB_EINVAL   ? "java/lang/IllegalArgumentException"
                      : rc == IFB_ESTATE ? "java/lang/IllegalStateException"
                      : rc == IFB_ENOMEM ? "java/lang/OutOfMemoryError"
                                         : "java/lang/RuntimeException";
    jclass c = (*env)->FindClass(env, cls);
    if (c) (*env)->ThrowNew(env, c, ifb_last_error());
}
static void throw_oom(JNIEnv *env, const char *what) {
    jclass c = (*env)->FindClass(env, "java/lang/OutOfMemoryError");
    if (c) (*env)->ThrowNew(env, c, what);
}

/* copies of Java arrays in native memory (NULL + pending exception on failure) */
#define COPY_IN(fn_name, jarr_t, c_t, Region)                                            \
    static c_t *fn_name(JNIEnv *env, jarr_t a, jsize *len_out) {                        \
        jsize n = (*env)->GetArrayLength(env, a);                                       \
        c_t *p = (c_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(c_t));                  \
        if (!p) { throw_oom(env, "ifb200_jni: native copy of a node table"); return NULL; } \
        (*env)->Region(env, a, 0, n, p);                                                \
        if ((*env)->ExceptionCheck(env)) { free(p); return NULL; }                      \
        if (len_out) *len_out = n;                                                      \
        return p;                                                                       \
    }
COPY_IN(copy_ints, jintArray, jint, GetIntArrayRegion)
COPY_IN(copy_longs, jlongArray, jlong, GetLongArrayRegion)
COPY_IN(copy_floats, jfloatArray, jfloat, GetFloatArrayRegion)
COPY_IN(copy_doubles, jdoubleArray, jdouble, GetDoubleArrayRegion)

#define JFN(name) Java_com_linkedin_relevance_isolationforest_gpu_NativeForest_00024_##name

JNIEXPORT jobject JNICALL JFN(hostAlloc)(JNIEnv *env, jobject self, jlong bytes) {
    (void)self;
    void *p = NULL;
    int rc = ifb_host_alloc((size_t)bytes, &p);
    if (rc != IFB_OK) { throw_for(env, rc); return NULL; }
    jobject buf = (*env)->NewDirectByteBuffer(env, p, bytes);
    if (!buf) ifb_host_free(p);
    return buf;
}

JNIEXPORT void JNICALL JFN(hostFree)(JNIEnv *env, jobject self, jobject buf) {
    (void)self;
    void *p = (*env)->GetDirectBufferAddress(env, buf);
    if (p) ifb_host_free(p);
}

JNIEXPORT jlong JNICALL JFN(createStandard)(JNIEnv *env, jobject self, jint device, jintArray nodeOff, jintArray left,
                                            jintArray right, jintArray feature, jdoubleArray threshold,
                                            jlongArray numInstances, jint numSamples, jint totalNumFeatures) {
    (void)self;
    jsize nt = 0;
    jint *no = NULL, *l = NULL, *r = NULL, *f = NULL;
    jdouble *t = NULL;
    jlong *n = NULL;
    ifb_forest *forest = NULL;
    int rc = IFB_OK;
    if ((no = copy_ints(env, nodeOff, &nt)) && (l = copy_ints(env, left, NULL)) && (r = copy_ints(env, right, NULL)) &&
        (f = copy_ints(env, feature, NULL)) && (t = copy_doubles(env, threshold, NULL)) &&
        (n = copy_longs(env, numInstances, NULL))) {
        rc = ifb_forest_create_standard(device, (int32_t)(nt - 1), (const int32_t *)no, (const int32_t *)l,
                                        (const int32_t *)r, (const int32_t *)f, t, (const int64_t *)n, numSamples,
                                        totalNumFeatures, &forest);
        if (rc != IFB_OK) throw_for(env, rc);
    }
    free(no); free(l); free(r); free(f); free(t); free(n);
    return (jlong)(intptr_t)forest;
}

JNIEXPORT jlong JNICALL JFN(createExtended)(JNIEnv *env, jobject self, jint device, jintArray nodeOff, jintArray left,
                                            jintArray right, jlongArray numInstances, jdoubleArray offset,
                                            jlongArray hpOff, jintArray hpIdx, jfloatArray hpW, jint numSamples,
                                            jint totalNumFeatures) {
    (void)self;
    jsize nt = 0;
    jint *no = NULL, *l = NULL, *r = NULL, *hi = NULL;
    jlong *n = NULL, *ho = NULL;
    jdouble *o = NULL;
    jfloat *hw = NULL;
    ifb_forest *forest = NULL;
    int rc = IFB_OK;
    if ((no = copy_ints(env, nodeOff, &nt)) && (l = copy_ints(env, left, NULL)) && (r = copy_ints(env, right, NULL)) &&
        (n = copy_longs(env, numInstances, NULL)) && (o = copy_doubles(env, offset, NULL)) &&
        (ho = copy_longs(env, hpOff, NULL)) && (hi = copy_ints(env, hpIdx, NULL)) && (hw = copy_floats(env, hpW, NULL))) {
        rc = ifb_forest_create_extended(device, (int32_t)(nt - 1), (const int32_t *)no, (const int32_t *)l,
                                        (const int32_t *)r, (const int64_t *)n, o, (const int64_t *)ho,
                                        (const int32_t *)hi, hw, numSamples, totalNumFeatures, &forest);
        if (rc != IFB_OK) throw_for(env, rc);
    }
    free(no); free(l); free(r); free(n); free(o); free(ho); free(hi); free(hw);
    return (jlong)(intptr_t)forest;
}

JNIEXPORT void JNICALL JFN(destroy)(JNIEnv *env, jobject self, jlong handle) {
    (void)env; (void)self;
    ifb_forest_destroy((ifb_forest *)(intptr_t)handle);
}

/* {extended, device, numTrees, numSamples, totalNumFeatures, maxFeatureIndex, maxDepth, maxNnz, numNodes, numHpEntries} */
JNIEXPORT jlongArray JNICALL JFN(info)(JNIEnv *env, jobject self, jlong handle) {
    (void)self;
    ifb_forest_info i;
    int rc = ifb_forest_get_info((const ifb_forest *)(intptr_t)handle, &i);
    if (rc != IFB_OK) { throw_for(env, rc); return NULL; }
    const jlong v[10] = {i.extended, i.device, i.num_trees, i.num_samples, i.total_num_features, i.max_feature_index,
                         i.max_depth, i.max_nnz, i.num_nodes, i.num_hp_entries};
    jlongArray out = (*env)->NewLongArray(env, 10);
    if (out) (*env)->SetLongArrayRegion(env, out, 0, 10, v);
    return out;
}

/* exportTables: the node tables in the persisted layout (what MLWriter.save writes and what
 * IsolationForestModelReadWrite.scala:179-205 buildTreeFromNodes consumes), as an Object[]:
 *   standard: {int[] nodeOff, int[] left, int[] right, int[] feature, double[] threshold, long[] numInstances}
 *   extended: {int[] nodeOff, int[] left, int[] right, long[] numInstances, double[] offset, long[] hpOff, int[] hpIdx, float[] hpW} */
JNIEXPORT jobjectArray JNICALL JFN(exportTables)(JNIEnv *env, jobject self, jlong handle) {
    (void)self;
    const ifb_forest *f = (const ifb_forest *)(intptr_t)handle;
    ifb_forest_info i;
    int rc = ifb_forest_get_info(f, &i);
    if (rc != IFB_OK) { throw_for(env, rc); return NULL; }
    const size_t T = (size_t)i.num_trees, N = (size_t)i.num_nodes, H = (size_t)i.num_hp_entries;
    if (N > 0x7fffffffu || H > 0x7fffffffu) { throw_oom(env, "ifb200_jni: forest too large for Java arrays"); return NULL; }
    int32_t *no = (int32_t *)malloc((T + 1) * 4), *l = (int32_t *)malloc((N + 1) * 4), *r = (int32_t *)malloc((N + 1) * 4);
    int64_t *n = (int64_t *)malloc((N + 1) * 8);
    int32_t *feat = NULL, *hi = NULL;
    double *thr = NULL, *off = NULL;
    int64_t *ho = NULL;
    float *hw = NULL;
    jobjectArray out = NULL;
    int ok = no && l && r && n;
    if (i.extended) {
        off = (double *)malloc((N + 1) * 8); ho = (int64_t *)malloc((N + 2) * 8);
        hi = (int32_t *)malloc((H + 1) * 4); hw = (float *)malloc((H + 1) * 4);
        ok = ok && off && ho && hi && hw;
    } else {
        feat = (int32_t *)malloc((N + 1) * 4); thr = (double *)malloc((N + 1) * 8);
        ok = ok && feat && thr;
    }
    if (!ok) {
        throw_oom(env, "ifb200_jni: exportTables");
    } else if ((rc = ifb_forest_export(f, no, l, r, feat, thr, n, off, ho, hi, hw)) != IFB_OK) {
        throw_for(env, rc);
    } else {
        jclass obj = (*env)->FindClass(env, "java/lang/Object");
        out = obj ? (*env)->NewObjectArray(env, i.extended ? 8 : 6, obj, NULL) : NULL;
        jsize k = 0;
#define PUT(New, Set, jt, ptr, len)                                                  \
        if (out && !(*env)->ExceptionCheck(env)) {                                   \
            jarray a = (*env)->New(env, (jsize)(len));                               \
            if (a) { (*env)->Set(env, a, 0, (jsize)(len), (const jt *)(ptr)); (*env)->SetObjectArrayElement(env, out, k, a); } \
            k++;                                                                     \
        }
        PUT(NewIntArray, SetIntArrayRegion, jint, no, T + 1)
        PUT(NewIntArray, SetIntArrayRegion, jint, l, N)
        PUT(NewIntArray, SetIntArrayRegion, jint, r, N)
        if (i.extended) {
            PUT(NewLongArray, SetLongArrayRegion, jlong, n, N)
            PUT(NewDoubleArray, SetDoubleArrayRegion, jdouble, off, N)
            PUT(NewLongArray, SetLongArrayRegion, jlong, ho, N + 1)
            PUT(NewIntArray, SetIntArrayRegion, jint, hi, H)
            PUT(NewFloatArray, SetFloatArrayRegion, jfloat, hw, H)
        } else {
            PUT(NewIntArray, SetIntArrayRegion, jint, feat, N)
            PUT(NewDoubleArray, SetDoubleArrayRegion, jdouble, thr, N)
            PUT(NewLongArray, SetLongArrayRegion, jlong, n, N)
        }
#undef PUT
        if ((*env)->ExceptionCheck(env)) out = NULL;
    }
    free(no); free(l); free(r); free(n); free(feat); free(thr); free(off); free(ho); free(hi); free(hw);
    return out;
}

JNIEXPORT void JNICALL JFN(scoreHost)(JNIEnv *env, jobject self, jlong handle, jobject x, jlong nRows, jint d, jlong ld,
                                      jint layout, jobject scores) {
    (void)self;
    const float *px = (const float *)(*env)->GetDirectBufferAddress(env, x);
    double *ps = (double *)(*env)->GetDirectBufferAddress(env, scores);
    if (!px || !ps || (*env)->GetDirectBufferCapacity(env, scores) < nRows * 8) {
        jclass c = (*env)->FindClass(env, "java/lang/IllegalArgumentException");
        if (c) (*env)->ThrowNew(env, c, "scoreHost needs direct ByteBuffers (x: f32 matrix, scores: >= 8*nRows bytes)");
        return;
    }
    int rc = ifb_score_host((const ifb_forest *)(intptr_t)handle, px, nRows, d, ld, layout, ps, NULL, NULL);
    if (rc != IFB_OK) throw_for(env, rc);
}

JNIEXPORT jlong JNICALL JFN(fitHost)(JNIEnv *env, jobject self, jint device, jobject x, jlong nRows, jint d, jlong ld,
                                     jint layout, jint numEstimators, jint numSamples, jint numFeatures,
                                     jboolean bootstrap, jlong randomSeed, jint numPartitions, jint extensionLevel,
                                     jint treeBegin, jint treeEnd) {
    (void)self;
    ifb_fit_params p;
    memset(&p, 0, sizeof p);
    p.num_estimators = numEstimators; p.num_samples = numSamples; p.num_features = numFeatures;
    p.bootstrap = bootstrap ? 1 : 0; p.random_seed = randomSeed; p.num_partitions = numPartitions;
    p.extension_level = extensionLevel; p.tree_begin = treeBegin; p.tree_end = treeEnd;
    const float *px = (const float *)(*env)->GetDirectBufferAddress(env, x);
    if (!px) {
        jclass c = (*env)->FindClass(env, "java/lang/IllegalArgumentException");
        if (c) (*env)->ThrowNew(env, c, "fitHost needs a direct ByteBuffer");
        return 0;
    }
    ifb_forest *forest = NULL;
    int rc = ifb_fit_host(device, px, nRows, d, ld, layout, &p, &forest);
    if (rc != IFB_OK) { throw_for(env, rc); return 0; }
    return (jlong)(intptr_t)forest;
}

/* exact order statistic of a score vector on the device (threshold step, SharedTrainLogic.scala:191-198): scores is a
 * direct buffer of host doubles; returns {value, observedFractionGe} */
JNIEXPORT jdoubleArray JNICALL JFN(quantileHost)(JNIEnv *env, jobject self, jint device, jobject scores, jlong n, jdouble q) {
    (void)self;
    const double *hs = (const double *)(*env)->GetDirectBufferAddress(env, scores);
    void *ds = NULL;
    double v[2] = {0.0, 0.0};
    int rc = hs ? ifb_device_alloc(device, (size_t)n * 8, &ds) : IFB_EINVAL;
    if (rc == IFB_OK) rc = ifb_copy_to_device(device, ds, hs, (size_t)n * 8);
    if (rc == IFB_OK) rc = ifb_quantile_device(device, (const double *)ds, n, q, &v[0], &v[1], NULL);
    if (ds) ifb_device_free(device, ds);
    if (rc != IFB_OK) { throw_for(env, rc); return NULL; }
    jdoubleArray out = (*env)->NewDoubleArray(env, 2);
    if (out) (*env)->SetDoubleArrayRegion(env, out, 0, 2, v);
    return out;
}

/* multi-GPU: the 128-byte NCCL id one executor creates and Spark broadcasts; one communicator per executor GPU */
JNIEXPORT jbyteArray JNICALL JFN(commUniqueId)(JNIEnv *env, jobject self) {
    (void)self;
    jbyte id[128];
    int rc = ifb_comm_unique_id(id);
    if (rc != IFB_OK) { throw_for(env, rc); return NULL; }
    jbyteArray out = (*env)->NewByteArray(env, 128);
    if (out) (*env)->SetByteArrayRegion(env, out, 0, 128, id);
    return out;
}

JNIEXPORT jlong JNICALL JFN(commInit)(JNIEnv *env, jobject self, jint device, jint world, jint rank, jbyteArray id) {
    (void)self;
    jbyte buf[128];
    if ((*env)->GetArrayLength(env, id) != 128) {
        jclass c = (*env)->FindClass(env, "java/lang/IllegalArgumentException");
        if (c) (*env)->ThrowNew(env, c, "the communicator id must be 128 bytes");
        return 0;
    }
    (*env)->GetByteArrayRegion(env, id, 0, 128, buf);
    if ((*env)->ExceptionCheck(env)) return 0;
    ifb_comm *c = NULL;
    int rc = ifb_comm_init(device, world, rank, buf, &c);
    if (rc != IFB_OK) { throw_for(env, rc); return 0; }
    return (jlong)(intptr_t)c;
}

JNIEXPORT void JNICALL JFN(commDestroy)(JNIEnv *env, jobject self, jlong comm) {
    (void)env; (void)self;
    ifb_comm_destroy((ifb_comm *)(intptr_t)comm);
}

/* Tree-sharded transform of a host batch: every executor passes the SAME rows and its own forest shard; mode 0 returns
 * all scores on every rank.  Host staging through ifb_device_alloc / ifb_copy_*; the collective runs inside the library. */
JNIEXPORT void JNICALL JFN(scoreShardedHost)(JNIEnv *env, jobject self, jlong handle, jlong comm, jint device, jobject x,
                                             jlong nRows, jint d, jint totalNumTrees, jobject scores) {
    (void)self;
    const float *px = (const float *)(*env)->GetDirectBufferAddress(env, x);
    double *ps = (double *)(*env)->GetDirectBufferAddress(env, scores);
    void *dx = NULL, *dsc = NULL;
    int64_t b = 0, e = 0;
    int rc = (px && ps) ? ifb_device_alloc(device, (size_t)nRows * (size_t)d * 4, &dx) : IFB_EINVAL;
    if (rc == IFB_OK) rc = ifb_device_alloc(device, (size_t)nRows * 8, &dsc);
    if (rc == IFB_OK) rc = ifb_copy_to_device(device, dx, px, (size_t)nRows * (size_t)d * 4);
    if (rc == IFB_OK)
        rc = ifb_score_sharded((const ifb_forest *)(intptr_t)handle, (ifb_comm *)(intptr_t)comm, (const float *)dx, nRows, d,
                               d, IFB_ROW_MAJOR, totalNumTrees, IFB_SHARD_ALLREDUCE, (double *)dsc, &b, &e, NULL);
    if (rc == IFB_OK) rc = ifb_copy_to_host(device, ps, dsc, (size_t)nRows * 8);
    if (dx) ifb_device_free(device, dx);
    if (dsc) ifb_device_free(device, dsc);
    if (rc != IFB_OK) throw_for(env, rc);
}
