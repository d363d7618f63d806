/*
 * dpf_jni.c — thin JNI shim between the reference's Scala facades and libdpf_b200 (include/dpf.h).
 *
 * JVM side: `object NativeDPF` in package mclab.deploy declares these natives; the facades
 * DensevectorRDFInit / SparsevectorRDFInit keep their public signatures and call them instead of
 * RandomDrawTreeMap.put / getSimilarWithStepWiseFaster.  Arrays are flattened on the JVM side
 * (Array[DenseVector] -> double[n*d]); large data (n*d >= 2^31) goes through direct ByteBuffers.
 * Errors: a non-zero status becomes a java.lang.RuntimeException carrying dpf_last_error().
 *
 * No JDK exists in the build image: `make -C jni check` compiles this file against jni/jni_stub/jni.h only to keep
 * it syntactically honest; a real build uses -I$JAVA_HOME/include -I$JAVA_HOME/include/linux.
 */
#include <jni.h>
#include <stdlib.h>

#include "../include/dpf.h"

static void throw_dpf(JNIEnv* env, dpf_handle h, int rc) {
    jclass cls = (*env)->FindClass(env, "java/lang/RuntimeException");
    if (cls) (*env)->ThrowNew(env, cls, h ? dpf_last_error(h) : dpf_strerror(rc));
}

/* long create(int[] cfg)  — cfg = {device,d,L,k,pb,bucketBits,dirNodeSize,bufferOverflow,family,typeOfIndex,selfExclude,rank,world}
 * replaces DensevectorRDFInit.initializeRDFHashMap (DensevectorRDFInit.scala:50-118) */
JNIEXPORT jlong JNICALL Java_mclab_deploy_NativeDPF_create(JNIEnv* env, jclass cls, jintArray jcfg) {
    (void)cls;
    jint* c = (jint*)(*env)->GetPrimitiveArrayCritical(env, jcfg, 0);
    dpf_config cfg = {DPF_ABI_VERSION, c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], c[9], c[10], c[11], c[12]};
    (*env)->ReleasePrimitiveArrayCritical(env, jcfg, c, JNI_ABORT);
    dpf_handle h = 0;
    int rc = dpf_create(&cfg, &h);
    if (rc != DPF_OK) { throw_dpf(env, 0, rc); return 0; }
    return (jlong)(intptr_t)h;
}

JNIEXPORT void JNICALL Java_mclab_deploy_NativeDPF_destroy(JNIEnv* env, jclass cls, jlong h) {
    (void)env; (void)cls;
    dpf_destroy((dpf_handle)(intptr_t)h);
}

/* void setFamily(long h, double[] A, int P, int[] chainIdx, double[] b, int[] w)
 * A / chainIdx come from LSH.tableIndexGenerators (LSH.scala:27): distinct functions + per-table indices */
JNIEXPORT void JNICALL Java_mclab_deploy_NativeDPF_setFamily(JNIEnv* env, jclass cls, jlong jh, jdoubleArray jA, jint P,
                                                              jintArray jchain, jdoubleArray jb, jintArray jw) {
    (void)cls;
    dpf_handle h = (dpf_handle)(intptr_t)jh;
    double* A = (double*)(*env)->GetPrimitiveArrayCritical(env, jA, 0);
    jint* chain = (jint*)(*env)->GetPrimitiveArrayCritical(env, jchain, 0);
    double* b = jb ? (double*)(*env)->GetPrimitiveArrayCritical(env, jb, 0) : 0;
    jint* w = jw ? (jint*)(*env)->GetPrimitiveArrayCritical(env, jw, 0) : 0;
    int rc = dpf_set_family(h, A, P, chain, b, w);
    if (w) (*env)->ReleasePrimitiveArrayCritical(env, jw, w, JNI_ABORT);
    if (b) (*env)->ReleasePrimitiveArrayCritical(env, jb, b, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jchain, chain, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jA, A, JNI_ABORT);
    if (rc != DPF_OK) throw_dpf(env, h, rc);
}

/* void setPartitioners(long h, double[] Ap) — the L LocalitySensitivePartitioner function sets (Partitioner.scala:31) */
JNIEXPORT void JNICALL Java_mclab_deploy_NativeDPF_setPartitioners(JNIEnv* env, jclass cls, jlong jh, jdoubleArray jAp) {
    (void)cls;
    dpf_handle h = (dpf_handle)(intptr_t)jh;
    double* Ap = (double*)(*env)->GetPrimitiveArrayCritical(env, jAp, 0);
    int rc = dpf_set_partitioners(h, Ap);
    (*env)->ReleasePrimitiveArrayCritical(env, jAp, Ap, JNI_ABORT);
    if (rc != DPF_OK) throw_dpf(env, h, rc);
}

/* void setPartitionersPStable(long h, double[] Ap, double[] b, int[] w) — the same when mclab.lsh.name = pStable: the
 * partitioner chains are PStableHashChains (DensevectorRDFInit.scala:63-70), b / w their offsets and widths (L x pb) */
JNIEXPORT void JNICALL Java_mclab_deploy_NativeDPF_setPartitionersPStable(JNIEnv* env, jclass cls, jlong jh, jdoubleArray jAp,
                                                                          jdoubleArray jb, jintArray jw) {
    (void)cls;
    dpf_handle h = (dpf_handle)(intptr_t)jh;
    double* Ap = (double*)(*env)->GetPrimitiveArrayCritical(env, jAp, 0);
    double* b = (double*)(*env)->GetPrimitiveArrayCritical(env, jb, 0);
    jint* w = (jint*)(*env)->GetPrimitiveArrayCritical(env, jw, 0);
    int rc = dpf_set_partitioners_pstable(h, Ap, b, w);
    (*env)->ReleasePrimitiveArrayCritical(env, jw, w, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jb, b, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jAp, Ap, JNI_ABORT);
    if (rc != DPF_OK) throw_dpf(env, h, rc);
}

/* void fitDense(long h, double[] X, long n) — newMultiThreadFit / newFastFit after parsing (DensevectorRDFInit.scala:127-206) */
JNIEXPORT void JNICALL Java_mclab_deploy_NativeDPF_setStoreMode(JNIEnv* env, jclass cls, jlong jh, jint mode) {
    (void)cls;
    dpf_handle h = (dpf_handle)(intptr_t)jh;
    int rc = dpf_set_store_mode(h, mode);     /* DPF_STORE_AUTO / DPF_STORE_F64_ONLY / DPF_STORE_NARROWEST, before fit */
    if (rc != DPF_OK) throw_dpf(env, h, rc);
}

JNIEXPORT void JNICALL Java_mclab_deploy_NativeDPF_fitDense(JNIEnv* env, jclass cls, jlong jh, jdoubleArray jX, jlong n) {
    (void)cls;
    dpf_handle h = (dpf_handle)(intptr_t)jh;
    double* X = (double*)(*env)->GetPrimitiveArrayCritical(env, jX, 0);
    int rc = dpf_fit_dense(h, X, n);
    (*env)->ReleasePrimitiveArrayCritical(env, jX, X, JNI_ABORT);
    if (rc != DPF_OK) throw_dpf(env, h, rc);
}

/* same through a direct ByteBuffer, for n*d >= 2^31 doubles (JVM array limit) */
JNIEXPORT void JNICALL Java_mclab_deploy_NativeDPF_fitDenseDirect(JNIEnv* env, jclass cls, jlong jh, jobject buf, jlong n) {
    (void)cls;
    dpf_handle h = (dpf_handle)(intptr_t)jh;
    int rc = dpf_fit_dense(h, (const double*)(*env)->GetDirectBufferAddress(env, buf), n);
    if (rc != DPF_OK) throw_dpf(env, h, rc);
}

/* void fitCsr(long h, long[] indptr, int[] indices, double[] values, long n) (SparsevectorRDFInit.scala:124-203) */
JNIEXPORT void JNICALL Java_mclab_deploy_NativeDPF_fitCsr(JNIEnv* env, jclass cls, jlong jh, jlongArray jptr,
                                                           jintArray jidx, jdoubleArray jval, jlong n) {
    (void)cls;
    dpf_handle h = (dpf_handle)(intptr_t)jh;
    jlong* ptr = (jlong*)(*env)->GetPrimitiveArrayCritical(env, jptr, 0);
    jint* idx = (jint*)(*env)->GetPrimitiveArrayCritical(env, jidx, 0);
    double* val = (double*)(*env)->GetPrimitiveArrayCritical(env, jval, 0);
    int rc = dpf_fit_csr(h, (const int64_t*)ptr, idx, val, n);
    (*env)->ReleasePrimitiveArrayCritical(env, jval, val, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jidx, idx, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jptr, ptr, JNI_ABORT);
    if (rc != DPF_OK) throw_dpf(env, h, rc);
}

/* int[] queryCandidatesDense(long h, double[] Q, int[] qids, int steps, long[] offsetsOut)
 * = NewMultiThreadQueryBatch (DensevectorRDFInit.scala:335-360): CSR of sorted unique ids; the Scala side rebuilds
 * Array[Set[AnyRef]] from (offsets, ids). */
JNIEXPORT jintArray JNICALL Java_mclab_deploy_NativeDPF_queryCandidatesDense(JNIEnv* env, jclass cls, jlong jh,
                                                                              jdoubleArray jQ, jintArray jqids, jint steps,
                                                                              jlongArray joff) {
    (void)cls;
    dpf_handle h = (dpf_handle)(intptr_t)jh;
    const jsize nq = (*env)->GetArrayLength(env, joff) - 1;
    int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nq + 1));
    int64_t cap = 1 << 20, total = 0;
    int32_t* ids = (int32_t*)malloc(sizeof(int32_t) * (size_t)cap);
    int rc;
    for (;;) {
        double* Q = (double*)(*env)->GetPrimitiveArrayCritical(env, jQ, 0);
        jint* qids = jqids ? (jint*)(*env)->GetPrimitiveArrayCritical(env, jqids, 0) : 0;
        rc = dpf_query_candidates_dense(h, Q, nq, qids, steps, DPF_PROBE_DENSE, off, ids, cap, &total);
        if (qids) (*env)->ReleasePrimitiveArrayCritical(env, jqids, qids, JNI_ABORT);
        (*env)->ReleasePrimitiveArrayCritical(env, jQ, Q, JNI_ABORT);
        if (rc != DPF_ERR_CAPACITY) break;
        cap = total;
        free(ids);
        ids = (int32_t*)malloc(sizeof(int32_t) * (size_t)cap);
    }
    jintArray out = 0;
    if (rc == DPF_OK) {
        out = (*env)->NewIntArray(env, (jsize)total);
        (*env)->SetIntArrayRegion(env, out, 0, (jsize)total, (const jint*)ids);
        (*env)->SetLongArrayRegion(env, joff, 0, nq + 1, (const jlong*)off);
    } else {
        throw_dpf(env, h, rc);
    }
    free(ids);
    free(off);
    return out;
}

/* int[] queryCandidatesById(long h, int[] qids, int steps, long[] offsetsOut)
 * = SparsevectorRDFInit.NewMultiThreadQueryBatch(queryArray, steps, threads) (SparsevectorRDFInit.scala:324-348) */
JNIEXPORT jintArray JNICALL Java_mclab_deploy_NativeDPF_queryCandidatesById(JNIEnv* env, jclass cls, jlong jh, jintArray jqids,
                                                                             jint steps, jlongArray joff) {
    (void)cls;
    dpf_handle h = (dpf_handle)(intptr_t)jh;
    const jsize nq = (*env)->GetArrayLength(env, jqids);
    int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nq + 1));
    int64_t cap = 1 << 20, total = 0;
    int32_t* ids = (int32_t*)malloc(sizeof(int32_t) * (size_t)cap);
    int rc;
    for (;;) {
        jint* qids = (jint*)(*env)->GetPrimitiveArrayCritical(env, jqids, 0);
        rc = dpf_query_candidates_by_id(h, qids, nq, steps, off, ids, cap, &total);
        (*env)->ReleasePrimitiveArrayCritical(env, jqids, qids, JNI_ABORT);
        if (rc != DPF_ERR_CAPACITY) break;
        cap = total;
        free(ids);
        ids = (int32_t*)malloc(sizeof(int32_t) * (size_t)cap);
    }
    jintArray out = 0;
    if (rc == DPF_OK) {
        out = (*env)->NewIntArray(env, (jsize)total);
        (*env)->SetIntArrayRegion(env, out, 0, (jsize)total, (const jint*)ids);
        (*env)->SetLongArrayRegion(env, joff, 0, nq + 1, (const jlong*)off);
    } else {
        throw_dpf(env, h, rc);
    }
    free(ids);
    free(off);
    return out;
}

/* void queryTopK(long h, double[] Q, int[] qids, int steps, int topK, int metric, int[] idsOut, double[] scoresOut)
 * = the gather / dgemv / argsort of topKAndPrecisionScore (DensevectorRDFInit.scala:472-507) */
JNIEXPORT void JNICALL Java_mclab_deploy_NativeDPF_queryTopK(JNIEnv* env, jclass cls, jlong jh, jdoubleArray jQ,
                                                              jintArray jqids, jint steps, jint topk, jint metric,
                                                              jintArray jids, jdoubleArray jsc) {
    (void)cls;
    dpf_handle h = (dpf_handle)(intptr_t)jh;
    const jsize nq = (*env)->GetArrayLength(env, jids) / topk;
    double* Q = (double*)(*env)->GetPrimitiveArrayCritical(env, jQ, 0);
    jint* qids = jqids ? (jint*)(*env)->GetPrimitiveArrayCritical(env, jqids, 0) : 0;
    jint* ids = (jint*)(*env)->GetPrimitiveArrayCritical(env, jids, 0);
    double* sc = (double*)(*env)->GetPrimitiveArrayCritical(env, jsc, 0);
    int rc = dpf_query_topk_dense(h, Q, nq, qids, steps, DPF_PROBE_DENSE, topk, metric, ids, sc);
    (*env)->ReleasePrimitiveArrayCritical(env, jsc, sc, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, jids, ids, 0);
    if (qids) (*env)->ReleasePrimitiveArrayCritical(env, jqids, qids, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jQ, Q, JNI_ABORT);
    if (rc != DPF_OK) throw_dpf(env, h, rc);
}

/* int[] hashDense(long h, double[] X, long n) — LSH.calculateIndex for a batch (LSH.scala:135-166): L x n keys */
JNIEXPORT jintArray JNICALL Java_mclab_deploy_NativeDPF_hashDense(JNIEnv* env, jclass cls, jlong jh, jdoubleArray jX, jlong n,
                                                                   jint L) {
    (void)cls;
    dpf_handle h = (dpf_handle)(intptr_t)jh;
    int32_t* keys = (int32_t*)malloc(sizeof(int32_t) * (size_t)(L * n));
    double* X = (double*)(*env)->GetPrimitiveArrayCritical(env, jX, 0);
    int rc = dpf_hash_dense(h, X, n, keys, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, jX, X, JNI_ABORT);
    jintArray out = 0;
    if (rc == DPF_OK) {
        out = (*env)->NewIntArray(env, (jsize)(L * n));
        (*env)->SetIntArrayRegion(env, out, 0, (jsize)(L * n), (const jint*)keys);
    } else {
        throw_dpf(env, h, rc);
    }
    free(keys);
    return out;
}

/* long remove(long h, int[] ids) — RandomDrawTreeMap.remove for a batch of ids in every table
 * (RandomDrawTreeMap.java:1817-1932); returns the (table, id) entries removed */
JNIEXPORT jlong JNICALL Java_mclab_deploy_NativeDPF_remove(JNIEnv* env, jclass cls, jlong jh, jintArray jids) {
    (void)cls;
    dpf_handle h = (dpf_handle)(intptr_t)jh;
    const jsize m = (*env)->GetArrayLength(env, jids);
    jint* ids = (jint*)(*env)->GetPrimitiveArrayCritical(env, jids, 0);
    int64_t gone = 0;
    int rc = dpf_remove(h, ids, m, &gone);
    (*env)->ReleasePrimitiveArrayCritical(env, jids, ids, JNI_ABORT);
    if (rc != DPF_OK) throw_dpf(env, h, rc);
    return (jlong)gone;
}

/* ---- multi-GPU: one handle per GPU, NCCL inside the library (Partitioner.scala:27-64 mapped onto the GPUs) ---------- */
/* byte[] commUniqueId() — rank 0; the caller hands the 128 bytes to the other ranks */
JNIEXPORT jbyteArray JNICALL Java_mclab_deploy_NativeDPF_commUniqueId(JNIEnv* env, jclass cls) {
    (void)cls;
    uint8_t id[DPF_COMM_ID_BYTES];
    int rc = dpf_comm_unique_id(id);
    if (rc != DPF_OK) { throw_dpf(env, 0, rc); return 0; }
    jbyteArray out = (*env)->NewByteArray(env, DPF_COMM_ID_BYTES);
    (*env)->SetByteArrayRegion(env, out, 0, DPF_COMM_ID_BYTES, (const jbyte*)id);
    return out;
}

/* void commInit(long h, byte[] id) — collective over the `world` ranks of the handle's configuration */
JNIEXPORT void JNICALL Java_mclab_deploy_NativeDPF_commInit(JNIEnv* env, jclass cls, jlong jh, jbyteArray jid) {
    (void)cls;
    dpf_handle h = (dpf_handle)(intptr_t)jh;
    uint8_t id[DPF_COMM_ID_BYTES];
    (*env)->GetByteArrayRegion(env, jid, 0, DPF_COMM_ID_BYTES, (jbyte*)id);
    int rc = dpf_comm_init(h, id);
    if (rc != DPF_OK) throw_dpf(env, h, rc);
}

/* void fitDenseSharded(long h, double[] X, long n) — collective newMultiThreadFit: every rank hashes n / world vectors */
JNIEXPORT void JNICALL Java_mclab_deploy_NativeDPF_fitDenseSharded(JNIEnv* env, jclass cls, jlong jh, jdoubleArray jX, jlong n) {
    (void)cls;
    dpf_handle h = (dpf_handle)(intptr_t)jh;
    double* X = (double*)(*env)->GetPrimitiveArrayCritical(env, jX, 0);
    int rc = dpf_fit_dense_sharded(h, X, n);
    (*env)->ReleasePrimitiveArrayCritical(env, jX, X, JNI_ABORT);
    if (rc != DPF_OK) throw_dpf(env, h, rc);
}

/* void queryTopKAll(long h, double[] Q, int[] qids, int steps, int topK, int metric, int[] idsOut, double[] scoresOut)
 * — collective: the same queries on every rank, the merged global top k on every rank */
JNIEXPORT void JNICALL Java_mclab_deploy_NativeDPF_queryTopKAll(JNIEnv* env, jclass cls, jlong jh, jdoubleArray jQ,
                                                                 jintArray jqids, jint steps, jint topk, jint metric,
                                                                 jintArray jids, jdoubleArray jsc) {
    (void)cls;
    dpf_handle h = (dpf_handle)(intptr_t)jh;
    const jsize nq = (*env)->GetArrayLength(env, jids) / topk;
    double* Q = (double*)(*env)->GetPrimitiveArrayCritical(env, jQ, 0);
    jint* qids = jqids ? (jint*)(*env)->GetPrimitiveArrayCritical(env, jqids, 0) : 0;
    jint* ids = (jint*)(*env)->GetPrimitiveArrayCritical(env, jids, 0);
    double* sc = (double*)(*env)->GetPrimitiveArrayCritical(env, jsc, 0);
    int rc = dpf_query_topk_dense_all(h, Q, nq, qids, steps, DPF_PROBE_DENSE, topk, metric, ids, sc);
    (*env)->ReleasePrimitiveArrayCritical(env, jsc, sc, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, jids, ids, 0);
    if (qids) (*env)->ReleasePrimitiveArrayCritical(env, jqids, qids, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jQ, Q, JNI_ABORT);
    if (rc != DPF_OK) throw_dpf(env, h, rc);
}
