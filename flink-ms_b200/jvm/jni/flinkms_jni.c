/*
 * flinkms_jni.c -- thin JNI shim: de.tub.it4bi.gpu.Native -> the C-ABI of include/flinkms.h.
 * Mechanical by design: unwrap direct ByteBuffers (GetDirectBufferAddress), call fms_*, turn a non-zero status
 * into java.lang.RuntimeException(fms_last_error).  No pointer is retained past a call.
 *
 * NOT COMPILED HERE (the image has no JDK / jni.h).  Build where a JDK exists:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../../include flinkms_jni.c \
 *       -L.. -lflinkms_b200 -Wl,-rpath,'$ORIGIN' -o libflinkms_jni.so
 */
#include <jni.h>
#include <stdint.h>

#include "flinkms.h"

static void *buf(JNIEnv *env, jobject b) { return b ? (*env)->GetDirectBufferAddress(env, b) : NULL; }

static int check(JNIEnv *env, fms_ctx *ctx, int rc) {
    if (rc == FMS_OK) return 0;
    jclass ex = (*env)->FindClass(env, "java/lang/RuntimeException");
    if (ex) (*env)->ThrowNew(env, ex, fms_last_error(ctx));
    return rc;
}

JNIEXPORT jlong JNICALL Java_de_tub_it4bi_gpu_Native_create(JNIEnv *env, jclass cls, jintArray devices) {
    (void)cls;
    fms_ctx *ctx = NULL;
    int dev[64];
    const jsize n = (*env)->GetArrayLength(env, devices);
    if (n < 1 || n > 64) { check(env, NULL, FMS_ERR_INVALID_ARG); return 0; }
    jint tmp[64];
    (*env)->GetIntArrayRegion(env, devices, 0, n, tmp);
    for (jsize k = 0; k < n; ++k) dev[k] = (int)tmp[k];
    if (check(env, NULL, fms_create(&ctx, dev, (int)n))) return 0;      /* n > 1: a group ctx */
    return (jlong)(intptr_t)ctx;
}

/* chunked upload (direct ByteBuffers are capped at 2 GiB) */
JNIEXPORT void JNICALL Java_de_tub_it4bi_gpu_Native_alsStreamBegin(JNIEnv *env, jclass cls, jlong c, jlong nnz) {
    (void)cls;
    fms_ctx *ctx = (fms_ctx *)(intptr_t)c;
    check(env, ctx, fms_als_stream_begin(ctx, nnz));
}
JNIEXPORT void JNICALL Java_de_tub_it4bi_gpu_Native_alsStreamAppend(JNIEnv *env, jclass cls, jlong c, jobject user, jobject item,
                                                                    jobject rating, jlong count) {
    (void)cls;
    fms_ctx *ctx = (fms_ctx *)(intptr_t)c;
    check(env, ctx, fms_als_stream_append(ctx, buf(env, user), buf(env, item), buf(env, rating), count));
}
JNIEXPORT jlongArray JNICALL Java_de_tub_it4bi_gpu_Native_alsStreamFit(JNIEnv *env, jclass cls, jlong c, jint rank, jint iterations,
                                                                       jdouble lambda, jlong seed) {
    (void)cls;
    fms_ctx *ctx = (fms_ctx *)(intptr_t)c;
    int64_t n[2] = {0, 0};
    if (check(env, ctx, fms_als_stream_fit(ctx, rank, iterations, lambda, seed, &n[0], &n[1]))) return NULL;
    jlongArray out = (*env)->NewLongArray(env, 2);
    jlong v[2] = {(jlong)n[0], (jlong)n[1]};
    if (out) (*env)->SetLongArrayRegion(env, out, 0, 2, v);
    return out;
}
JNIEXPORT void JNICALL Java_de_tub_it4bi_gpu_Native_svmStreamBegin(JNIEnv *env, jclass cls, jlong c, jlong n, jint dim) {
    (void)cls;
    fms_ctx *ctx = (fms_ctx *)(intptr_t)c;
    check(env, ctx, fms_svm_stream_begin(ctx, n, dim));
}
JNIEXPORT void JNICALL Java_de_tub_it4bi_gpu_Native_svmStreamAppend(JNIEnv *env, jclass cls, jlong c, jobject x, jint xDtype, jobject y,
                                                                    jobject blockId, jlong rows) {
    (void)cls;
    fms_ctx *ctx = (fms_ctx *)(intptr_t)c;
    check(env, ctx, fms_svm_stream_append(ctx, buf(env, x), xDtype, buf(env, y), buf(env, blockId), rows));
}
JNIEXPORT void JNICALL Java_de_tub_it4bi_gpu_Native_svmStreamFit(JNIEnv *env, jclass cls, jlong c, jint blocks, jint iterations,
                                                                 jlong localIterations, jdouble regularization, jdouble stepsize,
                                                                 jlong seed, jint rngMode, jobject w) {
    (void)cls;
    fms_ctx *ctx = (fms_ctx *)(intptr_t)c;
    check(env, ctx, fms_svm_stream_fit(ctx, blocks, iterations, localIterations, regularization, stepsize, seed, rngMode, buf(env, w)));
}

JNIEXPORT void JNICALL Java_de_tub_it4bi_gpu_Native_destroy(JNIEnv *env, jclass cls, jlong ctx) {
    (void)env; (void)cls;
    fms_destroy((fms_ctx *)(intptr_t)ctx);
}

JNIEXPORT jlongArray JNICALL Java_de_tub_it4bi_gpu_Native_alsFit(JNIEnv *env, jclass cls, jlong c, jobject user, jobject item,
                                                                 jobject rating, jlong nnz, jint rank, jint iterations,
                                                                 jdouble lambda, jlong seed) {
    (void)cls;
    fms_ctx *ctx = (fms_ctx *)(intptr_t)c;
    int64_t n[2] = {0, 0};
    if (check(env, ctx, fms_als_fit(ctx, buf(env, user), buf(env, item), buf(env, rating), nnz, rank, iterations, lambda, seed,
                                    &n[0], &n[1]))) return NULL;
    jlongArray out = (*env)->NewLongArray(env, 2);
    jlong v[2] = {(jlong)n[0], (jlong)n[1]};
    (*env)->SetLongArrayRegion(env, out, 0, 2, v);
    return out;
}

JNIEXPORT void JNICALL Java_de_tub_it4bi_gpu_Native_alsGetFactors(JNIEnv *env, jclass cls, jlong c, jint which, jobject ids,
                                                                  jobject factors) {
    (void)cls;
    fms_ctx *ctx = (fms_ctx *)(intptr_t)c;
    check(env, ctx, fms_als_get_factors(ctx, which, buf(env, ids), buf(env, factors)));
}

JNIEXPORT void JNICALL Java_de_tub_it4bi_gpu_Native_alsSetModel(JNIEnv *env, jclass cls, jlong c, jobject uids, jobject uf,
                                                                jlong nu, jobject iids, jobject itf, jlong ni, jint rank) {
    (void)cls;
    fms_ctx *ctx = (fms_ctx *)(intptr_t)c;
    check(env, ctx, fms_als_set_model(ctx, buf(env, uids), buf(env, uf), nu, buf(env, iids), buf(env, itf), ni, rank));
}

JNIEXPORT void JNICALL Java_de_tub_it4bi_gpu_Native_alsPredict(JNIEnv *env, jclass cls, jlong c, jobject u, jobject i, jlong n,
                                                               jobject out, jobject found) {
    (void)cls;
    fms_ctx *ctx = (fms_ctx *)(intptr_t)c;
    check(env, ctx, fms_als_predict(ctx, buf(env, u), buf(env, i), n, buf(env, out), buf(env, found)));
}

JNIEXPORT jdoubleArray JNICALL Java_de_tub_it4bi_gpu_Native_alsMse(JNIEnv *env, jclass cls, jlong c, jobject u, jobject i, jobject rating,
                                                                 jlong n) {
    (void)cls;
    fms_ctx *ctx = (fms_ctx *)(intptr_t)c;
    double mse = 0.0;
    int64_t used = 0;
    if (check(env, ctx, fms_als_mse(ctx, buf(env, u), buf(env, i), buf(env, rating), n, &mse, &used))) return NULL;
    jdouble r[2] = {mse, (jdouble)used};
    jdoubleArray out = (*env)->NewDoubleArray(env, 2);
    if (out) (*env)->SetDoubleArrayRegion(env, out, 0, 2, r);
    return out;
}

JNIEXPORT void JNICALL Java_de_tub_it4bi_gpu_Native_alsSgdStep(JNIEnv *env, jclass cls, jlong c, jobject user, jobject item, jobject rating,
                                                               jlong n, jdouble lr, jdouble ureg, jdouble ireg, jobject umean, jobject imean,
                                                               jboolean apply, jobject outUser, jobject outItem) {
    (void)cls;
    fms_ctx *ctx = (fms_ctx *)(intptr_t)c;
    check(env, ctx, fms_als_sgd_step(ctx, buf(env, user), buf(env, item), buf(env, rating), n, lr, ureg, ireg, buf(env, umean),
                                     buf(env, imean), apply ? 1 : 0, buf(env, outUser), buf(env, outItem)));
}

JNIEXPORT jlong JNICALL Java_de_tub_it4bi_gpu_Native_parseRatingsCsv(JNIEnv *env, jclass cls, jlong c, jobject text, jlong nbytes,
                                                                     jbyte delimiter, jboolean ignoreFirst, jobject user, jobject item,
                                                                     jobject rating, jlong capacity) {
    (void)cls;
    fms_ctx *ctx = (fms_ctx *)(intptr_t)c;
    int64_t rows = 0;
    if (check(env, ctx, fms_parse_ratings_csv(ctx, buf(env, text), nbytes, (char)delimiter, ignoreFirst ? 1 : 0, buf(env, user),
                                              buf(env, item), buf(env, rating), capacity, &rows))) return -1;
    return rows;
}

JNIEXPORT jlongArray JNICALL Java_de_tub_it4bi_gpu_Native_parseLibSvm(JNIEnv *env, jclass cls, jlong c, jobject text, jlong nbytes,
                                                                      jobject label, jlong rowCapacity, jobject csrPtr, jobject csrIdx,
                                                                      jobject csrVal, jlong nnzCapacity) {
    (void)cls;
    fms_ctx *ctx = (fms_ctx *)(intptr_t)c;
    int64_t rows = 0, nnz = 0;
    int32_t dim = 0;
    if (check(env, ctx, fms_parse_libsvm(ctx, buf(env, text), nbytes, buf(env, label), rowCapacity, buf(env, csrPtr), buf(env, csrIdx),
                                         buf(env, csrVal), nnzCapacity, &rows, &nnz, &dim))) return NULL;
    jlong r[3] = {rows, nnz, dim};
    jlongArray out = (*env)->NewLongArray(env, 3);
    if (out) (*env)->SetLongArrayRegion(env, out, 0, 3, r);
    return out;
}

JNIEXPORT void JNICALL Java_de_tub_it4bi_gpu_Native_svmFit(JNIEnv *env, jclass cls, jlong c, jobject x, jint xDtype, jobject csrPtr,
                                                           jobject csrIdx, jobject y, jlong n, jint dim, jobject blockId, jint blocks,
                                                           jint iterations, jlong localIterations, jdouble regularization,
                                                           jdouble stepsize, jlong seed, jint rngMode, jobject w) {
    (void)cls;
    fms_ctx *ctx = (fms_ctx *)(intptr_t)c;
    check(env, ctx, fms_svm_fit(ctx, buf(env, x), xDtype, buf(env, csrPtr), buf(env, csrIdx), buf(env, y), n, dim, buf(env, blockId),
                                blocks, iterations, localIterations, regularization, stepsize, seed, rngMode, buf(env, w)));
}

JNIEXPORT void JNICALL Java_de_tub_it4bi_gpu_Native_svmPredict(JNIEnv *env, jclass cls, jlong c, jobject x, jint xDtype,
                                                               jobject csrPtr, jobject csrIdx, jlong n, jint dim, jobject w,
                                                               jdouble threshold, jboolean decision, jobject out) {
    (void)cls;
    fms_ctx *ctx = (fms_ctx *)(intptr_t)c;
    check(env, ctx, fms_svm_predict(ctx, buf(env, x), xDtype, buf(env, csrPtr), buf(env, csrIdx), n, dim, buf(env, w), threshold,
                                    decision ? 1 : 0, buf(env, out)));
}
