package de.tub.it4bi.gpu;

import java.nio.ByteBuffer;

/**
 * JNI binding of libflinkms_b200.so (C-ABI: include/flinkms.h).  One instance == one fms_ctx (one GPU, or a group of GPUs); not
 * thread-safe (create one per Flink task thread, like the reference's QueryClientHelper, SGD.java:137).
 * All bulk arguments are DIRECT ByteBuffers in native byte order; the shim never retains their addresses.
 * A non-zero fms_status is turned into a RuntimeException carrying fms_last_error().
 *
 * NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no JDK (SURVEY.md, "READ THIS FIRST" item 3).
 */
public final class Native implements AutoCloseable {
    static { System.loadLibrary("flinkms_jni"); }   // flinkms_jni links libflinkms_b200.so

    private long ctx;                                // fms_ctx*

    public Native(int device) { ctx = create(new int[] {device}); }
    /** One process, several GPUs of one NVLink node: a GROUP ctx (fms_create with n_devices > 1).  The whole-fit calls shard
     *  the work inside the library; everything that reads the model is served by devices[0]. */
    public Native(int[] devices) { ctx = create(devices); }
    /** Devices from the system property flinkms.devices ("0" or "0,1,2,3"; default "0"). */
    public static Native fromProperties() {
        String[] t = System.getProperty("flinkms.devices", "0").split(",");
        int[] d = new int[t.length];
        for (int k = 0; k < t.length; ++k) d[k] = Integer.parseInt(t[k].trim());
        return new Native(d);
    }

    @Override public void close() { if (ctx != 0) { destroy(ctx); ctx = 0; } }

    // ---- ALS (replaces als.fit / factorsOption, ALSImpl.scala:52-57)
    /** user/item: nnz int64, rating: nnz float64.  Returns {nUsers, nItems}. */
    public long[] alsFit(ByteBuffer user, ByteBuffer item, ByteBuffer rating, long nnz, int rank, int iterations,
                         double lambda, long seed) { return alsFit(ctx, user, item, rating, nnz, rank, iterations, lambda, seed); }
    /** ids: n int64 (ascending), factors: n*rank float64 row-major. which: 0 = users, 1 = items. */
    public void alsGetFactors(int which, ByteBuffer ids, ByteBuffer factors) { alsGetFactors(ctx, which, ids, factors); }
    public void alsSetModel(ByteBuffer userIds, ByteBuffer userFactors, long nUsers, ByteBuffer itemIds, ByteBuffer itemFactors,
                            long nItems, int rank) { alsSetModel(ctx, userIds, userFactors, nUsers, itemIds, itemFactors, nItems, rank); }
    /** out: n float64, found: n bytes (0 = an id is not in the model). */
    public void alsPredict(ByteBuffer u, ByteBuffer i, long n, ByteBuffer out, ByteBuffer found) { alsPredict(ctx, u, i, n, out, found); }

    // ---- chunked upload: a direct ByteBuffer holds < 2 GiB, BASELINE configs[1] is 2.4 GB of triples and configs[3] 41 GB of rows
    public void alsStreamBegin(long nnzTotal) { alsStreamBegin(ctx, nnzTotal); }
    public void alsStreamAppend(ByteBuffer user, ByteBuffer item, ByteBuffer rating, long count) { alsStreamAppend(ctx, user, item, rating, count); }
    /** Returns {nUsers, nItems}. */
    public long[] alsStreamFit(int rank, int iterations, double lambda, long seed) { return alsStreamFit(ctx, rank, iterations, lambda, seed); }
    public void svmStreamBegin(long nTotal, int dim) { svmStreamBegin(ctx, nTotal, dim); }
    /** xRows: rows x dim dense values, xDtype 0 = f32, 1 = f64; y: rows float64; blockId: rows int32. */
    public void svmStreamAppend(ByteBuffer xRows, int xDtype, ByteBuffer y, ByteBuffer blockId, long rows) { svmStreamAppend(ctx, xRows, xDtype, y, blockId, rows); }
    public void svmStreamFit(int blocks, int iterations, long localIterations, double regularization, double stepsize, long seed,
                             int rngMode, ByteBuffer w) { svmStreamFit(ctx, blocks, iterations, localIterations, regularization, stepsize, seed, rngMode, w); }

    // ---- SVM (replaces svm.fit / weightsOption, SVMImpl.scala:29-34)
    /** xDtype: 0 dense f32, 1 dense f64, 2 CSR (values f64, csrPtr int64, csrIdx int32).  w: dim float64 out. */
    public void svmFit(ByteBuffer x, int xDtype, ByteBuffer csrPtr, ByteBuffer csrIdx, ByteBuffer y, long n, int dim,
                       ByteBuffer blockId, int blocks, int iterations, long localIterations, double regularization,
                       double stepsize, long seed, int rngMode, ByteBuffer w) {
        svmFit(ctx, x, xDtype, csrPtr, csrIdx, y, n, dim, blockId, blocks, iterations, localIterations, regularization,
               stepsize, seed, rngMode, w);
    }
    public void svmPredict(ByteBuffer x, int xDtype, ByteBuffer csrPtr, ByteBuffer csrIdx, long n, int dim, ByteBuffer w,
                           double threshold, boolean outputDecisionFunction, ByteBuffer out) {
        svmPredict(ctx, x, xDtype, csrPtr, csrIdx, n, dim, w, threshold, outputDecisionFunction, out);
    }

    /** als-ms evaluation MSE.java:122-160: {mse, ratings scored}. */
    public double[] alsMse(ByteBuffer u, ByteBuffer i, ByteBuffer rating, long n) { return alsMse(ctx, u, i, rating, n); }

    /** One micro-batch of the als-ms online SGD (SGD.java:160-217); outUser / outItem receive n x rank doubles. */
    public void alsSgdStep(ByteBuffer user, ByteBuffer item, ByteBuffer rating, long n, double learningRate, double userReg,
                           double itemReg, ByteBuffer userMean, ByteBuffer itemMean, boolean apply, ByteBuffer outUser, ByteBuffer outItem) {
        alsSgdStep(ctx, user, item, rating, n, learningRate, userReg, itemReg, userMean, itemMean, apply, outUser, outItem);
    }

    /** env.readCsvFile[(Int, Int, Double)] on the GPU (ALSImpl.scala:29-32); pass null outputs to size them. Returns rows. */
    public long parseRatingsCsv(ByteBuffer text, long nbytes, char delimiter, boolean ignoreFirstLine, ByteBuffer user, ByteBuffer item,
                                ByteBuffer rating, long capacity) {
        return parseRatingsCsv(ctx, text, nbytes, (byte) delimiter, ignoreFirstLine, user, item, rating, capacity);
    }

    /** env.readLibSVM on the GPU (SVMImpl.scala:21); pass null outputs to size them. Returns {rows, nnz, dim}. */
    public long[] parseLibSvm(ByteBuffer text, long nbytes, ByteBuffer label, long rowCapacity, ByteBuffer csrPtr, ByteBuffer csrIdx,
                              ByteBuffer csrVal, long nnzCapacity) {
        return parseLibSvm(ctx, text, nbytes, label, rowCapacity, csrPtr, csrIdx, csrVal, nnzCapacity);
    }

    private static native long create(int[] devices);
    private static native void alsStreamBegin(long ctx, long nnzTotal);
    private static native void alsStreamAppend(long ctx, ByteBuffer user, ByteBuffer item, ByteBuffer rating, long count);
    private static native long[] alsStreamFit(long ctx, int rank, int iterations, double lambda, long seed);
    private static native void svmStreamBegin(long ctx, long nTotal, int dim);
    private static native void svmStreamAppend(long ctx, ByteBuffer xRows, int xDtype, ByteBuffer y, ByteBuffer blockId, long rows);
    private static native void svmStreamFit(long ctx, int blocks, int iterations, long localIterations, double regularization,
                                            double stepsize, long seed, int rngMode, ByteBuffer w);
    private static native void destroy(long ctx);
    private static native long[] alsFit(long ctx, ByteBuffer user, ByteBuffer item, ByteBuffer rating, long nnz, int rank,
                                        int iterations, double lambda, long seed);
    private static native void alsGetFactors(long ctx, int which, ByteBuffer ids, ByteBuffer factors);
    private static native void alsSetModel(long ctx, ByteBuffer userIds, ByteBuffer userFactors, long nUsers, ByteBuffer itemIds,
                                           ByteBuffer itemFactors, long nItems, int rank);
    private static native void alsPredict(long ctx, ByteBuffer u, ByteBuffer i, long n, ByteBuffer out, ByteBuffer found);
    private static native double[] alsMse(long ctx, ByteBuffer u, ByteBuffer i, ByteBuffer rating, long n);
    private static native void alsSgdStep(long ctx, ByteBuffer user, ByteBuffer item, ByteBuffer rating, long n, double learningRate,
                                          double userReg, double itemReg, ByteBuffer userMean, ByteBuffer itemMean, boolean apply,
                                          ByteBuffer outUser, ByteBuffer outItem);
    private static native long parseRatingsCsv(long ctx, ByteBuffer text, long nbytes, byte delimiter, boolean ignoreFirstLine,
                                               ByteBuffer user, ByteBuffer item, ByteBuffer rating, long capacity);
    private static native long[] parseLibSvm(long ctx, ByteBuffer text, long nbytes, ByteBuffer label, long rowCapacity,
                                             ByteBuffer csrPtr, ByteBuffer csrIdx, ByteBuffer csrVal, long nnzCapacity);
    private static native void svmFit(long ctx, ByteBuffer x, int xDtype, ByteBuffer csrPtr, ByteBuffer csrIdx, ByteBuffer y,
                                      long n, int dim, ByteBuffer blockId, int blocks, int iterations, long localIterations,
                                      double regularization, double stepsize, long seed, int rngMode, ByteBuffer w);
    private static native void svmPredict(long ctx, ByteBuffer x, int xDtype, ByteBuffer csrPtr, ByteBuffer csrIdx, long n,
                                          int dim, ByteBuffer w, double threshold, boolean outputDecisionFunction, ByteBuffer out);
}
