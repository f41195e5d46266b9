/*
 * flinkms.h -- C-ABI of libflinkms_b200.so: the B200-native replacement for the FlinkML 1.3.1 hot paths
 * that mmziyad/flink-ms calls (ALS.fit/predict, SVM.fit/predict, queryable-state lookup-and-dot).
 *
 * The reference defines no FFI of its own; each entry point below names the reference call site whose
 * work it takes over (paths relative to the reference root, per-file line numbers).  A JVM host binds
 * these through the JNI shim shown in INTEGRATION.md; tests bind them through ctypes.
 *
 * Conventions
 *   - return 0 (FMS_OK) or a negative fms_status; the message is fms_last_error(ctx) (ctx-owned, valid
 *     until the next call on that ctx).  The JNI shim turns non-zero into a RuntimeException.
 *   - caller owns every in/out HOST buffer; pointers are never retained past the call.
 *   - the library owns all device memory inside fms_ctx.  A ctx is bound to ONE GPU, or -- fms_create with n_devices > 1 --
 *     is a group that drives several GPUs from one process (see fms_create).
 *   - a ctx is NOT thread-safe (the reference calls fit once from the driver thread, ALSImpl.scala:52,
 *     and uses one lookup helper per operator thread, SGD.java:137, MSE.java:112).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns FMS_ERR_CUDA.
 *   - ids are int64 (FlinkML widens Int to Long), values are IEEE double at the ABI like the reference
 *     types ((Int,Int,Double) ALSImpl.scala:29; Array[Double] ALSImpl.scala:83; DenseVector.data
 *     SVMImpl.scala:34).  Device arithmetic is fp32 (documented in DESIGN.md) unless noted.
 */
#ifndef FLINKMS_H
#define FLINKMS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FMS_API __attribute__((visibility("default")))
#else
#define FMS_API
#endif

typedef struct fms_ctx fms_ctx;

typedef enum fms_status {
    FMS_OK = 0,
    FMS_ERR_INVALID_ARG = -1,  /* null pointer, negative size, id out of range ...                     */
    FMS_ERR_CUDA = -2,         /* CUDA runtime/driver error, or no device                               */
    FMS_ERR_STATE = -3,        /* e.g. predict before fit ("The ALS model has not been fitted.")        */
    FMS_ERR_UNSUPPORTED = -4,  /* configuration outside what the kernels cover (rank > 128, ...)        */
    FMS_ERR_OOM = -5           /* device or host allocation failed                                      */
} fms_status;

/* x_dtype of the SVM design matrix */
enum { FMS_X_DENSE_F32 = 0, FMS_X_DENSE_F64 = 1, FMS_X_CSR_F64 = 2 };
/* SVM RNG schedule (SURVEY.md Appendix A.2 "determinism traps") */
enum { FMS_RNG_PER_BLOCK = 0, FMS_RNG_SHARED_SUBTASK = 1 };
/* which factor matrix */
enum { FMS_USERS = 0, FMS_ITEMS = 1 };
/* ALS solver selection: AUTO picks the tensor-core kernel when the rank supports it */
/* TC = tensor-core kernel, producer flavour chosen per half-step (TMA gather4 for long rows, 16-byte LDGSTS copies for
 * short ones); TC_TMA / TC_LDGSTS force one flavour (tests, experiments). */
enum { FMS_ALS_KERNEL_AUTO = 0, FMS_ALS_KERNEL_GENERIC = 1, FMS_ALS_KERNEL_TC = 2, FMS_ALS_KERNEL_TC_TMA = 3, FMS_ALS_KERNEL_TC_LDGSTS = 4 };
/* SVM sweep selection */
enum { FMS_SVM_KERNEL_AUTO = 0, FMS_SVM_KERNEL_GENERIC = 1, FMS_SVM_KERNEL_WARP = 2 };

/* ------------------------------------------------------------------------------------------------
 * context
 * ---------------------------------------------------------------------------------------------- */
/* n_devices == 1: a ctx bound to one GPU (device_ids may be NULL => current device); every entry point below works on it.
 * n_devices  > 1: a GROUP ctx -- ONE process (the JVM of a Flink task manager, a C program) drives n GPUs of one NVLink node:
 *   the whole-fit entry points (fms_als_fit, fms_svm_fit) shard the rows / blocks over the devices inside the library (triples
 *   and rows travel 1/n per device over PCIe, factor rows and the SVM delta between the GPUs over NVLink, device-side event
 *   barriers between half-steps; no NCCL, no torch), and every entry point that reads the fitted model (get_factors, predict,
 *   mse, sgd_step, get_weights, svm_predict, the parsers) is served by the first device, which holds the full model.  The
 *   The chunked-upload entry points (fms_*_stream_*) work on a group too.  The
 *   step-level entry points (fms_als_load .. fms_als_planes_*, fms_svm_load .. fms_svm_apply_delta) return
 *   FMS_ERR_UNSUPPORTED on a group ctx: they exist for hosts that run one process per GPU.
 *   This is how the reference's `setBlocks` / Flink parallelism (ALSImpl.scala:39-41, SVMImpl.scala:25) maps onto one box. */
FMS_API int fms_create(fms_ctx **out, const int *device_ids, int n_devices);
FMS_API void fms_destroy(fms_ctx *ctx);
FMS_API const char *fms_last_error(const fms_ctx *ctx);
FMS_API int fms_version(void);
/* raw cudaStream_t all of this ctx's kernels are launched on (for CUDA-event timing by the host) */
FMS_API void *fms_stream(fms_ctx *ctx);
FMS_API int fms_synchronize(fms_ctx *ctx);
/* number of kernel launches issued by this ctx since creation (bench.py's gpu_launches) */
FMS_API int64_t fms_launch_count(const fms_ctx *ctx);

/* ------------------------------------------------------------------------------------------------
 * ALS -- replaces `als.fit(inputDS, parameters)` (flink-als/.../ALSImpl.scala:35-52) and the
 * `factorsOption` read-out (ALSImpl.scala:55-57).
 * ---------------------------------------------------------------------------------------------- */
/* Whole fit from host COO triples: createBlockInformation (sorted distinct ids, by-user and by-item CSR),
 * items <- Random(id ^ seed).nextDouble x rank, `iterations` x {users, items} half-steps, final users
 * half-step.  ALS-WR regularisation lambda * n_u.  Results stay in ctx. */
FMS_API int fms_als_fit(fms_ctx *ctx, const int64_t *user, const int64_t *item, const double *rating, int64_t nnz,
                        int rank, int iterations, double lambda, int64_t seed,
                        /*out*/ int64_t *n_users, int64_t *n_items);
/* (ids ascending, row-major N x rank doubles).  Either out pointer may be NULL. */
FMS_API int fms_als_get_factors(fms_ctx *ctx, int which, int64_t *ids, double *factors);
/* The same for the rows [row_begin, row_end) of a side (dense row index = position of the id in ascending order):
 * what one rank of a sharded run emits for its own shard [S * g, S * (g + 1)) -- those rows are current on that
 * rank without the final all-gather of the fp32 matrices. */
FMS_API int fms_als_get_factor_rows(fms_ctx *ctx, int which, int64_t row_begin, int64_t row_end, int64_t *ids, double *factors);
/* Install an externally trained model (e.g. parsed `id,U|I,f1;f2;...` rows, ALSImpl.scala:84) for predict. */
FMS_API int fms_als_set_model(fms_ctx *ctx, const int64_t *user_ids, const double *user_factors, int64_t n_users,
                              const int64_t *item_ids, const double *item_factors, int64_t n_items, int rank);
/* ALS.predict / the serve-side lookup-and-dot (flink-queryable-client/.../ALSPredictRandom.java:63-92,
 * ALSPredict.java:65-83): out[k] = U[u_k] . I[i_k] (fp64 accumulate); found[k] = 0 when either id is not in
 * the model (the reference reports "Factors do not exist" and skips; out[k] is then 0). */
FMS_API int fms_als_predict(fms_ctx *ctx, const int64_t *u, const int64_t *i, int64_t n, double *out, uint8_t *found);
/* MSE.java:122-160 (als-ms evaluation): mean squared error of the model on held-out ratings, missing ids
 * skipped; *n_used receives the number of ratings scored. */
FMS_API int fms_als_mse(fms_ctx *ctx, const int64_t *u, const int64_t *i, const double *rating, int64_t n,
                        double *mse, int64_t *n_used);

/* Online SGD step of the als-ms job (als-ms/.../qs/SGD.java:160-217) for a micro-batch of n ratings against the
 * model resident in the ctx (fitted, or installed with fms_als_set_model).  For rating k: u = U[user_k] (user_mean
 * when the id is not in the model, SGD.java:171), v = I[item_k] (item_mean); e = r - u.v; u' = u + lr (e v - reg_u u);
 * v' = v + lr (e u - reg_i v) -- both from the ORIGINAL u, v (SGD.java:199-207; SGDV0 differs).  out_user / out_item
 * (n x rank doubles) receive the two records the reference emits per rating.  Every record reads the model as it was
 * before the batch (what the reference sees while a batch's updates are still travelling through Kafka); with
 * apply != 0 the rows of known ids are then overwritten, the LAST record of an id winning as in the Kafka consumer
 * (ALSKafkaConsumer.java:78-92); records of unknown ids are returned but not inserted.  fp64, Java operation order,
 * no fused multiply-add: rows are bit-identical to the Java arithmetic on the same fp64 model.  Missing mean vectors
 * are an error like the reference's IllegalStateException. */
FMS_API int fms_als_sgd_step(fms_ctx *ctx, const int64_t *user, const int64_t *item, const double *rating, int64_t n,
                             double learning_rate, double user_reg, double item_reg, const double *user_mean,
                             const double *item_mean, int apply, double *out_user, double *out_item);

/* --- N4: GPU-side ingest of the text the reference jobs read before training.  `text` may be a host or a device
 *     pointer; outputs likewise.  Call once with all output pointers NULL to size the buffers (*n_rows, *nnz). --- */
/* env.readCsvFile[(Int, Int, Double)](input, fieldDelimiter, ignoreFirstLine) (ALSImpl.scala:22-32): one
 * `user<delim>item<delim>rating` record per line (\n or \r\n), empty lines skipped, further columns ignored.
 * Integers as Integer.parseInt, doubles as Double.parseDouble for decimal literals (correctly rounded: Clinger fast
 * path + Eisel-Lemire).  A malformed line -> FMS_ERR_INVALID_ARG, a literal the parser cannot round exactly (exact
 * half-way cases, subnormals, hex floats) -> FMS_ERR_UNSUPPORTED, both naming the line; no CPU fallback. */
FMS_API int fms_parse_ratings_csv(fms_ctx *ctx, const char *text, int64_t nbytes, char delimiter, int ignore_first_line,
                                  int64_t *user, int64_t *item, double *rating, int64_t capacity, int64_t *n_rows);
/* env.readLibSVM(training) (SVMImpl.scala:21; FlinkML MLUtils.readLibSVM): `label idx:val idx:val ...`, '#' starts
 * a comment, blank lines skipped, indices 1-based in the file and 0-based in csr_idx, *dim = largest index seen.
 * Output is the CSR triple fms_svm_fit / fms_svm_load take (FMS_X_CSR_F64) plus the labels. */
FMS_API int fms_parse_libsvm(fms_ctx *ctx, const char *text, int64_t nbytes, double *label, int64_t row_capacity,
                             int64_t *csr_ptr, int32_t *csr_idx, double *csr_val, int64_t nnz_capacity, int64_t *n_rows,
                             int64_t *nnz, int32_t *dim);

/* --- step-level ALS API: data resident in HBM, used by bench.py (`value`) and by the one-process-per-GPU
 *     sharded driver (flink-ms_b200/dist.py), which all-gathers the factor shards between half-steps. --- */
/* Upload + index the ratings.  Every rank holds full factor replicas, the sorted distinct ids of both sides and the CSR of
 * the rows it SOLVES: [shard_index*S, (shard_index+1)*S) with S = ceil(N / shard_count) (rows of other shards are empty).  user / item / rating may be host OR
 * device pointers (a sharded host uploads 1/G of the triples per rank and all-gathers them over NVLink first). */
FMS_API int fms_als_load(fms_ctx *ctx, const int64_t *user, const int64_t *item, const double *rating, int64_t nnz,
                         int shard_index, int shard_count, /*out*/ int64_t *n_users, int64_t *n_items);
FMS_API int fms_als_init(fms_ctx *ctx, int rank, int64_t seed);
/* One updateFactors pass over this rank's shard: which = FMS_USERS solves users from items. */
FMS_API int fms_als_half_step(fms_ctx *ctx, int which, double lambda);
/* The same update restricted to chunk `chunk` of `n_chunks` equal row ranges of this rank's shard (c = ceil(S /
 * n_chunks) rows each, identical on every rank), so a sharded host can all-gather chunk k on a second stream while
 * chunk k+1 is being solved.  All chunks 0 .. n_chunks-1 together equal one fms_als_half_step. */
FMS_API int fms_als_half_step_chunk(fms_ctx *ctx, int which, double lambda, int chunk, int n_chunks);
/* Device view of a factor matrix for collectives: fp32 row-major, rows_padded = S * shard_count rows of
 * `rank` floats; shard rows are [shard_begin, shard_end). */
FMS_API int fms_als_factors_device(fms_ctx *ctx, int which, void **dev_ptr, int64_t *rows_padded,
                                   int64_t *shard_begin, int64_t *shard_end);
/* Tell the library a factor matrix was rewritten from outside (after an all-gather) so derived device
 * state (split-precision gather planes of the tensor-core kernel) is refreshed. */
FMS_API int fms_als_factors_updated(fms_ctx *ctx, int which);
/* Tensor-core path (rank 64): device view of the fp16 split-precision gather planes of a side.  The planes are
 * interleaved: rows_padded rows of 2 * rank halves, the `rank` hi halves of a factor row followed by its `rank` lo
 * halves (one 256-byte burst per gathered row); *hi is the base, *lo = *hi + rank halves, shard layout as for the
 * factor matrix.  A sharded host may all-gather the PLANES ([rows_padded, 2 * rank] fp16, one tensor) between
 * half-steps instead of the fp32 matrix (same bytes, no re-split afterwards) and call fms_als_planes_updated; it
 * must then all-gather the fp32 matrix once before reading factors out. */
FMS_API int fms_als_planes_device(fms_ctx *ctx, int which, void **hi, void **lo, int64_t *rows_padded);
/* Peer-to-peer exchange instead of a collective (one process per GPU on one NVLink node): every rank exports the CUDA
 * IPC handle (64 bytes) of its plane buffer for a side, the host gathers the handles in rank order and hands the table
 * to every rank.  From then on the half-step kernel stores each solved row's plane row into EVERY replica (posted
 * NVLink writes from the solver warps), so no all-gather follows the half-step: the host only needs a barrier across
 * ranks (and fms_als_planes_updated) before the next half-step.  n = 0 unmaps the peers again.  Mappings die with
 * the buffers: re-exchange handles after fms_als_load / fms_als_init with a different geometry. */
FMS_API int fms_als_planes_ipc_handle(fms_ctx *ctx, int which, void *handle64);
FMS_API int fms_als_planes_open_peers(fms_ctx *ctx, int which, const void *handles, int n);
FMS_API int fms_als_planes_updated(fms_ctx *ctx, int which);
/* With the replicas of BOTH sides mapped, the half-step kernels can order themselves across the ranks: the last CTA of a rank's
 * half-step publishes an epoch into flag words kept in every replica (system-scope release after the rank's remote stores), and a
 * half-step first waits until all ranks published the epoch of the side it gathers from.  enable != 0 switches that on; the host
 * then needs NO barrier or collective between half-steps (only fms_als_planes_updated).  Every rank must issue the same sequence
 * of half-step calls, and ONE host-side barrier across the ranks is required between fms_als_planes_open_peers and the first
 * half-step. */
FMS_API int fms_als_planes_flag_sync(fms_ctx *ctx, int enable);
/* number of replicas currently mapped for a side (0 = none): lets a host skip the handle exchange when the mappings of
 * an earlier fit are still valid (same plane buffers, same world size) */
FMS_API int fms_als_planes_peer_count(fms_ctx *ctx, int which, int *n_peers);
/* number of rows whose Cholesky hit a non-positive pivot in the last half-step (dposv info != 0; FlinkML
 * ignores it, so do we -- exposed for tests). */
FMS_API int fms_als_last_info(fms_ctx *ctx, int64_t *bad_rows);
FMS_API int fms_als_set_kernel(fms_ctx *ctx, int kernel);

/* ------------------------------------------------------------------------------------------------
 * SVM -- replaces `svm.fit(trainingDS)` (flink-svm/.../SVMImpl.scala:24-29) and `weightsOption`
 * (SVMImpl.scala:31-35).  CoCoA outer loop, SDCA local solver, hinge loss.
 * ---------------------------------------------------------------------------------------------- */
/* block_id[i] in [0, blocks): CoCoA block of row i (FlinkMLTools.block assigns by source subtask, an
 * input-split property, hence explicit).  local_iterations <= 0 => H = n_k (one epoch per round).
 * scaling = stepsize / blocks (configured count).  w_out: dim doubles. */
FMS_API int fms_svm_fit(fms_ctx *ctx, const void *X, int x_dtype, const int64_t *csr_ptr, const int32_t *csr_idx,
                        const double *y, int64_t n, int32_t dim, const int32_t *block_id, int blocks, int iterations,
                        int64_t local_iterations, double regularization, double stepsize, int64_t seed, int rng_mode,
                        /*out*/ double *w_out);
/* SVM.predict / serve-side SVMPredictRandom.java:68-88: raw = x.w (fp64); decision function or
 * (raw > threshold ? 1.0 : -1.0), STRICT '>'.  w == NULL uses the last fitted / installed weights. */
FMS_API int fms_svm_predict(fms_ctx *ctx, const void *X, int x_dtype, const int64_t *csr_ptr, const int32_t *csr_idx,
                            int64_t n, int32_t dim, const double *w, double threshold, int output_decision_function,
                            double *out);
FMS_API int fms_svm_set_weights(fms_ctx *ctx, const double *w, int32_t dim);
FMS_API int fms_svm_get_weights(fms_ctx *ctx, double *w, int32_t dim);
FMS_API int fms_svm_get_alpha(fms_ctx *ctx, double *alpha, int64_t n);

/* --- chunked upload for hosts whose buffers are capped (a Java direct ByteBuffer holds < 2 GiB; BASELINE configs[1] is 2.4 GB of
 *     triples, configs[3] 41 GB of rows): the same whole fits, fed in pieces -- begin, append x k, fit.  The pieces are staged in
 *     HBM as they arrive (ALS: the COO triples; SVM: dense fp32 rows + labels + block ids), the host never holds the whole input
 *     in one buffer, and the results are identical to the one-shot calls.  On a group ctx the triples are staged on the first
 *     device and the SVM rows in host memory owned by the library (they are scattered over the devices by block at fit time). --- */
FMS_API int fms_als_stream_begin(fms_ctx *ctx, int64_t nnz_total);
FMS_API int fms_als_stream_append(fms_ctx *ctx, const int64_t *user, const int64_t *item, const double *rating, int64_t count);
FMS_API int fms_als_stream_fit(fms_ctx *ctx, int rank, int iterations, double lambda, int64_t seed,
                               /*out*/ int64_t *n_users, int64_t *n_items);
/* X_rows: `rows` dense rows of `dim` values, x_dtype FMS_X_DENSE_F32 or FMS_X_DENSE_F64 (stored as fp32 on the device either way). */
FMS_API int fms_svm_stream_begin(fms_ctx *ctx, int64_t n_total, int32_t dim);
FMS_API int fms_svm_stream_append(fms_ctx *ctx, const void *X_rows, int x_dtype, const double *y, const int32_t *block_id, int64_t rows);
FMS_API int fms_svm_stream_fit(fms_ctx *ctx, int blocks, int iterations, int64_t local_iterations, double regularization,
                               double stepsize, int64_t seed, int rng_mode, /*out*/ double *w_out);

/* --- step-level SVM API (HBM-resident; sharded: this rank sweeps blocks k with k % shard_count ==
 *     shard_index, the host all-reduces the delta buffer, every rank applies it). --- */
/* X_is_device != 0: X (dense fp32 only) already lives on this ctx's GPU and is used in place. */
FMS_API int fms_svm_load(fms_ctx *ctx, const void *X, int x_dtype, int X_is_device, const int64_t *csr_ptr,
                         const int32_t *csr_idx, const double *y, int64_t n, int32_t dim, const int32_t *block_id,
                         int blocks, int shard_index, int shard_count, double regularization, double stepsize,
                         int64_t seed, int rng_mode);
/* localDualMethod over this rank's blocks: fills the delta buffer with sum_k dw_k (fp32, dim floats). */
FMS_API int fms_svm_sweep(fms_ctx *ctx, int64_t local_iterations);
FMS_API int fms_svm_delta_device(fms_ctx *ctx, void **dev_ptr, int32_t *dim);
/* w += (stepsize / blocks) * delta */
FMS_API int fms_svm_apply_delta(fms_ctx *ctx);
/* One process per GPU WITHOUT a collective library between the rounds (what fms_als_planes_* is for ALS): every rank maps every
 * rank's exchange buffer (CUDA IPC); fms_svm_round_exchange = fms_svm_sweep + ONE kernel that pushes the rank's delta-w into all
 * buffers over NVLink, publishes an epoch flag, waits for every rank's flag, sums in rank order and applies the step -- the
 * all-reduce and SVM.scala's `weights += scaling * sum` in one launch.  Protocol: every rank calls _ipc_handle(n_ranks), the host
 * exchanges the 64-byte handles, every rank calls _open_peers with all n handles in rank order, then ONE host barrier, then the
 * same number of fms_svm_round_exchange calls on every rank.  fms_svm_exchange_status reports a timed-out flag wait. */
FMS_API int fms_svm_exchange_ipc_handle(fms_ctx *ctx, int n_ranks, void *handle64);
FMS_API int fms_svm_exchange_open_peers(fms_ctx *ctx, const void *handles, int n_ranks, int my_rank);
FMS_API int fms_svm_round_exchange(fms_ctx *ctx, int64_t local_iterations);
FMS_API int fms_svm_exchange_status(fms_ctx *ctx, int *timed_out);
FMS_API int fms_svm_set_kernel(fms_ctx *ctx, int kernel);

/* ------------------------------------------------------------------------------------------------
 * java.util.Random on the device (bit-exact index selection): fills out[count] with the stream
 * `new Random(seed ^ block).nextInt(bound)` exactly as the SDCA kernel draws it.  Test hook.
 * ---------------------------------------------------------------------------------------------- */
FMS_API int fms_debug_index_stream(fms_ctx *ctx, int64_t seed, int block, int32_t bound, int64_t count, int32_t *out);
/* Test hook for the tensor-core ALS kernel (rank 64): runs it over rows [row_begin, row_end) of side `which` and
 * returns, per row, the 64x64 system matrix (Gramian + lambda*n on the diagonal, row-major) followed by the 64
 * right-hand-side values, exactly as the solver warps receive them.  Side effect: those rows are also solved. */
FMS_API int fms_debug_als_gramian(fms_ctx *ctx, int which, double lambda, int64_t row_begin, int64_t row_end, float *out);

#ifdef __cplusplus
}
#endif
#endif /* FLINKMS_H */
