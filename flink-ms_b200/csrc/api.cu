// api.cu -- the C-ABI of include/flinkms.h: argument checking, host<->device staging, kernel sequencing.
// No CPU fallback anywhere: every compute entry point needs the ctx's GPU.
#include <algorithm>
#include <new>

#include "common.cuh"

using namespace fms;

static thread_local std::string g_create_error;

#define CHECK_CTX(ctx) \
    do {               \
        if (!(ctx)) return FMS_ERR_INVALID_ARG; \
        cudaError_t e_ = cudaSetDevice((ctx)->device); \
        if (e_ != cudaSuccess) return fail((ctx), FMS_ERR_CUDA, "cudaSetDevice(%d): %s", (ctx)->device, cudaGetErrorString(e_)); \
    } while (0)

extern "C" {

static int als_fit_loaded(fms_ctx *ctx, int rank, int iterations, double lambda, int64_t seed, int64_t nnz);
static int svm_run_rounds(fms_ctx *ctx, int iterations, int64_t local_iterations, int32_t dim, double *w_out);

FMS_API int fms_version(void) { return 110; }

FMS_API int fms_create(fms_ctx **out, const int *device_ids, int n_devices) {
    if (!out) return FMS_ERR_INVALID_ARG;
    *out = nullptr;
    if (n_devices < 1 || n_devices > 64 || (n_devices > 1 && !device_ids)) {
        g_create_error = "fms_create: n_devices must be in 1 .. 64 (and device_ids given when it is > 1)";
        return FMS_ERR_INVALID_ARG;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        g_create_error = std::string("fms_create: no CUDA device (") + cudaGetErrorString(e) + ") -- libflinkms_b200 has no CPU fallback";
        cudaGetLastError();
        return FMS_ERR_CUDA;
    }
    if (n_devices > 1) {
        // a GROUP: one single-GPU member per device, sequenced by the parent (csrc/group.cu)
        fms_ctx *P = new (std::nothrow) fms_ctx();
        if (!P) return FMS_ERR_OOM;
        P->device = device_ids[0];
        const int rc = group_create(P, device_ids, n_devices);
        if (rc != FMS_OK) {
            g_create_error = P->err;
            group_destroy(P);
            delete P;
            return rc;
        }
        *out = P;
        return FMS_OK;
    }
    int dev = 0;
    if (device_ids) dev = device_ids[0];
    else if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    if (dev < 0 || dev >= count) {
        g_create_error = "fms_create: device id out of range";
        return FMS_ERR_INVALID_ARG;
    }
    if ((e = cudaSetDevice(dev)) != cudaSuccess) {
        g_create_error = std::string("fms_create: cudaSetDevice: ") + cudaGetErrorString(e);
        return FMS_ERR_CUDA;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess) {
        g_create_error = std::string("fms_create: cudaGetDeviceProperties: ") + cudaGetErrorString(e);
        return FMS_ERR_CUDA;
    }
    if (prop.major != 10) {
        g_create_error = "fms_create: libflinkms_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major) +
                         std::to_string(prop.minor);
        return FMS_ERR_UNSUPPORTED;
    }
    fms_ctx *ctx = new (std::nothrow) fms_ctx();
    if (!ctx) return FMS_ERR_OOM;
    ctx->device = dev;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = std::string("fms_create: cudaStreamCreate: ") + cudaGetErrorString(e);
        delete ctx;
        return FMS_ERR_CUDA;
    }
    *out = ctx;
    return FMS_OK;
}

FMS_API void fms_destroy(fms_ctx *ctx) {
    if (!ctx) return;
    if (!ctx->members.empty()) { group_destroy(ctx); delete ctx; return; }
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    als_tc_close_peers(ctx, 0);
    als_tc_close_peers(ctx, 1);
    if (ctx->stream) {
        cudaStreamDestroy(ctx->stream);
    }
    delete ctx;
}

FMS_API const char *fms_last_error(const fms_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
// group ctx: entry points that work on the fitted model are served by member 0 (every member holds the full model after a fit)
#define GROUP_FORWARD(ctx, call)                                         \
    do {                                                                 \
        if ((ctx) && !(ctx)->members.empty()) {                          \
            fms_ctx *m0 = (ctx)->members[0];                             \
            cudaSetDevice(m0->device);                                   \
            const int rc_ = (call);                                      \
            if (rc_ != FMS_OK) (ctx)->err = m0->err;                     \
            return rc_;                                                  \
        }                                                                \
    } while (0)
#define GROUP_UNSUPPORTED(ctx, name)                                     \
    do {                                                                 \
        if ((ctx) && !(ctx)->members.empty())                            \
            return fail((ctx), FMS_ERR_UNSUPPORTED, name ": step-level entry points need a single-device ctx (a group ctx runs whole fits)"); \
    } while (0)

FMS_API void *fms_stream(fms_ctx *ctx) { return !ctx ? nullptr : (ctx->members.empty() ? (void *)ctx->stream : (void *)ctx->members[0]->stream); }
FMS_API int64_t fms_launch_count(const fms_ctx *ctx) {
    if (!ctx) return 0;
    int64_t n = ctx->launches;
    for (const fms_ctx *m : ctx->members) n += m->launches;
    return n;
}
FMS_API int fms_synchronize(fms_ctx *ctx) {
    if (ctx && !ctx->members.empty()) {
        for (fms_ctx *m : ctx->members) {
            const int rc = fms_synchronize(m);
            if (rc != FMS_OK) { ctx->err = m->err; return rc; }
        }
        return FMS_OK;
    }
    CHECK_CTX(ctx);
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMS_OK;
}

// ====================================================================================================== ALS
FMS_API int fms_als_load(fms_ctx *ctx, const int64_t *user, const int64_t *item, const double *rating, int64_t nnz,
                         int shard_index, int shard_count, int64_t *n_users, int64_t *n_items) {
    GROUP_UNSUPPORTED(ctx, "fms_als_load");
    CHECK_CTX(ctx);
    if (nnz < 0 || (nnz > 0 && (!user || !item || !rating))) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_load: null input or nnz < 0");
    if (shard_count < 1 || shard_index < 0 || shard_index >= shard_count)
        return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_load: bad shard %d/%d", shard_index, shard_count);
    if (nnz >= ((int64_t)1 << 31) * 8) return fail(ctx, FMS_ERR_UNSUPPORTED, "fms_als_load: nnz too large");
    NvtxRange nvtx_("fms.als.load (H2D + createBlockInformation)");
    AlsState &A = ctx->als;
    A.loaded = A.inited = false;
    A.model_external = false;
    FMS_TRY(als_build_csr(ctx, user, item, rating, nnz, shard_index, shard_count));
    A.stage_rb[0] = A.stage_rb[1] = -1;                // stage layout of the tensor-core path is rebuilt lazily
    A.stage_zero_row[0] = A.stage_zero_row[1] = -1;
    A.shard_index = shard_index;
    A.shard_count = shard_count;
    A.shard_rows[0] = (A.n_users + shard_count - 1) / shard_count;
    A.shard_rows[1] = (A.n_items + shard_count - 1) / shard_count;
    A.loaded = true;
    if (n_users) *n_users = A.n_users;
    if (n_items) *n_items = A.n_items;
    return FMS_OK;
}

FMS_API int fms_als_init(fms_ctx *ctx, int rank, int64_t seed) {
    GROUP_UNSUPPORTED(ctx, "fms_als_init");
    CHECK_CTX(ctx);
    AlsState &A = ctx->als;
    if (!A.loaded) return fail(ctx, FMS_ERR_STATE, "fms_als_init: no ratings loaded");
    if (rank < 1) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_init: rank %d < 1", rank);
    if (rank > 128) return fail(ctx, FMS_ERR_UNSUPPORTED, "fms_als_init: rank %d > 128 is not supported", rank);
    A.rank = rank;
    FMS_TRY(als_launch_init(ctx, seed));
    A.inited = true;
    return FMS_OK;
}

FMS_API int fms_als_set_kernel(fms_ctx *ctx, int kernel) {
    if (ctx) for (fms_ctx *m : ctx->members) { const int rc_ = fms_als_set_kernel(m, kernel); if (rc_ != FMS_OK) { ctx->err = m->err; return rc_; } }
    if (!ctx) return FMS_ERR_INVALID_ARG;
    if (kernel < FMS_ALS_KERNEL_AUTO || kernel > FMS_ALS_KERNEL_TC_LDGSTS) return fail(ctx, FMS_ERR_INVALID_ARG, "bad ALS kernel id %d", kernel);
    ctx->als.kernel = kernel;
    return FMS_OK;
}

static int als_half_step_rows(fms_ctx *ctx, int which, double lambda, int chunk, int n_chunks, const char *fn) {
    AlsState &A = ctx->als;
    if (!A.loaded || !A.inited) return fail(ctx, FMS_ERR_STATE, "%s: load + init first", fn);
    if (which != FMS_USERS && which != FMS_ITEMS) return fail(ctx, FMS_ERR_INVALID_ARG, "%s: which = %d", fn, which);
    if (n_chunks < 1 || chunk < 0 || chunk >= n_chunks) return fail(ctx, FMS_ERR_INVALID_ARG, "%s: chunk %d of %d", fn, chunk, n_chunks);
    const int64_t n = which == FMS_USERS ? A.n_users : A.n_items;
    const int64_t S = A.shard_rows[which];
    const int64_t c = (S + n_chunks - 1) / n_chunks;                       // rows per chunk, the same on every rank
    const int64_t sb = S * A.shard_index;
    const int64_t rb = std::min<int64_t>(n, sb + std::min<int64_t>(S, c * chunk));
    const int64_t re = std::min<int64_t>(n, sb + std::min<int64_t>(S, c * (chunk + 1)));
    NvtxRange nvtx_(which == FMS_USERS ? "fms.als.half_step users" : "fms.als.half_step items");
    if (chunk == 0) FMS_CUDA(ctx, cudaMemsetAsync(A.info.p, 0, sizeof(int), ctx->stream));
    bool tc = false;
    if (A.kernel != FMS_ALS_KERNEL_GENERIC) tc = als_tc_supported(ctx, A.rank) != 0;
    if (A.kernel >= FMS_ALS_KERNEL_TC && !tc) return fail(ctx, FMS_ERR_UNSUPPORTED, "tensor-core ALS kernel does not cover rank %d", A.rank);
    A.fac64_valid = false;
    A.plane_valid[which] = false;                      // the tensor-core kernel re-validates the planes it writes itself
    if (tc) FMS_TRY(als_launch_half_step_tc(ctx, which, (float)lambda, rb, re));
    else FMS_TRY(als_launch_half_step_generic(ctx, which, (float)lambda, rb, re));
    A.last_bad = -1;
    return FMS_OK;
}

FMS_API int fms_als_half_step(fms_ctx *ctx, int which, double lambda) {
    GROUP_UNSUPPORTED(ctx, "fms_als_half_step");
    CHECK_CTX(ctx);
    return als_half_step_rows(ctx, which, lambda, 0, 1, "fms_als_half_step");
}

FMS_API int fms_als_half_step_chunk(fms_ctx *ctx, int which, double lambda, int chunk, int n_chunks) {
    GROUP_UNSUPPORTED(ctx, "fms_als_half_step_chunk");
    CHECK_CTX(ctx);
    return als_half_step_rows(ctx, which, lambda, chunk, n_chunks, "fms_als_half_step_chunk");
}

FMS_API int fms_als_last_info(fms_ctx *ctx, int64_t *bad_rows) {
    GROUP_FORWARD(ctx, fms_als_last_info(m0, bad_rows));
    CHECK_CTX(ctx);
    AlsState &A = ctx->als;
    if (!A.inited) return fail(ctx, FMS_ERR_STATE, "fms_als_last_info: no half-step run");
    int h = 0;
    FMS_CUDA(ctx, cudaMemcpyAsync(&h, A.info.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h >= (1 << 20)) {       // set by a half-step whose in-kernel wait for a peer rank's epoch flag timed out (fms_als_planes_flag_sync)
        if (bad_rows) *bad_rows = h & ((1 << 20) - 1);
        return fail(ctx, FMS_ERR_STATE, "ALS half-step: a peer rank never published its epoch (in-kernel flag wait timed out)");
    }
    if (bad_rows) *bad_rows = h;
    return FMS_OK;
}

FMS_API int fms_als_factors_device(fms_ctx *ctx, int which, void **dev_ptr, int64_t *rows_padded, int64_t *shard_begin,
                                   int64_t *shard_end) {
    GROUP_UNSUPPORTED(ctx, "fms_als_factors_device");
    if (!ctx) return FMS_ERR_INVALID_ARG;
    AlsState &A = ctx->als;
    if (!A.inited) return fail(ctx, FMS_ERR_STATE, "fms_als_factors_device: init first");
    if (which != 0 && which != 1) return fail(ctx, FMS_ERR_INVALID_ARG, "which = %d", which);
    const int64_t n = which == 0 ? A.n_users : A.n_items;
    if (dev_ptr) *dev_ptr = A.fac[which].p;
    if (rows_padded) *rows_padded = A.shard_rows[which] * A.shard_count;
    if (shard_begin) *shard_begin = std::min<int64_t>(n, A.shard_rows[which] * A.shard_index);
    if (shard_end) *shard_end = std::min<int64_t>(n, A.shard_rows[which] * (A.shard_index + 1));
    return FMS_OK;
}

FMS_API int fms_als_factors_updated(fms_ctx *ctx, int which) {
    GROUP_UNSUPPORTED(ctx, "fms_als_factors_updated");
    if (!ctx) return FMS_ERR_INVALID_ARG;
    if (which != 0 && which != 1) return fail(ctx, FMS_ERR_INVALID_ARG, "which = %d", which);
    ctx->als.fac64_valid = false;
    ctx->als.plane_valid[which] = false;
    return FMS_OK;
}

FMS_API int fms_als_planes_device(fms_ctx *ctx, int which, void **hi, void **lo, int64_t *rows_padded) {
    GROUP_UNSUPPORTED(ctx, "fms_als_planes_device");
    CHECK_CTX(ctx);
    AlsState &A = ctx->als;
    if (!A.inited) return fail(ctx, FMS_ERR_STATE, "fms_als_planes_device: init first");
    if (which != 0 && which != 1) return fail(ctx, FMS_ERR_INVALID_ARG, "which = %d", which);
    if (!als_tc_supported(ctx, A.rank)) return fail(ctx, FMS_ERR_UNSUPPORTED, "no gather planes at rank %d", A.rank);
    return als_tc_planes(ctx, which, hi, lo, rows_padded);
}

FMS_API int fms_als_planes_ipc_handle(fms_ctx *ctx, int which, void *handle64) {
    GROUP_UNSUPPORTED(ctx, "fms_als_planes_ipc_handle");
    CHECK_CTX(ctx);
    AlsState &A = ctx->als;
    if (!A.inited) return fail(ctx, FMS_ERR_STATE, "fms_als_planes_ipc_handle: init first");
    if ((which != FMS_USERS && which != FMS_ITEMS) || !handle64) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_planes_ipc_handle: bad arguments");
    if (!als_tc_supported(ctx, A.rank)) return fail(ctx, FMS_ERR_UNSUPPORTED, "no gather planes at rank %d", A.rank);
    return als_tc_ipc_handle(ctx, which, handle64);
}

FMS_API int fms_als_planes_open_peers(fms_ctx *ctx, int which, const void *handles, int n) {
    GROUP_UNSUPPORTED(ctx, "fms_als_planes_open_peers");
    CHECK_CTX(ctx);
    AlsState &A = ctx->als;
    if (!A.inited) return fail(ctx, FMS_ERR_STATE, "fms_als_planes_open_peers: init first");
    if (which != FMS_USERS && which != FMS_ITEMS) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_planes_open_peers: which = %d", which);
    if (!als_tc_supported(ctx, A.rank)) return fail(ctx, FMS_ERR_UNSUPPORTED, "no gather planes at rank %d", A.rank);
    if (n == 0) { als_tc_close_peers(ctx, which); return FMS_OK; }
    if (!handles) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_planes_open_peers: handles is null");
    return als_tc_open_peers(ctx, which, handles, n);
}

FMS_API int fms_als_planes_peer_count(fms_ctx *ctx, int which, int *n_peers) {
    GROUP_UNSUPPORTED(ctx, "fms_als_planes_peer_count");
    if (!ctx) return FMS_ERR_INVALID_ARG;
    if ((which != 0 && which != 1) || !n_peers) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_planes_peer_count: bad arguments");
    *n_peers = ctx->als.peer_n[which];
    return FMS_OK;
}

FMS_API int fms_als_planes_flag_sync(fms_ctx *ctx, int enable) {
    GROUP_UNSUPPORTED(ctx, "fms_als_planes_flag_sync");
    if (!ctx) return FMS_ERR_INVALID_ARG;
    ctx->als.flag_sync = enable != 0;
    return FMS_OK;
}

FMS_API int fms_als_planes_updated(fms_ctx *ctx, int which) {
    GROUP_UNSUPPORTED(ctx, "fms_als_planes_updated");
    if (!ctx) return FMS_ERR_INVALID_ARG;
    if (which != 0 && which != 1) return fail(ctx, FMS_ERR_INVALID_ARG, "which = %d", which);
    ctx->als.plane_valid[which] = true;      // every shard of the planes is current; the fp32 matrix may still be stale
    return FMS_OK;
}

FMS_API int fms_als_fit(fms_ctx *ctx, const int64_t *user, const int64_t *item, const double *rating, int64_t nnz, int rank,
                        int iterations, double lambda, int64_t seed, int64_t *n_users, int64_t *n_items) {
    if (ctx && !ctx->members.empty()) {
        if (iterations < 0) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_fit: iterations %d < 0", iterations);
        if (rank < 1) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_fit: rank %d < 1", rank);
        return group_als_fit(ctx, user, item, rating, nnz, rank, iterations, lambda, seed, n_users, n_items);
    }
    CHECK_CTX(ctx);
    if (iterations < 0) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_fit: iterations %d < 0", iterations);
    if (rank < 1) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_fit: rank %d < 1", rank);
    NvtxRange nvtx_("fms.als.fit");
    // FMS_LOG_JSON=1: one JSON line per iteration on stderr (costs a stream synchronisation per iteration, off by default)
    FMS_TRY(fms_als_load(ctx, user, item, rating, nnz, 0, 1, n_users, n_items));
    return als_fit_loaded(ctx, rank, iterations, lambda, seed, nnz);
}

// init + iterate(T){ users; items } + final users on the ratings already indexed in the ctx
static int als_fit_loaded(fms_ctx *ctx, int rank, int iterations, double lambda, int64_t seed, int64_t nnz) {
    static const bool log_json = getenv("FMS_LOG_JSON") != nullptr;
    FMS_TRY(fms_als_init(ctx, rank, seed));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (log_json) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
    for (int t = 0; t < iterations; ++t) {
        if (log_json) cudaEventRecord(e0, ctx->stream);
        FMS_TRY(fms_als_half_step(ctx, FMS_USERS, lambda));
        FMS_TRY(fms_als_half_step(ctx, FMS_ITEMS, lambda));
        if (log_json) {
            cudaEventRecord(e1, ctx->stream);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            int64_t bad = 0;
            fms_als_last_info(ctx, &bad);
            fprintf(stderr, "{\"event\": \"als_iteration\", \"iteration\": %d, \"ms\": %.4f, \"ratings_per_s\": %.4g, \"rank\": %d, \"bad_rows\": %lld}\n", t, ms,
                    ms > 0 ? (double)nnz / (ms * 1e-3) : 0.0, rank, (long long)bad);
        }
    }
    if (log_json) { cudaEventDestroy(e0); cudaEventDestroy(e1); }
    FMS_TRY(fms_als_half_step(ctx, FMS_USERS, lambda));   // the extra user half-step after the loop
    // dposv's info: FlinkML ignores it and so does the fit, but a model with broken-down rows (NaN factors: a system that
    // is not positive definite, or |rating| / |factor| beyond what the kernel's operand format holds) must not leave
    // silently -- the count stays readable through fms_als_last_info and the message through fms_last_error.
    int bad = 0;
    FMS_CUDA(ctx, cudaMemcpyAsync(&bad, ctx->als.info.p, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (bad) fail(ctx, FMS_OK, "fms_als_fit: %d row(s) of the last half-step broke down (non-positive pivot, dposv info != 0) or left the +-65504 range of the rank-64 kernel's fp16 gather planes (fms_als_set_kernel(FMS_ALS_KERNEL_GENERIC) has no such range)", bad);
    else ctx->err.clear();
    return FMS_OK;
}

FMS_API int fms_als_get_factors(fms_ctx *ctx, int which, int64_t *ids, double *factors) {
    GROUP_FORWARD(ctx, fms_als_get_factors(m0, which, ids, factors));
    CHECK_CTX(ctx);
    AlsState &A = ctx->als;
    if (!A.inited && !A.model_external) return fail(ctx, FMS_ERR_STATE, "The ALS model has not been fitted.");
    if (which != 0 && which != 1) return fail(ctx, FMS_ERR_INVALID_ARG, "which = %d", which);
    const int64_t n = which == 0 ? A.n_users : A.n_items;
    if (ids && n) FMS_CUDA(ctx, cudaMemcpyAsync(ids, A.ids[which].p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (factors && n) {
        FMS_TRY(als_ensure_fac64(ctx));
        FMS_CUDA(ctx, cudaMemcpyAsync(factors, A.fac64[which].p, (size_t)n * A.rank * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMS_OK;
}

FMS_API int fms_als_get_factor_rows(fms_ctx *ctx, int which, int64_t row_begin, int64_t row_end, int64_t *ids, double *factors) {
    GROUP_FORWARD(ctx, fms_als_get_factor_rows(m0, which, row_begin, row_end, ids, factors));
    CHECK_CTX(ctx);
    AlsState &A = ctx->als;
    if (!A.inited && !A.model_external) return fail(ctx, FMS_ERR_STATE, "The ALS model has not been fitted.");
    if (which != 0 && which != 1) return fail(ctx, FMS_ERR_INVALID_ARG, "which = %d", which);
    const int64_t n = which == 0 ? A.n_users : A.n_items;
    if (row_begin < 0 || row_end < row_begin || row_end > n) return fail(ctx, FMS_ERR_INVALID_ARG, "rows [%lld, %lld) of %lld", (long long)row_begin, (long long)row_end, (long long)n);
    const int64_t m = row_end - row_begin;
    if (ids && m) FMS_CUDA(ctx, cudaMemcpyAsync(ids, A.ids[which].p + row_begin, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (factors && m) {
        FMS_TRY(als_ensure_fac64(ctx));
        FMS_CUDA(ctx, cudaMemcpyAsync(factors, A.fac64[which].p + row_begin * A.rank, (size_t)m * A.rank * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMS_OK;
}

FMS_API int fms_als_set_model(fms_ctx *ctx, const int64_t *user_ids, const double *user_factors, int64_t n_users,
                              const int64_t *item_ids, const double *item_factors, int64_t n_items, int rank) {
    GROUP_FORWARD(ctx, fms_als_set_model(m0, user_ids, user_factors, n_users, item_ids, item_factors, n_items, rank));
    CHECK_CTX(ctx);
    if (n_users < 0 || n_items < 0 || rank < 1) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_set_model: bad sizes");
    if ((n_users && (!user_ids || !user_factors)) || (n_items && (!item_ids || !item_factors)))
        return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_set_model: null input");
    for (int64_t k = 1; k < n_users; ++k) if (user_ids[k] <= user_ids[k - 1]) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_set_model: user ids must be strictly ascending");
    for (int64_t k = 1; k < n_items; ++k) if (item_ids[k] <= item_ids[k - 1]) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_set_model: item ids must be strictly ascending");
    AlsState &A = ctx->als;
    A.loaded = A.inited = false;
    A.rank = rank;
    A.n_users = n_users;
    A.n_items = n_items;
    const int64_t n[2] = {n_users, n_items};
    const int64_t *idp[2] = {user_ids, item_ids};
    const double *fp[2] = {user_factors, item_factors};
    for (int s = 0; s < 2; ++s) {
        FMS_CUDA(ctx, A.ids[s].alloc((size_t)n[s]));
        FMS_CUDA(ctx, A.fac64[s].alloc((size_t)n[s] * rank));
        if (n[s]) {
            FMS_CUDA(ctx, cudaMemcpyAsync(A.ids[s].p, idp[s], (size_t)n[s] * 8, cudaMemcpyHostToDevice, ctx->stream));
            FMS_CUDA(ctx, cudaMemcpyAsync(A.fac64[s].p, fp[s], (size_t)n[s] * rank * 8, cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    A.fac64_valid = true;
    A.model_external = true;
    return FMS_OK;
}

__global__ void k_sq_err(const double *__restrict__ pred, const uint8_t *__restrict__ found, const double *__restrict__ rating,
                         int64_t n, double *__restrict__ acc /* [0]=sum, [1]=count */) {
    double s = 0.0, c = 0.0;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        if (found[k]) { const double e = rating[k] - pred[k]; s += e * e; c += 1.0; }
    s = warp_sum(s);
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&acc[0], s); atomicAdd(&acc[1], c); }
}

static int predict_common(fms_ctx *ctx, const int64_t *u, const int64_t *i, int64_t n, DevBuf<int64_t> &du, DevBuf<int64_t> &di,
                          DevBuf<double> &dout, DevBuf<uint8_t> &dfound) {
    AlsState &A = ctx->als;
    if (!A.inited && !A.model_external) return fail(ctx, FMS_ERR_STATE, "The ALS model has not been fitted.");
    if (n < 0 || (n && (!u || !i))) return fail(ctx, FMS_ERR_INVALID_ARG, "predict: null input or n < 0");
    FMS_TRY(als_ensure_fac64(ctx));
    FMS_CUDA(ctx, du.alloc((size_t)n));
    FMS_CUDA(ctx, di.alloc((size_t)n));
    FMS_CUDA(ctx, dout.alloc((size_t)n));
    FMS_CUDA(ctx, dfound.alloc((size_t)n));
    if (n) {
        FMS_CUDA(ctx, cudaMemcpyAsync(du.p, u, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        FMS_CUDA(ctx, cudaMemcpyAsync(di.p, i, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    return als_predict_device(ctx, du.p, di.p, n, dout.p, dfound.p);
}

FMS_API int fms_als_predict(fms_ctx *ctx, const int64_t *u, const int64_t *i, int64_t n, double *out, uint8_t *found) {
    GROUP_FORWARD(ctx, fms_als_predict(m0, u, i, n, out, found));
    CHECK_CTX(ctx);
    if (n && !out) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_predict: out is null");
    DevBuf<int64_t> &du = ctx->als.q_u, &di = ctx->als.q_i;
    DevBuf<double> &dout = ctx->als.q_out;
    DevBuf<uint8_t> &dfound = ctx->als.q_found;
    FMS_TRY(predict_common(ctx, u, i, n, du, di, dout, dfound));
    if (n) {
        FMS_CUDA(ctx, cudaMemcpyAsync(out, dout.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (found) FMS_CUDA(ctx, cudaMemcpyAsync(found, dfound.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMS_OK;
}

FMS_API int fms_als_mse(fms_ctx *ctx, const int64_t *u, const int64_t *i, const double *rating, int64_t n, double *mse,
                        int64_t *n_used) {
    GROUP_FORWARD(ctx, fms_als_mse(m0, u, i, rating, n, mse, n_used));
    CHECK_CTX(ctx);
    if (n && !rating) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_mse: rating is null");
    DevBuf<int64_t> &du = ctx->als.q_u, &di = ctx->als.q_i;
    DevBuf<double> &dout = ctx->als.q_out, &dr = ctx->als.q_rating;
    DevBuf<double> acc;
    DevBuf<uint8_t> &dfound = ctx->als.q_found;
    FMS_TRY(predict_common(ctx, u, i, n, du, di, dout, dfound));
    FMS_CUDA(ctx, dr.alloc((size_t)n));
    FMS_CUDA(ctx, acc.alloc(2));
    FMS_CUDA(ctx, cudaMemsetAsync(acc.p, 0, 16, ctx->stream));
    double h[2] = {0, 0};
    if (n) {
        FMS_CUDA(ctx, cudaMemcpyAsync(dr.p, rating, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        k_sq_err<<<(unsigned)std::min<int64_t>((n + 255) / 256, 4096), 256, 0, ctx->stream>>>(dout.p, dfound.p, dr.p, n, acc.p);
        FMS_LAUNCH_CHECK(ctx);
        FMS_CUDA(ctx, cudaMemcpyAsync(h, acc.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    }
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (mse) *mse = h[1] > 0 ? h[0] / h[1] : 0.0;
    if (n_used) *n_used = (int64_t)h[1];
    return FMS_OK;
}

FMS_API int fms_als_sgd_step(fms_ctx *ctx, const int64_t *user, const int64_t *item, const double *rating, int64_t n, double learning_rate,
                             double user_reg, double item_reg, const double *user_mean, const double *item_mean, int apply,
                             double *out_user, double *out_item) {
    GROUP_FORWARD(ctx, fms_als_sgd_step(m0, user, item, rating, n, learning_rate, user_reg, item_reg, user_mean, item_mean, apply, out_user, out_item));
    CHECK_CTX(ctx);
    AlsState &A = ctx->als;
    if (!A.inited && !A.model_external) return fail(ctx, FMS_ERR_STATE, "The ALS model has not been fitted.");
    if (n < 0 || (n && (!user || !item || !rating || !out_user || !out_item))) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_sgd_step: null argument or n < 0");
    if (!user_mean || !item_mean) return fail(ctx, FMS_ERR_INVALID_ARG, "Unable to load User mean or item mean factors.");   // SGD.java:150-152
    if (n == 0) return FMS_OK;
    FMS_TRY(als_ensure_fac64(ctx));
    const int d = A.rank;
    DevBuf<int64_t> du, di;
    DevBuf<double> dr, dmean, dou, doi;
    FMS_CUDA(ctx, du.alloc((size_t)n));
    FMS_CUDA(ctx, di.alloc((size_t)n));
    FMS_CUDA(ctx, dr.alloc((size_t)n));
    FMS_CUDA(ctx, dmean.alloc((size_t)2 * d));
    FMS_CUDA(ctx, dou.alloc((size_t)n * d));
    FMS_CUDA(ctx, doi.alloc((size_t)n * d));
    FMS_CUDA(ctx, cudaMemcpyAsync(du.p, user, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMS_CUDA(ctx, cudaMemcpyAsync(di.p, item, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMS_CUDA(ctx, cudaMemcpyAsync(dr.p, rating, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMS_CUDA(ctx, cudaMemcpyAsync(dmean.p, user_mean, (size_t)d * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMS_CUDA(ctx, cudaMemcpyAsync(dmean.p + d, item_mean, (size_t)d * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMS_TRY(als_sgd_step_device(ctx, du.p, di.p, dr.p, n, learning_rate, user_reg, item_reg, dmean.p, dmean.p + d, apply, dou.p, doi.p));
    FMS_CUDA(ctx, cudaMemcpyAsync(out_user, dou.p, (size_t)n * d * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMS_CUDA(ctx, cudaMemcpyAsync(out_item, doi.p, (size_t)n * d * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMS_OK;
}

// ====================================================================================================== N4 ingest
FMS_API int fms_parse_ratings_csv(fms_ctx *ctx, const char *text, int64_t nbytes, char delimiter, int ignore_first_line, int64_t *user,
                                  int64_t *item, double *rating, int64_t capacity, int64_t *n_rows) {
    GROUP_FORWARD(ctx, fms_parse_ratings_csv(m0, text, nbytes, delimiter, ignore_first_line, user, item, rating, capacity, n_rows));
    CHECK_CTX(ctx);
    if (nbytes < 0 || (nbytes && !text) || capacity < 0) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_parse_ratings_csv: bad arguments");
    if (delimiter == '\n' || delimiter == '\r') return fail(ctx, FMS_ERR_INVALID_ARG, "fms_parse_ratings_csv: the field delimiter cannot be a line break");
    return ingest_ratings_csv(ctx, text, nbytes, delimiter, ignore_first_line, user, item, rating, capacity, n_rows);
}

FMS_API int fms_parse_libsvm(fms_ctx *ctx, const char *text, int64_t nbytes, double *label, int64_t row_capacity, int64_t *csr_ptr,
                             int32_t *csr_idx, double *csr_val, int64_t nnz_capacity, int64_t *n_rows, int64_t *nnz, int32_t *dim) {
    GROUP_FORWARD(ctx, fms_parse_libsvm(m0, text, nbytes, label, row_capacity, csr_ptr, csr_idx, csr_val, nnz_capacity, n_rows, nnz, dim));
    CHECK_CTX(ctx);
    if (nbytes < 0 || (nbytes && !text) || row_capacity < 0 || nnz_capacity < 0) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_parse_libsvm: bad arguments");
    return ingest_libsvm(ctx, text, nbytes, label, row_capacity, csr_ptr, csr_idx, csr_val, nnz_capacity, n_rows, nnz, dim);
}

// ====================================================================================================== SVM
__global__ void k_f64_to_f32_chunk(const double *__restrict__ in, int64_t n, float *__restrict__ out) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) out[k] = (float)in[k];
}

static int upload_f64_as_f32(fms_ctx *ctx, const double *h, int64_t n, float *d_out) {
    const int64_t CH = (int64_t)1 << 24;   // 128 MB staging
    DevBuf<double> stage;
    FMS_CUDA(ctx, stage.alloc((size_t)std::min<int64_t>(n, CH)));
    for (int64_t o = 0; o < n; o += CH) {
        const int64_t c = std::min<int64_t>(CH, n - o);
        FMS_CUDA(ctx, cudaMemcpyAsync(stage.p, h + o, (size_t)c * 8, cudaMemcpyHostToDevice, ctx->stream));
        k_f64_to_f32_chunk<<<(unsigned)((c + 255) / 256), 256, 0, ctx->stream>>>(stage.p, c, d_out + o);
        FMS_LAUNCH_CHECK(ctx);
        FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return FMS_OK;
}

FMS_API int fms_svm_set_kernel(fms_ctx *ctx, int kernel) {
    if (ctx) for (fms_ctx *m : ctx->members) { const int rc_ = fms_svm_set_kernel(m, kernel); if (rc_ != FMS_OK) { ctx->err = m->err; return rc_; } }
    if (!ctx) return FMS_ERR_INVALID_ARG;
    if (kernel < FMS_SVM_KERNEL_AUTO || kernel > FMS_SVM_KERNEL_WARP) return fail(ctx, FMS_ERR_INVALID_ARG, "bad SVM kernel id %d", kernel);
    ctx->svm.kernel = kernel;
    return FMS_OK;
}

FMS_API int fms_svm_load(fms_ctx *ctx, const void *X, int x_dtype, int X_is_device, const int64_t *csr_ptr, const int32_t *csr_idx,
                         const double *y, int64_t n, int32_t dim, const int32_t *block_id, int blocks, int shard_index,
                         int shard_count, double regularization, double stepsize, int64_t seed, int rng_mode) {
    GROUP_UNSUPPORTED(ctx, "fms_svm_load");
    CHECK_CTX(ctx);
    NvtxRange nvtx_("fms.svm.load (group rows by block, H2D)");
    SvmState &S = ctx->svm;
    S.loaded = false;
    if (n < 0 || dim < 1 || blocks < 1) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_load: n=%lld dim=%d blocks=%d", (long long)n, dim, blocks);
    if (n > 0 && (!X || !y || !block_id)) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_load: null input");
    if (x_dtype < FMS_X_DENSE_F32 || x_dtype > FMS_X_CSR_F64) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_load: x_dtype %d", x_dtype);
    if (x_dtype == FMS_X_CSR_F64 && n > 0 && (!csr_ptr || !csr_idx)) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_load: CSR arrays missing");
    if (X_is_device && x_dtype != FMS_X_DENSE_F32) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_load: device X must be dense fp32");
    if (shard_count < 1 || shard_index < 0 || shard_index >= shard_count) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_load: bad shard");
    if (rng_mode != FMS_RNG_PER_BLOCK && rng_mode != FMS_RNG_SHARED_SUBTASK) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_load: rng_mode %d", rng_mode);
    if (n >= ((int64_t)1 << 31)) return fail(ctx, FMS_ERR_UNSUPPORTED, "fms_svm_load: n >= 2^31 (Random.nextInt bound is an int)");
    if (!(regularization > 0.0)) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_load: regularization must be > 0");
    cudaStream_t st = ctx->stream;
    S.n = n; S.n_total = n; S.dim = dim; S.blocks = blocks; S.shard_index = shard_index; S.shard_count = shard_count;
    S.lambda = regularization; S.stepsize = stepsize; S.seed = seed; S.rng_mode = rng_mode;
    // ---- group rows by block (arrival order inside a block)
    std::vector<int64_t> cnt((size_t)blocks + 1, 0);
    for (int64_t i = 0; i < n; ++i) {
        if (block_id[i] < 0 || block_id[i] >= blocks) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_load: block_id[%lld]=%d out of range", (long long)i, block_id[i]);
        cnt[(size_t)block_id[i] + 1]++;
    }
    for (int k = 0; k < blocks; ++k) cnt[(size_t)k + 1] += cnt[(size_t)k];
    std::vector<int64_t> rows((size_t)n), fill(cnt.begin(), cnt.end() - 1);
    for (int64_t i = 0; i < n; ++i) rows[(size_t)fill[(size_t)block_id[i]]++] = i;
    S.h_blk_ptr = cnt;
    FMS_CUDA(ctx, S.rows.alloc((size_t)n));
    FMS_CUDA(ctx, S.blk_ptr.alloc((size_t)blocks + 1));
    if (n) FMS_CUDA(ctx, cudaMemcpyAsync(S.rows.p, rows.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    std::vector<int32_t> rows32(rows.begin(), rows.end());
    FMS_CUDA(ctx, S.rows32.alloc((size_t)n));
    if (n) FMS_CUDA(ctx, cudaMemcpyAsync(S.rows32.p, rows32.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    FMS_CUDA(ctx, cudaMemcpyAsync(S.blk_ptr.p, cnt.data(), ((size_t)blocks + 1) * 8, cudaMemcpyHostToDevice, st));
    S.local_blocks.clear();
    for (int k = 0; k < blocks; ++k) if (k % shard_count == shard_index) S.local_blocks.push_back(k);
    if (rng_mode == FMS_RNG_SHARED_SUBTASK) { S.local_blocks.clear(); for (int k = 0; k < blocks; ++k) S.local_blocks.push_back(k); }
    const size_t n_local = S.local_blocks.size();
    FMS_CUDA(ctx, S.d_local_blocks.alloc(n_local));
    if (n_local) FMS_CUDA(ctx, cudaMemcpyAsync(S.d_local_blocks.p, S.local_blocks.data(), n_local * 4, cudaMemcpyHostToDevice, st));
    std::vector<uint64_t> rng((size_t)blocks);
    for (int k = 0; k < blocks; ++k) rng[(size_t)k] = JRand::scramble((long long)(seed ^ (int64_t)k));
    FMS_CUDA(ctx, S.rng.alloc((size_t)blocks));
    FMS_CUDA(ctx, cudaMemcpyAsync(S.rng.p, rng.data(), (size_t)blocks * 8, cudaMemcpyHostToDevice, st));
    // ---- labels
    FMS_CUDA(ctx, S.y.alloc((size_t)n));
    if (n) FMS_TRY(upload_f64_as_f32(ctx, y, n, S.y.p));
    // ---- design matrix
    S.X = nullptr;
    if (x_dtype == FMS_X_CSR_F64) {
        S.x_kind = FMS_X_CSR_F64;
        const int64_t nnz = n ? csr_ptr[n] : 0;
        for (int64_t p = 0; p < nnz; ++p) if (csr_idx[p] < 0 || csr_idx[p] >= dim) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_load: csr_idx[%lld]=%d outside [0,%d)", (long long)p, csr_idx[p], dim);
        FMS_CUDA(ctx, S.csr_ptr.alloc((size_t)n + 1));
        FMS_CUDA(ctx, S.csr_idx.alloc((size_t)nnz));
        FMS_CUDA(ctx, S.csr_val.alloc((size_t)nnz));
        if (n) FMS_CUDA(ctx, cudaMemcpyAsync(S.csr_ptr.p, csr_ptr, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
        if (nnz) {
            FMS_CUDA(ctx, cudaMemcpyAsync(S.csr_idx.p, csr_idx, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
            FMS_TRY(upload_f64_as_f32(ctx, (const double *)X, nnz, S.csr_val.p));
        }
        S.X_own.release();
    } else {
        S.x_kind = FMS_X_DENSE_F32;
        const int64_t cnt_x = n * (int64_t)dim;
        if (X_is_device) {
            S.X = (const float *)X;
            S.X_own.release();
        } else {
            FMS_CUDA(ctx, S.X_own.alloc((size_t)cnt_x));
            if (cnt_x) {
                if (x_dtype == FMS_X_DENSE_F32) FMS_CUDA(ctx, cudaMemcpyAsync(S.X_own.p, X, (size_t)cnt_x * 4, cudaMemcpyHostToDevice, st));
                else FMS_TRY(upload_f64_as_f32(ctx, (const double *)X, cnt_x, S.X_own.p));
            }
            S.X = S.X_own.p;
        }
    }
    // ---- state
    FMS_CUDA(ctx, S.alpha.alloc((size_t)n));
    FMS_CUDA(ctx, cudaMemsetAsync(S.alpha.p, 0, (size_t)std::max<int64_t>(n, 1) * 4, st));
    FMS_CUDA(ctx, S.w.alloc((size_t)dim));
    FMS_CUDA(ctx, cudaMemsetAsync(S.w.p, 0, (size_t)dim * 4, st));
    FMS_CUDA(ctx, S.delta.alloc((size_t)dim));
    FMS_CUDA(ctx, cudaMemsetAsync(S.delta.p, 0, (size_t)dim * 4, st));
    FMS_CUDA(ctx, S.dw_blocks.alloc(std::max<size_t>(n_local, 1) * (size_t)dim));
    FMS_CUDA(ctx, cudaStreamSynchronize(st));
    S.w64_valid = false;
    S.loaded = true;
    return FMS_OK;
}

FMS_API int fms_svm_sweep(fms_ctx *ctx, int64_t local_iterations) {
    GROUP_UNSUPPORTED(ctx, "fms_svm_sweep");
    CHECK_CTX(ctx);
    if (!ctx->svm.loaded) return fail(ctx, FMS_ERR_STATE, "fms_svm_sweep: load first");
    return svm_launch_sweep(ctx, local_iterations);
}

FMS_API int fms_svm_delta_device(fms_ctx *ctx, void **dev_ptr, int32_t *dim) {
    GROUP_UNSUPPORTED(ctx, "fms_svm_delta_device");
    if (!ctx) return FMS_ERR_INVALID_ARG;
    if (!ctx->svm.loaded) return fail(ctx, FMS_ERR_STATE, "fms_svm_delta_device: load first");
    if (dev_ptr) *dev_ptr = ctx->svm.delta.p;
    if (dim) *dim = ctx->svm.dim;
    return FMS_OK;
}

FMS_API int fms_svm_apply_delta(fms_ctx *ctx) {
    GROUP_UNSUPPORTED(ctx, "fms_svm_apply_delta");
    CHECK_CTX(ctx);
    if (!ctx->svm.loaded) return fail(ctx, FMS_ERR_STATE, "fms_svm_apply_delta: load first");
    return svm_launch_apply(ctx);
}

FMS_API int fms_svm_exchange_ipc_handle(fms_ctx *ctx, int n_ranks, void *handle64) {
    GROUP_UNSUPPORTED(ctx, "fms_svm_exchange_ipc_handle");
    CHECK_CTX(ctx);
    return svm_exchange_handle(ctx, n_ranks, handle64);
}

FMS_API int fms_svm_exchange_open_peers(fms_ctx *ctx, const void *handles, int n_ranks, int my_rank) {
    GROUP_UNSUPPORTED(ctx, "fms_svm_exchange_open_peers");
    CHECK_CTX(ctx);
    return svm_exchange_open(ctx, handles, n_ranks, my_rank);
}

FMS_API int fms_svm_round_exchange(fms_ctx *ctx, int64_t local_iterations) {
    GROUP_UNSUPPORTED(ctx, "fms_svm_round_exchange");
    CHECK_CTX(ctx);
    if (!ctx->svm.loaded) return fail(ctx, FMS_ERR_STATE, "fms_svm_round_exchange: load first");
    NvtxRange nvtx_("fms.svm.round_exchange");
    FMS_TRY(svm_launch_sweep(ctx, local_iterations));
    return svm_launch_exchange_apply(ctx);
}

FMS_API int fms_svm_exchange_status(fms_ctx *ctx, int *timed_out) {
    GROUP_UNSUPPORTED(ctx, "fms_svm_exchange_status");
    CHECK_CTX(ctx);
    int v = 0;
    if (ctx->svm.xchg_err.p) {
        FMS_CUDA(ctx, cudaMemcpyAsync(&v, ctx->svm.xchg_err.p, sizeof v, cudaMemcpyDeviceToHost, ctx->stream));
        FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (timed_out) *timed_out = v;
    if (v) return fail(ctx, FMS_ERR_STATE, "fms_svm_round_exchange: a peer rank never published its epoch (flag wait timed out)");
    return FMS_OK;
}

__global__ void k_widen(const float *__restrict__ in, int64_t n, double *__restrict__ out) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) out[k] = (double)in[k];
}

static int svm_ensure_w64(fms_ctx *ctx) {
    SvmState &S = ctx->svm;
    if (S.w64_valid) return FMS_OK;
    if (!S.loaded) return fail(ctx, FMS_ERR_STATE, "The SVM model has not been trained.");
    FMS_CUDA(ctx, S.w64.alloc((size_t)S.dim));
    k_widen<<<(S.dim + 255) / 256, 256, 0, ctx->stream>>>(S.w.p, S.dim, S.w64.p);
    FMS_LAUNCH_CHECK(ctx);
    S.w64_valid = true;
    S.w64_dim = S.dim;
    return FMS_OK;
}

FMS_API int fms_svm_get_weights(fms_ctx *ctx, double *w, int32_t dim) {
    GROUP_FORWARD(ctx, fms_svm_get_weights(m0, w, dim));
    CHECK_CTX(ctx);
    SvmState &S = ctx->svm;
    FMS_TRY(svm_ensure_w64(ctx));
    if (!w || dim != S.w64_dim) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_get_weights: dim %d != model dim %d", dim, S.w64_dim);
    FMS_CUDA(ctx, cudaMemcpyAsync(w, S.w64.p, (size_t)dim * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMS_OK;
}

FMS_API int fms_svm_set_weights(fms_ctx *ctx, const double *w, int32_t dim) {
    GROUP_FORWARD(ctx, fms_svm_set_weights(m0, w, dim));
    CHECK_CTX(ctx);
    if (!w || dim < 1) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_set_weights: null or dim < 1");
    SvmState &S = ctx->svm;
    FMS_CUDA(ctx, S.w64.alloc((size_t)dim));
    FMS_CUDA(ctx, cudaMemcpyAsync(S.w64.p, w, (size_t)dim * 8, cudaMemcpyHostToDevice, ctx->stream));
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    S.w64_valid = true;
    S.w64_dim = dim;
    return FMS_OK;
}

FMS_API int fms_svm_get_alpha(fms_ctx *ctx, double *alpha, int64_t n) {
    GROUP_UNSUPPORTED(ctx, "fms_svm_get_alpha");
    CHECK_CTX(ctx);
    SvmState &S = ctx->svm;
    if (!S.loaded) return fail(ctx, FMS_ERR_STATE, "fms_svm_get_alpha: load first");
    if (n != S.n || (n && !alpha)) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_get_alpha: n mismatch");
    if (!n) return FMS_OK;
    DevBuf<double> tmp;
    FMS_CUDA(ctx, tmp.alloc((size_t)n));
    k_widen<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(S.alpha.p, n, tmp.p);
    FMS_LAUNCH_CHECK(ctx);
    FMS_CUDA(ctx, cudaMemcpyAsync(alpha, tmp.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMS_OK;
}

FMS_API int fms_svm_fit(fms_ctx *ctx, const void *X, int x_dtype, const int64_t *csr_ptr, const int32_t *csr_idx, const double *y,
                        int64_t n, int32_t dim, const int32_t *block_id, int blocks, int iterations, int64_t local_iterations,
                        double regularization, double stepsize, int64_t seed, int rng_mode, double *w_out) {
    if (ctx && !ctx->members.empty()) {
        if (iterations < 0) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_fit: iterations < 0");
        if (!w_out) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_fit: w_out is null");
        return group_svm_fit(ctx, X, x_dtype, csr_ptr, csr_idx, y, n, dim, block_id, blocks, iterations, local_iterations, regularization,
                             stepsize, seed, rng_mode, w_out);
    }
    CHECK_CTX(ctx);
    if (iterations < 0) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_fit: iterations < 0");
    if (!w_out) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_fit: w_out is null");
    FMS_TRY(fms_svm_load(ctx, X, x_dtype, 0, csr_ptr, csr_idx, y, n, dim, block_id, blocks, 0, 1, regularization, stepsize, seed, rng_mode));
    return svm_run_rounds(ctx, iterations, local_iterations, dim, w_out);
}

static int svm_run_rounds(fms_ctx *ctx, int iterations, int64_t local_iterations, int32_t dim, double *w_out) {
    NvtxRange nvtx_("fms.svm.fit rounds");
    for (int t = 0; t < iterations; ++t) {
        NvtxRange nvtx_r("fms.svm.round (sweep + reduce + apply)");
        FMS_TRY(svm_launch_sweep(ctx, local_iterations));
        FMS_TRY(svm_launch_apply(ctx));
    }
    return fms_svm_get_weights(ctx, w_out, dim);
}

// ====================================================================================================== chunked upload
FMS_API int fms_als_stream_begin(fms_ctx *ctx, int64_t nnz_total) {
    GROUP_FORWARD(ctx, fms_als_stream_begin(m0, nnz_total));            // a group stages the triples on its first device
    CHECK_CTX(ctx);
    if (nnz_total < 0 || nnz_total >= ((int64_t)1 << 31) * 8) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_stream_begin: nnz_total %lld", (long long)nnz_total);
    auto &T = ctx->staged;
    FMS_CUDA(ctx, T.u.alloc((size_t)nnz_total));
    FMS_CUDA(ctx, T.i.alloc((size_t)nnz_total));
    FMS_CUDA(ctx, T.r.alloc((size_t)nnz_total));
    T.nnz_cap = nnz_total;
    T.nnz_fill = 0;
    return FMS_OK;
}

FMS_API int fms_als_stream_append(fms_ctx *ctx, const int64_t *user, const int64_t *item, const double *rating, int64_t count) {
    GROUP_FORWARD(ctx, fms_als_stream_append(m0, user, item, rating, count));
    CHECK_CTX(ctx);
    auto &T = ctx->staged;
    if (T.nnz_cap < 0) return fail(ctx, FMS_ERR_STATE, "fms_als_stream_append: begin first");
    if (count < 0 || (count > 0 && (!user || !item || !rating))) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_stream_append: null input or count < 0");
    if (T.nnz_fill + count > T.nnz_cap) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_stream_append: %lld + %lld ratings exceed the announced %lld", (long long)T.nnz_fill, (long long)count, (long long)T.nnz_cap);
    if (count) {
        FMS_CUDA(ctx, cudaMemcpyAsync(T.u.p + T.nnz_fill, user, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
        FMS_CUDA(ctx, cudaMemcpyAsync(T.i.p + T.nnz_fill, item, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
        FMS_CUDA(ctx, cudaMemcpyAsync(T.r.p + T.nnz_fill, rating, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
        FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));       // the caller may reuse its buffer for the next piece
    }
    T.nnz_fill += count;
    return FMS_OK;
}

FMS_API int fms_als_stream_fit(fms_ctx *ctx, int rank, int iterations, double lambda, int64_t seed, int64_t *n_users, int64_t *n_items) {
    if (ctx && !ctx->members.empty()) {
        fms_ctx *m0 = ctx->members[0];
        auto &T0 = m0->staged;
        if (T0.nnz_cap < 0) return fail(ctx, FMS_ERR_STATE, "fms_als_stream_fit: begin first");
        if (T0.nnz_fill != T0.nnz_cap) return fail(ctx, FMS_ERR_STATE, "fms_als_stream_fit: %lld of the announced %lld ratings appended", (long long)T0.nnz_fill, (long long)T0.nnz_cap);
        if (iterations < 0 || rank < 1) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_stream_fit: rank %d iterations %d", rank, iterations);
        const int rc = group_als_fit(ctx, T0.u.p, T0.i.p, T0.r.p, T0.nnz_fill, rank, iterations, lambda, seed, n_users, n_items);   // device pointers
        cudaSetDevice(m0->device);
        T0.u.release(); T0.i.release(); T0.r.release();
        T0.nnz_cap = -1; T0.nnz_fill = 0;
        return rc;
    }
    CHECK_CTX(ctx);
    auto &T = ctx->staged;
    if (T.nnz_cap < 0) return fail(ctx, FMS_ERR_STATE, "fms_als_stream_fit: begin first");
    if (T.nnz_fill != T.nnz_cap) return fail(ctx, FMS_ERR_STATE, "fms_als_stream_fit: %lld of the announced %lld ratings appended", (long long)T.nnz_fill, (long long)T.nnz_cap);
    if (iterations < 0) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_stream_fit: iterations %d < 0", iterations);
    if (rank < 1) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_als_stream_fit: rank %d < 1", rank);
    const int64_t nnz = T.nnz_fill;
    int rc = fms_als_load(ctx, T.u.p, T.i.p, T.r.p, nnz, 0, 1, n_users, n_items);        // device pointers
    T.u.release(); T.i.release(); T.r.release();
    T.nnz_cap = -1; T.nnz_fill = 0;
    if (rc != FMS_OK) return rc;
    return als_fit_loaded(ctx, rank, iterations, lambda, seed, nnz);
}

FMS_API int fms_svm_stream_begin(fms_ctx *ctx, int64_t n_total, int32_t dim) {
    if (!ctx) return FMS_ERR_INVALID_ARG;
    if (n_total < 0 || dim < 1 || n_total >= ((int64_t)1 << 31)) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_stream_begin: n_total %lld dim %d", (long long)n_total, dim);
    if (!ctx->members.empty()) {
        // a group scatters the rows over its members by block at fit time: stage them in host memory inside the library
        auto &T = ctx->staged;
        try {
            T.hX.clear(); T.y.clear(); T.bid.clear();
            T.hX.reserve((size_t)n_total * (size_t)dim);          // reserved up front: the appends never allocate (no exception may cross the C-ABI)
            T.y.reserve((size_t)n_total); T.bid.reserve((size_t)n_total);
        } catch (const std::bad_alloc &) { return fail(ctx, FMS_ERR_OOM, "fms_svm_stream_begin: host staging of %lld x %d floats", (long long)n_total, dim); }
        T.n_cap = n_total; T.n_fill = 0; T.dim = dim;
        return FMS_OK;
    }
    CHECK_CTX(ctx);
    auto &T = ctx->staged;
    FMS_CUDA(ctx, T.X.alloc((size_t)n_total * (size_t)dim));
    try {
        T.y.clear(); T.bid.clear();
        T.y.reserve((size_t)n_total); T.bid.reserve((size_t)n_total);
    } catch (const std::bad_alloc &) { return fail(ctx, FMS_ERR_OOM, "fms_svm_stream_begin: host staging of %lld labels", (long long)n_total); }
    T.n_cap = n_total; T.n_fill = 0; T.dim = dim;
    return FMS_OK;
}

FMS_API int fms_svm_stream_append(fms_ctx *ctx, const void *X_rows, int x_dtype, const double *y, const int32_t *block_id, int64_t rows) {
    if (!ctx) return FMS_ERR_INVALID_ARG;
    auto &T = ctx->staged;
    if (T.n_cap < 0) return fail(ctx, FMS_ERR_STATE, "fms_svm_stream_append: begin first");
    if (rows < 0 || (rows > 0 && (!X_rows || !y || !block_id))) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_stream_append: null input or rows < 0");
    if (x_dtype != FMS_X_DENSE_F32 && x_dtype != FMS_X_DENSE_F64) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_stream_append: rows must be dense fp32 or fp64");
    if (T.n_fill + rows > T.n_cap) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_stream_append: %lld + %lld rows exceed the announced %lld", (long long)T.n_fill, (long long)rows, (long long)T.n_cap);
    if (!ctx->members.empty()) {
        const size_t cnt = (size_t)rows * (size_t)T.dim;
        if (x_dtype == FMS_X_DENSE_F32) T.hX.insert(T.hX.end(), static_cast<const float *>(X_rows), static_cast<const float *>(X_rows) + cnt);
        else { const double *xd = static_cast<const double *>(X_rows); for (size_t q = 0; q < cnt; ++q) T.hX.push_back((float)xd[q]); }
        T.y.insert(T.y.end(), y, y + rows);
        T.bid.insert(T.bid.end(), block_id, block_id + rows);
        T.n_fill += rows;
        return FMS_OK;
    }
    CHECK_CTX(ctx);
    if (rows) {
        float *dst = T.X.p + (size_t)T.n_fill * (size_t)T.dim;
        const int64_t cnt = rows * (int64_t)T.dim;
        if (x_dtype == FMS_X_DENSE_F32) {
            FMS_CUDA(ctx, cudaMemcpyAsync(dst, X_rows, (size_t)cnt * 4, cudaMemcpyHostToDevice, ctx->stream));
            FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        } else {
            FMS_TRY(upload_f64_as_f32(ctx, static_cast<const double *>(X_rows), cnt, dst));
        }
        T.y.insert(T.y.end(), y, y + rows);
        T.bid.insert(T.bid.end(), block_id, block_id + rows);
    }
    T.n_fill += rows;
    return FMS_OK;
}

FMS_API int fms_svm_stream_fit(fms_ctx *ctx, int blocks, int iterations, int64_t local_iterations, double regularization, double stepsize,
                               int64_t seed, int rng_mode, double *w_out) {
    if (!ctx) return FMS_ERR_INVALID_ARG;
    auto &T = ctx->staged;
    if (T.n_cap < 0) return fail(ctx, FMS_ERR_STATE, "fms_svm_stream_fit: begin first");
    if (T.n_fill != T.n_cap) return fail(ctx, FMS_ERR_STATE, "fms_svm_stream_fit: %lld of the announced %lld rows appended", (long long)T.n_fill, (long long)T.n_cap);
    if (iterations < 0) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_stream_fit: iterations < 0");
    if (!w_out) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_stream_fit: w_out is null");
    if (!ctx->members.empty()) {
        const int rc = group_svm_fit(ctx, T.hX.data(), FMS_X_DENSE_F32, nullptr, nullptr, T.y.data(), T.n_fill, T.dim, T.bid.data(), blocks, iterations,
                                     local_iterations, regularization, stepsize, seed, rng_mode, w_out);
        T.n_cap = -1;
        std::vector<float>().swap(T.hX); std::vector<double>().swap(T.y); std::vector<int32_t>().swap(T.bid);
        return rc;
    }
    CHECK_CTX(ctx);
    const int64_t n = T.n_fill;
    // the staged rows are used in place (they stay alive in the ctx until the next stream_begin)
    FMS_TRY(fms_svm_load(ctx, T.X.p, FMS_X_DENSE_F32, 1, nullptr, nullptr, T.y.data(), n, T.dim, T.bid.data(), blocks, 0, 1, regularization,
                         stepsize, seed, rng_mode));
    T.n_cap = -1;
    std::vector<double>().swap(T.y);
    std::vector<int32_t>().swap(T.bid);
    return svm_run_rounds(ctx, iterations, local_iterations, T.dim, w_out);
}

FMS_API int fms_svm_predict(fms_ctx *ctx, const void *X, int x_dtype, const int64_t *csr_ptr, const int32_t *csr_idx, int64_t n,
                            int32_t dim, const double *w, double threshold, int output_decision_function, double *out) {
    GROUP_FORWARD(ctx, fms_svm_predict(m0, X, x_dtype, csr_ptr, csr_idx, n, dim, w, threshold, output_decision_function, out));
    CHECK_CTX(ctx);
    SvmState &S = ctx->svm;
    if (n < 0 || dim < 1 || (n && (!X || !out))) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_predict: bad arguments");
    if (x_dtype < FMS_X_DENSE_F32 || x_dtype > FMS_X_CSR_F64) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_predict: x_dtype %d", x_dtype);
    if (x_dtype == FMS_X_CSR_F64 && n && (!csr_ptr || !csr_idx)) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_predict: CSR arrays missing");
    DevBuf<double> dw_in;
    const double *d_w = nullptr;
    if (w) {
        FMS_CUDA(ctx, dw_in.alloc((size_t)dim));
        FMS_CUDA(ctx, cudaMemcpyAsync(dw_in.p, w, (size_t)dim * 8, cudaMemcpyHostToDevice, ctx->stream));
        d_w = dw_in.p;
    } else {
        if (!S.w64_valid && !S.loaded) return fail(ctx, FMS_ERR_STATE, "The SVM model has not been trained. Call first fit before calling the predict operation.");
        FMS_TRY(svm_ensure_w64(ctx));
        if (dim != S.w64_dim) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_predict: dim %d != model dim %d", dim, S.w64_dim);
        d_w = S.w64.p;
    }
    if (!n) return FMS_OK;
    DevBuf<unsigned char> dX;
    DevBuf<int64_t> dptr;
    DevBuf<int32_t> didx;
    DevBuf<double> dout;
    FMS_CUDA(ctx, dout.alloc((size_t)n));
    size_t xbytes;
    if (x_dtype == FMS_X_CSR_F64) {
        const int64_t nnz = csr_ptr[n];
        xbytes = (size_t)nnz * 8;
        FMS_CUDA(ctx, dptr.alloc((size_t)n + 1));
        FMS_CUDA(ctx, didx.alloc((size_t)nnz));
        FMS_CUDA(ctx, cudaMemcpyAsync(dptr.p, csr_ptr, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (nnz) FMS_CUDA(ctx, cudaMemcpyAsync(didx.p, csr_idx, (size_t)nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        xbytes = (size_t)n * dim * (x_dtype == FMS_X_DENSE_F32 ? 4 : 8);
    }
    FMS_CUDA(ctx, dX.alloc(xbytes));
    if (xbytes) FMS_CUDA(ctx, cudaMemcpyAsync(dX.p, X, xbytes, cudaMemcpyHostToDevice, ctx->stream));
    FMS_TRY(svm_predict_device(ctx, dX.p, x_dtype, dptr.p, didx.p, n, dim, d_w, threshold, output_decision_function, dout.p));
    FMS_CUDA(ctx, cudaMemcpyAsync(out, dout.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMS_OK;
}

// ====================================================================================================== debug
__global__ void k_index_stream(long long seed, int block, int bound, long long count, int32_t *out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        JRand r;
        r.s = JRand::scramble(seed ^ (long long)block);
        for (long long i = 0; i < count; ++i) out[i] = r.next_int_bound(bound);
    }
}

FMS_API int fms_debug_index_stream(fms_ctx *ctx, int64_t seed, int block, int32_t bound, int64_t count, int32_t *out) {
    GROUP_FORWARD(ctx, fms_debug_index_stream(m0, seed, block, bound, count, out));
    CHECK_CTX(ctx);
    if (bound < 1 || count < 0 || (count && !out)) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_debug_index_stream: bad arguments");
    if (!count) return FMS_OK;
    DevBuf<int32_t> d;
    FMS_CUDA(ctx, d.alloc((size_t)count));
    k_index_stream<<<1, 32, 0, ctx->stream>>>((long long)seed, block, bound, (long long)count, d.p);
    FMS_LAUNCH_CHECK(ctx);
    FMS_CUDA(ctx, cudaMemcpyAsync(out, d.p, (size_t)count * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMS_OK;
}

FMS_API int fms_debug_als_gramian(fms_ctx *ctx, int which, double lambda, int64_t row_begin, int64_t row_end, float *out) {
    GROUP_UNSUPPORTED(ctx, "fms_debug_als_gramian");
    CHECK_CTX(ctx);
    AlsState &A = ctx->als;
    if (!A.loaded || !A.inited) return fail(ctx, FMS_ERR_STATE, "fms_debug_als_gramian: load + init first");
    if (which != 0 && which != 1) return fail(ctx, FMS_ERR_INVALID_ARG, "which = %d", which);
    const int64_t n = which == 0 ? A.n_users : A.n_items;
    if (row_begin < 0 || row_end > n || row_end < row_begin || !out) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_debug_als_gramian: bad row range");
    if (!als_tc_supported(ctx, A.rank)) return fail(ctx, FMS_ERR_UNSUPPORTED, "tensor-core ALS kernel does not cover rank %d", A.rank);
    const size_t per = (size_t)A.rank * A.rank + A.rank;
    DevBuf<float> d;
    FMS_CUDA(ctx, d.alloc((size_t)(row_end - row_begin) * per));
    FMS_CUDA(ctx, cudaMemsetAsync(A.info.p, 0, sizeof(int), ctx->stream));
    FMS_TRY(als_tc_dump(ctx, which, (float)lambda, row_begin, row_end, d.p));
    FMS_CUDA(ctx, cudaMemcpyAsync(out, d.p, (size_t)(row_end - row_begin) * per * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    A.fac64_valid = false;
    return FMS_OK;
}

}  // extern "C"
