// common.cuh -- context, error plumbing and small device helpers shared by every translation unit of
// libflinkms_b200.so.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>            // header-only: ranges are no-ops unless a profiler injects the NVTX library

#include "../../include/flinkms.h"

namespace fms {

// --------------------------------------------------------------------------------------------- buffers
// Device allocation owned by the ctx; freed in the destructor (stream-ordered frees are not needed:
// every (re)allocation happens between synchronised phases of fit/load).
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    size_t bytes() const { return n * sizeof(T); }
};

// --------------------------------------------------------------------------------------------- ALS state
struct AlsState {
    bool loaded = false, inited = false;
    int64_t nnz = 0, n_users = 0, n_items = 0;
    int shard_index = 0, shard_count = 1;
    int64_t shard_rows[2] = {0, 0};          // S = ceil(N / shard_count) per side
    int rank = 0;
    int kernel = FMS_ALS_KERNEL_AUTO;
    DevBuf<int64_t> ids[2];                  // sorted distinct ids (createBlockInformation)
    // serving scratch (grow-only): query ids, results, found flags -- a lookup batch must not pay cudaMalloc / cudaFree
    DevBuf<int64_t> q_u, q_i;
    DevBuf<double> q_out, q_rating;
    DevBuf<uint8_t> q_found;
    DevBuf<int64_t> ptr[2];                  // CSR row pointers  (side 0: by user, side 1: by item)
    DevBuf<int32_t> col[2];                  // dense partner index, ascending inside a row
    DevBuf<float> val[2];                    // rating
    DevBuf<float> fac[2];                    // factor matrices, fp32 row-major, rows_padded x rank
    DevBuf<double> fac64[2];                 // fp64 copies for the predict path (lazy / external model)
    bool fac64_valid = false;
    bool model_external = false;             // installed via fms_als_set_model (no CSR)
    DevBuf<int> info;                        // [0] = rows with non-positive pivot in the last half-step
    int64_t last_bad = 0;
    // tensor-core path: split-precision gather planes (hi | lo fp16) per side, refreshed after each update
    DevBuf<uint16_t> plane[2];
    bool plane_valid[2] = {false, false};
    int64_t plane_rows[2] = {0, 0};
    // peer-to-peer exchange: device table of every rank's plane base for a side (own pointer at own index), filled from
    // CUDA IPC handles; the solver warps then store each new plane row into all replicas over NVLink
    DevBuf<unsigned long long> peer_tab[2];
    void *peer_open[2][64] = {};
    int peer_n[2] = {0, 0};
    // in-kernel cross-rank ordering of the half-steps (als_tc.cu FlagSync): epochs per side, last-CTA counter
    int plane_epoch[2] = {0, 0};
    bool flag_sync = false;                  // host opted in (fms_als_planes_flag_sync): it then needs no barrier between half-steps
    DevBuf<int> sync_ctr;
    bool sync_ctr_zeroed = false;
    // tensor-core path: ratings re-laid in 16-rating stages per row (padded), see als_tc.cu
    DevBuf<int32_t> stage_ptr[2];            // per row: first stage (exclusive scan of ceil(n_row / 16)), N + 1 entries
    DevBuf<int32_t> colp[2];                 // 16 * total_stages partner indices (padding -> zero row of the planes)
    DevBuf<float> valp[2];                   // 16 * total_stages ratings (padding -> 0)
    DevBuf<int32_t> cta_rows[2];             // row range boundaries per CTA (balanced by stage count)
    int64_t total_stages[2] = {0, 0};
    int stage_grid[2] = {0, 0};
    int64_t stage_rb[2] = {-1, -1}, stage_re[2] = {-1, -1};
    int stage_zero_row[2] = {-1, -1};
    int stage_krows[2] = {0, 0};             // ratings per ring slot the stage layout was padded for (16 x KS)
};

// --------------------------------------------------------------------------------------------- SVM state
struct SvmState {
    bool loaded = false;
    int64_t n = 0;
    int64_t n_total = 0;                     // size of the WHOLE training set (lambda * n of `maximize`); == n unless a group ctx sharded the rows
    int32_t dim = 0;
    int blocks = 0;
    int shard_index = 0, shard_count = 1;
    int x_kind = FMS_X_DENSE_F32;            // device layout: dense fp32 or CSR (fp32 values)
    double lambda = 1.0, stepsize = 1.0;
    int64_t seed = 0;
    int rng_mode = FMS_RNG_PER_BLOCK;
    int kernel = FMS_SVM_KERNEL_AUTO;
    const float *X = nullptr;                // dense rows (owned buffer or caller's device pointer)
    DevBuf<float> X_own;
    DevBuf<int64_t> csr_ptr;
    DevBuf<int32_t> csr_idx;
    DevBuf<float> csr_val;
    DevBuf<float> y;
    DevBuf<float> alpha, alpha_tmp;          // persistent alpha, working copy (stepsize != 1)
    DevBuf<int64_t> rows;                    // row ids grouped by block, arrival order inside a block
    DevBuf<int32_t> rows32;                  // same, 32-bit (n < 2^31): what the warp sweep reads
    DevBuf<int64_t> blk_ptr;                 // blocks + 1 offsets into rows
    std::vector<int64_t> h_blk_ptr;
    DevBuf<uint64_t> rng;                    // per-block java.util.Random state (48-bit), persists across rounds
    DevBuf<float> w;                         // dim
    DevBuf<float> dw_blocks;                 // blocks_local x dim partial deltas
    DevBuf<float> delta;                     // dim: sum over this rank's blocks
    DevBuf<float> red_partial;               // two-level reduction scratch (chunks x dim)
    DevBuf<double> w64;                      // weights for the predict path
    bool w64_valid = false;
    int32_t w64_dim = 0;
    std::vector<int> local_blocks;           // block ids swept by this rank
    DevBuf<int32_t> d_local_blocks;
    DevBuf<unsigned long long> peer_delta;   // group ctx: device table of every member's delta buffer (fused all-reduce + apply)
    // one process per GPU: exchange buffer [2 epoch parities][n ranks][dim] floats + 64 flag words, mapped into every rank (CUDA IPC)
    DevBuf<float> xchg;
    DevBuf<unsigned long long> xchg_tab;     // device table: every rank's exchange buffer
    DevBuf<int> xchg_err;                    // set by a kernel whose flag wait timed out (a peer died)
    void *xchg_open[64] = {};
    int xchg_n = 0, xchg_rank = 0, xchg_epoch = 0, xchg_dim = 0;
};

}  // namespace fms

struct fms_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    fms::AlsState als;
    fms::SvmState svm;
    // chunked upload (fms_*_stream_*): inputs staged in HBM piece by piece before a whole fit
    struct {
        fms::DevBuf<int64_t> u, i;
        fms::DevBuf<double> r;
        int64_t nnz_cap = -1, nnz_fill = 0;
        fms::DevBuf<float> X;
        std::vector<double> y;
        std::vector<int32_t> bid;
        int64_t n_cap = -1, n_fill = 0;
        int32_t dim = 0;
        std::vector<float> hX;               // group ctx: rows staged on the host (they are scattered over the members at fit time)
    } staged;
    // fms_create(ids, n > 1): this ctx is a GROUP -- one single-GPU member ctx per device (csrc/group.cu); compute state lives
    // in the members, the parent only sequences them
    std::vector<fms_ctx *> members;
    std::vector<cudaEvent_t> member_ev;
};

namespace fms {

inline int fail(fms_ctx *ctx, int code, const char *fmt, ...) __attribute__((format(printf, 3, 4)));
inline int fail(fms_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

// NVTX range for the phases of a fit (the reference has no tracing beyond wall-clock prints)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

#define FMS_CUDA(ctx, expr)                                                                              \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return ::fms::fail((ctx), e__ == cudaErrorMemoryAllocation ? FMS_ERR_OOM : FMS_ERR_CUDA,      \
                               "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));   \
    } while (0)

#define FMS_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != FMS_OK) return rc__; \
    } while (0)

#define FMS_LAUNCH_CHECK(ctx)                                   \
    do {                                                        \
        ++(ctx)->launches;                                      \
        FMS_CUDA((ctx), cudaGetLastError());                    \
    } while (0)

// --------------------------------------------------------------------------------------------- java.util.Random
// 48-bit LCG of the JDK specification; bit-exact on the device.
struct JRand {
    unsigned long long s;
    __host__ __device__ static unsigned long long scramble(long long seed) {
        return ((unsigned long long)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1);
    }
    __host__ __device__ int next(int bits) {
        s = (s * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
        return (int)(long long)(s >> (48 - bits));
    }
    __host__ __device__ int next_int_bound(int bound) {
        int r = next(31);
        int m = bound - 1;
        if ((bound & m) == 0) return (int)(((long long)bound * (long long)r) >> 31);
        for (int u = r;; u = next(31)) {
            r = u % bound;
            if ((int)((unsigned)u - (unsigned)r + (unsigned)m) >= 0) return r;
        }
    }
    // Same stream as next_int_bound, with the `u % bound` done by Lemire's exact fast-modulo
    // (magic = 2^64 / bound + 1, exact for 32-bit operands): bit-identical results, no integer division.
    __host__ __device__ static unsigned long long magic_for(int bound) { return 0xFFFFFFFFFFFFFFFFULL / (unsigned long long)bound + 1ULL; }
    __device__ int next_int_bound_fast(int bound, unsigned long long magic) {
        int r = next(31);
        const int m = bound - 1;
        if ((bound & m) == 0) return (int)(((long long)bound * (long long)r) >> 31);
        for (int u = r;; u = next(31)) {
            r = (int)__umul64hi(magic * (unsigned long long)(unsigned)u, (unsigned long long)bound);
            if ((int)((unsigned)u - (unsigned)r + (unsigned)m) >= 0) return r;
        }
    }
    __host__ __device__ double next_double() {
        long long hi = (long long)next(26);
        long long lo = (long long)next(27);
        return (double)((hi << 27) + lo) * 0x1.0p-53;
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- internal entry points implemented across translation units ----
int als_build_csr(fms_ctx *ctx, const int64_t *h_user, const int64_t *h_item, const double *h_rating, int64_t nnz, int shard_index,
                  int shard_count);
int als_launch_init(fms_ctx *ctx, int64_t seed);
int als_launch_half_step_generic(fms_ctx *ctx, int which, float lambda, int64_t row_begin, int64_t row_end);
int als_tc_supported(const fms_ctx *ctx, int rank);
int als_wide_supported(const fms_ctx *ctx, int rank);                  // 64 < rank <= 128: FFMA Gramians + HMMA solver (als_wide.cu)
int als_launch_half_step_wide(fms_ctx *ctx, int which, float lambda, int64_t row_begin, int64_t row_end);
int als_launch_half_step_tc(fms_ctx *ctx, int which, float lambda, int64_t row_begin, int64_t row_end);
int als_sgd_step_device(fms_ctx *ctx, const int64_t *d_u, const int64_t *d_i, const double *d_r, int64_t n, double lr, double ureg,
                        double ireg, const double *d_umean, const double *d_imean, int apply, double *d_out_u, double *d_out_i);
int ingest_ratings_csv(fms_ctx *ctx, const char *text, int64_t nbytes, char delimiter, int ignore_first_line, int64_t *user, int64_t *item,
                       double *rating, int64_t capacity, int64_t *n_rows);
int ingest_libsvm(fms_ctx *ctx, const char *text, int64_t nbytes, double *label, int64_t row_capacity, int64_t *csr_ptr, int32_t *csr_idx,
                  double *csr_val, int64_t nnz_capacity, int64_t *n_rows, int64_t *nnz, int32_t *dim);
int als_tc_refresh_planes(fms_ctx *ctx, int which);
int als_tc_dump(fms_ctx *ctx, int which, float lambda, int64_t rb, int64_t re, float *d_dump);
int als_tc_planes(fms_ctx *ctx, int s, void **hi, void **lo, int64_t *rows_padded);
int als_tc_ipc_handle(fms_ctx *ctx, int s, void *handle64);
int als_tc_open_peers(fms_ctx *ctx, int s, const void *handles, int n);
void als_tc_close_peers(fms_ctx *ctx, int s);
int als_predict_device(fms_ctx *ctx, const int64_t *d_u, const int64_t *d_i, int64_t n, double *d_out, uint8_t *d_found);
int als_ensure_fac64(fms_ctx *ctx);
int group_create(fms_ctx *P, const int *device_ids, int n);
void group_destroy(fms_ctx *P);
int group_als_fit(fms_ctx *P, const int64_t *user, const int64_t *item, const double *rating, int64_t nnz, int rank, int iterations,
                  double lambda, int64_t seed, int64_t *n_users, int64_t *n_items);
int group_svm_fit(fms_ctx *P, const void *X, int x_dtype, const int64_t *csr_ptr, const int32_t *csr_idx, const double *y, int64_t n,
                  int32_t dim, const int32_t *block_id, int blocks, int iterations, int64_t local_iterations, double regularization,
                  double stepsize, int64_t seed, int rng_mode, double *w_out);
int svm_launch_sweep(fms_ctx *ctx, int64_t local_iterations);
int svm_launch_reduce_delta(fms_ctx *ctx);
int svm_launch_apply(fms_ctx *ctx);
int svm_exchange_handle(fms_ctx *ctx, int n_ranks, void *handle64);
int svm_exchange_open(fms_ctx *ctx, const void *handles, int n_ranks, int my_rank);
void svm_exchange_close(fms_ctx *ctx);
int svm_launch_exchange_apply(fms_ctx *ctx);
int svm_predict_device(fms_ctx *ctx, const void *d_X, int x_kind, const int64_t *d_ptr, const int32_t *d_idx, int64_t n,
                       int32_t dim, const double *d_w, double threshold, int decision, double *d_out);

}  // namespace fms
