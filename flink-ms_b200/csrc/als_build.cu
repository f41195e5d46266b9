// als_build.cu -- K8: ratings COO -> (sorted distinct ids, by-user CSR, by-item CSR) on the device.
//
// Takes over FlinkML's createBlockInformation (invoked from ALS.fit at
// flink-als/src/main/scala/de/tub/it4bi/ALSImpl.scala:52): per side the element ids are the sorted
// distinct ids, and the ratings are regrouped by target row with the partner index ascending inside a
// row (duplicate (u,i) pairs stay separate ratings in input order, as a DataSet would keep them).
// One-off integer sort/scan work (CUB radix sort) -- not part of the per-iteration hot loop.
#include <algorithm>

#include <cub/cub.cuh>

#include "common.cuh"

namespace fms {

__global__ void k_dense_index(const int64_t *__restrict__ raw, int64_t n, const int64_t *__restrict__ ids, int64_t n_ids,
                              int32_t *__restrict__ out) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    int64_t key = raw[k];
    int64_t lo = 0, hi = n_ids;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (ids[mid] < key) lo = mid + 1; else hi = mid;
    }
    out[k] = (int32_t)lo;
}

__global__ void k_make_keys(const int32_t *__restrict__ row, const int32_t *__restrict__ colv, int64_t n,
                            uint64_t *__restrict__ keys) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) keys[k] = ((uint64_t)(uint32_t)row[k] << 32) | (uint32_t)colv[k];
}
// sharded build: flag the ratings whose row this rank solves
__global__ void k_make_keys_flags(const int32_t *__restrict__ row, const int32_t *__restrict__ colv, int64_t n, int32_t rb, int32_t re,
                                  uint64_t *__restrict__ keys, uint8_t *__restrict__ flags) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) {
        const int32_t r = row[k];
        keys[k] = ((uint64_t)(uint32_t)r << 32) | (uint32_t)colv[k];
        flags[k] = (r >= rb && r < re) ? 1 : 0;
    }
}

__global__ void k_f64_to_f32(const double *__restrict__ in, int64_t n, float *__restrict__ out) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) out[k] = (float)in[k];
}

__global__ void k_split_keys(const uint64_t *__restrict__ keys, int64_t n, int32_t *__restrict__ col) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) col[k] = (int32_t)(uint32_t)(keys[k] & 0xffffffffu);
}

// ptr[r] = first position whose row >= r  (r in [0, n_rows])
__global__ void k_row_ptr(const uint64_t *__restrict__ keys, int64_t n, int64_t n_rows, int64_t *__restrict__ ptr) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    uint64_t key = (uint64_t)r << 32;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    ptr[r] = lo;
}

static inline int bits_for(int64_t n) {
    int b = 1;
    while (b < 32 && ((int64_t)1 << b) < n) ++b;
    return b;
}
static inline unsigned grid_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// sorted distinct ids of `raw` (device, length n) -> out (allocated here), count -> *n_out
static int unique_ids(fms_ctx *ctx, const int64_t *d_raw, int64_t n, DevBuf<int64_t> &out, int64_t *n_out,
                      DevBuf<int64_t> &scratch_sorted, DevBuf<unsigned char> &temp) {
    cudaStream_t st = ctx->stream;
    FMS_CUDA(ctx, scratch_sorted.alloc((size_t)n));
    FMS_CUDA(ctx, out.alloc((size_t)n));
    DevBuf<int64_t> d_count;
    FMS_CUDA(ctx, d_count.alloc(1));
    size_t tb = 0;
    FMS_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tb, d_raw, scratch_sorted.p, n, 0, 64, st));
    size_t tb2 = 0;
    FMS_CUDA(ctx, cub::DeviceSelect::Unique(nullptr, tb2, scratch_sorted.p, out.p, d_count.p, n, st));
    FMS_CUDA(ctx, temp.alloc(tb > tb2 ? tb : tb2));
    size_t t = temp.n;
    FMS_CUDA(ctx, cub::DeviceRadixSort::SortKeys(temp.p, t, d_raw, scratch_sorted.p, n, 0, 64, st));
    ctx->launches += 1;
    t = temp.n;
    FMS_CUDA(ctx, cub::DeviceSelect::Unique(temp.p, t, scratch_sorted.p, out.p, d_count.p, n, st));
    ctx->launches += 1;
    int64_t h = 0;
    FMS_CUDA(ctx, cudaMemcpyAsync(&h, d_count.p, sizeof h, cudaMemcpyDeviceToHost, st));
    FMS_CUDA(ctx, cudaStreamSynchronize(st));
    *n_out = h;
    return FMS_OK;
}

// rb / re: the rows this rank solves.  With more than one shard only THEIR ratings are sorted into the CSR (rows outside
// get empty ranges): SURVEY.md 8(e) "rank g holds the by-user CSR for its users and the by-item CSR for its items" -- the
// replicated build was what kept the end-to-end time from scaling in round 1.
static int build_side(fms_ctx *ctx, int side, const int32_t *d_row, const int32_t *d_col, const float *d_val, int64_t nnz,
                      int64_t n_rows, int64_t n_cols, DevBuf<uint64_t> &keys_in, DevBuf<uint64_t> &keys_out,
                      DevBuf<unsigned char> &temp, int64_t rb, int64_t re) {
    AlsState &A = ctx->als;
    cudaStream_t st = ctx->stream;
    const int T = 256;
    FMS_CUDA(ctx, A.ptr[side].alloc((size_t)n_rows + 1));
    DevBuf<float> val_sel;
    if (nnz > 0 && (rb > 0 || re < n_rows)) {
        DevBuf<uint8_t> flags;
        DevBuf<int64_t> d_cnt;
        FMS_CUDA(ctx, flags.alloc((size_t)nnz));
        FMS_CUDA(ctx, d_cnt.alloc(1));
        FMS_CUDA(ctx, val_sel.alloc((size_t)nnz));
        k_make_keys_flags<<<grid_for(nnz, T), T, 0, st>>>(d_row, d_col, nnz, (int32_t)rb, (int32_t)re, keys_out.p, flags.p);
        FMS_LAUNCH_CHECK(ctx);
        size_t tb = 0, tb2 = 0;
        FMS_CUDA(ctx, cub::DeviceSelect::Flagged(nullptr, tb, keys_out.p, flags.p, keys_in.p, d_cnt.p, nnz, st));
        FMS_CUDA(ctx, cub::DeviceSelect::Flagged(nullptr, tb2, d_val, flags.p, val_sel.p, d_cnt.p, nnz, st));
        FMS_CUDA(ctx, temp.alloc(tb > tb2 ? tb : tb2));
        size_t t = temp.n;
        FMS_CUDA(ctx, cub::DeviceSelect::Flagged(temp.p, t, keys_out.p, flags.p, keys_in.p, d_cnt.p, nnz, st));
        t = temp.n;
        FMS_CUDA(ctx, cub::DeviceSelect::Flagged(temp.p, t, d_val, flags.p, val_sel.p, d_cnt.p, nnz, st));
        ctx->launches += 2;
        int64_t h = 0;
        FMS_CUDA(ctx, cudaMemcpyAsync(&h, d_cnt.p, sizeof h, cudaMemcpyDeviceToHost, st));
        FMS_CUDA(ctx, cudaStreamSynchronize(st));
        nnz = h;
        d_val = val_sel.p;
        FMS_CUDA(ctx, A.col[side].alloc((size_t)nnz));
        FMS_CUDA(ctx, A.val[side].alloc((size_t)nnz));
    } else {
        FMS_CUDA(ctx, A.col[side].alloc((size_t)nnz));
        FMS_CUDA(ctx, A.val[side].alloc((size_t)nnz));
        if (nnz > 0) {
            k_make_keys<<<grid_for(nnz, T), T, 0, st>>>(d_row, d_col, nnz, keys_in.p);
            FMS_LAUNCH_CHECK(ctx);
        }
    }
    if (nnz > 0) {
        const int end_bit = 32 + bits_for(n_rows);
        size_t tb = 0;
        FMS_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tb, keys_in.p, keys_out.p, d_val, A.val[side].p, nnz, 0,
                                                      end_bit, st));
        FMS_CUDA(ctx, temp.alloc(tb));
        size_t t = temp.n;
        FMS_CUDA(ctx, cub::DeviceRadixSort::SortPairs(temp.p, t, keys_in.p, keys_out.p, d_val, A.val[side].p, nnz, 0,
                                                      end_bit, st));
        ctx->launches += 1;
        k_split_keys<<<grid_for(nnz, T), T, 0, st>>>(keys_out.p, nnz, A.col[side].p);
        FMS_LAUNCH_CHECK(ctx);
    }
    k_row_ptr<<<grid_for(n_rows + 1, T), T, 0, st>>>(keys_out.p, nnz, n_rows, A.ptr[side].p);
    FMS_LAUNCH_CHECK(ctx);
    (void)n_cols;
    return FMS_OK;
}

// h_user / h_item / h_rating: host OR device pointers (unified addressing decides the copy direction)
int als_build_csr(fms_ctx *ctx, const int64_t *h_user, const int64_t *h_item, const double *h_rating, int64_t nnz, int shard_index,
                  int shard_count) {
    AlsState &A = ctx->als;
    cudaStream_t st = ctx->stream;
    const int T = 256;
    DevBuf<int64_t> d_user, d_item, sorted;
    DevBuf<double> d_r64;
    DevBuf<float> d_r32;
    DevBuf<int32_t> du, di;
    DevBuf<uint64_t> keys_in, keys_out;
    DevBuf<unsigned char> temp;
    FMS_CUDA(ctx, d_user.alloc((size_t)nnz));
    FMS_CUDA(ctx, d_item.alloc((size_t)nnz));
    FMS_CUDA(ctx, d_r64.alloc((size_t)nnz));
    FMS_CUDA(ctx, d_r32.alloc((size_t)nnz));
    FMS_CUDA(ctx, cudaMemcpyAsync(d_user.p, h_user, (size_t)nnz * 8, cudaMemcpyDefault, st));
    FMS_CUDA(ctx, cudaMemcpyAsync(d_item.p, h_item, (size_t)nnz * 8, cudaMemcpyDefault, st));
    FMS_CUDA(ctx, cudaMemcpyAsync(d_r64.p, h_rating, (size_t)nnz * 8, cudaMemcpyDefault, st));
    if (nnz > 0) {
        k_f64_to_f32<<<grid_for(nnz, T), T, 0, st>>>(d_r64.p, nnz, d_r32.p);
        FMS_LAUNCH_CHECK(ctx);
    }
    int64_t nu = 0, ni = 0;
    if (nnz > 0) {
        FMS_TRY(unique_ids(ctx, d_user.p, nnz, A.ids[0], &nu, sorted, temp));
        FMS_TRY(unique_ids(ctx, d_item.p, nnz, A.ids[1], &ni, sorted, temp));
    } else {
        FMS_CUDA(ctx, A.ids[0].alloc(1));
        FMS_CUDA(ctx, A.ids[1].alloc(1));
    }
    sorted.release();
    d_r64.release();
    if (nu >= (int64_t)1 << 31 || ni >= (int64_t)1 << 31)
        return fail(ctx, FMS_ERR_UNSUPPORTED, "more than 2^31 distinct ids on one side");
    A.n_users = nu;
    A.n_items = ni;
    A.nnz = nnz;
    FMS_CUDA(ctx, du.alloc((size_t)nnz));
    FMS_CUDA(ctx, di.alloc((size_t)nnz));
    if (nnz > 0) {
        k_dense_index<<<grid_for(nnz, T), T, 0, st>>>(d_user.p, nnz, A.ids[0].p, nu, du.p);
        FMS_LAUNCH_CHECK(ctx);
        k_dense_index<<<grid_for(nnz, T), T, 0, st>>>(d_item.p, nnz, A.ids[1].p, ni, di.p);
        FMS_LAUNCH_CHECK(ctx);
    }
    FMS_CUDA(ctx, cudaStreamSynchronize(st));
    d_user.release();
    d_item.release();
    FMS_CUDA(ctx, keys_in.alloc((size_t)nnz));
    FMS_CUDA(ctx, keys_out.alloc((size_t)nnz));
    const int64_t su = (nu + shard_count - 1) / shard_count, si = (ni + shard_count - 1) / shard_count;
    const int64_t ub = std::min<int64_t>(nu, su * shard_index), ue = std::min<int64_t>(nu, su * (shard_index + 1));
    const int64_t ib = std::min<int64_t>(ni, si * shard_index), ie = std::min<int64_t>(ni, si * (shard_index + 1));
    FMS_TRY(build_side(ctx, 0, du.p, di.p, d_r32.p, nnz, nu, ni, keys_in, keys_out, temp, ub, ue));
    FMS_TRY(build_side(ctx, 1, di.p, du.p, d_r32.p, nnz, ni, nu, keys_in, keys_out, temp, ib, ie));
    FMS_CUDA(ctx, cudaStreamSynchronize(st));
    return FMS_OK;
}

}  // namespace fms
