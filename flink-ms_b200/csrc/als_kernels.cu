// als_kernels.cu -- K2 (factor init), K1-generic (gather + Gramian + Cholesky, any rank <= 128, CUDA cores),
// K3 (predict gather-dot).  The tensor-core K1 for rank 64/128 lives in als_tc.cu.
//
// K1 takes over the FlinkML `updateFactors` coGroup body (entered
// from flink-als/src/main/scala/de/tub/it4bi/ALSImpl.scala:52): per target row u
//     A_u = sum_k y_k y_k^T + lambda * n_u * I      (ALS-WR weighting)
//     b_u = sum_k r_k y_k
//     x_u = A_u^{-1} b_u                             (Cholesky, LAPACK dposv in the reference)
// K2 takes over `randomFactors` (items <- Random(id ^ seed).nextDouble() x rank).
// K3 takes over ALS.predict / the serve-side dot (flink-queryable-client/.../ALSPredictRandom.java:79-92).
#include "common.cuh"

namespace fms {

// ------------------------------------------------------------------------------------------------ K2
__global__ void k_als_init(const int64_t *__restrict__ ids, int64_t n, int rank, long long seed, float *__restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    JRand r;
    r.s = JRand::scramble((long long)ids[i] ^ seed);
    float *o = out + i * rank;
    for (int k = 0; k < rank; ++k) o[k] = (float)r.next_double();
}

int als_launch_init(fms_ctx *ctx, int64_t seed) {
    AlsState &A = ctx->als;
    const int64_t rows_u = A.shard_rows[0] * A.shard_count, rows_i = A.shard_rows[1] * A.shard_count;
    FMS_CUDA(ctx, A.fac[0].alloc((size_t)rows_u * A.rank));
    FMS_CUDA(ctx, A.fac[1].alloc((size_t)rows_i * A.rank));
    FMS_CUDA(ctx, cudaMemsetAsync(A.fac[0].p, 0, (size_t)rows_u * A.rank * sizeof(float), ctx->stream));
    FMS_CUDA(ctx, cudaMemsetAsync(A.fac[1].p, 0, (size_t)rows_i * A.rank * sizeof(float), ctx->stream));
    FMS_CUDA(ctx, A.info.alloc(1));
    if (A.n_items > 0) {
        k_als_init<<<(unsigned)((A.n_items + 127) / 128), 128, 0, ctx->stream>>>(A.ids[1].p, A.n_items, A.rank,
                                                                                 (long long)seed, A.fac[1].p);
        FMS_LAUNCH_CHECK(ctx);
    }
    A.fac64_valid = false;
    A.plane_valid[0] = A.plane_valid[1] = false;
    return FMS_OK;
}

// ------------------------------------------------------------------------------------------------ K1 generic
// One CTA (256 threads as a 16 x 16 grid) per target row; each thread owns a TM x TM register tile of the
// DP x DP Gramian (DP = 16 * TM >= rank).  Gathered partner rows are staged CH at a time in shared memory.
// The Cholesky runs CTA-wide in shared memory.  Correct for any rank <= 128 and any row length; it is the
// device-side reference for the tensor-core kernel and the path for ranks that kernel does not cover.
template <int TM>
__global__ void __launch_bounds__(256) k_als_rows_generic(const int64_t *__restrict__ ptr, const int32_t *__restrict__ col,
                                                          const float *__restrict__ val, const float *__restrict__ P,
                                                          float *__restrict__ T, int d, float lambda, int64_t row_begin,
                                                          int64_t row_end, int *__restrict__ info) {
    constexpr int DP = 16 * TM;
    constexpr int CH = 32;
    constexpr int LD = DP + 1;
    extern __shared__ float smem[];
    float *tile = smem;                 // CH x DP
    float *rt = tile + CH * DP;         // CH
    float *A = rt + CH;                 // DP x LD (lower triangle used)
    float *b = A + DP * LD;             // DP
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

    for (int64_t row = row_begin + blockIdx.x; row < row_end; row += gridDim.x) {
        const int64_t s = ptr[row], e = ptr[row + 1];
        float acc[TM][TM];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TM; ++j) acc[i][j] = 0.f;
        float bacc = 0.f;
        for (int64_t k0 = s; k0 < e; k0 += CH) {
            const int cn = (int)((e - k0) < CH ? (e - k0) : CH);
            for (int idx = tid; idx < cn * DP; idx += 256) {
                const int r = idx / DP, j = idx - r * DP;
                const int32_t c = col[k0 + r];
                tile[idx] = j < d ? P[(int64_t)c * d + j] : 0.f;
            }
            if (tid < cn) rt[tid] = val[k0 + tid];
            __syncthreads();
            for (int r = 0; r < cn; ++r) {
                float a[TM], bb[TM];
#pragma unroll
                for (int i = 0; i < TM; ++i) a[i] = tile[r * DP + ty * TM + i];
#pragma unroll
                for (int j = 0; j < TM; ++j) bb[j] = tile[r * DP + tx * TM + j];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TM; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
                if (tid < DP) bacc = fmaf(rt[r], tile[r * DP + tid], bacc);
            }
            __syncthreads();
        }
        const float reg = lambda * (float)(e - s);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TM; ++j) {
                const int gi = ty * TM + i, gj = tx * TM + j;
                float v = acc[i][j];
                if (gi == gj) v += (gi < d) ? reg : 1.f;   // padded diagonal = 1 keeps the factorisation defined
                A[gi * LD + gj] = v;
            }
        if (tid < DP) b[tid] = bacc;
        __syncthreads();
        // --- Cholesky A = L L^T on the leading d x d block (lower), CTA-wide
        bool bad = false;
        for (int j = 0; j < d; ++j) {
            const float ajj = A[j * LD + j];
            if (!(ajj > 0.f)) bad = true;
            const float ljj = sqrtf(ajj);
            __syncthreads();
            for (int i = j + tid; i < d; i += 256) A[i * LD + j] = (i == j) ? ljj : A[i * LD + j] / ljj;
            __syncthreads();
            const int m = d - j - 1;
            for (int idx = tid; idx < m * m; idx += 256) {
                const int i = j + 1 + idx / m, k = j + 1 + idx % m;
                if (k <= i) A[i * LD + k] = fmaf(-A[i * LD + j], A[k * LD + j], A[i * LD + k]);
            }
            __syncthreads();
        }
        // --- L z = b ; L^T x = z  (warp 0; lanes split each dot product)
        if (tid < 32) {
            for (int i = 0; i < d; ++i) {
                float p = 0.f;
                for (int k = tid; k < i; k += 32) p = fmaf(A[i * LD + k], b[k], p);
                p = warp_sum(p);
                if (tid == 0) b[i] = (b[i] - p) / A[i * LD + i];
                __syncwarp();
            }
            for (int i = d - 1; i >= 0; --i) {
                float p = 0.f;
                for (int k = i + 1 + tid; k < d; k += 32) p = fmaf(A[k * LD + i], b[k], p);
                p = warp_sum(p);
                if (tid == 0) b[i] = (b[i] - p) / A[i * LD + i];
                __syncwarp();
            }
            if (tid == 0 && bad) atomicAdd(info, 1);
        }
        __syncthreads();
        if (tid < d) T[row * d + tid] = b[tid];
        __syncthreads();
    }
}

template <int TM>
static int launch_generic(fms_ctx *ctx, int which, float lambda, int64_t rb, int64_t re) {
    AlsState &A = ctx->als;
    constexpr int DP = 16 * TM;
    const size_t smem = (size_t)(32 * DP + 32 + DP * (DP + 1) + DP) * sizeof(float);
    FMS_CUDA(ctx, cudaFuncSetAttribute(k_als_rows_generic<TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t rows = re - rb;
    if (rows <= 0) return FMS_OK;
    const int64_t max_grid = (int64_t)ctx->sm_count * 16;
    const unsigned grid = (unsigned)(rows < max_grid ? rows : max_grid);
    k_als_rows_generic<TM><<<grid, 256, smem, ctx->stream>>>(A.ptr[which].p, A.col[which].p, A.val[which].p,
                                                             A.fac[1 - which].p, A.fac[which].p, A.rank, lambda, rb, re,
                                                             A.info.p);
    FMS_LAUNCH_CHECK(ctx);
    return FMS_OK;
}

int als_launch_half_step_generic(fms_ctx *ctx, int which, float lambda, int64_t rb, int64_t re) {
    const int d = ctx->als.rank;
    if (als_wide_supported(ctx, d)) return als_launch_half_step_wide(ctx, which, lambda, rb, re);
    if (d <= 16) return launch_generic<1>(ctx, which, lambda, rb, re);
    if (d <= 32) return launch_generic<2>(ctx, which, lambda, rb, re);
    if (d <= 64) return launch_generic<4>(ctx, which, lambda, rb, re);
    if (d <= 128) return launch_generic<8>(ctx, which, lambda, rb, re);
    return fail(ctx, FMS_ERR_UNSUPPORTED, "ALS rank %d > 128 is not supported", d);
}

// ------------------------------------------------------------------------------------------------ K3 predict
__global__ void k_f32_to_f64(const float *__restrict__ in, int64_t n, double *__restrict__ out) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) out[k] = (double)in[k];
}

int als_ensure_fac64(fms_ctx *ctx) {
    AlsState &A = ctx->als;
    if (A.fac64_valid) return FMS_OK;
    const int64_t n[2] = {A.n_users, A.n_items};
    for (int s = 0; s < 2; ++s) {
        const int64_t cnt = n[s] * A.rank;
        FMS_CUDA(ctx, A.fac64[s].alloc((size_t)cnt));
        if (cnt > 0) {
            k_f32_to_f64<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(A.fac[s].p, cnt, A.fac64[s].p);
            FMS_LAUNCH_CHECK(ctx);
        }
    }
    A.fac64_valid = true;
    return FMS_OK;
}

// 8 lanes per (u,i) pair: binary-search both ids (lane 0 of the group), then a strided fp64 dot.
__global__ void k_als_predict(const int64_t *__restrict__ qu, const int64_t *__restrict__ qi, int64_t n,
                              const int64_t *__restrict__ uids, int64_t nu, const int64_t *__restrict__ iids, int64_t ni,
                              const double *__restrict__ U, const double *__restrict__ I, int d, double *__restrict__ out,
                              uint8_t *__restrict__ found) {
    const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~7);
    if (g >= n) return;
    int64_t a = -1, b = -1;
    if (sub == 0) {
        int64_t key = qu[g], lo = 0, hi = nu;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (uids[mid] < key) lo = mid + 1; else hi = mid; }
        a = (lo < nu && uids[lo] == key) ? lo : -1;
        key = qi[g]; lo = 0; hi = ni;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (iids[mid] < key) lo = mid + 1; else hi = mid; }
        b = (lo < ni && iids[lo] == key) ? lo : -1;
    }
    const int src = (threadIdx.x & 31) & ~7;
    a = __shfl_sync(gmask, a, src);
    b = __shfl_sync(gmask, b, src);
    const bool ok = a >= 0 && b >= 0;
    double s = 0.0;
    if (ok) {
        const double *x = U + a * d, *y = I + b * d;
        for (int j = sub; j < d; j += 8) s = fma(x[j], y[j], s);
    }
    s += __shfl_xor_sync(gmask, s, 4);
    s += __shfl_xor_sync(gmask, s, 2);
    s += __shfl_xor_sync(gmask, s, 1);
    if (sub == 0) {
        out[g] = ok ? s : 0.0;
        if (found) found[g] = ok ? 1 : 0;
    }
}

int als_predict_device(fms_ctx *ctx, const int64_t *d_u, const int64_t *d_i, int64_t n, double *d_out, uint8_t *d_found) {
    AlsState &A = ctx->als;
    if (n <= 0) return FMS_OK;
    const int64_t threads = n * 8;
    k_als_predict<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(d_u, d_i, n, A.ids[0].p, A.n_users, A.ids[1].p,
                                                                              A.n_items, A.fac64[0].p, A.fac64[1].p, A.rank,
                                                                              d_out, d_found);
    FMS_LAUNCH_CHECK(ctx);
    return FMS_OK;
}

// ------------------------------------------------------------------------------------------------ N2 online SGD step
// als-ms SGD.java:160-217.  One warp per rating; every record reads the model as it was before the batch.  All
// arithmetic is fp64 with the reference's operation order and NO fused multiply-add (Java rounds every operation), so the
// emitted rows are bit-identical to the Java code's for the same fp64 model.
__device__ __forceinline__ int64_t find_id(const int64_t *__restrict__ ids, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (ids[mid] < key) lo = mid + 1; else hi = mid; }
    return (lo < n && ids[lo] == key) ? lo : -1;
}

__global__ void k_als_sgd(const int64_t *__restrict__ qu, const int64_t *__restrict__ qi, const double *__restrict__ qr, int64_t n,
                          const int64_t *__restrict__ uids, int64_t nu, const int64_t *__restrict__ iids, int64_t ni,
                          const double *__restrict__ U, const double *__restrict__ I, const double *__restrict__ umean,
                          const double *__restrict__ imean, int d, double lr, double ureg, double ireg, double *__restrict__ out_u,
                          double *__restrict__ out_i, int64_t *__restrict__ row_u, int64_t *__restrict__ row_i,
                          int *__restrict__ win_u, int *__restrict__ win_i) {
    const int64_t k = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= n) return;
    int64_t a = -1, b = -1;
    if (lane == 0) { a = find_id(uids, nu, qu[k]); b = find_id(iids, ni, qi[k]); }
    a = __shfl_sync(0xffffffffu, a, 0);
    b = __shfl_sync(0xffffffffu, b, 0);
    const double *x = a >= 0 ? U + a * d : umean, *y = b >= 0 ? I + b * d : imean;   // cold start: the MEAN rows (SGD.java:171-172)
    double xs[4], ys[4];                                               // rank <= 128: four components per lane
    double pred = 0.0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int j = lane + 32 * t;
        xs[t] = j < d ? x[j] : 0.0;
        ys[t] = j < d ? y[j] : 0.0;
    }
    for (int j = 0; j < d; ++j) {                                      // prediction += u[j] * v[j], in index order
        const double p = __dmul_rn(xs[j >> 5], ys[j >> 5]);
        pred = __dadd_rn(pred, __shfl_sync(0xffffffffu, p, j & 31));
    }
    const double err = __dsub_rn(qr[k], pred);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int j = lane + 32 * t;
        if (j < d) {
            out_u[k * d + j] = __dadd_rn(xs[t], __dmul_rn(lr, __dsub_rn(__dmul_rn(err, ys[t]), __dmul_rn(ureg, xs[t]))));
            out_i[k * d + j] = __dadd_rn(ys[t], __dmul_rn(lr, __dsub_rn(__dmul_rn(err, xs[t]), __dmul_rn(ireg, ys[t]))));
        }
    }
    if (lane == 0) {
        row_u[k] = a; row_i[k] = b;
        if (win_u && a >= 0) atomicMax(win_u + a, (int)k);            // the last record of an id wins, as in the Kafka consumer
        if (win_i && b >= 0) atomicMax(win_i + b, (int)k);
    }
}

__global__ void k_als_sgd_apply(int64_t n, int d, const double *__restrict__ out, const int64_t *__restrict__ row, const int *__restrict__ win,
                                double *__restrict__ F64, float *__restrict__ F32) {
    const int64_t k = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= n) return;
    const int64_t r = row[k];
    if (r < 0 || win[r] != (int)k) return;
    for (int j = lane; j < d; j += 32) {
        const double v = out[k * d + j];
        F64[r * d + j] = v;
        if (F32) F32[r * d + j] = (float)v;
    }
}

int als_sgd_step_device(fms_ctx *ctx, const int64_t *d_u, const int64_t *d_i, const double *d_r, int64_t n, double lr, double ureg,
                        double ireg, const double *d_umean, const double *d_imean, int apply, double *d_out_u, double *d_out_i) {
    AlsState &A = ctx->als;
    if (n <= 0) return FMS_OK;
    if (n >= ((int64_t)1 << 31)) return fail(ctx, FMS_ERR_UNSUPPORTED, "sgd batch too large");
    // k_als_sgd keeps four components per lane (rank <= 128); fms_als_set_model accepts any rank, so refuse here
    if (A.rank > 128) return fail(ctx, FMS_ERR_UNSUPPORTED, "fms_als_sgd_step: rank %d > 128 is not supported", A.rank);
    DevBuf<int64_t> row_u, row_i;
    DevBuf<int> win_u, win_i;
    FMS_CUDA(ctx, row_u.alloc((size_t)n));
    FMS_CUDA(ctx, row_i.alloc((size_t)n));
    if (apply) {
        FMS_CUDA(ctx, win_u.alloc((size_t)A.n_users + 1));
        FMS_CUDA(ctx, win_i.alloc((size_t)A.n_items + 1));
        FMS_CUDA(ctx, cudaMemsetAsync(win_u.p, 0xff, win_u.bytes(), ctx->stream));
        FMS_CUDA(ctx, cudaMemsetAsync(win_i.p, 0xff, win_i.bytes(), ctx->stream));
    }
    const unsigned grid = (unsigned)((n * 32 + 255) / 256);
    k_als_sgd<<<grid, 256, 0, ctx->stream>>>(d_u, d_i, d_r, n, A.ids[0].p, A.n_users, A.ids[1].p, A.n_items, A.fac64[0].p, A.fac64[1].p,
                                             d_umean, d_imean, A.rank, lr, ureg, ireg, d_out_u, d_out_i, row_u.p, row_i.p,
                                             apply ? win_u.p : nullptr, apply ? win_i.p : nullptr);
    FMS_LAUNCH_CHECK(ctx);
    if (apply) {
        const bool f32 = A.inited;                                     // a fitted (fp32) model is kept in step with its fp64 copy
        k_als_sgd_apply<<<grid, 256, 0, ctx->stream>>>(n, A.rank, d_out_u, row_u.p, win_u.p, A.fac64[0].p, f32 ? A.fac[0].p : nullptr);
        FMS_LAUNCH_CHECK(ctx);
        k_als_sgd_apply<<<grid, 256, 0, ctx->stream>>>(n, A.rank, d_out_i, row_i.p, win_i.p, A.fac64[1].p, f32 ? A.fac[1].p : nullptr);
        FMS_LAUNCH_CHECK(ctx);
        A.plane_valid[0] = A.plane_valid[1] = false;
    }
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMS_OK;
}

}  // namespace fms
