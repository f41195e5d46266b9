// svm_kernels.cu -- K5-generic (CoCoA block sweep, SDCA local solver), K6 (delta reduce + weight update),
// K7 (predict).  The register-resident warp-per-block sweep for dense D = 1024-class inputs is svm_warp.cu.
//
// K5 takes over FlinkML's `localDualMethod` + `maximize` (entered from
// flink-svm/src/main/scala/de/tub/it4bi/SVMImpl.scala:29).  Per block, H strictly sequential steps:
//     idx = r.nextInt(n_k)                       (java.util.Random, bit-exact)
//     g   = (y * (x . w) - 1) * lambda * n
//     a'  = clip(a - g / (x . x), 0, 1)   (1 if x . x == 0);  skipped when g == 0
//     w  += x * y * (a' - a) / (lambda n) ;  dw += same ;  alpha[idx] = a'
// K6 takes over `w.union(dws.map(_ * scaling)).reduce(_ + _)` (row a14), summed in ascending block order.
// K7 takes over SVM.predict / SVMPredictRandom.java:68-88 (strict `raw > threshold`).
#include "common.cuh"

namespace fms {

__device__ __forceinline__ float block_sum2(float a, float &b_inout, float *red /* 2*32 floats */) {
    // reduces (a, b) across a 256-thread CTA; every thread returns the totals
    float b = b_inout;
    a = warp_sum(a);
    b = warp_sum(b);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();                    // protect red[] from the previous use
    if (l == 0) { red[w] = a; red[32 + w] = b; }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    float ta = 0.f, tb = 0.f;
    for (int i = 0; i < nw; ++i) { ta += red[i]; tb += red[32 + i]; }   // fixed order -> deterministic
    b_inout = tb;
    return ta;
}

// One CTA per CoCoA block.  w_local and dw live in shared memory (dim floats each).
// x_kind: FMS_X_DENSE_F32 (X row-major n x dim) or FMS_X_CSR_F64 (device values are fp32).
__global__ void __launch_bounds__(256) k_svm_sweep_generic(
    const float *__restrict__ X, const int64_t *__restrict__ cptr, const int32_t *__restrict__ cidx,
    const float *__restrict__ cval, int x_kind, const float *__restrict__ y, float *__restrict__ alpha,
    const int64_t *__restrict__ rows, const int64_t *__restrict__ blk_ptr, const int32_t *__restrict__ local_blocks,
    unsigned long long *__restrict__ rng_state, int rng_shared, const float *__restrict__ w0, float *__restrict__ dw_out,
    int dim, long long local_iterations, float lambda_n, float inv_lambda_n) {
    extern __shared__ float sm[];
    float *w = sm, *dw = sm + dim, *red = sm + 2 * dim;
    const int lb = blockIdx.x;
    const int blk = local_blocks[lb];
    const int64_t r0 = blk_ptr[blk], nk = blk_ptr[blk + 1] - r0;
    float *dwo = dw_out + (int64_t)lb * dim;
    for (int j = threadIdx.x; j < dim; j += blockDim.x) { w[j] = w0[j]; dw[j] = 0.f; }
    __syncthreads();
    if (nk > 0) {
        JRand rng;
        rng.s = rng_state[rng_shared ? 0 : blk];
        const long long H = local_iterations > 0 ? local_iterations : (long long)nk;
        for (long long it = 0; it < H; ++it) {
            const int idx = rng.next_int_bound((int)nk);        // every thread draws the same stream
            const int64_t row = rows[r0 + idx];
            float dot = 0.f, q = 0.f;
            if (x_kind == FMS_X_DENSE_F32) {
                const float *x = X + row * (int64_t)dim;
                for (int j = threadIdx.x; j < dim; j += blockDim.x) { const float v = x[j]; dot = fmaf(v, w[j], dot); q = fmaf(v, v, q); }
            } else {
                for (int64_t p = cptr[row] + threadIdx.x; p < cptr[row + 1]; p += blockDim.x) {
                    const float v = cval[p]; dot = fmaf(v, w[cidx[p]], dot); q = fmaf(v, v, q);
                }
            }
            const float yy = y[row], a = alpha[row];             // read before the barriers inside block_sum2
            dot = block_sum2(dot, q, red);
            const float g = (yy * dot - 1.f) * lambda_n;
            if (g != 0.f) {                                       // uniform across the CTA
                const float an = (q != 0.f) ? fminf(fmaxf(a - g / q, 0.f), 1.f) : 1.f;
                const float da = an - a;
                const float c = yy * da * inv_lambda_n;      // the same fp32 reciprocal in every sweep kernel (svm_warp.cu)
                if (x_kind == FMS_X_DENSE_F32) {
                    const float *x = X + row * (int64_t)dim;
                    for (int j = threadIdx.x; j < dim; j += blockDim.x) { const float t = c * x[j]; w[j] += t; dw[j] += t; }
                } else {
                    // column indices inside a row are distinct (a SparseVector), so no write conflicts
                    for (int64_t p = cptr[row] + threadIdx.x; p < cptr[row + 1]; p += blockDim.x) {
                        const float t = c * cval[p]; w[cidx[p]] += t; dw[cidx[p]] += t;
                    }
                }
                if (threadIdx.x == 0) alpha[row] = a + da;        // tmpAlpha(idx) += deltaAlpha
            }
            __syncthreads();                                      // w/alpha visible to the next step
        }
        if (threadIdx.x == 0 && !rng_shared) rng_state[blk] = rng.s;
    }
    for (int j = threadIdx.x; j < dim; j += blockDim.x) dwo[j] = dw[j];
}

// rng_shared sweeps must run block after block on one stream: a single CTA walks all blocks in order.
__global__ void __launch_bounds__(256) k_svm_sweep_generic_shared(
    const float *__restrict__ X, const int64_t *__restrict__ cptr, const int32_t *__restrict__ cidx,
    const float *__restrict__ cval, int x_kind, const float *__restrict__ y, float *__restrict__ alpha,
    const int64_t *__restrict__ rows, const int64_t *__restrict__ blk_ptr, int n_blocks,
    unsigned long long *__restrict__ rng_state, const float *__restrict__ w0, float *__restrict__ dw_out, int dim,
    long long local_iterations, float lambda_n, float inv_lambda_n) {
    extern __shared__ float sm[];
    float *w = sm, *dw = sm + dim, *red = sm + 2 * dim;
    JRand rng;
    rng.s = rng_state[0];
    for (int blk = 0; blk < n_blocks; ++blk) {
        const int64_t r0 = blk_ptr[blk], nk = blk_ptr[blk + 1] - r0;
        float *dwo = dw_out + (int64_t)blk * dim;
        __syncthreads();
        for (int j = threadIdx.x; j < dim; j += blockDim.x) { w[j] = w0[j]; dw[j] = 0.f; }
        __syncthreads();
        const long long H = nk > 0 ? (local_iterations > 0 ? local_iterations : (long long)nk) : 0;
        for (long long it = 0; it < H; ++it) {
            const int idx = rng.next_int_bound((int)nk);
            const int64_t row = rows[r0 + idx];
            float dot = 0.f, q = 0.f;
            if (x_kind == FMS_X_DENSE_F32) {
                const float *x = X + row * (int64_t)dim;
                for (int j = threadIdx.x; j < dim; j += blockDim.x) { const float v = x[j]; dot = fmaf(v, w[j], dot); q = fmaf(v, v, q); }
            } else {
                for (int64_t p = cptr[row] + threadIdx.x; p < cptr[row + 1]; p += blockDim.x) {
                    const float v = cval[p]; dot = fmaf(v, w[cidx[p]], dot); q = fmaf(v, v, q);
                }
            }
            const float yy = y[row], a = alpha[row];             // read before the barriers inside block_sum2
            dot = block_sum2(dot, q, red);
            const float g = (yy * dot - 1.f) * lambda_n;
            if (g != 0.f) {
                const float an = (q != 0.f) ? fminf(fmaxf(a - g / q, 0.f), 1.f) : 1.f;
                const float da = an - a;
                const float c = yy * da * inv_lambda_n;      // the same fp32 reciprocal in every sweep kernel (svm_warp.cu)
                if (x_kind == FMS_X_DENSE_F32) {
                    const float *x = X + row * (int64_t)dim;
                    for (int j = threadIdx.x; j < dim; j += blockDim.x) { const float t = c * x[j]; w[j] += t; dw[j] += t; }
                } else {
                    for (int64_t p = cptr[row] + threadIdx.x; p < cptr[row + 1]; p += blockDim.x) {
                        const float t = c * cval[p]; w[cidx[p]] += t; dw[cidx[p]] += t;
                    }
                }
                if (threadIdx.x == 0) alpha[row] = a + da;
            }
            __syncthreads();
        }
        for (int j = threadIdx.x; j < dim; j += blockDim.x) dwo[j] = dw[j];
    }
    if (threadIdx.x == 0) rng_state[0] = rng.s;
}

// delta[j] = sum over local blocks of dw_blocks[k][j], in a FIXED two-level order (deterministic): the blocks are cut
// into RED_CHUNKS contiguous chunks summed in ascending order by one thread each, then the chunk sums are added in
// ascending chunk order.  (A single thread walking thousands of blocks per column was latency-bound.)
constexpr int RED_CHUNKS = 64;
__global__ void k_svm_reduce_partial(const float *__restrict__ dw_blocks, int n_local, int dim, float *__restrict__ partial) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (j >= dim) return;
    const int per = (n_local + RED_CHUNKS - 1) / RED_CHUNKS;
    const int k0 = c * per, k1 = min(n_local, k0 + per);
    float s = 0.f;
    for (int k = k0; k < k1; ++k) s += dw_blocks[(int64_t)k * dim + j];
    partial[(int64_t)c * dim + j] = s;
}
__global__ void k_svm_reduce_delta(const float *__restrict__ partial, int dim, float *__restrict__ delta) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= dim) return;
    float s = 0.f;
#pragma unroll 8
    for (int c = 0; c < RED_CHUNKS; ++c) s += partial[(int64_t)c * dim + j];
    delta[j] = s;
}

__global__ void k_svm_apply(float *__restrict__ w, const float *__restrict__ delta, int dim, float scaling) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < dim) w[j] = fmaf(delta[j], scaling, w[j]);
}

// alpha += stepsize * (alpha_tmp - alpha)   (alpha_block += deltaAlphas * stepsize)
__global__ void k_svm_alpha_commit(float *__restrict__ alpha, const float *__restrict__ tmp, int64_t n, float stepsize) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) alpha[i] = fmaf(stepsize, tmp[i] - alpha[i], alpha[i]);
}

int svm_warp_supported(const fms_ctx *ctx);
int svm_launch_sweep_warp(fms_ctx *ctx, float *alpha_work, int64_t local_iterations);

int svm_launch_sweep(fms_ctx *ctx, int64_t local_iterations) {
    SvmState &S = ctx->svm;
    cudaStream_t st = ctx->stream;
    const int n_local = (int)S.local_blocks.size();
    float *alpha_work = S.alpha.p;
    if (S.stepsize != 1.0) {
        FMS_CUDA(ctx, S.alpha_tmp.alloc((size_t)S.n));
        FMS_CUDA(ctx, cudaMemcpyAsync(S.alpha_tmp.p, S.alpha.p, (size_t)S.n * sizeof(float), cudaMemcpyDeviceToDevice, st));
        alpha_work = S.alpha_tmp.p;
    }
    const float lambda_n = (float)(S.lambda * (double)S.n_total), inv_lambda_n = (float)(1.0 / (S.lambda * (double)S.n_total));
    bool use_warp = false;
    if (S.kernel != FMS_SVM_KERNEL_GENERIC && S.rng_mode == FMS_RNG_PER_BLOCK) use_warp = svm_warp_supported(ctx) != 0;
    if (S.kernel == FMS_SVM_KERNEL_WARP && !use_warp)
        return fail(ctx, FMS_ERR_UNSUPPORTED, "warp SVM sweep needs dense fp32 X, dim %% 128 == 0, dim <= 1024, per-block RNG");
    if (use_warp) {
        FMS_TRY(svm_launch_sweep_warp(ctx, alpha_work, local_iterations));
    } else {
        const size_t smem = (size_t)(2 * S.dim + 64) * sizeof(float);
        if (smem > ctx->smem_optin) return fail(ctx, FMS_ERR_UNSUPPORTED, "SVM dim %d needs %zu B shared memory", S.dim, smem);
        if (S.rng_mode == FMS_RNG_SHARED_SUBTASK) {
            if (S.shard_count != 1) return fail(ctx, FMS_ERR_UNSUPPORTED, "shared-subtask RNG cannot be sharded");
            FMS_CUDA(ctx, cudaFuncSetAttribute(k_svm_sweep_generic_shared, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_svm_sweep_generic_shared<<<1, 256, smem, st>>>(S.X, S.csr_ptr.p, S.csr_idx.p, S.csr_val.p, S.x_kind, S.y.p, alpha_work,
                                                             S.rows.p, S.blk_ptr.p, S.blocks, (unsigned long long *)S.rng.p, S.w.p,
                                                             S.dw_blocks.p, S.dim, (long long)local_iterations, lambda_n, inv_lambda_n);
            FMS_LAUNCH_CHECK(ctx);
        } else if (n_local > 0) {
            FMS_CUDA(ctx, cudaFuncSetAttribute(k_svm_sweep_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_svm_sweep_generic<<<n_local, 256, smem, st>>>(S.X, S.csr_ptr.p, S.csr_idx.p, S.csr_val.p, S.x_kind, S.y.p, alpha_work,
                                                            S.rows.p, S.blk_ptr.p, S.d_local_blocks.p,
                                                            (unsigned long long *)S.rng.p, 0, S.w.p, S.dw_blocks.p, S.dim,
                                                            (long long)local_iterations, lambda_n, inv_lambda_n);
            FMS_LAUNCH_CHECK(ctx);
        }
    }
    if (S.stepsize != 1.0 && S.n > 0) {
        k_svm_alpha_commit<<<(unsigned)((S.n + 255) / 256), 256, 0, st>>>(S.alpha.p, S.alpha_tmp.p, S.n, (float)S.stepsize);
        FMS_LAUNCH_CHECK(ctx);
    }
    return svm_launch_reduce_delta(ctx);
}

int svm_launch_reduce_delta(fms_ctx *ctx) {
    SvmState &S = ctx->svm;
    const int n_local = (int)S.local_blocks.size();
    FMS_CUDA(ctx, S.red_partial.alloc((size_t)RED_CHUNKS * S.dim));
    k_svm_reduce_partial<<<dim3((S.dim + 127) / 128, RED_CHUNKS), 128, 0, ctx->stream>>>(S.dw_blocks.p, n_local, S.dim, S.red_partial.p);
    FMS_LAUNCH_CHECK(ctx);
    k_svm_reduce_delta<<<(S.dim + 127) / 128, 128, 0, ctx->stream>>>(S.red_partial.p, S.dim, S.delta.p);
    FMS_LAUNCH_CHECK(ctx);
    return FMS_OK;
}

int svm_launch_apply(fms_ctx *ctx) {
    SvmState &S = ctx->svm;
    k_svm_apply<<<(S.dim + 255) / 256, 256, 0, ctx->stream>>>(S.w.p, S.delta.p, S.dim, (float)(S.stepsize / (double)S.blocks));
    FMS_LAUNCH_CHECK(ctx);
    S.w64_valid = false;
    return FMS_OK;
}

// ------------------------------------------------------------------------------------------------ K6' exchange + apply, one process per GPU
// The delta-w all-reduce and the apply step as ONE kernel per rank over IPC-mapped peer memory (what k_svm_allreduce_apply is
// for the single-process group ctx): every rank pushes its dim floats into slot [epoch parity][rank] of EVERY rank's exchange
// buffer (NVLink stores), publishes the epoch into flag word [rank] of every buffer (system-scope release), waits until all flag
// words of its OWN buffer reached the epoch, sums the slots in rank order (identical on every rank) and applies the step.
// Two parities suffice: a rank rewrites a slot two rounds later, after an apply that waited for flags every peer set only after
// finishing its previous round's reads.
__device__ __forceinline__ int ld_acquire_sys_i(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_i(int *p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__global__ void __launch_bounds__(1024) k_svm_exchange_apply(float *__restrict__ w, const float *__restrict__ delta, float *const *__restrict__ peers, int n,
                                                             int me, int dim, int epoch, float scaling, int *__restrict__ err) {
    const size_t slot = ((size_t)(epoch & 1) * n + me) * dim, flags_at = (size_t)2 * n * dim;
    for (int j = threadIdx.x; j < dim; j += blockDim.x) {
        const float v = delta[j];
        for (int p = 0; p < n; ++p) peers[p][slot + j] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        for (int p = 0; p < n; ++p) st_release_sys_i(reinterpret_cast<int *>(peers[p] + flags_at) + me, epoch);
    }
    if (threadIdx.x < n) {
        const int *f = reinterpret_cast<const int *>(peers[me] + flags_at) + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys_i(f) < epoch) {
            if (clock64() - t0 > 8000000000LL) { atomicExch(err, 1); break; }     // ~4 s: a peer is gone; do not hang the GPU
        }
    }
    __syncthreads();
    const float *mine = peers[me] + (size_t)(epoch & 1) * n * dim;
    for (int j = threadIdx.x; j < dim; j += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < n; ++r) s += __ldcv(mine + (size_t)r * dim + j);
        w[j] = fmaf(s, scaling, w[j]);
    }
}

static size_t xchg_floats(int n_ranks, int dim) { return (size_t)2 * n_ranks * dim + 64; }

int svm_exchange_handle(fms_ctx *ctx, int n_ranks, void *handle64) {
    SvmState &S = ctx->svm;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    if (!S.loaded) return fail(ctx, FMS_ERR_STATE, "fms_svm_exchange_ipc_handle: load first");
    if (n_ranks < 2 || n_ranks > 64 || !handle64) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_exchange_ipc_handle: n_ranks %d", n_ranks);
    svm_exchange_close(ctx);
    S.xchg.release();                                            // a fresh allocation per geometry: handles name allocations
    FMS_CUDA(ctx, S.xchg.alloc(xchg_floats(n_ranks, S.dim)));
    FMS_CUDA(ctx, cudaMemsetAsync(S.xchg.p, 0, xchg_floats(n_ranks, S.dim) * sizeof(float), ctx->stream));
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    S.xchg_dim = S.dim;
    cudaIpcMemHandle_t h;
    FMS_CUDA(ctx, cudaIpcGetMemHandle(&h, S.xchg.p));
    memcpy(handle64, &h, sizeof h);
    return FMS_OK;
}

void svm_exchange_close(fms_ctx *ctx) {
    SvmState &S = ctx->svm;
    for (int r = 0; r < S.xchg_n; ++r) {
        if (S.xchg_open[r]) cudaIpcCloseMemHandle(S.xchg_open[r]);
        S.xchg_open[r] = nullptr;
    }
    S.xchg_n = 0;
    S.xchg_epoch = 0;
}

int svm_exchange_open(fms_ctx *ctx, const void *handles, int n_ranks, int my_rank) {
    SvmState &S = ctx->svm;
    if (!S.xchg.p || S.xchg_dim != S.dim) return fail(ctx, FMS_ERR_STATE, "fms_svm_exchange_open_peers: fms_svm_exchange_ipc_handle first");
    if (n_ranks < 2 || n_ranks > 64 || my_rank < 0 || my_rank >= n_ranks || !handles) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_svm_exchange_open_peers: rank %d of %d", my_rank, n_ranks);
    svm_exchange_close(ctx);
    unsigned long long tab[64];
    for (int r = 0; r < n_ranks; ++r) {
        if (r == my_rank) { tab[r] = (unsigned long long)(uintptr_t)S.xchg.p; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * (size_t)r, sizeof h);
        void *pp = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&pp, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            S.xchg_n = r;
            svm_exchange_close(ctx);
            return fail(ctx, FMS_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        }
        S.xchg_open[r] = pp;
        tab[r] = (unsigned long long)(uintptr_t)pp;
    }
    S.xchg_n = n_ranks;
    S.xchg_rank = my_rank;
    S.xchg_epoch = 0;
    FMS_CUDA(ctx, S.xchg_tab.alloc(64));
    FMS_CUDA(ctx, S.xchg_err.alloc(1));
    FMS_CUDA(ctx, cudaMemcpyAsync(S.xchg_tab.p, tab, sizeof(unsigned long long) * n_ranks, cudaMemcpyHostToDevice, ctx->stream));
    FMS_CUDA(ctx, cudaMemsetAsync(S.xchg_err.p, 0, sizeof(int), ctx->stream));
    // fresh flag words; the HOST must put one barrier across the ranks between this call and the first round
    FMS_CUDA(ctx, cudaMemsetAsync(S.xchg.p + (size_t)2 * n_ranks * S.dim, 0, 64 * sizeof(int), ctx->stream));
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMS_OK;
}

int svm_launch_exchange_apply(fms_ctx *ctx) {
    SvmState &S = ctx->svm;
    if (S.xchg_n < 2) return fail(ctx, FMS_ERR_STATE, "fms_svm_round_exchange: fms_svm_exchange_open_peers first");
    if (S.xchg_dim != S.dim) return fail(ctx, FMS_ERR_STATE, "fms_svm_round_exchange: the exchange buffers were set up for dim %d, the loaded problem has %d", S.xchg_dim, S.dim);
    ++S.xchg_epoch;
    k_svm_exchange_apply<<<1, 1024, 0, ctx->stream>>>(S.w.p, S.delta.p, reinterpret_cast<float *const *>(S.xchg_tab.p), S.xchg_n, S.xchg_rank, S.dim,
                                                       S.xchg_epoch, (float)(S.stepsize / (double)S.blocks), S.xchg_err.p);
    FMS_LAUNCH_CHECK(ctx);
    S.w64_valid = false;
    return FMS_OK;
}

// ------------------------------------------------------------------------------------------------ K7 predict
// One warp per query row, fp64 accumulate.  d_X: dense fp32 / dense fp64 / CSR fp64 values.
__global__ void k_svm_predict(const void *__restrict__ Xv, int x_kind, const int64_t *__restrict__ cptr,
                              const int32_t *__restrict__ cidx, int64_t n, int dim, const double *__restrict__ w,
                              double threshold, int decision, double *__restrict__ out) {
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    double s = 0.0;
    if (x_kind == FMS_X_DENSE_F32) {
        const float *x = (const float *)Xv + row * (int64_t)dim;
        for (int j = lane; j < dim; j += 32) s = fma((double)x[j], w[j], s);
    } else if (x_kind == FMS_X_DENSE_F64) {
        const double *x = (const double *)Xv + row * (int64_t)dim;
        for (int j = lane; j < dim; j += 32) s = fma(x[j], w[j], s);
    } else {
        const double *v = (const double *)Xv;
        for (int64_t p = cptr[row] + lane; p < cptr[row + 1]; p += 32) {
            const int32_t c = cidx[p];
            if (c >= 0 && c < dim) s = fma(v[p], w[c], s);       // feature missing from the model contributes nothing
        }
    }
    s = warp_sum(s);
    if (lane == 0) out[row] = decision ? s : (s > threshold ? 1.0 : -1.0);
}

int svm_predict_device(fms_ctx *ctx, const void *d_X, int x_kind, const int64_t *d_ptr, const int32_t *d_idx, int64_t n,
                       int32_t dim, const double *d_w, double threshold, int decision, double *d_out) {
    if (n <= 0) return FMS_OK;
    const int64_t threads = n * 32;
    k_svm_predict<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(d_X, x_kind, d_ptr, d_idx, n, dim, d_w, threshold,
                                                                              decision, d_out);
    FMS_LAUNCH_CHECK(ctx);
    return FMS_OK;
}

}  // namespace fms
