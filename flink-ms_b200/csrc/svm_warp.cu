// svm_warp.cu -- K5 fast path: register-resident warp-per-block SDCA sweep for dense fp32 rows, dim = 128*NV4 <= 1024.
//
// Same arithmetic and the same java.util.Random index stream as k_svm_sweep_generic (svm_kernels.cu; takes over
// FlinkML `localDualMethod`/`maximize`, SURVEY.md Appendix A.2, entered from flink-svm/.../SVMImpl.scala:29), laid
// out for HBM streaming:
//   * one warp owns one CoCoA block for the whole round; its private copy of w and its dw live in registers
//     (lane l holds floats [128 v + 4 l, +4) for v < NV4), so a step touches HBM only for the row, y and alpha;
//   * the sampled-row sequence does not depend on w, so rows are fetched STAGES steps ahead by the TMA engine
//     (cp.async.bulk global -> shared, mbarrier complete_tx) into a per-warp ring; y / alpha / row ids are
//     prefetched in registers one and two steps ahead;
//   * the only cross-lane traffic per step is one fused 2-value butterfly (x.w and x.x).
// Algorithmic bytes per step: 4*dim (row) + 4 (y) + 8 (alpha read+write).  HBM-bound; no tensor cores.
#include "common.cuh"

namespace fms {

namespace {

constexpr int STAGES = 4;
constexpr int WARPS = 4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}

template <int NV4>
__global__ void __launch_bounds__(WARPS * 32) k_svm_sweep_warp(
    const float *__restrict__ X, const float *__restrict__ y, float *alpha, const int32_t *__restrict__ rows,
    const int64_t *__restrict__ blk_ptr, const int32_t *__restrict__ local_blocks, int n_local,
    unsigned long long *__restrict__ rng_state, const float *__restrict__ w0, float *__restrict__ dw_out,
    long long local_iterations, float lambda_n) {
    constexpr int D = 128 * NV4;
    constexpr uint32_t ROW_BYTES = D * 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lb = blockIdx.x * WARPS + warp;
    if (lb >= n_local) return;                                   // warps are independent: no CTA-wide barrier below
    float *ring = reinterpret_cast<float *>(smem_raw) + (size_t)warp * STAGES * D;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)WARPS * STAGES * ROW_BYTES) + warp * STAGES;

    const int blk = local_blocks[lb];
    const int64_t r0 = blk_ptr[blk];
    const int nk = (int)(blk_ptr[blk + 1] - r0);
    float4 *dwo = reinterpret_cast<float4 *>(dw_out + (size_t)lb * D);

    float4 w[NV4], dw[NV4];
#pragma unroll
    for (int v = 0; v < NV4; ++v) {
        w[v] = reinterpret_cast<const float4 *>(w0)[v * 32 + lane];
        dw[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (nk > 0) {
        const long long H = local_iterations > 0 ? local_iterations : (long long)nk;
        const int32_t *rws = rows + r0;
        uint64_t policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&bars[s]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();

        // ---- producer side: row ids STAGES (+1 in flight) steps ahead of the consumer
        JRand rng_p, rng_c;
        rng_p.s = rng_c.s = rng_state[blk];
        for (int s = 0; s < STAGES; ++s) {                        // prologue: steps 0 .. STAGES-1
            if (s < H) {
                const int row = rws[rng_p.next_int_bound(nk)];
                if (lane == 0) {
                    mbar_expect_tx(smem_u32(&bars[s]), ROW_BYTES);
                    bulk_g2s(smem_u32(ring + s * D), X + (size_t)row * D, ROW_BYTES, smem_u32(&bars[s]), policy);
                }
            }
        }
        int row_pending = (STAGES < H) ? rws[rng_p.next_int_bound(nk)] : 0;      // row of step STAGES

        // ---- consumer side scalar pipeline: (row, y, alpha) of step t, t+1; row of t+2
        int row_cur = rws[rng_c.next_int_bound(nk)];
        float y_cur = y[row_cur];
        float a_cur = __shfl_sync(0xffffffffu, lane == 0 ? alpha[row_cur] : 0.f, 0);
        int row_n1 = (1 < H) ? rws[rng_c.next_int_bound(nk)] : row_cur;
        int row_n2 = (2 < H) ? rws[rng_c.next_int_bound(nk)] : row_cur;

        for (long long base = 0; base < H; base += STAGES) {
            const uint32_t parity = (uint32_t)((base / STAGES) & 1);
#pragma unroll
            for (int s = 0; s < STAGES; ++s) {
                const long long t = base + s;
                if (t >= H) break;
                // prefetch the scalars of step t+1 (issued before this step's alpha store: see the fix-up below)
                const float y_n1 = y[row_n1];
                float a_n1 = (lane == 0) ? alpha[row_n1] : 0.f;
                // wait for this step's row
                const uint32_t bar = smem_u32(&bars[s]);
                while (!mbar_try_wait(bar, parity)) {}
                const float4 *xs = reinterpret_cast<const float4 *>(ring + s * D);
                float4 x[NV4];
#pragma unroll
                for (int v = 0; v < NV4; ++v) x[v] = xs[v * 32 + lane];
                float d0 = 0.f, d1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
                for (int v = 0; v < NV4; ++v) {
                    d0 = fmaf(x[v].x, w[v].x, d0); d1 = fmaf(x[v].y, w[v].y, d1);
                    d0 = fmaf(x[v].z, w[v].z, d0); d1 = fmaf(x[v].w, w[v].w, d1);
                    q0 = fmaf(x[v].x, x[v].x, q0); q1 = fmaf(x[v].y, x[v].y, q1);
                    q0 = fmaf(x[v].z, x[v].z, q0); q1 = fmaf(x[v].w, x[v].w, q1);
                }
                float dot = d0 + d1, q = q0 + q1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    dot += __shfl_xor_sync(0xffffffffu, dot, o);
                    q += __shfl_xor_sync(0xffffffffu, q, o);
                }
                // the row buffer is free again: refill it with the row of step t + STAGES
                __syncwarp();
                if (t + STAGES < H && lane == 0) {
                    mbar_expect_tx(bar, ROW_BYTES);
                    bulk_g2s(smem_u32(ring + s * D), X + (size_t)row_pending * D, ROW_BYTES, bar, policy);
                }
                if (t + STAGES + 1 < H) row_pending = rws[rng_p.next_int_bound(nk)];
                // maximize()
                const float g = (y_cur * dot - 1.f) * lambda_n;
                float a_new = a_cur;
                if (g != 0.f) {
                    a_new = (q != 0.f) ? fminf(fmaxf(a_cur - g / q, 0.f), 1.f) : 1.f;
                    const float c = y_cur * (a_new - a_cur) / lambda_n;
#pragma unroll
                    for (int v = 0; v < NV4; ++v) {
                        float t0 = c * x[v].x, t1 = c * x[v].y, t2 = c * x[v].z, t3 = c * x[v].w;
                        w[v].x += t0; w[v].y += t1; w[v].z += t2; w[v].w += t3;
                        dw[v].x += t0; dw[v].y += t1; dw[v].z += t2; dw[v].w += t3;
                    }
                    if (lane == 0) alpha[row_cur] = a_new;
                }
                // rotate the scalar pipeline; alpha[row_n1] was read before the store above
                a_n1 = __shfl_sync(0xffffffffu, a_n1, 0);
                if (row_n1 == row_cur) a_n1 = a_new;
                row_cur = row_n1; y_cur = y_n1; a_cur = a_n1;
                row_n1 = row_n2;
                if (t + 3 < H) row_n2 = rws[rng_c.next_int_bound(nk)];
            }
        }
        if (lane == 0) rng_state[blk] = rng_c.s;
    }
#pragma unroll
    for (int v = 0; v < NV4; ++v) dwo[v * 32 + lane] = dw[v];
}

template <int NV4>
int launch(fms_ctx *ctx, float *alpha_work, int64_t local_iterations) {
    SvmState &S = ctx->svm;
    const int n_local = (int)S.local_blocks.size();
    if (n_local == 0) return FMS_OK;
    const size_t smem = (size_t)WARPS * STAGES * 128 * NV4 * 4 + WARPS * STAGES * 8;
    FMS_CUDA(ctx, cudaFuncSetAttribute(k_svm_sweep_warp<NV4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float lambda_n = (float)(S.lambda * (double)S.n);
    k_svm_sweep_warp<NV4><<<(n_local + WARPS - 1) / WARPS, WARPS * 32, smem, ctx->stream>>>(
        S.X, S.y.p, alpha_work, S.rows32.p, S.blk_ptr.p, S.d_local_blocks.p, n_local, (unsigned long long *)S.rng.p, S.w.p,
        S.dw_blocks.p, (long long)local_iterations, lambda_n);
    FMS_LAUNCH_CHECK(ctx);
    return FMS_OK;
}

}  // namespace

int svm_warp_supported(const fms_ctx *ctx) {
    const SvmState &S = ctx->svm;
    return S.x_kind == FMS_X_DENSE_F32 && S.dim % 128 == 0 && S.dim <= 1024 && S.rng_mode == FMS_RNG_PER_BLOCK &&
           ((uintptr_t)S.X % 16) == 0 && S.rows32.p != nullptr;
}

int svm_launch_sweep_warp(fms_ctx *ctx, float *alpha_work, int64_t local_iterations) {
    switch (ctx->svm.dim / 128) {
        case 1: return launch<1>(ctx, alpha_work, local_iterations);
        case 2: return launch<2>(ctx, alpha_work, local_iterations);
        case 3: return launch<3>(ctx, alpha_work, local_iterations);
        case 4: return launch<4>(ctx, alpha_work, local_iterations);
        case 5: return launch<5>(ctx, alpha_work, local_iterations);
        case 6: return launch<6>(ctx, alpha_work, local_iterations);
        case 7: return launch<7>(ctx, alpha_work, local_iterations);
        case 8: return launch<8>(ctx, alpha_work, local_iterations);
    }
    return fail(ctx, FMS_ERR_UNSUPPORTED, "warp SVM sweep: dim %d", ctx->svm.dim);
}

}  // namespace fms
