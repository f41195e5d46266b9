// svm_warp.cu -- K5 fast path: register-resident warp-per-block SDCA sweep for dense fp32 rows, dim = 128*NV4 <= 1024.
//
// Same arithmetic and the same java.util.Random index stream as k_svm_sweep_generic (svm_kernels.cu; takes over
// FlinkML `localDualMethod`/`maximize`, SURVEY.md Appendix A.2, entered from flink-svm/.../SVMImpl.scala:29), laid
// out for HBM streaming:
//   * one warp owns one CoCoA block for the whole round; its private copy of w and its dw live in registers
//     (lane l holds floats [128 v + 4 l, +4) for v < NV4), so a step touches HBM only for the row, y and alpha;
//   * the sampled-row sequence does not depend on w, so rows are fetched STAGES steps ahead by the TMA engine
//     (cp.async.bulk global -> shared, mbarrier complete_tx) into a per-warp ring; the step's label and alpha are
//     fetched with it (STAGES steps ahead, patched when an earlier in-flight step rewrites the same alpha) and
//     the row ids are drawn three further steps ahead, so no scalar load sits on the critical path;
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
    long long local_iterations, float lambda_n, float inv_lambda_n) {
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

        // ---- one RNG stream, drawn RD steps ahead of the row fetches, which run STAGES steps ahead of the maths.
        // Every per-step scalar (row id, y, alpha) is therefore in flight for >= 3 steps: nothing on the critical path
        // waits for DRAM.  Draws are guarded by the step count so the saved RNG state is the state after exactly H draws.
        JRand rng;
        rng.s = rng_state[blk];
        long long drawn = 0;
        const unsigned long long magic = JRand::magic_for(nk);
        auto draw = [&]() -> int {
            if (drawn >= H) return 0;
            ++drawn;
            return rws[rng.next_int_bound_fast(nk, magic)];
        };
        int rq[STAGES];                                           // row id of the step that owns ring stage s
        float yq[STAGES], aq[STAGES];                             // its label and (lane 0) its alpha
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {                        // prologue: steps 0 .. STAGES-1
            rq[s] = 0; yq[s] = 0.f; aq[s] = 0.f;
            if (s < H) {
                rq[s] = draw();
                if (lane == 0) {
                    mbar_expect_tx(smem_u32(&bars[s]), ROW_BYTES);
                    bulk_g2s(smem_u32(ring + s * D), X + (size_t)rq[s] * D, ROW_BYTES, smem_u32(&bars[s]), policy);
                }
                yq[s] = y[rq[s]];
                aq[s] = (lane == 0) ? alpha[rq[s]] : 0.f;
            }
        }
        int rp0 = draw(), rp1 = draw(), rp2 = draw();             // rows of steps STAGES, STAGES+1, STAGES+2

        for (long long base = 0; base < H; base += STAGES) {
            const uint32_t parity = (uint32_t)((base / STAGES) & 1);
#pragma unroll
            for (int s = 0; s < STAGES; ++s) {
                const long long t = base + s;
                if (t >= H) break;
                const uint32_t bar = smem_u32(&bars[s]);
                while (!mbar_try_wait(bar, parity)) {}
                const float4 *xs = reinterpret_cast<const float4 *>(ring + s * D);
                float4 x[NV4];
#pragma unroll
                for (int v = 0; v < NV4; ++v) x[v] = xs[v * 32 + lane];
                float2 dd = make_float2(0.f, 0.f), qq = make_float2(0.f, 0.f);   // packed partials of x.w and x.x
#pragma unroll
                for (int v = 0; v < NV4; ++v) {
                    const float2 xa = make_float2(x[v].x, x[v].y), xb = make_float2(x[v].z, x[v].w);
                    dd = __ffma2_rn(xa, make_float2(w[v].x, w[v].y), dd); dd = __ffma2_rn(xb, make_float2(w[v].z, w[v].w), dd);
                    qq = __ffma2_rn(xa, xa, qq); qq = __ffma2_rn(xb, xb, qq);
                }
                float dot = dd.x + dd.y, q = qq.x + qq.y;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    dot += __shfl_xor_sync(0xffffffffu, dot, o);
                    q += __shfl_xor_sync(0xffffffffu, q, o);
                }
                const int row_cur = rq[s];
                const float y_cur = yq[s];
                const float a_cur = __shfl_sync(0xffffffffu, aq[s], 0);
                // the row buffer is free again: refill it with the row of step t + STAGES
                const bool more = t + STAGES < H;
                const int row_new = rp0;
                __syncwarp();
                if (more && lane == 0) {
                    mbar_expect_tx(bar, ROW_BYTES);
                    bulk_g2s(smem_u32(ring + s * D), X + (size_t)row_new * D, ROW_BYTES, bar, policy);
                }
                // maximize()
                const float g = (y_cur * dot - 1.f) * lambda_n;
                float a_new = a_cur;
                if (g != 0.f) {
                    a_new = (q != 0.f) ? fminf(fmaxf(a_cur - g / q, 0.f), 1.f) : 1.f;     // IEEE divide (it decides the clip), as k_svm_sweep_generic
                    const float c = y_cur * (a_new - a_cur) * inv_lambda_n;      // same expression as k_svm_sweep_generic
                    const float2 cc = make_float2(c, c);
#pragma unroll
                    for (int v = 0; v < NV4; ++v) {
                        const float2 xa = make_float2(x[v].x, x[v].y), xb = make_float2(x[v].z, x[v].w);
                        const float2 wa = __ffma2_rn(cc, xa, make_float2(w[v].x, w[v].y)), wb = __ffma2_rn(cc, xb, make_float2(w[v].z, w[v].w));
                        const float2 da = __ffma2_rn(cc, xa, make_float2(dw[v].x, dw[v].y)), db = __ffma2_rn(cc, xb, make_float2(dw[v].z, dw[v].w));
                        w[v] = make_float4(wa.x, wa.y, wb.x, wb.y);
                        dw[v] = make_float4(da.x, da.y, db.x, db.y);
                    }
                    if (lane == 0) alpha[row_cur] = a_new;
                }
                // alpha values already prefetched for the in-flight steps were read before this store
#pragma unroll
                for (int s2 = 0; s2 < STAGES; ++s2)
                    if (s2 != s && rq[s2] == row_cur) aq[s2] = a_new;
                // scalars of step t + STAGES (loads issued after the store above: same-lane program order)
                if (more) {
                    rq[s] = row_new;
                    yq[s] = y[row_new];
                    aq[s] = (lane == 0) ? alpha[row_new] : 0.f;
                }
                rp0 = rp1; rp1 = rp2; rp2 = draw();
            }
        }
        if (lane == 0) rng_state[blk] = rng.s;
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
    const float lambda_n = (float)(S.lambda * (double)S.n_total);
    k_svm_sweep_warp<NV4><<<(n_local + WARPS - 1) / WARPS, WARPS * 32, smem, ctx->stream>>>(
        S.X, S.y.p, alpha_work, S.rows32.p, S.blk_ptr.p, S.d_local_blocks.p, n_local, (unsigned long long *)S.rng.p, S.w.p,
        S.dw_blocks.p, (long long)local_iterations, lambda_n, (float)(1.0 / (S.lambda * (double)S.n_total)));
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
