// als_wide.cu -- ALS half-step for 64 < rank <= 128 (BASELINE configs[2] is rank 128): fp32 Gramians on the CUDA cores, the
// solve on the warp-level tensor path of chol_mma.cuh (solve_tiles<8>), the two overlapped inside one persistent CTA per SM.
//
// Why not the round-1 generic kernel: its CTA-wide column Cholesky (three __syncthreads per column, a div/mod per trailing
// entry) needs 1.57 s for configs[2]'s user half-step on one B200.  Why not tcgen05 like rank 64: a 128 x 128 fp32 system is
// 37 KB in the fragment-tile layout, so at most five are resident per SM, and the solver -- not the Gramian -- bounds the
// half-step (tools/micro/chol_mma_test 8: 22 400 cycles per system per SM with four solver warps = 77 ms per million systems);
// the Gramian of ~100-rating rows costs about a third of that on plain FFMA.  What this kernel does:
//   * a producer warp streams partner rows, 16 ratings per stage, through a 4-deep cp.async ring (mbarrier full / empty);
//   * warps 0..3 (128 threads for the 136 8 x 8 blocks of the lower block triangle) accumulate G = sum y y^T and
//     b = sum r y, then write -(G + lambda n I) in the fragment-tile layout and -b into one of NBUF system buffers;
//   * NSOLVE solver warps each solve one buffered system at a time (HMMA blocked Cholesky, fp16-packed products on a power-of-two scaled system) and store the fp32 row.
// Reference: the per-row normal equations of FlinkML ALS.updateFactors (oracle/flinkml_oracle.c:als_half_step cites the Scala).
#include "chol_mma.cuh"
#include "common.cuh"

namespace fms {
namespace wide {

constexpr int DP = 128, MT = 8;
constexpr int ACC_WARPS = 4, W_PROD = ACC_WARPS, W_SOLVE0 = ACC_WARPS + 1;
constexpr int NSOLVE = 5, NBUF = 5, CH = 16, NST = 4;
constexpr int THREADS = (ACC_WARPS + 1 + NSOLVE) * 32;
using TL = cm::Tiles<MT>;
constexpr int BUF_FLOATS = TL::FLOATS + 3 * DP;                              // tiles | bz | scratch | xs
constexpr int STAGE_FLOATS = CH * DP + CH;                                   // CH partner rows | CH ratings
constexpr int N_BARS = 2 * NST + 2 * NBUF;
constexpr size_t SMEM_BYTES = (size_t)(NST * STAGE_FLOATS + NBUF * BUF_FLOATS) * 4 + N_BARS * 8 + 32;     // + scale slots (2 x 3 floats)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(1000000u)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void acc_sync() { asm volatile("bar.sync 1, %0;" ::"n"(ACC_WARPS * 32) : "memory"); }

// float offset of element (k, i) of the fragment-tile layout; valid for every (k, i) inside a stored tile (also the entries
// above the diagonal of a diagonal tile, which the solver multiplies by exact zeros and therefore must be finite)
__device__ __forceinline__ int tile_at(int k, int i) {
    const int I = k >> 4, j = i >> 3;
    return (I * (I + 1) + j) * 128 + ((k & 7) * 4 + ((i & 7) >> 1)) * 4 + ((k >> 3) & 1) * 2 + (i & 1);
}
// A staged partner row keeps its 8-float block m as two 16-byte pieces at piece slots 2 m + (half ^ ((m >> 2) & 1)): eight
// consecutive blocks then start in eight different bank groups (a warp reads blocks bj .. bj + 7 in one LDS.128).
__device__ __forceinline__ int piece_slot(int m, int half) { return 2 * m + (half ^ ((m >> 2) & 1)); }

__global__ void __launch_bounds__(THREADS, 1) k_als_rows_wide(const int64_t *__restrict__ ptr, const int32_t *__restrict__ col,
                                                              const float *__restrict__ val, const float *__restrict__ P,
                                                              float *__restrict__ T, int d, float lambda, int64_t row_begin,
                                                              int64_t row_end, int *__restrict__ info) {
    extern __shared__ __align__(16) float smem[];
    float *stage = smem;                                     // NST x (CH x DP | CH)
    float *bufs = stage + NST * STAGE_FLOATS;                // NBUF x BUF_FLOATS
    const uint32_t bar0 = smem_u32(bufs + NBUF * BUF_FLOATS);
    volatile float *scale_slot = bufs + NBUF * BUF_FLOATS + 2 * N_BARS;      // two floats behind the barriers
    auto st_full = [&](int s) { return bar0 + 8u * s; };
    auto st_empty = [&](int s) { return bar0 + 8u * (NST + s); };
    auto bar_full = [&](int b) { return bar0 + 8u * (2 * NST + b); };
    auto bar_empty = [&](int b) { return bar0 + 8u * (2 * NST + NBUF + b); };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < NST * STAGE_FLOATS; i += THREADS) stage[i] = 0.f;     // columns d .. DP-1 stay zero for good
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(st_full(s), 33); mbar_init(st_empty(s), ACC_WARPS); }
        for (int b = 0; b < NBUF; ++b) { mbar_init(bar_full(b), 1); mbar_init(bar_empty(b), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t n_rows = row_end - row_begin;
    const int64_t n_jobs = n_rows > blockIdx.x ? (n_rows - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // rows b, b + grid, ...

    if (warp == W_PROD) {
        // ============================================================ producer warp: one continuous stream of 16-rating chunks
        // lane (r = lane / 2, hf = lane % 2) copies half of partner row r (d / 8 pieces of 16 bytes; d % 4 == 0, host-checked)
        const int r = lane >> 1, hf = lane & 1, q4 = d >> 2;
        uint32_t cc = 0;
        for (int64_t job = 0; job < n_jobs; ++job) {
            const int64_t row = row_begin + blockIdx.x + job * gridDim.x;
            const int64_t s = ptr[row], e = ptr[row + 1];
            for (int64_t k0 = s; k0 < e; k0 += CH, ++cc) {
                const int sg = (int)(cc % NST);
                mbar_wait(st_empty(sg), ((cc / NST) & 1) ^ 1u);
                float *st = stage + sg * STAGE_FLOATS;
                const bool live = k0 + r < e;
                if (live) {
                    const float *src = P + (int64_t)col[k0 + r] * d;
                    for (int p4 = hf; p4 < q4; p4 += 2) cp_async16(smem_u32(st + r * DP + 4 * piece_slot(p4 >> 1, p4 & 1)), src + 4 * p4);
                } else {
                    for (int p4 = hf; p4 < q4; p4 += 2) *reinterpret_cast<float4 *>(st + r * DP + 4 * piece_slot(p4 >> 1, p4 & 1)) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (hf == 0) st[CH * DP + r] = live ? val[k0 + r] : 0.f;
                cp_async_arrive_noinc(st_full(sg));
                __syncwarp();
                if (lane == 0) mbar_arrive(st_full(sg));      // releases the plain stores above
            }
        }
    } else if (warp < ACC_WARPS) {
        // ============================================================ Gramian warps
        // 136 lower 8 x 8 blocks on 128 threads (one full warp per SM sub-partition; a fifth, mostly empty warp would double the
        // instruction stream of its sub-partition): threads 0..119 own the strictly lower blocks (bi > bj, row-major), threads
        // 120..127 the diagonal blocks 0..7; row k of diagonal block m >= 8 is one extra 8-vector of thread (bi = m, bj = k < 8),
        // which already holds y[8 m .. 8 m + 7] for its own block.
        int bi, bj;
        if (tid < 120) {
            int t = tid;
            bi = 1;
            while (t >= bi) { t -= bi; ++bi; }
            bj = t;
        } else {
            bi = bj = tid - 120;
        }
        const bool has_diag = bi >= 8 && bj < 8 && bi > bj;
        const int odiag = 4 * piece_slot(bi, (bj >> 2) & 1) + (bj & 3);              // y[8 bi + bj] of a staged row
        const int oa0 = 4 * piece_slot(bi, 0), oa1 = 4 * piece_slot(bi, 1), ob0 = 4 * piece_slot(bj, 0), ob1 = 4 * piece_slot(bj, 1);
        const int ocol = 4 * piece_slot(tid >> 3, (tid >> 2) & 1) + (tid & 3);                     // column tid of a staged row
        uint32_t cc = 0;
        for (int64_t job = 0; job < n_jobs; ++job) {
            const int64_t row = row_begin + blockIdx.x + job * gridDim.x;
            const int64_t s = ptr[row], e = ptr[row + 1];
            float acc[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
            float bacc = 0.f;
            float dacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int64_t k0 = s; k0 < e; k0 += CH, ++cc) {
                const int sg = (int)(cc % NST);
                mbar_wait(st_full(sg), (cc / NST) & 1);
                const float *st = stage + sg * STAGE_FLOATS;
#pragma unroll 4
                for (int r = 0; r < CH; ++r) {
                    const float4 a0 = *reinterpret_cast<const float4 *>(st + r * DP + oa0);
                    const float4 a1 = *reinterpret_cast<const float4 *>(st + r * DP + oa1);
                    const float4 b0 = *reinterpret_cast<const float4 *>(st + r * DP + ob0);
                    const float4 b1 = *reinterpret_cast<const float4 *>(st + r * DP + ob1);
                    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
                    if (has_diag) {
                        const float ak = st[r * DP + odiag];
#pragma unroll
                        for (int j = 0; j < 8; ++j) dacc[j] = fmaf(ak, a[j], dacc[j]);
                    }
                }
#pragma unroll
                for (int r = 0; r < CH; ++r) bacc = fmaf(st[CH * DP + r], st[r * DP + ocol], bacc);
                __syncwarp();
                if (lane == 0) mbar_arrive(st_empty(sg));
            }
            // ---- hand the system to a solver warp
            const int bf = (int)(job % NBUF);
            mbar_wait(bar_empty(bf), (uint32_t)((job / NBUF) & 1) ^ 1u);
            float *S = bufs + (size_t)bf * BUF_FLOATS;
            const float reg = lambda * (float)(e - s);
            // fp16-packed solver: the matrix goes out scaled by the power of two that brings its first diagonal entry into [1, 2)
            // (thread 120 owns diagonal block 0), the right-hand side by that and by a power of two of its own (below).
            if (tid == 120) {
                const uint32_t ex = (__float_as_uint(acc[0][0] + reg) >> 23) & 0x1ffu;
                scale_slot[3 * (job & 1)] = (ex >= 1u && ex <= 253u) ? __uint_as_float((254u - ex) << 23) : 1.f;
            }
            if (warp == 0) {                                            // largest |b| among the first 32 components
                float m = fabsf(bacc);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                if (lane == 0) scale_slot[3 * (job & 1) + 1] = m;
            }
            acc_sync();
            const float sc = scale_slot[3 * (job & 1)];
            float sb = sc, inv_sb = 1.f;
            {
                // the right-hand side gets a second power of two that brings its own magnitude to ~8 (an fp16 lo part resolves
                // 6e-8 absolute); x is linear in b, the solver warp multiplies the inverse back in
                const uint32_t eb_ = (__float_as_uint(scale_slot[3 * (job & 1) + 1] * sc) >> 23) & 0xffu;
                if (eb_ >= 8u && eb_ <= 250u) { sb = sc * __uint_as_float((257u - eb_) << 23); inv_sb = __uint_as_float((eb_ - 3u) << 23); }
            }
            {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        const int k = 8 * bi + i, c0 = 8 * bj + j;
                        float v0 = -acc[i][j] * sc, v1 = -acc[i][j + 1] * sc;
                        if (k == c0) v0 -= k < d ? reg * sc : 1.f;     // padded diagonal = 1 keeps the factorisation defined
                        if (k == c0 + 1) v1 -= k < d ? reg * sc : 1.f;
                        *reinterpret_cast<float2 *>(S + tile_at(k, c0)) = make_float2(v0, v1);
                    }
                // the block right of the diagonal inside a diagonal 16 x 16 tile pair (rows 16 I .. +7, columns 16 I + 8 .. +15) is
                // stored too: the mirror image of block (2 I + 1, 2 I)
                if ((bi & 1) && bj == bi - 1) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int i = 0; i < 8; i += 2)
                            *reinterpret_cast<float2 *>(S + tile_at(8 * bj + j, 8 * bi + i)) = make_float2(-acc[i][j] * sc, -acc[i + 1][j] * sc);
                }
                if (has_diag) {                                        // row bj of diagonal block bi
                    const int k = 8 * bi + bj;
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        float v0 = -dacc[j] * sc, v1 = -dacc[j + 1] * sc;
                        if (bj == j) v0 -= k < d ? reg * sc : 1.f;
                        if (bj == j + 1) v1 -= k < d ? reg * sc : 1.f;
                        *reinterpret_cast<float2 *>(S + tile_at(k, 8 * bi + j)) = make_float2(v0, v1);
                    }
                }
            }
            S[TL::FLOATS + tid] = -bacc * sb;
            if (tid == 0) S[TL::FLOATS + DP] = inv_sb;              // first scratch word: the solver multiplies x by it
            acc_sync();
            if (tid == 0) mbar_arrive(bar_full(bf));
        }
    } else {
        // ============================================================ solver warps
        const int sv = warp - W_SOLVE0;
        for (int64_t job = sv; job < n_jobs; job += NSOLVE) {
            const int64_t row = row_begin + blockIdx.x + job * gridDim.x;
            const int bf = (int)(job % NBUF);
            float *S = bufs + (size_t)bf * BUF_FLOATS;
            float *bz = S + TL::FLOATS, *scratch = bz + DP, *xs = scratch + DP;
            mbar_wait(bar_full(bf), (uint32_t)((job / NBUF) & 1));
            float x[MT / 2];
            const float inv_sb = scratch[0];
            cm::solve_tiles_h<MT, true>(S, bz, scratch, xs, lane, x);      // rolled row-tile loops: the solver must fit the 32 KB instruction cache
#pragma unroll
            for (int q = 0; q < MT / 2; ++q) x[q] *= inv_sb;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty(bf));
            bool bad = false;
#pragma unroll
            for (int q = 0; q < MT / 2; ++q) {
                const int c = lane + 32 * q;
                if (c < d) {
                    T[row * d + c] = x[q];
                    bad = bad || !(fabsf(x[q]) <= 3.0e38f);
                }
            }
            if (__any_sync(0xffffffffu, bad) && lane == 0) atomicAdd(info, 1);
        }
    }
}

}  // namespace wide

int als_wide_supported(const fms_ctx *ctx, int rank) {
    static const bool off = getenv("FMS_ALS_NO_WIDE") != nullptr;            // experiments: fall back to the generic kernel
    return !off && rank > 64 && rank <= wide::DP && rank % 4 == 0 && ctx->smem_optin >= wide::SMEM_BYTES;
}

int als_launch_half_step_wide(fms_ctx *ctx, int which, float lambda, int64_t rb, int64_t re) {
    AlsState &A = ctx->als;
    const int64_t rows = re - rb;
    if (rows <= 0) return FMS_OK;
    FMS_CUDA(ctx, cudaFuncSetAttribute(wide::k_als_rows_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide::SMEM_BYTES));
    const unsigned grid = (unsigned)(rows < ctx->sm_count ? rows : ctx->sm_count);
    wide::k_als_rows_wide<<<grid, wide::THREADS, wide::SMEM_BYTES, ctx->stream>>>(A.ptr[which].p, A.col[which].p, A.val[which].p,
                                                                                  A.fac[1 - which].p, A.fac[which].p, A.rank, lambda, rb,
                                                                                  re, A.info.p);
    FMS_LAUNCH_CHECK(ctx);
    return FMS_OK;
}

}  // namespace fms
