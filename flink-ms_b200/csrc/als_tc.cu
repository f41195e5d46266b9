// als_tc.cu -- K1 on the 5th-generation tensor cores (rank 64): warp-specialised persistent kernel
//
//   loader warps   : gather partner rows (fp32, LDG.128), split each value into fp16 hi + fp16 lo*2^11, write both
//                    planes into a shared-memory ring in the UMMA MN-major / 128-byte-swizzle layout; per-stage
//                    partial of b = sum r*y
//   MMA warp       : one elected lane issues tcgen05.mma (kind::f16, M=64, N=64, K=16): D1 += hi hi^T,
//                    D2 += hi lo^T + lo hi^T into TMEM; the Gramian is D1 + 2^-11 D2 (error-compensated split,
//                    ~2^-22 relative: fp32-grade, needed for the 1e-4 parity target; plain fp16/tf32 is 2^-11);
//                    lanes 0-15 fold the per-stage b partials in stage order (deterministic)
//   epilogue warps : tcgen05.ld the accumulators (one TMEM lane quadrant per warp), add lambda*n_u on the
//                    diagonal, hand the matrix to a solver warp through shared memory
//   solver warps   : one warp per system, two matrix rows per lane in registers, right-looking Cholesky with the
//                    pivot column broadcast through shared memory, forward substitution fused (augmented column),
//                    butterfly back substitution; writes x_u
//
// Takes over the FlinkML `updateFactors` coGroup body (entered from
// flink-als/src/main/scala/de/tub/it4bi/ALSImpl.scala:52), same math as k_als_rows_generic (als_kernels.cu).
// Algorithmic bytes per half-step: nnz*(4 idx + 4 val + 4*64 row) + N*(4 ptr + 4*64 out)  (SURVEY.md 8d).
#include <cuda_fp16.h>

#include "common.cuh"

namespace fms {
namespace tc {

constexpr int D = 64;
constexpr int KROWS = 16;                         // gathered rows per stage == K of one tcgen05.mma (fp16)
constexpr int TILE_BYTES = KROWS * 128;           // one plane (hi or lo) of one stage, 2 swizzle atoms
constexpr int STAGE_BYTES = 2 * TILE_BYTES;
constexpr int NSTG = 16;                          // ring depth (64 KB)
constexpr int NACC = 4;                           // TMEM accumulator slots (128 columns each: D1 | D2)
constexpr int NGB = 4;                            // Gramian hand-off buffers == solver warps
constexpr int GLD = 68;                           // padded row pitch of a hand-off buffer (floats)
constexpr int N_EPI = 4, N_SOLVE = 4, N_LOAD = 3; // warps per role (+1 MMA warp)
constexpr int W_EPI0 = 0, W_SOLVE0 = 4, W_MMA = 8, W_LOAD0 = 9;
constexpr int NWARPS = 12;
constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;

struct Smem {
    // offsets in bytes from the 1024-aligned base
    static constexpr int ring = 0;
    static constexpr int bpart = ring + NSTG * STAGE_BYTES;                 // NSTG x 64 floats
    static constexpr int gbuf = bpart + NSTG * D * 4;                       // NGB x 64 x GLD floats
    static constexpr int bacc = gbuf + NGB * D * GLD * 4;                   // NACC x 64 floats (+ n)
    static constexpr int bsol = bacc + NACC * (D + 4) * 4;                  // NGB x (64 floats + n + row)
    static constexpr int colb = bsol + NGB * (D + 4) * 4;                   // NGB x 68 floats (pivot column)
    static constexpr int bars = colb + NGB * 72 * 4;                        // mbarriers
    static constexpr int n_bars = 2 * NSTG + 2 * NACC + 2 * NGB;
    static constexpr int tmem_ptr = bars + n_bars * 8;
    static constexpr int total = tmem_ptr + 16;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread i of the warp reads TMEM lane (quadrant base + i), columns [c, c+32)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor: MN-major, SWIZZLE_128B, 16 K-rows of 128 B (two 8-row atoms, SBO = 1024 B).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);            // start address
    d |= (uint64_t)((TILE_BYTES >> 4) & 0x3FFF) << 16;     // leading byte offset (next 64-wide MN chunk; unused: one chunk)
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;           // stride byte offset: next group of 8 K-rows
    d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
    return d;
}
// instruction descriptor: D=f32, A=B=f16, both MN-major, N=64, M=64
constexpr uint32_t IDESC = (1u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((64u >> 4) << 24);

__device__ __forceinline__ float rsqrt_refined(float x) {
    float r = rsqrtf(x);
    return r * fmaf(-0.5f * x * r, r, 1.5f);              // one Newton step: ~1 ulp
}

template <bool DUMP>
__global__ void __launch_bounds__(NWARPS * 32, 1) k_als_rows_tc64(
    const int64_t *__restrict__ ptr, const int32_t *__restrict__ col, const float *__restrict__ val,
    const float *__restrict__ P, float *__restrict__ T, float lambda, int64_t row_begin, int64_t row_end,
    int *__restrict__ info, float *__restrict__ dump) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms need 1024-B alignment
    const uint32_t sm_base = smem_u32(sm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = sm_base + Smem::bars;
    auto bar_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return bar0 + 8u * (NSTG + s); };
    auto bar_acc_full = [&](int a) { return bar0 + 8u * (2 * NSTG + a); };
    auto bar_acc_empty = [&](int a) { return bar0 + 8u * (2 * NSTG + NACC + a); };
    auto bar_g_full = [&](int g) { return bar0 + 8u * (2 * NSTG + 2 * NACC + g); };
    auto bar_g_empty = [&](int g) { return bar0 + 8u * (2 * NSTG + 2 * NACC + NGB + g); };
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(sm + Smem::tmem_ptr);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTG; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        for (int a = 0; a < NACC; ++a) { mbar_init(bar_acc_full(a), 2); mbar_init(bar_acc_empty(a), N_EPI); }
        for (int g = 0; g < NGB; ++g) { mbar_init(bar_g_full(g), N_EPI); mbar_init(bar_g_empty(g), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sm_base + Smem::tmem_ptr) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t first = row_begin + blockIdx.x;
    const int64_t stride = gridDim.x;

    if (warp >= W_LOAD0) {
        // ================================================================== loaders
        const int lw = warp - W_LOAD0;
        const int half = lane >> 4, sub = lane & 15;
        const float4 *P4 = reinterpret_cast<const float4 *>(P);
        uint32_t stage_ctr = 0;                                   // global stage counter of this CTA
        for (int64_t row = first; row < row_end; row += stride) {
            const int64_t s0 = ptr[row], e0 = ptr[row + 1];
            const int nst = (int)((e0 - s0 + KROWS - 1) / KROWS);
            for (int st = 0; st < nst; ++st, ++stage_ctr) {
                if ((int)(stage_ctr % N_LOAD) != lw) continue;
                const int slot = stage_ctr % NSTG;
                const int64_t k0 = s0 + (int64_t)st * KROWS;
                float4 v[8];
                float rt[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int64_t k = k0 + 2 * i + half;
                    const bool ok = k < e0;
                    const int32_t c = ok ? col[k] : 0;
                    rt[i] = ok ? val[k] : 0.f;
                    v[i] = ok ? P4[(int64_t)c * (D / 4) + sub] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                mbar_wait(bar_empty(slot), ((stage_ctr / NSTG) & 1) ^ 1);
                unsigned char *hi_t = sm + Smem::ring + slot * STAGE_BYTES;
                unsigned char *lo_t = hi_t + TILE_BYTES;
                float4 bp = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = 2 * i + half;
                    bp.x = fmaf(rt[i], v[i].x, bp.x); bp.y = fmaf(rt[i], v[i].y, bp.y);
                    bp.z = fmaf(rt[i], v[i].z, bp.z); bp.w = fmaf(rt[i], v[i].w, bp.w);
                    const __half2 h01 = __floats2half2_rn(v[i].x, v[i].y), h23 = __floats2half2_rn(v[i].z, v[i].w);
                    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                    const __half2 l01 = __floats2half2_rn((v[i].x - f01.x) * LO_SCALE, (v[i].y - f01.y) * LO_SCALE);
                    const __half2 l23 = __floats2half2_rn((v[i].z - f23.x) * LO_SCALE, (v[i].w - f23.y) * LO_SCALE);
                    // 128-byte swizzle: 16-byte chunk index ^= row & 7 (tiles are 1024-byte aligned)
                    const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)sub >> 1) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)sub & 1u) * 8u;
                    uint2 hv, lv;
                    hv.x = *reinterpret_cast<const uint32_t *>(&h01); hv.y = *reinterpret_cast<const uint32_t *>(&h23);
                    lv.x = *reinterpret_cast<const uint32_t *>(&l01); lv.y = *reinterpret_cast<const uint32_t *>(&l23);
                    *reinterpret_cast<uint2 *>(hi_t + off) = hv;
                    *reinterpret_cast<uint2 *>(lo_t + off) = lv;
                }
                bp.x += __shfl_xor_sync(0xffffffffu, bp.x, 16); bp.y += __shfl_xor_sync(0xffffffffu, bp.y, 16);
                bp.z += __shfl_xor_sync(0xffffffffu, bp.z, 16); bp.w += __shfl_xor_sync(0xffffffffu, bp.w, 16);
                if (half == 0) reinterpret_cast<float4 *>(sm + Smem::bpart + slot * D * 4)[sub] = bp;
                fence_proxy_async();                              // generic-proxy stores -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full(slot));
            }
        }
    } else if (warp == W_MMA) {
        // ================================================================== MMA issuer (+ b reduction on lanes 0-15)
        uint32_t stage_ctr = 0, job = 0;
        for (int64_t row = first; row < row_end; row += stride, ++job) {
            const int64_t s0 = ptr[row], e0 = ptr[row + 1];
            const int nst = (int)((e0 - s0 + KROWS - 1) / KROWS);
            const int acc = job % NACC;
            mbar_wait(bar_acc_empty(acc), ((job / NACC) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d1 = tmem_base + (uint32_t)acc * 128u, d2 = d1 + 64u;
            float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int st = 0; st < nst; ++st, ++stage_ctr) {
                const int slot = stage_ctr % NSTG;
                mbar_wait(bar_full(slot), (stage_ctr / NSTG) & 1);
                tc_fence_after();
                if (lane < 16) {
                    const float4 p = reinterpret_cast<const float4 *>(sm + Smem::bpart + slot * D * 4)[lane];
                    bsum.x += p.x; bsum.y += p.y; bsum.z += p.z; bsum.w += p.w;
                }
                __syncwarp();
                if (lane == 0) {
                    const uint32_t hi_a = sm_base + Smem::ring + slot * STAGE_BYTES;
                    const uint64_t dh = make_desc(hi_a), dl = make_desc(hi_a + TILE_BYTES);
                    tc_mma_f16(d1, dh, dh, IDESC, st > 0 ? 1u : 0u);      // D1 (+)= hi hi^T
                    tc_mma_f16(d2, dh, dl, IDESC, st > 0 ? 1u : 0u);      // D2 (+)= hi lo^T
                    tc_mma_f16(d2, dl, dh, IDESC, 1u);                    // D2  += lo hi^T
                    tc_commit(bar_empty(slot));                           // ring slot free once these MMAs retire
                }
                __syncwarp();
            }
            float *bo = reinterpret_cast<float *>(sm + Smem::bacc + acc * (D + 4) * 4);
            if (lane < 16) reinterpret_cast<float4 *>(bo)[lane] = bsum;
            if (lane == 16) { bo[D] = (float)(e0 - s0); reinterpret_cast<int *>(bo)[D + 1] = (int)(row - row_begin); }
            __syncwarp();
            if (lane == 0) {
                tc_commit(bar_acc_full(acc));                             // arrives when every MMA of this row has retired
                mbar_arrive(bar_acc_full(acc));                           // releases the b / n stores above
            }
            __syncwarp();
        }
    } else if (warp < W_SOLVE0) {
        // ================================================================== epilogue: TMEM -> registers -> hand-off buffer
        const int q = warp - W_EPI0;                              // TMEM lane quadrant == warp index % 4
        uint32_t job = 0;
        for (int64_t row = first; row < row_end; row += stride, ++job) {
            const int acc = job % NACC, g = job % NGB;
            mbar_wait(bar_acc_full(acc), (job / NACC) & 1);
            tc_fence_after();
            mbar_wait(bar_g_empty(g), ((job / NGB) & 1) ^ 1);
            const float *bi = reinterpret_cast<const float *>(sm + Smem::bacc + acc * (D + 4) * 4);
            float *bo = reinterpret_cast<float *>(sm + Smem::bsol + g * (D + 4) * 4);
            float *G = reinterpret_cast<float *>(sm + Smem::gbuf + g * D * GLD * 4);
            const float reg = lambda * bi[D];
            const int my_row = 16 * q + lane;                     // valid for lane < 16
            const uint32_t tq = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)acc * 128u;
#pragma unroll
            for (int c0 = 0; c0 < D; c0 += 32) {
                float d1[32], d2[32];
                tc_ld32(tq + c0, d1);
                tc_ld32(tq + 64 + c0, d2);
                if (lane < 16) {
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        float4 o;
                        o.x = fmaf(d2[c + 0], LO_INV, d1[c + 0]); o.y = fmaf(d2[c + 1], LO_INV, d1[c + 1]);
                        o.z = fmaf(d2[c + 2], LO_INV, d1[c + 2]); o.w = fmaf(d2[c + 3], LO_INV, d1[c + 3]);
                        const int cc = c0 + c;
                        if (my_row == cc + 0) o.x += reg;
                        if (my_row == cc + 1) o.y += reg;
                        if (my_row == cc + 2) o.z += reg;
                        if (my_row == cc + 3) o.w += reg;
                        *reinterpret_cast<float4 *>(G + my_row * GLD + cc) = o;
                    }
                }
            }
            if (q == 0) {
                bo[lane] = bi[lane]; bo[lane + 32] = bi[lane + 32];
                if (lane < 2) bo[D + lane] = bi[D + lane];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar_acc_empty(acc)); mbar_arrive(bar_g_full(g)); }
        }
    } else {
        // ================================================================== solvers: one warp per system
        const int g = warp - W_SOLVE0;
        float *cb = reinterpret_cast<float *>(sm + Smem::colb + g * 72 * 4);
        uint32_t use = 0;
        uint32_t job = 0;
        for (int64_t row = first; row < row_end; row += stride, ++job) {
            if ((int)(job % NGB) != g) continue;
            mbar_wait(bar_g_full(g), use & 1);
            const float *G = reinterpret_cast<const float *>(sm + Smem::gbuf + g * D * GLD * 4);
            const float *bi = reinterpret_cast<const float *>(sm + Smem::bsol + g * (D + 4) * 4);
            float a0[D + 1], a1[D + 1];                           // rows `lane` and `lane + 32`, augmented with b
#pragma unroll
            for (int c = 0; c < D; c += 4) {
                const float4 x0 = *reinterpret_cast<const float4 *>(G + lane * GLD + c);
                const float4 x1 = *reinterpret_cast<const float4 *>(G + (lane + 32) * GLD + c);
                a0[c] = x0.x; a0[c + 1] = x0.y; a0[c + 2] = x0.z; a0[c + 3] = x0.w;
                a1[c] = x1.x; a1[c + 1] = x1.y; a1[c + 2] = x1.z; a1[c + 3] = x1.w;
            }
            a0[D] = bi[lane];
            a1[D] = bi[lane + 32];
            if (DUMP) {
                float *o = dump + (int64_t)(row - row_begin) * (D * D + D);
#pragma unroll
                for (int c = 0; c < D; ++c) { o[lane * D + c] = a0[c]; o[(lane + 32) * D + c] = a1[c]; }
                o[D * D + lane] = a0[D]; o[D * D + lane + 32] = a1[D];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_g_empty(g));           // every lane has its rows in registers
            ++use;
            bool bad = false;
            // ---- right-looking Cholesky, forward substitution fused as column D
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const float pj = (j < 32) ? a0[j] : a1[j];
                const float ajj = __shfl_sync(0xffffffffu, pj, j & 31);
                bad |= !(ajj > 0.f);
                const float inv = rsqrt_refined(ajj);
                a0[j] *= inv;                                     // L[l][j]   (row j itself: sqrt(ajj))
                a1[j] *= inv;                                     // L[l+32][j]
                const float bj = (j < 32) ? a0[D] : a1[D];
                const float zj = __shfl_sync(0xffffffffu, bj, j & 31) * inv;
                if ((j & 31) == lane) { if (j < 32) a0[D] = zj; else a1[D] = zj; }
                if (j + 1 < D) {
                    cb[lane] = a0[j];
                    cb[lane + 32] = a1[j];
                    __syncwarp();
                    // trailing update of my two rows: a[i][k] -= L[i][j] * L[k][j], k = j+1 .. 63
                    const float l0 = a0[j], l1 = a1[j];
#pragma unroll
                    for (int k4 = (j + 1) & ~3; k4 < D; k4 += 4) {
                        const float4 c4 = *reinterpret_cast<const float4 *>(cb + k4);
                        if (k4 + 0 > j) { a0[k4 + 0] = fmaf(-l0, c4.x, a0[k4 + 0]); a1[k4 + 0] = fmaf(-l1, c4.x, a1[k4 + 0]); }
                        if (k4 + 1 > j) { a0[k4 + 1] = fmaf(-l0, c4.y, a0[k4 + 1]); a1[k4 + 1] = fmaf(-l1, c4.y, a1[k4 + 1]); }
                        if (k4 + 2 > j) { a0[k4 + 2] = fmaf(-l0, c4.z, a0[k4 + 2]); a1[k4 + 2] = fmaf(-l1, c4.z, a1[k4 + 2]); }
                        if (k4 + 3 > j) { a0[k4 + 3] = fmaf(-l0, c4.w, a0[k4 + 3]); a1[k4 + 3] = fmaf(-l1, c4.w, a1[k4 + 3]); }
                    }
                    __syncwarp();
                }
                // b_i -= L[i][j] * z_j for the rows below j
                if (lane > j) a0[D] = fmaf(-a0[j], zj, a0[D]);
                if (lane + 32 > j) a1[D] = fmaf(-a1[j], zj, a1[D]);
            }
            // ---- back substitution L^T x = z
            float x0 = 0.f, x1 = 0.f;
#pragma unroll
            for (int j = D - 1; j >= 0; --j) {
                float part = 0.f;
                if (lane > j) part = a0[j] * x0;
                if (lane + 32 > j) part = fmaf(a1[j], x1, part);
                part = warp_sum(part);
                if ((j & 31) == lane) {
                    if (j < 32) x0 = (a0[D] - part) / a0[j];
                    else x1 = (a1[D] - part) / a1[j];
                }
            }
            T[row * D + lane] = x0;
            T[row * D + lane + 32] = x1;
            if (bad && lane == 0) atomicAdd(info, 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

}  // namespace tc

int als_tc_supported(const fms_ctx *ctx, int rank) {
    return rank == tc::D && ctx->smem_optin >= (size_t)tc::Smem::total + 1024;
}

int als_tc_refresh_planes(fms_ctx *, int) { return FMS_OK; }

static int launch_tc(fms_ctx *ctx, int which, float lambda, int64_t rb, int64_t re, float *dump) {
    AlsState &A = ctx->als;
    if (re <= rb) return FMS_OK;
    const size_t smem = (size_t)tc::Smem::total + 1024;
    const int64_t rows = re - rb;
    const unsigned grid = (unsigned)(rows < ctx->sm_count ? rows : ctx->sm_count);
    if (dump) {
        FMS_CUDA(ctx, cudaFuncSetAttribute(tc::k_als_rows_tc64<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc::k_als_rows_tc64<true><<<grid, tc::NWARPS * 32, smem, ctx->stream>>>(A.ptr[which].p, A.col[which].p, A.val[which].p,
                                                                                 A.fac[1 - which].p, A.fac[which].p, lambda, rb, re,
                                                                                 A.info.p, dump);
    } else {
        FMS_CUDA(ctx, cudaFuncSetAttribute(tc::k_als_rows_tc64<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc::k_als_rows_tc64<false><<<grid, tc::NWARPS * 32, smem, ctx->stream>>>(A.ptr[which].p, A.col[which].p, A.val[which].p,
                                                                                  A.fac[1 - which].p, A.fac[which].p, lambda, rb, re,
                                                                                  A.info.p, nullptr);
    }
    FMS_LAUNCH_CHECK(ctx);
    return FMS_OK;
}

int als_launch_half_step_tc(fms_ctx *ctx, int which, float lambda, int64_t rb, int64_t re) {
    return launch_tc(ctx, which, lambda, rb, re, nullptr);
}

// test hook: Gramian + rhs of rows [rb, re) exactly as the solver warps receive them (row-major 64x64 + 64 floats each)
int als_tc_dump(fms_ctx *ctx, int which, float lambda, int64_t rb, int64_t re, float *d_dump) {
    return launch_tc(ctx, which, lambda, rb, re, d_dump);
}

}  // namespace fms
