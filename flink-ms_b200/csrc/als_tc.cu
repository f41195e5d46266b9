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
constexpr int NGB = 8;                            // Gramian hand-off buffers: two per solver warp
constexpr int GLD = 68;                           // padded row pitch of a hand-off buffer (floats)
constexpr int N_EPI = 4, N_SOLVE = 4, N_LOAD = 7; // warps per role (+1 MMA warp)
constexpr int W_EPI0 = 0, W_SOLVE0 = 4, W_MMA = 8, W_LOAD0 = 9;
constexpr int NWARPS = 16;
constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;

struct Smem {
    // offsets in bytes from the 1024-aligned base
    static constexpr int ring = 0;
    static constexpr int bpart = ring + NSTG * STAGE_BYTES;                 // NSTG x 64 floats
    static constexpr int gbuf = bpart + NSTG * D * 4;                       // NGB x 64 x GLD floats
    static constexpr int bacc = gbuf + NGB * D * GLD * 4;                   // NACC x 64 floats (+ n)
    static constexpr int bsol = bacc + NACC * (D + 4) * 4;                  // NGB x (64 floats + n + row)
    static constexpr int colb = bsol + NGB * (D + 4) * 4;                   // N_SOLVE x 64 floats (1 / L_jj)
    static constexpr int lring = colb + N_SOLVE * D * 4;                    // N_LOAD x 4 x 32 words (index ring)
    static constexpr int bars = lring + N_LOAD * 4 * 32 * 4;                // mbarriers
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
// long waits (epilogue / solver hand-offs): back off so the spinning warp does not eat issue slots
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    for (;;) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        __nanosleep(100);
    }
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
        // Loader warp lw owns the CTA's stages lw, lw + N_LOAD, ...  Three-deep software pipeline per warp:
        //   index  (IA stages ahead): lanes 0-15 read col/val of the stage's 16 ratings, park them in a small
        //                              shared ring and prefetch the partner rows into L2
        //   gather (1 stage ahead)   : 8 x LDG.128 per lane into one of two register buffers
        //   convert (current)        : fp32 -> fp16 hi / lo planes in UMMA layout, b partial, signal the MMA warp
        const int lw = warp - W_LOAD0;
        const int half = lane >> 4, sub = lane & 15;
        const float4 *P4 = reinterpret_cast<const float4 *>(P);
        constexpr int IA = 3, IRING = 4;
        int *iring = reinterpret_cast<int *>(sm + Smem::lring + lw * IRING * 32 * 4);   // [IRING][16 idx | 16 rating bits]

        struct It { int64_t row, s0, e0; uint32_t base, nst; };
        auto it_init = [&](It &it) {
            it.row = first; it.base = 0; it.nst = 0; it.s0 = it.e0 = 0;
            if (it.row < row_end) { it.s0 = ptr[it.row]; it.e0 = ptr[it.row + 1]; it.nst = (uint32_t)((it.e0 - it.s0 + KROWS - 1) / KROWS); }
        };
        // position the iterator on global stage counter c; false when the CTA has no such stage
        auto it_seek = [&](It &it, uint32_t c) -> bool {
            while (it.row < row_end && c >= it.base + it.nst) {
                it.base += it.nst;
                it.row += stride;
                if (it.row < row_end) { it.s0 = ptr[it.row]; it.e0 = ptr[it.row + 1]; it.nst = (uint32_t)((it.e0 - it.s0 + KROWS - 1) / KROWS); }
            }
            return it.row < row_end;
        };
        It it_i, it_g;
        it_init(it_i); it_init(it_g);

        auto do_index = [&](uint32_t n) {
            const uint32_t c = (uint32_t)lw + n * N_LOAD;
            if (!it_seek(it_i, c)) return;
            const int64_t k = it_i.s0 + (int64_t)(c - it_i.base) * KROWS + sub;
            if (half == 0) {
                const bool ok = k < it_i.e0;
                const int32_t ci = ok ? col[k] : -1;
                const float rv = ok ? val[k] : 0.f;
                int *slot = iring + (n % IRING) * 32;
                const int pos = (sub & 1) * 8 + (sub >> 1);       // rows of one half-warp become contiguous
                slot[pos] = ci;
                slot[16 + pos] = __float_as_int(rv);
                if (ok) {
                    const char *rp = reinterpret_cast<const char *>(P + (int64_t)ci * D);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(rp));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + 128));
                }
            }
        };
        auto do_gather = [&](uint32_t n, float4 (&v)[8], float (&rt)[8]) -> bool {
            const uint32_t c = (uint32_t)lw + n * N_LOAD;
            if (!it_seek(it_g, c)) return false;
            __syncwarp();                                          // ring entries written by lanes 0-15 of this warp
            const int *slot = iring + (n % IRING) * 32;
            const int4 i0 = *reinterpret_cast<const int4 *>(slot + half * 8), i1 = *reinterpret_cast<const int4 *>(slot + half * 8 + 4);
            const int4 r0 = *reinterpret_cast<const int4 *>(slot + 16 + half * 8), r1 = *reinterpret_cast<const int4 *>(slot + 16 + half * 8 + 4);
            const int ci[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
            const int ri[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                rt[i] = __int_as_float(ri[i]);
                v[i] = ci[i] >= 0 ? P4[(int64_t)ci[i] * (D / 4) + sub] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            return true;
        };
        auto do_convert = [&](uint32_t n, const float4 (&v)[8], const float (&rt)[8]) {
            const uint32_t c = (uint32_t)lw + n * N_LOAD;
            const int slot = c % NSTG;
            mbar_wait(bar_empty(slot), ((c / NSTG) & 1) ^ 1);
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
            fence_proxy_async();                                  // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full(slot));
        };

        float4 va[8], vb[8];
        float ra[8], rb[8];
        for (uint32_t n = 0; n < IA; ++n) do_index(n);
        bool have_a = do_gather(0, va, ra), have_b = false;
        for (uint32_t n = 0; have_a; n += 2) {
            do_index(n + IA);
            have_b = do_gather(n + 1, vb, rb);
            do_convert(n, va, ra);
            if (!have_b) break;
            do_index(n + 1 + IA);
            have_a = do_gather(n + 2, va, ra);
            do_convert(n + 1, vb, rb);
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
        // Lane l owns matrix rows l and l+32 in registers.  Right-looking Cholesky in 16 rank-4 panel steps written as
        // two LOOPS (panels over columns 0..31, then 32..63) so the code stays instruction-cache resident: the
        // register window is rotated by 4 columns per panel (folded into the last FFMA2 of the panel, no moves).
        // The finished L columns go back into the hand-off buffer (column-major, pitch GLD) where the trailing
        // updates (broadcast LDS.128 + packed FFMA2) and the back substitution read them.  Forward substitution is
        // fused (b is carried as an extra column).
        const int sv = warp - W_SOLVE0;
        float *invd = reinterpret_cast<float *>(sm + Smem::colb + sv * D * 4);
        uint32_t job = 0;
        for (int64_t row = first; row < row_end; row += stride, ++job) {
            if ((int)(job % N_SOLVE) != sv) continue;
            const int g = job % NGB;
            mbar_wait_sleep(bar_g_full(g), (job / NGB) & 1);
            float *G = reinterpret_cast<float *>(sm + Smem::gbuf + g * D * GLD * 4);
            const float *bi = reinterpret_cast<const float *>(sm + Smem::bsol + g * (D + 4) * 4);
            float2 A0[16], A1a[16], A1b[16];                      // row l cols 0..31 | row l+32 cols 0..31 | row l+32 cols 32..63
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 x0 = *reinterpret_cast<const float4 *>(G + lane * GLD + 4 * c);
                const float4 x1 = *reinterpret_cast<const float4 *>(G + (lane + 32) * GLD + 4 * c);
                const float4 x2 = *reinterpret_cast<const float4 *>(G + (lane + 32) * GLD + 32 + 4 * c);
                A0[2 * c] = make_float2(x0.x, x0.y); A0[2 * c + 1] = make_float2(x0.z, x0.w);
                A1a[2 * c] = make_float2(x1.x, x1.y); A1a[2 * c + 1] = make_float2(x1.z, x1.w);
                A1b[2 * c] = make_float2(x2.x, x2.y); A1b[2 * c + 1] = make_float2(x2.z, x2.w);
            }
            float b0 = bi[lane], b1 = bi[lane + 32];
            if (DUMP) {
                float *o = dump + (int64_t)(row - row_begin) * (D * D + D);
                for (int c = 0; c < D; ++c) { o[lane * D + c] = G[lane * GLD + c]; o[(lane + 32) * D + c] = G[(lane + 32) * GLD + c]; }
                o[D * D + lane] = b0; o[D * D + lane + 32] = b1;
            }
            __syncwarp();                                         // every lane holds its rows: the buffer becomes L storage
            float *Ls = G;                                        // Ls[j * GLD + i] = L[i][j]
            bool bad = false;

            // ---------------- panels over columns 0..31 (pivot rows live in A0 / b0 of lanes j0..j0+3)
#pragma unroll 1
            for (int j0 = 0; j0 < 32; j0 += 4) {
                const float d00 = __shfl_sync(0xffffffffu, A0[0].x, j0);
                const float d10 = __shfl_sync(0xffffffffu, A0[0].x, j0 + 1), d11 = __shfl_sync(0xffffffffu, A0[0].y, j0 + 1);
                const float d20 = __shfl_sync(0xffffffffu, A0[0].x, j0 + 2), d21 = __shfl_sync(0xffffffffu, A0[0].y, j0 + 2);
                const float d22 = __shfl_sync(0xffffffffu, A0[1].x, j0 + 2);
                const float d30 = __shfl_sync(0xffffffffu, A0[0].x, j0 + 3), d31 = __shfl_sync(0xffffffffu, A0[0].y, j0 + 3);
                const float d32 = __shfl_sync(0xffffffffu, A0[1].x, j0 + 3), d33 = __shfl_sync(0xffffffffu, A0[1].y, j0 + 3);
                const float zb0 = __shfl_sync(0xffffffffu, b0, j0), zb1 = __shfl_sync(0xffffffffu, b0, j0 + 1);
                const float zb2 = __shfl_sync(0xffffffffu, b0, j0 + 2), zb3 = __shfl_sync(0xffffffffu, b0, j0 + 3);
                // 4x4 Cholesky of the diagonal block, redundantly in every lane
                bad |= !(d00 > 0.f);
                const float i0 = rsqrt_refined(d00);
                const float l10 = d10 * i0, l20 = d20 * i0, l30 = d30 * i0;
                const float t11 = fmaf(-l10, l10, d11);
                bad |= !(t11 > 0.f);
                const float i1 = rsqrt_refined(t11);
                const float l21 = fmaf(-l20, l10, d21) * i1, l31 = fmaf(-l30, l10, d31) * i1;
                const float t22 = fmaf(-l21, l21, fmaf(-l20, l20, d22));
                bad |= !(t22 > 0.f);
                const float i2 = rsqrt_refined(t22);
                const float l32 = fmaf(-l31, l21, fmaf(-l30, l20, d32)) * i2;
                const float t33 = fmaf(-l32, l32, fmaf(-l31, l31, fmaf(-l30, l30, d33)));
                bad |= !(t33 > 0.f);
                const float i3 = rsqrt_refined(t33);
                // forward substitution for the panel's right-hand side
                const float z0 = zb0 * i0;
                const float z1 = fmaf(-z0, l10, zb1) * i1;
                const float z2 = fmaf(-z1, l21, fmaf(-z0, l20, zb2)) * i2;
                const float z3 = fmaf(-z2, l32, fmaf(-z1, l31, fmaf(-z0, l30, zb3))) * i3;
                // my two rows of the panel: x = A[i][J] * Ld^-T
                const float p0 = A0[0].x * i0;
                const float p1 = fmaf(-p0, l10, A0[0].y) * i1;
                const float p2 = fmaf(-p1, l21, fmaf(-p0, l20, A0[1].x)) * i2;
                const float p3 = fmaf(-p2, l32, fmaf(-p1, l31, fmaf(-p0, l30, A0[1].y))) * i3;
                const float q0 = A1a[0].x * i0;
                const float q1 = fmaf(-q0, l10, A1a[0].y) * i1;
                const float q2 = fmaf(-q1, l21, fmaf(-q0, l20, A1a[1].x)) * i2;
                const float q3 = fmaf(-q2, l32, fmaf(-q1, l31, fmaf(-q0, l30, A1a[1].y))) * i3;
                // right-hand side: rows below the panel get b_i -= sum_c L[i][c] z_c; pivot rows keep z
                const float nb0 = fmaf(-p3, z3, fmaf(-p2, z2, fmaf(-p1, z1, fmaf(-p0, z0, b0))));
                b1 = fmaf(-q3, z3, fmaf(-q2, z2, fmaf(-q1, z1, fmaf(-q0, z0, b1))));
                const int rel = lane - j0;
                b0 = rel == 0 ? z0 : rel == 1 ? z1 : rel == 2 ? z2 : rel == 3 ? z3 : rel > 3 ? nb0 : b0;   // finished rows keep z
                // publish the four L columns (column-major) and 1 / L_jj
                float *Lc = Ls + j0 * GLD;
                Lc[lane] = p0; Lc[GLD + lane] = p1; Lc[2 * GLD + lane] = p2; Lc[3 * GLD + lane] = p3;
                Lc[lane + 32] = q0; Lc[GLD + lane + 32] = q1; Lc[2 * GLD + lane + 32] = q2; Lc[3 * GLD + lane + 32] = q3;
                if (lane == 0) *reinterpret_cast<float4 *>(invd + j0) = make_float4(i0, i1, i2, i3);
                __syncwarp();
                // trailing update; the window (positions 2..15 = columns j0+4..j0+31) shifts down by two float2
                const float pm[4] = {-p0, -p1, -p2, -p3}, qm[4] = {-q0, -q1, -q2, -q3};
                float2 N0[16], N1[16];
#pragma unroll
                for (int w = 0; w < 14; ++w) { N0[w] = A0[w + 2]; N1[w] = A1a[w + 2]; }
                N0[14] = N0[15] = N1[14] = N1[15] = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float2 pp = make_float2(pm[c], pm[c]), qq = make_float2(qm[c], qm[c]);
                    const float *col_c = Lc + c * GLD + j0 + 4;                       // L[j0+4 ..][j0+c]
#pragma unroll
                    for (int w4 = 0; w4 < 7; ++w4) {
                        const float4 lv = *reinterpret_cast<const float4 *>(col_c + 4 * w4);
                        const float2 la = make_float2(lv.x, lv.y), lb = make_float2(lv.z, lv.w);
                        N0[2 * w4] = __ffma2_rn(pp, la, N0[2 * w4]); N0[2 * w4 + 1] = __ffma2_rn(pp, lb, N0[2 * w4 + 1]);
                        N1[2 * w4] = __ffma2_rn(qq, la, N1[2 * w4]); N1[2 * w4 + 1] = __ffma2_rn(qq, lb, N1[2 * w4 + 1]);
                    }
                    const float *col_s = Lc + c * GLD + 32;                           // L[32 ..][j0+c]
#pragma unroll
                    for (int w4 = 0; w4 < 8; ++w4) {
                        const float4 lv = *reinterpret_cast<const float4 *>(col_s + 4 * w4);
                        A1b[2 * w4] = __ffma2_rn(qq, make_float2(lv.x, lv.y), A1b[2 * w4]);
                        A1b[2 * w4 + 1] = __ffma2_rn(qq, make_float2(lv.z, lv.w), A1b[2 * w4 + 1]);
                    }
                }
#pragma unroll
                for (int w = 0; w < 16; ++w) { A0[w] = N0[w]; A1a[w] = N1[w]; }
            }
            // ---------------- panels over columns 32..63 (pivot rows live in A1b / b1 of lanes j0-32 .. +3)
#pragma unroll 1
            for (int j0 = 32; j0 < 64; j0 += 4) {
                const int src = j0 - 32;
                const float d00 = __shfl_sync(0xffffffffu, A1b[0].x, src);
                const float d10 = __shfl_sync(0xffffffffu, A1b[0].x, src + 1), d11 = __shfl_sync(0xffffffffu, A1b[0].y, src + 1);
                const float d20 = __shfl_sync(0xffffffffu, A1b[0].x, src + 2), d21 = __shfl_sync(0xffffffffu, A1b[0].y, src + 2);
                const float d22 = __shfl_sync(0xffffffffu, A1b[1].x, src + 2);
                const float d30 = __shfl_sync(0xffffffffu, A1b[0].x, src + 3), d31 = __shfl_sync(0xffffffffu, A1b[0].y, src + 3);
                const float d32 = __shfl_sync(0xffffffffu, A1b[1].x, src + 3), d33 = __shfl_sync(0xffffffffu, A1b[1].y, src + 3);
                const float zb0 = __shfl_sync(0xffffffffu, b1, src), zb1 = __shfl_sync(0xffffffffu, b1, src + 1);
                const float zb2 = __shfl_sync(0xffffffffu, b1, src + 2), zb3 = __shfl_sync(0xffffffffu, b1, src + 3);
                bad |= !(d00 > 0.f);
                const float i0 = rsqrt_refined(d00);
                const float l10 = d10 * i0, l20 = d20 * i0, l30 = d30 * i0;
                const float t11 = fmaf(-l10, l10, d11);
                bad |= !(t11 > 0.f);
                const float i1 = rsqrt_refined(t11);
                const float l21 = fmaf(-l20, l10, d21) * i1, l31 = fmaf(-l30, l10, d31) * i1;
                const float t22 = fmaf(-l21, l21, fmaf(-l20, l20, d22));
                bad |= !(t22 > 0.f);
                const float i2 = rsqrt_refined(t22);
                const float l32 = fmaf(-l31, l21, fmaf(-l30, l20, d32)) * i2;
                const float t33 = fmaf(-l32, l32, fmaf(-l31, l31, fmaf(-l30, l30, d33)));
                bad |= !(t33 > 0.f);
                const float i3 = rsqrt_refined(t33);
                const float z0 = zb0 * i0;
                const float z1 = fmaf(-z0, l10, zb1) * i1;
                const float z2 = fmaf(-z1, l21, fmaf(-z0, l20, zb2)) * i2;
                const float z3 = fmaf(-z2, l32, fmaf(-z1, l31, fmaf(-z0, l30, zb3))) * i3;
                const float q0 = A1b[0].x * i0;
                const float q1 = fmaf(-q0, l10, A1b[0].y) * i1;
                const float q2 = fmaf(-q1, l21, fmaf(-q0, l20, A1b[1].x)) * i2;
                const float q3 = fmaf(-q2, l32, fmaf(-q1, l31, fmaf(-q0, l30, A1b[1].y))) * i3;
                const float nb1 = fmaf(-q3, z3, fmaf(-q2, z2, fmaf(-q1, z1, fmaf(-q0, z0, b1))));
                const int rel = lane - src;
                b1 = rel == 0 ? z0 : rel == 1 ? z1 : rel == 2 ? z2 : rel == 3 ? z3 : rel > 3 ? nb1 : b1;
                float *Lc = Ls + j0 * GLD;
                Lc[lane + 32] = q0; Lc[GLD + lane + 32] = q1; Lc[2 * GLD + lane + 32] = q2; Lc[3 * GLD + lane + 32] = q3;
                if (lane == 0) *reinterpret_cast<float4 *>(invd + j0) = make_float4(i0, i1, i2, i3);
                __syncwarp();
                const float qm[4] = {-q0, -q1, -q2, -q3};
                float2 N1[16];
#pragma unroll
                for (int w = 0; w < 14; ++w) N1[w] = A1b[w + 2];
                N1[14] = N1[15] = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float2 qq = make_float2(qm[c], qm[c]);
                    const float *col_c = Lc + c * GLD + j0 + 4;
#pragma unroll
                    for (int w4 = 0; w4 < 7; ++w4) {
                        // columns beyond 63 fall into the next L column / padding: finite garbage that is never used
                        const float4 lv = *reinterpret_cast<const float4 *>(col_c + 4 * w4);
                        N1[2 * w4] = __ffma2_rn(qq, make_float2(lv.x, lv.y), N1[2 * w4]);
                        N1[2 * w4 + 1] = __ffma2_rn(qq, make_float2(lv.z, lv.w), N1[2 * w4 + 1]);
                    }
                }
#pragma unroll
                for (int w = 0; w < 16; ++w) A1b[w] = N1[w];
            }
            __syncwarp();
            // ---------------- back substitution L^T x = z (column sweep): b0 / b1 hold z and turn into x
            float x0 = 0.f, x1 = 0.f;
#pragma unroll 1
            for (int j = D - 1; j >= 32; --j) {
                const float xj = __shfl_sync(0xffffffffu, b1, j - 32) * invd[j];
                if (lane == j - 32) x1 = xj;
                b0 = fmaf(-Ls[lane * GLD + j], xj, b0);                               // L[j][lane], lane < 32 <= j
                if (lane + 32 < j) b1 = fmaf(-Ls[(lane + 32) * GLD + j], xj, b1);
            }
#pragma unroll 1
            for (int j = 31; j >= 0; --j) {
                const float xj = __shfl_sync(0xffffffffu, b0, j) * invd[j];
                if (lane == j) x0 = xj;
                if (lane < j) b0 = fmaf(-Ls[lane * GLD + j], xj, b0);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_g_empty(g));           // L storage no longer needed
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
