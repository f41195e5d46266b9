// als_tc.cu -- K1 on the 5th-generation tensor cores (rank 64): warp-specialised persistent kernel, one CTA per SM.
//
// Data layout in HBM: besides the fp32 factor matrix every side keeps its fp16 "gather planes" hi = f16(x) and
// lo = f16((x - hi) * 2^11), interleaved as one 256-byte row per factor row (64 hi halves, then 64 lo halves).
// hi + lo * 2^-11 carries ~22 mantissa bits, so the Gramian  hi hi^T + 2^-11 (hi lo^T + lo hi^T)  is fp32-grade
// (plain fp16 / tf32 would be 2^-11: not enough for the 1e-4 parity target) while every product is exact in the fp32
// accumulator.  The ratings are pre-laid as 16-entry stages (rows padded with a zero plane row).
//
// 28 warps, register budgets set per role with setmaxnreg:
//   producers (6-7): stream the stage layout, issue TMA gather4 loads of the 16 partner rows (hi half and lo half of
//                    the plane row) straight into the UMMA MN-major / 128-byte-swizzle stage -- no registers, no
//                    per-rating SM instructions -- and write the ratings as two 16-byte column groups R, R' of a
//                    third operand chunk.  Stage = 6 KB, ring of 13, mbarrier full / empty.
//   MMA warp       : per stage two tcgen05.mma (kind::f16, M=64, K=16):
//                        D[:,   0:136] (+)= hi [hi | lo | R ]^T        (N = 136)
//                        D[:, 64:136]   += lo [hi |      R']^T        (N = 72, leading-dim offset skips `lo`)
//                    => cols 0-63 hi hi^T, 64-127 hi lo^T + lo hi^T, col 128 sum r_hi*hi, col 129 the 2^-11 terms of
//                    b = Y^T r.  Accumulators live in TMEM (2 slots x 256 columns).
//   epilogue (4)   : tcgen05.ld one TMEM lane quadrant each, G = D1 + 2^-11 D2 (+ lambda n on the diagonal) as a packed
//                    lower triangle, b, handed to a solver warp through one of 16 shared-memory buffers
//   solvers (16)   : one system per warp, ONE matrix row per lane at a time (32 data registers): 32x32 Cholesky of
//                    A11 in 8 rank-4 panels (pivot block by shuffle, factored redundantly per lane, packed FFMA2
//                    trailing updates against broadcast LDS.128, register window rotated inside the last update
//                    of a panel), private triangular solve per row for L21, syrk on the A22 row, the same 32x32
//                    loop for A22, fused forward substitution, two-half blocked back substitution; writes x (fp32)
//                    and its plane row -- into every rank's replica when peers are mapped (NVLink stores).
//
// Takes over the FlinkML `updateFactors` coGroup body (entered from
// flink-als/src/main/scala/de/tub/it4bi/ALSImpl.scala:52), same math as k_als_rows_generic (als_kernels.cu).
// Algorithmic bytes per half-step: nnz*(4 idx + 4 val + 4*64 row) + N*(4 ptr + 4*64 out)  (SURVEY.md 8d).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cub/cub.cuh>

#include "common.cuh"
#include "chol_mma.cuh"

namespace fms {
namespace tc {

constexpr int D = 64;
constexpr int KROWS = 16;                         // ratings per stage == K of one tcgen05.mma (fp16)
constexpr int TILE_BYTES = KROWS * 128;           // one 64-wide operand chunk of one stage (two swizzle atoms)
constexpr int STAGE_BYTES = 3 * TILE_BYTES;       // hi | lo | R (R' lives in the same chunk tile as R, one 16-byte column group over)
constexpr int NSTG = 13;                          // ring depth (78 KB, 52 KB of gathers in flight)
constexpr int NACC = 2;                           // TMEM accumulator slots
constexpr uint32_t ACC_COLS = 256u;               // TMEM columns per slot (136 used)
// MMA widths: the first covers hi | lo | 8 rating columns, the second hi | 8 rating columns.  Only two rating columns
// carry data; N = 136 / 72 instead of three / two full 64-wide chunks cuts a third of the tensor-pipe time per stage.
constexpr uint32_t MMA_N1 = 136u, MMA_N2 = 72u;
constexpr int N_EPI = 4, N_SOLVE = 16, N_PROD = 7;
constexpr int NGB = N_SOLVE;                      // one packed system / L buffer per solver warp
constexpr int W_EPI0 = 0, W_SOLVE0 = 4, W_MMA = 20, W_PROD0 = 21;
constexpr int NWARPS = 28;                        // 7 warpgroups: epilogue | 4 x solvers | MMA + producers x 2
constexpr int REGS_EPI = 88, REGS_SOLVE = 80, REGS_AUX = 40;    // setmaxnreg budgets; sum <= launch allocation (72 x 896)
// gather planes: one 256-byte row per factor row, hi halves then lo halves (one DRAM page burst serves both gathers)
constexpr int PLANE_PITCH = 2 * D;
constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;

// Packed lower-triangular column-major storage of a 64x64 system (and later of its Cholesky factor): column c holds
// rows 4*(c/4) .. 63, so every column start and every 4-row group inside it is 16-byte aligned.
constexpr int PK_FLOATS = 2176;
__host__ __device__ constexpr int pk_off(int c) {                  // offset of column c (its first stored row is 4*(c/4))
    return 256 * (c >> 2) - 8 * (c >> 2) * ((c >> 2) - 1) + (c & 3) * (64 - 4 * (c >> 2));
}
__host__ __device__ constexpr int pk_at(int i, int c) { return pk_off(c) + i - 4 * (c >> 2); }   // entry (i, c), i >= 4*(c/4)

struct Smem {   // byte offsets from the 1024-aligned base
    static constexpr int ring = 0;
    static constexpr int gbuf = ring + NSTG * STAGE_BYTES;                  // NGB x PK_FLOATS floats (+ slack for tail reads)
    static constexpr int bsol = gbuf + NGB * PK_FLOATS * 4 + 512;           // NGB x 64 floats (rhs)
    static constexpr int invd = bsol + NGB * D * 4;                         // N_SOLVE x 64 floats (1 / L_jj)
    static constexpr int bars = invd + N_SOLVE * D * 4;
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
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes instead of spinning
// (tight test_wait polling on the ring barriers measured slower: the polls compete with the barrier updates)
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
// non-blocking probe of a phase (about 80 cycles of latency that can sit under other work)
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMA gather4: four rows (row indices r0..r3, 64 halves each) of a 2-D fp16 tensor -> 512 contiguous swizzled bytes.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// Issues the hi-plane and lo-plane gathers of one row quartet from ONE elected lane (the operands must be valid in
// every lane: they come from warp shuffles), predicated instead of branched so the compiler emits no lane loop.
__device__ __forceinline__ void tma_gather4_pair(uint32_t dst_hi, uint32_t dst_lo, const CUtensorMap *map, uint32_t bar, int r0, int r1,
                                                 int r2, int r3) {
    asm volatile(
        "{\n"
        ".reg .pred pe;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "@pe cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%2, {%4, %6, %7, %8, %9}], [%3];\n"
        "@pe cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%1], [%2, {%5, %6, %7, %8, %9}], [%3];\n"
        "}\n" ::"r"(dst_hi),
        "r"(dst_lo), "l"(map), "r"(bar), "r"(0), "r"(D), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
}

// One stage worth of tensor-core work, issued by one elected lane without a divergent branch.
//   d        : TMEM address of the accumulator slot (column 0)
//   lo_*     : low words of the shared-memory descriptors (start address and leading-dim offset; the high word is
//              the same for all): hi chunk as A (also hi|lo|R as B: the same bits), lo chunk as A, hi|R' as B
__device__ __forceinline__ void tc_stage_mma(uint32_t d, uint32_t lo_hi, uint32_t lo_lo, uint32_t lo_b2, uint32_t idesc_n1, uint32_t idesc_n2,
                                             uint32_t accumulate, uint32_t bar_empty, uint32_t bar_acc, uint32_t last) {
    // descriptor high word 0x40004040: SBO 1024 B (bits 32-45), version 1 (bit 46), SWIZZLE_128B (bits 61-63)
    static_assert(((uint32_t)((1024 >> 4) & 0x3FFF) | (1u << 14) | (2u << 29)) == 0x40004040u, "descriptor high word");
    asm volatile(
        "{\n"
        ".reg .pred pe, pa, pl, pt;\n"
        ".reg .b32 d2;\n"
        ".reg .b64 ahi, alo, b2;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "setp.ne.b32 pa, %6, 0;\n"
        "setp.ne.b32 pt, %9, 0;\n"
        "and.pred pl, pt, pe;\n"
        "setp.eq.u32 pt, %0, %0;\n"
        "add.u32 d2, %0, 64;\n"
        "mov.b64 ahi, {%1, 0x40004040};\n"
        "mov.b64 alo, {%2, 0x40004040};\n"
        "mov.b64 b2, {%3, 0x40004040};\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], ahi, ahi, %4, pa;\n"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [d2], alo, b2, %5, pt;\n"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n"
        "@pl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n"
        "}\n" ::"r"(d),
        "r"(lo_hi), "r"(lo_lo), "r"(lo_b2), "r"(idesc_n1), "r"(idesc_n2), "r"(accumulate), "r"(bar_empty), "r"(bar_acc), "r"(last)
        : "memory");
}

// v2: one 16-rating sub-stage of a ring slot.  commit_empty: last sub-stage of the slot (releases it to the producers);
// last: last sub-stage of the row (hands the accumulator to the epilogue); n_mma: 2 (0 / 1 only in timing experiments).
__device__ __forceinline__ void tc_stage_mma2(uint32_t d, uint32_t lo_hi, uint32_t lo_lo, uint32_t lo_b2, uint32_t idesc_n1, uint32_t idesc_n2,
                                              uint32_t accumulate, uint32_t bar_empty, uint32_t bar_acc, uint32_t commit_empty, uint32_t last,
                                              uint32_t n_mma) {
    asm volatile(
        "{\n"
        ".reg .pred pe, pa, pl, pt, pc, p1, p2;\n"
        ".reg .b32 d2;\n"
        ".reg .b64 ahi, alo, b2;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "setp.ne.b32 pa, %6, 0;\n"
        "setp.ne.b32 pt, %10, 0;\n"
        "and.pred pl, pt, pe;\n"
        "setp.ne.b32 pt, %9, 0;\n"
        "and.pred pc, pt, pe;\n"
        "setp.ge.u32 pt, %11, 1;\n"
        "and.pred p1, pt, pe;\n"
        "setp.ge.u32 pt, %11, 2;\n"
        "and.pred p2, pt, pe;\n"
        "setp.eq.u32 pt, %0, %0;\n"
        "add.u32 d2, %0, 64;\n"
        "mov.b64 ahi, {%1, 0x40004040};\n"
        "mov.b64 alo, {%2, 0x40004040};\n"
        "mov.b64 b2, {%3, 0x40004040};\n"
        "@p1 tcgen05.mma.cta_group::1.kind::f16 [%0], ahi, ahi, %4, pa;\n"
        "@p2 tcgen05.mma.cta_group::1.kind::f16 [d2], alo, b2, %5, pt;\n"
        "@pc tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n"
        "@pl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n"
        "}\n" ::"r"(d),
        "r"(lo_hi), "r"(lo_lo), "r"(lo_b2), "r"(idesc_n1), "r"(idesc_n2), "r"(accumulate), "r"(bar_empty), "r"(bar_acc), "r"(commit_empty),
        "r"(last), "r"(n_mma)
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
__device__ __forceinline__ void tc_ld2(uint32_t taddr, float &a, float &b) {
    uint32_t r0, r1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    a = __uint_as_float(r0);
    b = __uint_as_float(r1);
}

// UMMA shared-memory descriptor: MN-major, SWIZZLE_128B; K = 16 rows = two 8-row atoms (stride 1024 B); successive
// 64-wide MN chunks are `lbo_bytes` apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
    return d;
}
// instruction descriptor: D = f32, A = B = f16, both MN-major, M = 64
__host__ __device__ constexpr uint32_t make_idesc(uint32_t n) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((64u >> 4) << 24);
}

__device__ __forceinline__ float rsqrt_refined(float x) {
    float r = rsqrtf(x);
    return r * fmaf(-0.5f * x * r, r, 1.5f);              // one Newton step: ~1 ulp
}
__device__ __forceinline__ void split_half(float x, __half &hi, __half &lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn((x - __half2float(hi)) * LO_SCALE);
}

// 4x4 Cholesky of the pivot block + forward substitution of the panel's rhs (redundantly in every lane)
struct Panel {
    float i0, i1, i2, i3, l10, l20, l30, l21, l31, l32, z0, z1, z2, z3;
};
// A non-positive pivot turns into NaN (rsqrt of a negative number, or 0 * inf in the refinement) and poisons the
// solution, which is how breakdown is detected afterwards (LAPACK's info != 0; FlinkML ignores it, we count it).
__device__ __forceinline__ Panel panel_factor(float d00, float d10, float d11, float d20, float d21, float d22, float d30, float d31,
                                              float d32, float d33, float zb0, float zb1, float zb2, float zb3) {
    Panel p;
    p.i0 = rsqrt_refined(d00);
    p.l10 = d10 * p.i0; p.l20 = d20 * p.i0; p.l30 = d30 * p.i0;
    const float t11 = fmaf(-p.l10, p.l10, d11);
    p.i1 = rsqrt_refined(t11);
    p.l21 = fmaf(-p.l20, p.l10, d21) * p.i1; p.l31 = fmaf(-p.l30, p.l10, d31) * p.i1;
    const float t22 = fmaf(-p.l21, p.l21, fmaf(-p.l20, p.l20, d22));
    p.i2 = rsqrt_refined(t22);
    p.l32 = fmaf(-p.l31, p.l21, fmaf(-p.l30, p.l20, d32)) * p.i2;
    const float t33 = fmaf(-p.l32, p.l32, fmaf(-p.l31, p.l31, fmaf(-p.l30, p.l30, d33)));
    p.i3 = rsqrt_refined(t33);
    p.z0 = zb0 * p.i0;
    p.z1 = fmaf(-p.z0, p.l10, zb1) * p.i1;
    p.z2 = fmaf(-p.z1, p.l21, fmaf(-p.z0, p.l20, zb2)) * p.i2;
    p.z3 = fmaf(-p.z2, p.l32, fmaf(-p.z1, p.l31, fmaf(-p.z0, p.l30, zb3))) * p.i3;
    return p;
}
// one matrix row against the factored pivot block: x = a * Ld^-T
__device__ __forceinline__ void panel_row(const Panel &p, float a0, float a1, float a2, float a3, float &x0, float &x1, float &x2, float &x3) {
    x0 = a0 * p.i0;
    x1 = fmaf(-x0, p.l10, a1) * p.i1;
    x2 = fmaf(-x1, p.l21, fmaf(-x0, p.l20, a2)) * p.i2;
    x3 = fmaf(-x2, p.l32, fmaf(-x1, p.l31, fmaf(-x0, p.l30, a3))) * p.i3;
}

__device__ __forceinline__ float fsel(int pred, float a, float b) {
    float r;
    asm("{\n.reg .pred p;\nsetp.ne.s32 p, %3, 0;\nselp.f32 %0, %1, %2, p;\n}\n" : "=f"(r) : "f"(a), "f"(b), "r"(pred));
    return r;
}
// ---- solver building blocks (one matrix row per lane) -----------------------------------------------------------
// Cholesky of one 32x32 diagonal block: lane l holds row (off + l) of the block, columns off .. off+31, in A[16]
// (float2, rotating window: positions 0..1 are the current panel) and its rhs entry in b.  Rank-4 panels; the 4x4
// pivot block is factored redundantly by every lane, its columns are published to the packed L storage and the
// trailing window is updated with packed FFMA2 from broadcast LDS.128.  Forward substitution is fused (b -> z).
__device__ __forceinline__ void chol32(float2 (&A)[16], float &b, float *Ls, float *invd, const int off, const int lane) {
#pragma unroll 1
    for (int jr = 0; jr < 32; jr += 4) {
        const int j0 = off + jr;
        const float d00 = __shfl_sync(0xffffffffu, A[0].x, jr);
        const float d10 = __shfl_sync(0xffffffffu, A[0].x, jr + 1), d11 = __shfl_sync(0xffffffffu, A[0].y, jr + 1);
        const float d20 = __shfl_sync(0xffffffffu, A[0].x, jr + 2), d21 = __shfl_sync(0xffffffffu, A[0].y, jr + 2);
        const float d22 = __shfl_sync(0xffffffffu, A[1].x, jr + 2);
        const float d30 = __shfl_sync(0xffffffffu, A[0].x, jr + 3), d31 = __shfl_sync(0xffffffffu, A[0].y, jr + 3);
        const float d32 = __shfl_sync(0xffffffffu, A[1].x, jr + 3), d33 = __shfl_sync(0xffffffffu, A[1].y, jr + 3);
        const float zb0 = __shfl_sync(0xffffffffu, b, jr), zb1 = __shfl_sync(0xffffffffu, b, jr + 1);
        const float zb2 = __shfl_sync(0xffffffffu, b, jr + 2), zb3 = __shfl_sync(0xffffffffu, b, jr + 3);
        const Panel P = panel_factor(d00, d10, d11, d20, d21, d22, d30, d31, d32, d33, zb0, zb1, zb2, zb3);
        float q0, q1, q2, q3;
        panel_row(P, A[0].x, A[0].y, A[1].x, A[1].y, q0, q1, q2, q3);
        const float nb = fmaf(-q3, P.z3, fmaf(-q2, P.z2, fmaf(-q1, P.z1, fmaf(-q0, P.z0, b))));
        const int rel = lane - jr;
        const float zlo = fsel(rel & 1, P.z1, P.z0), zhi = fsel(rel & 1, P.z3, P.z2);
        b = fsel(rel < 0, b, fsel(rel < 4, fsel(rel & 2, zhi, zlo), nb));                  // finished rows keep z; branch-free
        const int gq = j0 >> 2, len = 64 - j0;
        float *Lc = Ls + 256 * gq - 8 * gq * (gq - 1) - j0;                               // column j0, indexed by row
        if (rel >= 0) { Lc[off + lane] = q0; Lc[len + off + lane] = q1; Lc[2 * len + off + lane] = q2; Lc[3 * len + off + lane] = q3; }
        if (lane == 0) *reinterpret_cast<float4 *>(invd + j0) = make_float4(P.i0, P.i1, P.i2, P.i3);
        __syncwarp();
        const float qm[4] = {-q0, -q1, -q2, -q3};
        // window rows beyond the block read other columns / stale data: garbage that only lands in unused entries
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float2 qq = make_float2(qm[c], qm[c]);
            const float *col_c = Lc + c * len + j0 + 4;
#pragma unroll
            for (int w4 = 0; w4 < 7; ++w4) {
                const float4 lv = *reinterpret_cast<const float4 *>(col_c + 4 * w4);
                A[2 + 2 * w4] = __ffma2_rn(qq, make_float2(lv.x, lv.y), A[2 + 2 * w4]);
                A[3 + 2 * w4] = __ffma2_rn(qq, make_float2(lv.z, lv.w), A[3 + 2 * w4]);
            }
        }
        {
            const float2 qq = make_float2(qm[3], qm[3]);
            const float *col_c = Lc + 3 * len + j0 + 4;
#pragma unroll
            for (int w4 = 0; w4 < 7; ++w4) {                     // the fourth update writes two positions down: window rotation
                const float4 lv = *reinterpret_cast<const float4 *>(col_c + 4 * w4);
                A[2 * w4] = __ffma2_rn(qq, make_float2(lv.x, lv.y), A[2 + 2 * w4]);
                A[2 * w4 + 1] = __ffma2_rn(qq, make_float2(lv.z, lv.w), A[3 + 2 * w4]);
            }
        }
    }
}

// Back substitution with one 32x32 diagonal block of L^T (column sweep, four columns at a time): on entry z holds
// the block's rhs entry of this lane, on exit x.  Rows of the block are off + lane.
__device__ __forceinline__ void backsub32(float &z, const float *Ls, const float *invd, const int off, const int lane) {
    float x = 0.f;
    const int i = off + lane, gi = i >> 2;
    const float *coli = Ls + 256 * gi - 8 * gi * (gi - 1) + (i & 3) * (64 - 4 * gi) - 4 * gi;   // column i, indexed by row
#pragma unroll 1
    for (int jr = 28; jr >= 0; jr -= 4) {
        const int j0 = off + jr;
        const float z0 = __shfl_sync(0xffffffffu, z, jr), z1 = __shfl_sync(0xffffffffu, z, jr + 1);
        const float z2 = __shfl_sync(0xffffffffu, z, jr + 2), z3 = __shfl_sync(0xffffffffu, z, jr + 3);
        const int gq = j0 >> 2, len = 64 - j0;
        const float *Lc = Ls + 256 * gq - 8 * gq * (gq - 1);                             // column j0 starts at row j0
        const float4 c0 = *reinterpret_cast<const float4 *>(Lc);                         // L[j0..j0+3][j0]
        const float4 c1 = *reinterpret_cast<const float4 *>(Lc + len);
        const float4 c2 = *reinterpret_cast<const float4 *>(Lc + 2 * len);
        const float4 iv = *reinterpret_cast<const float4 *>(invd + j0);
        const float v3 = z3 * iv.w;
        const float v2 = fmaf(-c2.w, v3, z2) * iv.z;
        const float v1 = fmaf(-c1.w, v3, fmaf(-c1.z, v2, z1)) * iv.y;
        const float v0 = fmaf(-c0.w, v3, fmaf(-c0.z, v2, fmaf(-c0.y, v1, z0))) * iv.x;
        const int rel = lane - jr;
        if (rel >= 0 && rel < 4) x = rel == 0 ? v0 : rel == 1 ? v1 : rel == 2 ? v2 : v3;
        if (rel < 0) {                                           // rows above the panel: z_i -= sum_c L[j0+c][i] x_c
            const float4 lr = *reinterpret_cast<const float4 *>(coli + j0);
            z = fmaf(-lr.w, v3, fmaf(-lr.z, v2, fmaf(-lr.y, v1, fmaf(-lr.x, v0, z))));
        }
    }
    z = x;
}

// L21 = A21 L11^-T for row 32+lane (a private triangular solve, no cross-lane traffic) with the rhs update, publication
// of the row, then this lane's A22 row minus L21 L21^T.  On exit A holds row 32+lane of the updated A22.
__device__ __forceinline__ void solver_trsm_syrk(float2 (&A)[16], float &b1, float *Ls, const float *invd, const float *bx, const int lane) {
    // ---------------- trsm: row 32+l of L21 = A21 L11^-T, and b1 -= L21[row] . z1
#pragma unroll
    for (int w = 0; w < 16; ++w) {
        const int c = 2 * w;
        A[w].x = Ls[pk_off(c) - 4 * (c >> 2) + lane + 32];
        A[w].y = Ls[pk_off(c + 1) - 4 * ((c + 1) >> 2) + lane + 32];
    }
#pragma unroll
    for (int m = 0; m < 32; ++m) {
        float4 iv4, z4;
        if ((m & 3) == 0) { }
        iv4 = *reinterpret_cast<const float4 *>(invd + (m & ~3));
        z4 = *reinterpret_cast<const float4 *>(bx + (m & ~3));
        const float ivm = (m & 3) == 0 ? iv4.x : (m & 3) == 1 ? iv4.y : (m & 3) == 2 ? iv4.z : iv4.w;
        const float zm = (m & 3) == 0 ? z4.x : (m & 3) == 1 ? z4.y : (m & 3) == 2 ? z4.z : z4.w;
        float &am = (m & 1) ? A[m >> 1].y : A[m >> 1].x;
        const float xm = am * ivm;
        am = xm;
        b1 = fmaf(-xm, zm, b1);
        const float2 xx = make_float2(-xm, -xm);
        const float *colm = Ls + pk_off(m) - 4 * (m >> 2);                      // column m of L11, indexed by row
#pragma unroll
        for (int k4 = (m + 1) & ~3; k4 < 32; k4 += 4) {
            const float4 lv = *reinterpret_cast<const float4 *>(colm + k4);
            if (k4 > m) A[k4 >> 1] = __ffma2_rn(xx, make_float2(lv.x, lv.y), A[k4 >> 1]);
            else if (k4 + 1 > m) A[k4 >> 1].y = fmaf(-xm, lv.y, A[k4 >> 1].y);
            if (k4 + 2 > m) A[(k4 >> 1) + 1] = __ffma2_rn(xx, make_float2(lv.z, lv.w), A[(k4 >> 1) + 1]);
            else if (k4 + 3 > m) A[(k4 >> 1) + 1].y = fmaf(-xm, lv.w, A[(k4 >> 1) + 1].y);
        }
    }
#pragma unroll
    for (int w = 0; w < 16; ++w) {                         // publish row 32+l of L21
        const int c = 2 * w;
        Ls[pk_off(c) - 4 * (c >> 2) + lane + 32] = A[w].x;
        Ls[pk_off(c + 1) - 4 * ((c + 1) >> 2) + lane + 32] = A[w].y;
    }
    __syncwarp();
    // ---------------- syrk: A22[32+l][32+k] -= sum_{m<32} L21[l][m] L21[k][m]
#pragma unroll
    for (int w = 0; w < 16; ++w) {
        const int c = 32 + 2 * w;
        A[w].x = (c <= lane + 32) ? Ls[pk_off(c) - 4 * (c >> 2) + lane + 32] : 0.f;
        A[w].y = (c + 1 <= lane + 32) ? Ls[pk_off(c + 1) - 4 * ((c + 1) >> 2) + lane + 32] : 0.f;
    }
#pragma unroll 2
    for (int m = 0; m < 32; ++m) {
        const int gq = m >> 2, len = 64 - 4 * gq;
        const float *colm = Ls + 256 * gq - 8 * gq * (gq - 1) + (m & 3) * len - 4 * gq;   // column m, indexed by row
        const float own = -colm[lane + 32];
        const float2 oo = make_float2(own, own);
#pragma unroll
        for (int w4 = 0; w4 < 8; ++w4) {
            const float4 lv = *reinterpret_cast<const float4 *>(colm + 32 + 4 * w4);
            A[2 * w4] = __ffma2_rn(oo, make_float2(lv.x, lv.y), A[2 * w4]);
            A[2 * w4 + 1] = __ffma2_rn(oo, make_float2(lv.z, lv.w), A[2 * w4 + 1]);
        }
    }
}

template <bool DUMP, bool LDGSTS>
__global__ void __launch_bounds__(NWARPS * 32, 1) k_als_rows_tc64(
    const __grid_constant__ CUtensorMap map_pl, const int64_t *__restrict__ ptr,
    const int32_t *__restrict__ stage_ptr, const int32_t *__restrict__ colp, const float *__restrict__ valp,
    const int32_t *__restrict__ cta_rows, float *__restrict__ T, __half *__restrict__ T_pl, float lambda,
    int64_t row_begin, int *__restrict__ info, float *__restrict__ dump, __half *const *__restrict__ peers, int n_peers,
    const __half *__restrict__ S_pl) {
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

    // zero the ring once: the R / R' chunks only ever get their first four bytes per row rewritten
    for (int i = threadIdx.x; i < NSTG * STAGE_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sm + Smem::ring)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        // arrivals per full barrier: one expect_tx arrive (TMA gathers) or, with LDGSTS producers, every lane's copies
        // (cp.async.mbarrier.arrive.noinc) plus one release-arrive for the rating chunks
        for (int s = 0; s < NSTG; ++s) { mbar_init(bar_full(s), LDGSTS ? 33 : 1); mbar_init(bar_empty(s), 1); }
        for (int a = 0; a < NACC; ++a) { mbar_init(bar_acc_full(a), 1); mbar_init(bar_acc_empty(a), N_EPI); }
        for (int g = 0; g < NGB; ++g) { mbar_init(bar_g_full(g), N_EPI); mbar_init(bar_g_empty(g), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sm_base + Smem::tmem_ptr) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();                                           // the zero fill must be visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // this CTA owns the contiguous rows [first, last) == the stages [stage_ptr[first], stage_ptr[last])
    const int64_t first = cta_rows[blockIdx.x], last = cta_rows[blockIdx.x + 1];

    if (warp >= W_MMA) {
      // warpgroups 4-5: tensor-core issuer + TMA producers need few registers
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_AUX));
      if (warp >= W_PROD0) {
        // ================================================================== producers: stage stream -> TMA gather4 + rating chunks
        // The ratings of a side are pre-laid as 16-entry stages (colp / valp, rows padded with the planes' zero row),
        // so a producer warp just streams: stage s of this CTA lives at [16 s, 16 s + 16).
        const uint32_t pw = (uint32_t)(warp - W_PROD0);           // this warp owns local stages pw, pw + N_PROD, ...
        const int64_t s_lo = stage_ptr[first], s_hi = stage_ptr[last];
        const uint32_t n_local = (uint32_t)(s_hi - s_lo);
        const int32_t *cp = colp + s_lo * KROWS + (lane & 15);
        const float *vp = valp + s_lo * KROWS + (lane & 15);
        int idx0 = 0, idx1 = 0;
        float rt0 = 0.f, rt1 = 0.f;
        if (pw < n_local) { idx0 = cp[(int64_t)pw * KROWS]; rt0 = vp[(int64_t)pw * KROWS]; }
        if (pw + N_PROD < n_local) { idx1 = cp[(int64_t)(pw + N_PROD) * KROWS]; rt1 = vp[(int64_t)(pw + N_PROD) * KROWS]; }
        for (uint32_t c = pw; c < n_local; c += N_PROD) {
            int idx2 = 0;
            float rt2 = 0.f;
            if (c + 2 * N_PROD < n_local) { idx2 = cp[(int64_t)(c + 2 * N_PROD) * KROWS]; rt2 = vp[(int64_t)(c + 2 * N_PROD) * KROWS]; }
            const int slot = c % NSTG;
            const uint32_t stage = sm_base + Smem::ring + slot * STAGE_BYTES;
            if (LDGSTS) {
                // Short rows (a handful of stages per system): the gathered plane rows sit in L2 and the producers' own
                // instruction stream competes with the solver warps, so 16-byte asynchronous copies (LDGSTS: 8 per lane
                // and stage, a quarter of the instructions of the elected-lane TMA issue) are used instead.  Lane l moves
                // chunk (l & 7) of rows (l >> 3) + 4 k of both halves of the plane row into the 128B-swizzled tile
                // layout (chunk j of row r at j ^ (r & 7)); the MMA issuer adds the generic -> async proxy fence.
                const int rq = lane >> 3, jc = lane & 7;
                int ri[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) ri[k] = __shfl_sync(0xffffffffu, idx0, 4 * k + rq);
                mbar_wait(bar_empty(slot), ((c / NSTG) & 1) ^ 1);
                if (lane < 16) {
                    __half rh, rl;
                    split_half(rt0, rh, rl);
                    const uint32_t off = (uint32_t)lane * 128u + (((uint32_t)lane & 7u) << 4);
                    const uint32_t w0 = (uint32_t)__half_as_ushort(rh) | ((uint32_t)__half_as_ushort(rl) << 16);
                    const uint32_t w1 = (uint32_t)__half_as_ushort(rh) << 16;
                    *reinterpret_cast<uint32_t *>(sm + Smem::ring + slot * STAGE_BYTES + 2 * TILE_BYTES + off) = w0;
                    *reinterpret_cast<uint32_t *>(sm + Smem::ring + slot * STAGE_BYTES + 2 * TILE_BYTES + (off ^ 16u)) = w1;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int r = 4 * k + rq;
                    const uint32_t dst = stage + (uint32_t)r * 128u + (uint32_t)((jc ^ (r & 7)) << 4);
                    const char *src = reinterpret_cast<const char *>(S_pl) + (size_t)ri[k] * (PLANE_PITCH * 2) + (size_t)jc * 16;
                    cp_async16(dst, src);
                    cp_async16(dst + TILE_BYTES, src + D * 2);
                }
                cp_async_arrive_noinc(bar_full(slot));             // fires when this lane's eight copies have landed
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full(slot));        // releases the rating chunks written above
            } else {
                // every lane gets the 16 row indices of the stage (lanes 0..15 hold one each)
                int r[16];
    #pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = __shfl_sync(0xffffffffu, idx0, i);
                mbar_wait(bar_empty(slot), ((c / NSTG) & 1) ^ 1);
                // ratings: R row r = (r_hi, r_lo * 2^11, 0...), R' row r = (0, r_hi, 0...); 16-byte chunk 0 sits at chunk (r & 7)
                if (lane < 16) {
                    __half rh, rl;
                    split_half(rt0, rh, rl);
                    const uint32_t off = (uint32_t)lane * 128u + (((uint32_t)lane & 7u) << 4);
                    const uint32_t w0 = (uint32_t)__half_as_ushort(rh) | ((uint32_t)__half_as_ushort(rl) << 16);
                    const uint32_t w1 = (uint32_t)__half_as_ushort(rh) << 16;
                    *reinterpret_cast<uint32_t *>(sm + Smem::ring + slot * STAGE_BYTES + 2 * TILE_BYTES + off) = w0;
                    *reinterpret_cast<uint32_t *>(sm + Smem::ring + slot * STAGE_BYTES + 2 * TILE_BYTES + (off ^ 16u)) = w1;   // logical column group 1
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_expect_tx(bar_full(slot), 2 * TILE_BYTES);
                __syncwarp();
    #pragma unroll
                for (int t = 0; t < 4; ++t)
                    tma_gather4_pair(stage + t * 512, stage + TILE_BYTES + t * 512, &map_pl, bar_full(slot), r[4 * t], r[4 * t + 1],
                                     r[4 * t + 2], r[4 * t + 3]);
            }
            idx0 = idx1; rt0 = rt1;
            idx1 = idx2; rt1 = rt2;
        }
      } else {
        // ================================================================== MMA issuer
        // One warp, in-order, is the pacing agent of a gather-bound half-step (measured: ~290 cycles per stage when every
        // stage paid a blocking try_wait (~80 cycles even when the data is there), four descriptor builds and a modulo).
        // So the ring cursor (slot, phase, descriptor words, barrier addresses) advances incrementally and the probe
        // of the NEXT slot's full barrier is issued before this stage's MMAs, hiding its latency under them.
        uint32_t job = 0;
        const uint32_t idesc1 = make_idesc(MMA_N1), idesc2 = make_idesc(MMA_N2);
        constexpr uint32_t LBO_T = (uint32_t)(TILE_BYTES >> 4) << 16, LBO_B2 = (uint32_t)((2 * TILE_BYTES + 16) >> 4) << 16;
        const uint32_t enc0 = ((sm_base + Smem::ring) >> 4) & 0x3FFF;
        uint32_t slot = 0, ph = 0, enc = enc0, ready = 0;
        int nst = (first < last) ? stage_ptr[first + 1] - stage_ptr[first] : 0;
        for (int64_t row = first; row < last; ++row, ++job) {
            const int nst_next = (row + 1 < last) ? stage_ptr[row + 2] - stage_ptr[row + 1] : 0;
            const int acc = job % NACC;
            mbar_wait(bar_acc_empty(acc), ((job / NACC) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_base + (uint32_t)acc * ACC_COLS;
            for (int st = 0; st < nst; ++st) {
                if (!ready) mbar_wait(bar0 + 8u * slot, ph);
                tc_fence_after();
                if (LDGSTS) fence_proxy_async();                   // generic-proxy writes (cp.async, st.shared) -> tensor-core reads
                uint32_t nslot = slot + 1, nph = ph, nenc = enc + (STAGE_BYTES >> 4);
                if (nslot == NSTG) { nslot = 0; nph ^= 1u; nenc = enc0; }
                const uint32_t nready = mbar_test(bar0 + 8u * nslot, nph);
                tc_stage_mma(d, enc | LBO_T, (enc + (TILE_BYTES >> 4)) | LBO_T, enc | LBO_B2, idesc1, idesc2, st > 0 ? 1u : 0u,
                             bar0 + 8u * (NSTG + slot), bar_acc_full(acc), st == nst - 1 ? 1u : 0u);
                slot = nslot; ph = nph; enc = nenc; ready = nready;
            }
            nst = nst_next;
        }
      }
    } else if (warp < W_SOLVE0) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
        // ================================================================== epilogue: TMEM -> registers -> hand-off buffer
        const int q = warp - W_EPI0;                              // TMEM lane quadrant == warp index % 4
        uint32_t job = 0;
        for (int64_t row = first; row < last; ++row, ++job) {
            const int acc = job % NACC, g = job % NGB;
            const float reg = lambda * (float)(ptr[row + 1] - ptr[row]);
            mbar_wait(bar_acc_full(acc), (job / NACC) & 1);
            tc_fence_after();
            mbar_wait(bar_g_empty(g), ((job / NGB) & 1) ^ 1);
            float *bo = reinterpret_cast<float *>(sm + Smem::bsol + g * D * 4);
            float *G = reinterpret_cast<float *>(sm + Smem::gbuf + g * PK_FLOATS * 4);
            const int my_row = 16 * q + lane;                     // valid for lane < 16
            const uint32_t tq = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)acc * ACC_COLS;
#pragma unroll
            for (int c0 = 0; c0 < D; c0 += 32) {
                if (c0 > 16 * q + 15) break;                       // this quadrant's rows have no lower-triangle entry there
                float d1[32], d2[32];
                tc_ld32(tq + c0, d1);
                tc_ld32(tq + 64 + c0, d2);
                if (lane < 16) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int cc = c0 + c;
                        if (cc <= my_row) {                        // lower triangle only, packed column-major
                            float v = fmaf(d2[c], LO_INV, d1[c]);
                            if (cc == my_row) v += reg;
                            G[pk_off(cc) - 4 * (cc >> 2) + my_row] = v;
                        }
                    }
                }
            }
            float b_hi, b_lo;
            tc_ld2(tq + 128, b_hi, b_lo);
            if (lane < 16) bo[my_row] = fmaf(b_lo, LO_INV, b_hi);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar_acc_empty(acc)); mbar_arrive(bar_g_full(g)); }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SOLVE));
        // ================================================================== solvers: one warp per system
        // The system arrives packed (lower triangle, column-major) in this warp's private buffer, which then becomes the
        // storage of L.  With A = [A11 .; A21 A22] in 32x32 blocks and ONE matrix row per lane at a time:
        //   chol32(A11)  ->  L11, z1          (lane = row l, 32 registers of matrix data)
        //   trsm         ->  L21 = A21 L11^-T (lane = row 32+l; a private triangular solve per row: no shuffles, no barriers)
        //   syrk         ->  A22 -= L21 L21^T (lane = row 32+l)
        //   chol32(A22)  ->  L22, z2
        //   back substitution: x2 from L22^T, z1 -= L21^T x2, x1 from L11^T.
        const int sv = warp - W_SOLVE0;
        const int g = sv;
        float *invd = reinterpret_cast<float *>(sm + Smem::invd + sv * D * 4);
        float *Ls = reinterpret_cast<float *>(sm + Smem::gbuf + g * PK_FLOATS * 4);   // entry (i, c) at pk_at(i, c)
        float *bx = reinterpret_cast<float *>(sm + Smem::bsol + g * D * 4);           // rhs in; scratch for z1 / x2 afterwards
        uint32_t use = 0;
        for (int64_t row = first + sv; row < last; row += N_SOLVE) {
            mbar_wait(bar_g_full(g), use & 1);
            ++use;
            float b0 = bx[lane], b1 = bx[lane + 32];
            if (DUMP) {
                float *o = dump + (int64_t)(row - row_begin) * (D * D + D);
                for (int c = 0; c < D; ++c) {                      // symmetric fill from the packed lower triangle
                    const int i0 = lane, i1 = lane + 32;
                    o[i0 * D + c] = (c <= i0) ? Ls[pk_at(i0, c)] : Ls[pk_at(c, i0)];
                    o[i1 * D + c] = (c <= i1) ? Ls[pk_at(i1, c)] : Ls[pk_at(c, i1)];
                }
                o[D * D + lane] = b0; o[D * D + lane + 32] = b1;
            }
            float2 A[16];
            float bc = b0, z1 = 0.f;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                if (h == 0) {
                    // ---------------- A11: row l, columns 0..31
#pragma unroll
                    for (int w = 0; w < 16; ++w) {
                        const int c = 2 * w;
                        A[w].x = (c <= lane) ? Ls[pk_off(c) - 4 * (c >> 2) + lane] : 0.f;
                        A[w].y = (c + 1 <= lane) ? Ls[pk_off(c + 1) - 4 * ((c + 1) >> 2) + lane] : 0.f;
                    }
                    __syncwarp();
                } else {
                    solver_trsm_syrk(A, bc, Ls, invd, bx, lane);
                }
                chol32(A, bc, Ls, invd, 32 * h, lane);
                if (h == 0) {
                    z1 = bc;
                    bx[lane] = z1;                                 // z1, broadcast source for the trsm rhs update
                    __syncwarp();
                    bc = b1;
                }
            }
            b0 = z1; b1 = bc;
            // ---------------- back substitution
            backsub32(b1, Ls, invd, 32, lane);                     // b1 = x[32+l]
            bx[32 + lane] = b1;
            __syncwarp();
            {                                                      // z1[l] -= sum_i L21[i][l] x2[i]
                const int gi = lane >> 2;
                const float *coll = Ls + 256 * gi - 8 * gi * (gi - 1) + (lane & 3) * (64 - 4 * gi) - 4 * gi + 32;   // column l, rows 32..
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const float4 lv = *reinterpret_cast<const float4 *>(coll + 4 * t);
                    const float4 xv = *reinterpret_cast<const float4 *>(bx + 32 + 4 * t);
                    acc = __ffma2_rn(make_float2(lv.x, lv.y), make_float2(xv.x, xv.y), acc);
                    acc = __ffma2_rn(make_float2(lv.z, lv.w), make_float2(xv.z, xv.w), acc);
                }
                b0 -= acc.x + acc.y;
            }
            backsub32(b0, Ls, invd, 0, lane);                      // b0 = x[l]
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_g_empty(g));           // L storage / scratch no longer needed
            T[row * D + lane] = b0;
            T[row * D + lane + 32] = b1;
            __half h0, l0, h1, l1;
            split_half(b0, h0, l0);
            split_half(b1, h1, l1);
            // the plane row (256 B) is staged through this warp's scratch so that each lane stores 8 contiguous bytes --
            // into the local plane, or, when peers are mapped, into every rank's replica (posted NVLink writes): the
            // all-gather of the half-step happens while the other solver warps are still factorising
            __half *stg = reinterpret_cast<__half *>(invd);
            stg[lane] = h0; stg[lane + 32] = h1;
            stg[D + lane] = l0; stg[D + 32 + lane] = l1;
            __syncwarp();
            const uint2 prow = reinterpret_cast<const uint2 *>(stg)[lane];
            __syncwarp();
            const int64_t poff = row * PLANE_PITCH + 4 * lane;
            if (n_peers == 0) {
                *reinterpret_cast<uint2 *>(T_pl + poff) = prow;
            } else {
                for (int p = 0; p < n_peers; ++p) *reinterpret_cast<uint2 *>(peers[p] + poff) = prow;
            }
            const bool bad = !(fabsf(b0) <= 3.0e38f) || !(fabsf(b1) <= 3.0e38f);                 // NaN / inf: the factorisation broke down
            if (__any_sync(0xffffffffu, bad) && lane == 0) atomicAdd(info, 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// =====================================================================================================================
// Kernel v2 (round 2).  Same producers and MMA issuer; what changed is everything behind the accumulator:
//   * the epilogue warps read TMEM with tcgen05.ld.16x256b -- the register layout of an mma.sync accumulator -- and hand the
//     NEGATED system over as 16x8 fragment tiles (chol_mma.cuh "tile layout"): 4 FFMA + 1 STS.128 per tile instead of a
//     predicated scalar store per matrix entry (1 430 -> ~300 warp instructions per system);
//   * the solver warps factor it on the warp-level tensor path (cm::solve_tiles: blocked right-looking Cholesky, 3xTF32
//     HMMA for the triangular solves and trailing updates, forward substitution as one more matrix row);
//   * ring depth and the number of ACTIVE solver warps are template parameters: the long-row half-step (few systems, every
//     stage a DRAM gather) trades solver buffers for a deeper ring -- gathers in flight are what paces it (Little's law:
//     13 stages x 6 KB against ~2.5 us of loaded DRAM latency is one stage per ~200 cycles).
// 24 warps: epilogue 0-3 | solvers 4-15 | MMA 16 | producers 17-23.
// Warp roles are a template parameter set: NSW solver warps after the 4 epilogue warps, then the MMA issuer, then NP producers
// (the rest of the last warpgroup idles).  Register budgets (setmaxnreg, per warpgroup) must fit the launch allocation 768 x 80:
//   long rows : 12 solver warps x 104 + 4 x 72 + 8 x 40   (6 solver buffers in use: few systems, many stages)
//   short rows: 16 solver warps x  96 + 4 x 56 + 4 x 40   (the solvers pace this half-step: every warp that fits is one)
constexpr int V2_WARPS = 24, V2_W_SOLVE0 = 4;
constexpr int V2_REGS_AUX = 40;
using TL4 = cm::Tiles<4>;
constexpr int V2_BUF_FLOATS = TL4::FLOATS + 3 * D;                          // tiles | bz | invd | xs  per solver warp

template <int NSTG_, int NSA, int KS>
struct Smem2 {
    static constexpr int ring = 0;
    static constexpr int bufs = ring + NSTG_ * KS * STAGE_BYTES;
    static constexpr int bars = bufs + NSA * V2_BUF_FLOATS * 4;
    static constexpr int n_bars = 2 * NSTG_ + 2 * NACC + 2 * NSA;
    static constexpr int tmem_ptr = bars + n_bars * 8;
    static constexpr int total = tmem_ptr + 16;
};

__device__ __forceinline__ void tc_ld16x256_x4(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_ld16x256_x1(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <bool DUMP, bool LDGSTS, int NSTG_, int NSA, int KS, int NP, int NSW>
__global__ void __launch_bounds__(V2_WARPS * 32, 1) k_als_rows_tc64_v2(
    const __grid_constant__ CUtensorMap map_pl, const int64_t *__restrict__ ptr,
    const int32_t *__restrict__ stage_ptr, const int32_t *__restrict__ colp, const float *__restrict__ valp,
    const int32_t *__restrict__ cta_rows, float *__restrict__ T, __half *__restrict__ T_pl, float lambda,
    int64_t row_begin, int *__restrict__ info, float *__restrict__ dump, __half *const *__restrict__ peers, int n_peers,
    const __half *__restrict__ S_pl, int dbg) {
    using SM = Smem2<NSTG_, NSA, KS>;
    constexpr int V2_W_MMA = V2_W_SOLVE0 + NSW, V2_W_PROD0 = V2_W_MMA + 1;
    constexpr int V2_REGS_EPI = NSW > 12 ? 56 : 72, V2_REGS_SOLVE = NSW > 12 ? 96 : 104;
    static_assert(NSW % 4 == 0 && V2_W_PROD0 + NP <= V2_WARPS && NSA <= NSW, "warp roles");
    static_assert((N_EPI * V2_REGS_EPI + NSW * V2_REGS_SOLVE + (V2_WARPS - N_EPI - NSW) * V2_REGS_AUX) <= V2_WARPS * 80, "register budgets");
    // Parity waits tell adjacent phases apart only: a producer may run at most one ring revolution ahead of the issuer.  Its
    // previous stage was NP stages back and legitimately within NSTG_ of the issuer, so NSTG_ >= NP keeps it inside two
    // phases (with 6 slots and 7 producers the seventh stage's producer passed its wait two phases early: mbarrier overflow).
    static_assert(NSTG_ >= NP, "ring depth must cover the active producers");
    constexpr int STG = KS * STAGE_BYTES;                       // one ring slot: KS sub-stages of 16 ratings (one handshake per slot)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms need 1024-B alignment
    const uint32_t sm_base = smem_u32(sm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = sm_base + SM::bars;
    auto bar_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return bar0 + 8u * (NSTG_ + s); };
    auto bar_acc_full = [&](int a) { return bar0 + 8u * (2 * NSTG_ + a); };
    auto bar_acc_empty = [&](int a) { return bar0 + 8u * (2 * NSTG_ + NACC + a); };
    auto bar_g_full = [&](int g) { return bar0 + 8u * (2 * NSTG_ + 2 * NACC + g); };
    auto bar_g_empty = [&](int g) { return bar0 + 8u * (2 * NSTG_ + 2 * NACC + NSA + g); };
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(sm + SM::tmem_ptr);

    // zero the ring once: the R / R' chunks only ever get their first four bytes per row rewritten
    for (int i = threadIdx.x; i < NSTG_ * STG / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sm + SM::ring)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTG_; ++s) { mbar_init(bar_full(s), LDGSTS ? 33 : 1); mbar_init(bar_empty(s), 1); }
        for (int a = 0; a < NACC; ++a) { mbar_init(bar_acc_full(a), 1); mbar_init(bar_acc_empty(a), N_EPI); }
        for (int g = 0; g < NSA; ++g) { mbar_init(bar_g_full(g), N_EPI); mbar_init(bar_g_empty(g), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == V2_W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sm_base + SM::tmem_ptr) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();                                           // the zero fill must be visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t first = cta_rows[blockIdx.x], last = cta_rows[blockIdx.x + 1];

    if (warp >= V2_W_MMA) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(V2_REGS_AUX));
      if (warp >= V2_W_PROD0) {
        // ================================================================== producers: one ring slot = KS sub-stages of 16 ratings
        // (round 2: the mbarrier handshake per slot, not the gathers or the MMAs, paced the kernel -- with every MMA and every
        // gather switched off it still needed ~280 cycles per 16-rating stage -- so a slot now carries KS x 16 ratings)
        const uint32_t pw = (uint32_t)(warp - V2_W_PROD0);
        const int64_t s_lo = stage_ptr[first], s_hi = stage_ptr[last];
        const uint32_t n_local = pw < (uint32_t)NP ? (uint32_t)(s_hi - s_lo) : 0u;      // producers beyond NP stay idle
        const int32_t *cp = colp + s_lo * (KROWS * KS) + (lane & 15);
        const float *vp = valp + s_lo * (KROWS * KS) + (lane & 15);
        int idx0[KS], idx1[KS];
        float rt0[KS], rt1[KS];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            idx0[ks] = 0; rt0[ks] = 0.f;
            if (pw < n_local) { idx0[ks] = cp[((int64_t)pw * KS + ks) * KROWS]; rt0[ks] = vp[((int64_t)pw * KS + ks) * KROWS]; }
        }
        for (uint32_t c = pw; c < n_local; c += NP) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                idx1[ks] = 0; rt1[ks] = 0.f;
                if (c + NP < n_local) { idx1[ks] = cp[((int64_t)(c + NP) * KS + ks) * KROWS]; rt1[ks] = vp[((int64_t)(c + NP) * KS + ks) * KROWS]; }
            }
            const int slot = c % NSTG_;
            const uint32_t stage0 = sm_base + SM::ring + slot * STG;
            unsigned char *stage0_p = sm + SM::ring + slot * STG;
            mbar_wait(bar_empty(slot), ((c / NSTG_) & 1) ^ 1);
            // ratings: R row r = (r_hi, r_lo * 2^11, 0...), R' row r = (0, r_hi, 0...); 16-byte chunk 0 sits at chunk (r & 7)
            if (lane < 16) {
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    __half rh, rl;
                    split_half(rt0[ks], rh, rl);
                    const uint32_t off = (uint32_t)lane * 128u + (((uint32_t)lane & 7u) << 4);
                    const uint32_t w0 = (uint32_t)__half_as_ushort(rh) | ((uint32_t)__half_as_ushort(rl) << 16);
                    const uint32_t w1 = (uint32_t)__half_as_ushort(rh) << 16;
                    *reinterpret_cast<uint32_t *>(stage0_p + ks * STAGE_BYTES + 2 * TILE_BYTES + off) = w0;
                    *reinterpret_cast<uint32_t *>(stage0_p + ks * STAGE_BYTES + 2 * TILE_BYTES + (off ^ 16u)) = w1;
                }
            }
            if (LDGSTS) {
                // lane l moves chunk (l & 7) of rows (l >> 3) + 4 k of both halves of the plane row into the 128B-swizzled tile
                const int rq = lane >> 3, jc = lane & 7;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const uint32_t stage = stage0 + ks * STAGE_BYTES;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const int r = 4 * k4 + rq;
                        const int ri = __shfl_sync(0xffffffffu, idx0[ks], r);
                        const uint32_t dst = stage + (uint32_t)r * 128u + (uint32_t)((jc ^ (r & 7)) << 4);
                        const char *src = reinterpret_cast<const char *>(S_pl) + (size_t)ri * (PLANE_PITCH * 2) + (size_t)jc * 16;
                        cp_async16(dst, src);
                        cp_async16(dst + TILE_BYTES, src + D * 2);
                    }
                }
                cp_async_arrive_noinc(bar_full(slot));             // fires when all of this lane's copies have landed
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full(slot));        // releases the rating chunks written above
            } else {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_expect_tx(bar_full(slot), (dbg & 4) ? 0u : KS * 2 * TILE_BYTES);
                __syncwarp();
                if (!(dbg & 4)) {
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        const uint32_t stage = stage0 + ks * STAGE_BYTES;
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const int r0 = __shfl_sync(0xffffffffu, idx0[ks], 4 * t), r1 = __shfl_sync(0xffffffffu, idx0[ks], 4 * t + 1);
                            const int r2 = __shfl_sync(0xffffffffu, idx0[ks], 4 * t + 2), r3 = __shfl_sync(0xffffffffu, idx0[ks], 4 * t + 3);
                            tma_gather4_pair(stage + t * 512, stage + TILE_BYTES + t * 512, &map_pl, bar_full(slot), r0, r1, r2, r3);
                        }
                    }
                }
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) { idx0[ks] = idx1[ks]; rt0[ks] = rt1[ks]; }
        }
      } else {
        // ================================================================== MMA issuer: one handshake per ring slot, KS x 2 MMAs
        uint32_t job = 0;
        const uint32_t idesc1 = make_idesc(MMA_N1), idesc2 = make_idesc(MMA_N2);
        constexpr uint32_t LBO_T = (uint32_t)(TILE_BYTES >> 4) << 16, LBO_B2 = (uint32_t)((2 * TILE_BYTES + 16) >> 4) << 16;
        const uint32_t enc0 = ((sm_base + SM::ring) >> 4) & 0x3FFF;
        uint32_t slot = 0, ph = 0, enc = enc0, ready = 0;
        int nst = (first < last) ? stage_ptr[first + 1] - stage_ptr[first] : 0;
        for (int64_t row = first; row < last; ++row, ++job) {
            const int nst_next = (row + 1 < last) ? stage_ptr[row + 2] - stage_ptr[row + 1] : 0;
            const int acc = job % NACC;
            mbar_wait(bar_acc_empty(acc), ((job / NACC) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_base + (uint32_t)acc * ACC_COLS;
            for (int st = 0; st < nst; ++st) {
                if (!ready) mbar_wait(bar0 + 8u * slot, ph);
                tc_fence_after();
                if (LDGSTS) fence_proxy_async();                   // generic-proxy writes (cp.async, st.shared) -> tensor-core reads
                uint32_t nslot = slot + 1, nph = ph, nenc = enc + (STG >> 4);
                if (nslot == NSTG_) { nslot = 0; nph ^= 1u; nenc = enc0; }
                const uint32_t nready = (dbg & 2) ? 0u : mbar_test(bar0 + 8u * nslot, nph);
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const uint32_t e = enc + ks * (STAGE_BYTES >> 4);
                    tc_stage_mma2(d, e | LBO_T, (e + (TILE_BYTES >> 4)) | LBO_T, e | LBO_B2, idesc1, idesc2, (st > 0 || ks > 0) ? 1u : 0u,
                                  bar0 + 8u * (NSTG_ + slot), bar_acc_full(acc), ks == KS - 1 ? 1u : 0u, (ks == KS - 1 && st == nst - 1) ? 1u : 0u,
                                  (dbg & 8) ? 0u : ((dbg & 1) ? 1u : 2u));
                }
                slot = nslot; ph = nph; enc = nenc; ready = nready;
            }
            nst = nst_next;
        }
      }
    } else if (warp < V2_W_SOLVE0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(V2_REGS_EPI));
        // ================================================================== epilogue: TMEM -> fragment tiles of -(G + lambda n I), -b
        // M = 64 accumulator: rows 16 q .. 16 q + 15 live in the first 16 lanes of quadrant q == m-tile q of the system.
        const int q = warp;
        const int g8 = lane >> 2, t4 = lane & 3;
        const bool dg0 = g8 == 2 * t4, dg1 = g8 == 2 * t4 + 1;      // this lane's entries on the diagonal of a diagonal tile
        uint32_t job = 0;
        for (int64_t row = first; row < last; ++row, ++job) {
            const int acc = job % NACC, g = job % NSA;
            const float reg = lambda * (float)(ptr[row + 1] - ptr[row]);
            mbar_wait(bar_acc_full(acc), (job / NACC) & 1);
            tc_fence_after();
            mbar_wait(bar_g_empty(g), ((job / NSA) & 1) ^ 1);
            float *buf = reinterpret_cast<float *>(sm + SM::bufs) + (size_t)g * V2_BUF_FLOATS;
            float4 *tiles = reinterpret_cast<float4 *>(buf) + (q * (q + 1)) * 32 + lane;       // tile (q, j) at + 32 j
            const uint32_t tq = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)acc * ACC_COLS;
#pragma unroll
            for (int c0 = 0; c0 < 8; c0 += 4) {
                if (c0 > 2 * q + 1) break;
                float d1[16], d2[16];
                tc_ld16x256_x4(tq + 8 * c0, d1);
                tc_ld16x256_x4(tq + 64 + 8 * c0, d2);
                tc_wait_ld();
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = c0 + jj;
                    if (j <= 2 * q + 1) {
                        float4 v;
                        v.x = fmaf(-LO_INV, d2[4 * jj + 0], -d1[4 * jj + 0]);
                        v.y = fmaf(-LO_INV, d2[4 * jj + 1], -d1[4 * jj + 1]);
                        v.z = fmaf(-LO_INV, d2[4 * jj + 2], -d1[4 * jj + 2]);
                        v.w = fmaf(-LO_INV, d2[4 * jj + 3], -d1[4 * jj + 3]);
                        if (j == 2 * q) { v.x -= dg0 ? reg : 0.f; v.y -= dg1 ? reg : 0.f; }          // rows g   x cols 16 q + 2 t (+1)
                        if (j == 2 * q + 1) { v.z -= dg0 ? reg : 0.f; v.w -= dg1 ? reg : 0.f; }      // rows g+8 x cols 16 q + 8 + 2 t (+1)
                        tiles[32 * j] = v;
                    }
                }
            }
            float bb[4];
            tc_ld16x256_x1(tq + 128, bb);                          // column 128: sum r_hi y, column 129: the 2^-11 terms
            tc_wait_ld();
            if (t4 == 0) {
                float *bz = buf + TL4::FLOATS;
                bz[16 * q + g8] = -fmaf(bb[1], LO_INV, bb[0]);
                bz[16 * q + g8 + 8] = -fmaf(bb[3], LO_INV, bb[2]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar_acc_empty(acc)); mbar_arrive(bar_g_full(g)); }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(V2_REGS_SOLVE));
        // ================================================================== solvers: one system per warp on the HMMA path
        const int sv = warp - V2_W_SOLVE0;
        if (sv < NSA) {
            float *S = reinterpret_cast<float *>(sm + SM::bufs) + (size_t)sv * V2_BUF_FLOATS;
            float *bz = S + TL4::FLOATS, *invd = bz + D, *xs = invd + D;
            uint32_t use = 0;
            for (int64_t row = first + sv; row < last; row += NSA) {
                mbar_wait(bar_g_full(sv), use & 1);
                ++use;
                if (DUMP) {
                    float *o = dump + (int64_t)(row - row_begin) * (D * D + D);
                    for (int c = 0; c < D; ++c) {
                        const int i0 = lane, i1 = lane + 32;
                        o[i0 * D + c] = -S[c <= i0 ? TL4::at(i0, c) : TL4::at(c, i0)];
                        o[i1 * D + c] = -S[c <= i1 ? TL4::at(i1, c) : TL4::at(c, i1)];
                    }
                    o[D * D + lane] = -bz[lane]; o[D * D + lane + 32] = -bz[lane + 32];
                    __syncwarp();
                }
                float x[2] = {0.f, 0.f};
                if (!(dbg & 16)) cm::solve_tiles<4>(S, bz, invd, xs, lane, x);
                const float b0 = x[0], b1 = x[1];
                __half h0, l0, h1, l1;
                split_half(b0, h0, l0);
                split_half(b1, h1, l1);
                __syncwarp();
                // the plane row (256 B) is staged through this warp's scratch so that each lane stores 8 contiguous bytes
                __half *stg = reinterpret_cast<__half *>(invd);
                stg[lane] = h0; stg[lane + 32] = h1;
                stg[D + lane] = l0; stg[D + 32 + lane] = l1;
                __syncwarp();
                const uint2 prow = reinterpret_cast<const uint2 *>(stg)[lane];
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_g_empty(sv));          // tiles / scratch no longer needed
                T[row * D + lane] = b0;
                T[row * D + lane + 32] = b1;
                const int64_t poff = row * PLANE_PITCH + 4 * lane;
                if (n_peers == 0) {
                    *reinterpret_cast<uint2 *>(T_pl + poff) = prow;
                } else {
                    for (int p = 0; p < n_peers; ++p) *reinterpret_cast<uint2 *>(peers[p] + poff) = prow;
                }
                const bool bad = !(fabsf(b0) <= 3.0e38f) || !(fabsf(b1) <= 3.0e38f);             // NaN / inf: the factorisation broke down
                if (__any_sync(0xffffffffu, bad) && lane == 0) atomicAdd(info, 1);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == V2_W_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// fp32 factor rows -> gather planes (after init, after an all-gather, after the generic kernel wrote a side)
__global__ void k_split_planes(const float *__restrict__ f, int64_t n, __half *__restrict__ pl) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    __half h, l;
    split_half(f[i], h, l);
    const int64_t o = (i / D) * PLANE_PITCH + (i % D);
    pl[o] = h;
    pl[o + D] = l;
}

// ---- stage CSR build (one-off per load / shard geometry)
__global__ void k_stage_counts(const int64_t *__restrict__ ptr, int64_t n_rows, int32_t *__restrict__ cnt, int krows) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    cnt[r] = r < n_rows ? (int32_t)((ptr[r + 1] - ptr[r] + krows - 1) / krows) : 0;
}
__global__ void k_fill_stages(const int64_t *__restrict__ ptr, const int32_t *__restrict__ col, const float *__restrict__ val,
                              const int32_t *__restrict__ stage_ptr, int64_t n_rows, int zero_row, int32_t *__restrict__ colp,
                              float *__restrict__ valp, int krows) {
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const int64_t s = ptr[row], e = ptr[row + 1];
    const int64_t base = (int64_t)stage_ptr[row] * krows;
    const int64_t n_pad = (int64_t)(stage_ptr[row + 1] - stage_ptr[row]) * krows;
    for (int64_t t = lane; t < n_pad; t += 32) {
        const int64_t k = s + t;
        colp[base + t] = k < e ? col[k] : zero_row;
        valp[base + t] = k < e ? val[k] : 0.f;
    }
}
// cta_rows[b] = first row of CTA b: rows [rb, re) split into `grid` contiguous ranges of (nearly) equal stage count
__global__ void k_cta_partition(const int32_t *__restrict__ stage_ptr, int64_t rb, int64_t re, int grid, int32_t *__restrict__ cta_rows) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > grid) return;
    if (b == grid) { cta_rows[b] = (int32_t)re; return; }
    const int64_t s_lo = stage_ptr[rb], s_hi = stage_ptr[re];
    const int64_t target = s_lo + ((s_hi - s_lo) * b) / grid;
    int64_t lo = rb, hi = re;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (stage_ptr[mid] < target) lo = mid + 1; else hi = mid;
    }
    cta_rows[b] = (int32_t)lo;
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
        cudaGetLastError();
    }
    return fn;
}

int als_tc_supported(const fms_ctx *ctx, int rank) {
    constexpr size_t need = (size_t)(tc::Smem2<6, 6, 4>::total > tc::Smem::total ? tc::Smem2<6, 6, 4>::total : tc::Smem::total) + 1024;
    return rank == tc::D && ctx->smem_optin >= need && get_encode_fn() != nullptr;
}

// planes of side s: [hi: rows x 64 halves][lo: rows x 64 halves], rows = padded rows + 8 (row `padded` stays zero)
static int64_t plane_rows(const AlsState &A, int s) { return A.shard_rows[s] * A.shard_count + 8; }

static int ensure_planes(fms_ctx *ctx, int s) {
    AlsState &A = ctx->als;
    const size_t need = (size_t)plane_rows(A, s) * tc::D * 2;
    if (A.plane[s].n < need || !A.plane[s].p || A.plane_rows[s] != plane_rows(A, s)) {
        als_tc_close_peers(ctx, s);                              // exported handles / peer mappings die with the old buffer
        FMS_CUDA(ctx, A.plane[s].alloc(need));
        FMS_CUDA(ctx, cudaMemsetAsync(A.plane[s].p, 0, A.plane[s].n * sizeof(uint16_t), ctx->stream));
        A.plane_rows[s] = plane_rows(A, s);
        A.plane_valid[s] = false;
    }
    return FMS_OK;
}

int als_tc_refresh_planes(fms_ctx *ctx, int s) {
    AlsState &A = ctx->als;
    FMS_TRY(ensure_planes(ctx, s));
    const int64_t n = (A.shard_rows[s] * A.shard_count) * tc::D;
    if (n > 0) {
        __half *hi = reinterpret_cast<__half *>(A.plane[s].p);
        tc::k_split_planes<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(A.fac[s].p, n, hi);
        FMS_LAUNCH_CHECK(ctx);
    }
    A.plane_valid[s] = true;
    return FMS_OK;
}

static int make_map(fms_ctx *ctx, CUtensorMap *m, void *base, int64_t rows) {
    const cuuint64_t gdim[2] = {(cuuint64_t)tc::PLANE_PITCH, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)tc::PLANE_PITCH * 2};
    const cuuint32_t box[2] = {(cuuint32_t)tc::D, 1};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, FMS_ERR_CUDA, "cuTensorMapEncodeTiled failed: %d", (int)r);
    return FMS_OK;
}

static int ensure_stage_csr(fms_ctx *ctx, int which, int64_t rb, int64_t re, int zero_row, int grid, int krows) {
    AlsState &A = ctx->als;
    const int64_t n = which == 0 ? A.n_users : A.n_items;
    cudaStream_t st = ctx->stream;
    const bool refill = A.stage_zero_row[which] != zero_row || A.stage_rb[which] < 0 || A.stage_krows[which] != krows;
    if (refill) {
        DevBuf<int32_t> cnt;
        DevBuf<unsigned char> temp;
        FMS_CUDA(ctx, cnt.alloc((size_t)n + 1));
        FMS_CUDA(ctx, A.stage_ptr[which].alloc((size_t)n + 1));
        tc::k_stage_counts<<<(unsigned)((n + 256) / 256), 256, 0, st>>>(A.ptr[which].p, n, cnt.p, krows);
        FMS_LAUNCH_CHECK(ctx);
        size_t tb = 0;
        FMS_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, A.stage_ptr[which].p, n + 1, st));
        FMS_CUDA(ctx, temp.alloc(tb));
        FMS_CUDA(ctx, cub::DeviceScan::ExclusiveSum(temp.p, tb, cnt.p, A.stage_ptr[which].p, n + 1, st));
        ctx->launches += 1;
        int32_t total = 0;
        FMS_CUDA(ctx, cudaMemcpyAsync(&total, A.stage_ptr[which].p + n, sizeof total, cudaMemcpyDeviceToHost, st));
        FMS_CUDA(ctx, cudaStreamSynchronize(st));
        A.total_stages[which] = total;
        FMS_CUDA(ctx, A.colp[which].alloc((size_t)total * krows));
        FMS_CUDA(ctx, A.valp[which].alloc((size_t)total * krows));
        if (n > 0) {
            tc::k_fill_stages<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(A.ptr[which].p, A.col[which].p, A.val[which].p,
                                                                              A.stage_ptr[which].p, n, zero_row, A.colp[which].p,
                                                                              A.valp[which].p, krows);
            FMS_LAUNCH_CHECK(ctx);
        }
        A.stage_zero_row[which] = zero_row;
        A.stage_krows[which] = krows;
        A.stage_rb[which] = -2;                                    // force the partition below
    }
    if (A.stage_rb[which] != rb || A.stage_re[which] != re || A.stage_grid[which] != grid) {
        FMS_CUDA(ctx, A.cta_rows[which].alloc((size_t)grid + 1));
        tc::k_cta_partition<<<(grid + 1 + 127) / 128, 128, 0, st>>>(A.stage_ptr[which].p, rb, re, grid, A.cta_rows[which].p);
        FMS_LAUNCH_CHECK(ctx);
        A.stage_rb[which] = rb; A.stage_re[which] = re; A.stage_grid[which] = grid;
    }
    return FMS_OK;
}

static int launch_tc(fms_ctx *ctx, int which, float lambda, int64_t rb, int64_t re, float *dump) {
    AlsState &A = ctx->als;
    if (re <= rb) return FMS_OK;
    const int partner = 1 - which;
    FMS_TRY(ensure_planes(ctx, which));
    FMS_TRY(ensure_planes(ctx, partner));
    if (!A.plane_valid[partner]) FMS_TRY(als_tc_refresh_planes(ctx, partner));
    const int64_t prow = plane_rows(A, partner);
    __half *p_hi = reinterpret_cast<__half *>(A.plane[partner].p), *t_hi = reinterpret_cast<__half *>(A.plane[which].p);
    CUtensorMap m_pl;
    FMS_TRY(make_map(ctx, &m_pl, p_hi, prow));
    const int zero_row = (int)(A.shard_rows[partner] * A.shard_count);
    __half *const *peers = A.peer_n[which] > 0 ? reinterpret_cast<__half *const *>(A.peer_tab[which].p) : nullptr;
    const int n_peers = A.peer_n[which];
    const int64_t rows = re - rb;
    const int grid = (int)(rows < ctx->sm_count ? rows : ctx->sm_count);
    // producer flavour: LDGSTS for short rows (see the kernel), TMA gather4 for long ones
    const int64_t n_side = which == 0 ? A.n_users : A.n_items;
    const bool short_rows = n_side > 0 && A.nnz / n_side < 16 * tc::KROWS;
    static const int force_kernel = getenv("FMS_ALS_KERNEL") ? atoi(getenv("FMS_ALS_KERNEL")) : 0;     // experiments: 3 = TMA, 4 = LDGSTS producers
    const int kern = force_kernel ? force_kernel : A.kernel;
    const bool ldgsts = !dump && (kern == FMS_ALS_KERNEL_TC_LDGSTS || (kern != FMS_ALS_KERNEL_TC_TMA && short_rows));
    static const bool use_v1 = getenv("FMS_ALS_TC_V1") != nullptr;             // A/B switch: the round-1 kernel (CUDA-core solver)
    if (use_v1) {
        FMS_TRY(ensure_stage_csr(ctx, which, rb, re, zero_row, grid, tc::KROWS));
        const size_t smem = (size_t)tc::Smem::total + 1024;
#define FMS_TC_LAUNCH(DUMPF, LDG, DUMP_PTR)                                                                                          \
    do {                                                                                                                             \
        FMS_CUDA(ctx, cudaFuncSetAttribute(tc::k_als_rows_tc64<DUMPF, LDG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        tc::k_als_rows_tc64<DUMPF, LDG><<<grid, tc::NWARPS * 32, smem, ctx->stream>>>(                                                \
            m_pl, A.ptr[which].p, A.stage_ptr[which].p, A.colp[which].p, A.valp[which].p, A.cta_rows[which].p, A.fac[which].p, t_hi,  \
            lambda, rb, A.info.p, DUMP_PTR, peers, n_peers, p_hi);                                                                   \
    } while (0)
        if (dump) FMS_TC_LAUNCH(true, false, dump);
        else if (ldgsts) FMS_TC_LAUNCH(false, true, nullptr);
        else FMS_TC_LAUNCH(false, false, nullptr);
#undef FMS_TC_LAUNCH
    } else {
        // v2: ring slots of KS x 16 ratings (one producer / issuer handshake per slot), ring depth and active solver warps per
        // row-length regime: many short systems -> 16 ratings per slot, 7 slots, 3 producers, 16 solver warps; few long ones ->
        // 4 x 16 ratings per slot, 6 slots (144 KB of gathers in flight, 6 of the 7 producer warps active), 6 solver warps
        static const int dbg = getenv("FMS_TC_DBG") ? atoi(getenv("FMS_TC_DBG")) : 0;   // timing experiments (wrong results): see the kernel
        // (measured on cfg2: slots of 32 ratings raise the short-row half-step's no-solver floor from 8.8 to 10.7 ms -- every
        // system pays the padding to a multiple of the slot -- while slots of 64 cut the long-row half-step from 5.8 to 4.3 ms)
        const int ks = short_rows ? 1 : 4;
        FMS_TRY(ensure_stage_csr(ctx, which, rb, re, zero_row, grid, tc::KROWS * ks));
#define FMS_TC2_LAUNCH(DUMPF, LDG, NS, NA, KSV, NPV, NSWV, DUMP_PTR)                                                                    \
    do {                                                                                                                             \
        const size_t smem = (size_t)tc::Smem2<NS, NA, KSV>::total + 1024;                                                             \
        FMS_CUDA(ctx, cudaFuncSetAttribute(tc::k_als_rows_tc64_v2<DUMPF, LDG, NS, NA, KSV, NPV, NSWV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        tc::k_als_rows_tc64_v2<DUMPF, LDG, NS, NA, KSV, NPV, NSWV><<<grid, tc::V2_WARPS * 32, smem, ctx->stream>>>(                    \
            m_pl, A.ptr[which].p, A.stage_ptr[which].p, A.colp[which].p, A.valp[which].p, A.cta_rows[which].p, A.fac[which].p, t_hi,  \
            lambda, rb, A.info.p, DUMP_PTR, peers, n_peers, p_hi, dbg);                                                              \
    } while (0)
        //                                  DUMP   LDGSTS slots bufs KS NP NSW
        if (dump && short_rows) FMS_TC2_LAUNCH(true, false, 7, 16, 1, 3, 16, dump);
        else if (dump) FMS_TC2_LAUNCH(true, false, 6, 6, 4, 6, 12, dump);
        else if (short_rows && ldgsts) FMS_TC2_LAUNCH(false, true, 7, 16, 1, 3, 16, nullptr);
        else if (short_rows) FMS_TC2_LAUNCH(false, false, 7, 16, 1, 3, 16, nullptr);
        else if (ldgsts) FMS_TC2_LAUNCH(false, true, 6, 6, 4, 6, 12, nullptr);
        else FMS_TC2_LAUNCH(false, false, 6, 6, 4, 6, 12, nullptr);
#undef FMS_TC2_LAUNCH
    }
    FMS_LAUNCH_CHECK(ctx);
    // the kernel wrote the planes of its own rows; with one shard that is the whole side
    A.plane_valid[which] = (A.shard_count == 1);
    return FMS_OK;
}

int als_launch_half_step_tc(fms_ctx *ctx, int which, float lambda, int64_t rb, int64_t re) {
    return launch_tc(ctx, which, lambda, rb, re, nullptr);
}

// device views of the gather planes of side `s` (allocated on demand) for plane-level exchange between ranks
int als_tc_planes(fms_ctx *ctx, int s, void **hi, void **lo, int64_t *rows_padded) {
    AlsState &A = ctx->als;
    FMS_TRY(ensure_planes(ctx, s));
    __half *h = reinterpret_cast<__half *>(A.plane[s].p);
    if (hi) *hi = h;
    if (lo) *lo = h + tc::D;                                   // interleaved: row pitch 2 * rank halves, lo half follows hi
    if (rows_padded) *rows_padded = A.shard_rows[s] * A.shard_count;
    return FMS_OK;
}

// ---- peer-to-peer plane exchange (one process per GPU on one node): CUDA IPC handles in, device table of replicas out
void als_tc_close_peers(fms_ctx *ctx, int s) {
    AlsState &A = ctx->als;
    for (int r = 0; r < A.peer_n[s]; ++r) {
        if (A.peer_open[s][r]) cudaIpcCloseMemHandle(A.peer_open[s][r]);
        A.peer_open[s][r] = nullptr;
    }
    A.peer_n[s] = 0;
}

int als_tc_ipc_handle(fms_ctx *ctx, int s, void *handle64) {
    AlsState &A = ctx->als;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    FMS_TRY(ensure_planes(ctx, s));
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));                     // the zero fill of a fresh buffer is done
    cudaIpcMemHandle_t h;
    FMS_CUDA(ctx, cudaIpcGetMemHandle(&h, A.plane[s].p));
    memcpy(handle64, &h, sizeof h);
    return FMS_OK;
}

int als_tc_open_peers(fms_ctx *ctx, int s, const void *handles, int n) {
    AlsState &A = ctx->als;
    if (n != A.shard_count || n > 64) return fail(ctx, FMS_ERR_INVALID_ARG, "peer table: %d handles for %d shards", n, A.shard_count);
    FMS_TRY(ensure_planes(ctx, s));
    als_tc_close_peers(ctx, s);
    unsigned long long tab[64];
    for (int r = 0; r < n; ++r) {
        if (r == A.shard_index) { tab[r] = (unsigned long long)(uintptr_t)A.plane[s].p; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + (size_t)r * sizeof h, sizeof h);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            A.peer_n[s] = r;                                               // close what was opened so far
            als_tc_close_peers(ctx, s);
            return fail(ctx, FMS_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        }
        A.peer_open[s][r] = p;
        tab[r] = (unsigned long long)(uintptr_t)p;
    }
    A.peer_n[s] = n;
    FMS_CUDA(ctx, A.peer_tab[s].alloc(64));
    FMS_CUDA(ctx, cudaMemcpyAsync(A.peer_tab[s].p, tab, sizeof(unsigned long long) * n, cudaMemcpyHostToDevice, ctx->stream));
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FMS_OK;
}

// replicas inside ONE process (group ctx): the peers' plane buffers are ordinary device pointers (peer access enabled)
int als_tc_set_peers_direct(fms_ctx *ctx, int s, void *const *ptrs, int n) {
    AlsState &A = ctx->als;
    if (n != A.shard_count || n > 64) return fail(ctx, FMS_ERR_INVALID_ARG, "peer table: %d pointers for %d shards", n, A.shard_count);
    FMS_TRY(ensure_planes(ctx, s));
    als_tc_close_peers(ctx, s);
    unsigned long long tab[64];
    for (int r = 0; r < n; ++r) tab[r] = (unsigned long long)(uintptr_t)(r == A.shard_index ? (void *)A.plane[s].p : ptrs[r]);
    FMS_CUDA(ctx, A.peer_tab[s].alloc(64));
    FMS_CUDA(ctx, cudaMemcpyAsync(A.peer_tab[s].p, tab, sizeof(unsigned long long) * n, cudaMemcpyHostToDevice, ctx->stream));
    FMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    A.peer_n[s] = n;
    return FMS_OK;
}

// test hook: system matrix + rhs of rows [rb, re) exactly as the solver warps receive them (row-major 64x64 + 64 floats each)
int als_tc_dump(fms_ctx *ctx, int which, float lambda, int64_t rb, int64_t re, float *d_dump) {
    return launch_tc(ctx, which, lambda, rb, re, d_dump);
}

}  // namespace fms
