// als_tc.cu -- K1 on the 5th-generation tensor cores (rank 64): warp-specialised persistent kernel, one CTA per SM.
//
// Data layout in HBM: besides the fp32 factor matrix every side keeps its fp16 "gather planes" hi = f16(x) and
// lo = f16((x - hi) * 2^11), interleaved as one 256-byte row per factor row (64 hi halves, then 64 lo halves).
// hi + lo * 2^-11 carries ~22 mantissa bits, so the Gramian  hi hi^T + 2^-11 (hi lo^T + lo hi^T)  is fp32-grade
// (plain fp16 / tf32 would be 2^-11: not enough for the 1e-4 parity target) while every product is exact in the fp32
// accumulator.  The ratings are pre-laid as 16-entry stages (rows padded with a zero plane row).
//
// 24 warps (role counts are template parameters of the kernel), register budgets set per role with setmaxnreg:
//   producers (3-6): stream the stage layout and bring the 16 partner rows of a stage (hi half and lo half of the plane row)
//                    straight into the UMMA MN-major / 128-byte-swizzle stage, either by TMA gather4 from one elected lane (no
//                    registers, few instructions: the default beside 16 busy solver warps, i.e. for short rows) or by 16-byte
//                    LDGSTS copies (the default for long rows); results are bitwise equal.  The ratings go into two 16-byte column
//                    groups R, R' of a third operand chunk.  A ring SLOT is KS stages of 16 ratings (6 KB each): one
//                    mbarrier full / empty handshake per slot.
//   MMA warp       : per stage two tcgen05.mma (kind::f16, M=64, K=16):
//                        D[:,   0:136] (+)= hi [hi | lo | R ]^T        (N = 136)
//                        D[:, 64:136]   += lo [hi |      R']^T        (N = 72, leading-dim offset skips `lo`)
//                    => cols 0-63 hi hi^T, 64-127 hi lo^T + lo hi^T, col 128 sum r_hi*hi, col 129 the 2^-11 terms of
//                    b = Y^T r.  Accumulators live in TMEM (2 slots x 256 columns).
//   epilogue (4)   : tcgen05.ld.16x256b -- the register layout of an mma.sync accumulator -- one m-tile of 16 rows each;
//                    -(D1 + 2^-11 D2 + lambda n I) goes out as 16x8 fragment tiles (chol_mma.cuh), -b as a vector, into
//                    the shared-memory buffer of a solver warp -- both scaled by powers of two (matrix: first diagonal entry
//                    -> [1, 2); right-hand side: its own magnitude -> ~8) so that fp16 hi/lo operands keep 22 bits
//   solvers (12-16): one system per warp on the warp-level tensor path: blocked right-looking Cholesky, 8x8 diagonal
//                    blocks factored redundantly in registers, triangular solves and trailing updates as fp16-packed
//                    HMMA.16816 products (hi and lo parts along K: 2 instructions per full-precision product) on the
//                    fragment tiles, forward substitution as one more matrix row, panel-wise back substitution
//                    (cm::solve_tiles_h); writes x (fp32) and its plane row -- into every rank's replica when peers are
//                    mapped (NVLink stores).
// Round 1's solver (one matrix row per lane, FFMA2 against broadcast LDS.128) ran the short-row half-step at 94 % of the
// shared-memory data pipe; profiles/README.md has the numbers that led here.
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
constexpr int NACC = 2;                           // TMEM accumulator slots
constexpr uint32_t ACC_COLS = 256u;               // TMEM columns per slot (136 used)
// MMA widths: the first covers hi | lo | 8 rating columns, the second hi | 8 rating columns.  Only two rating columns
// carry data; N = 136 / 72 instead of three / two full 64-wide chunks cuts a third of the tensor-pipe time per stage.
constexpr uint32_t MMA_N1 = 136u, MMA_N2 = 72u;
constexpr int N_EPI = 4;
// gather planes: one 256-byte row per factor row, hi halves then lo halves (one DRAM page burst serves both gathers)
constexpr int PLANE_PITCH = 2 * D;
constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;

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

// One 16-rating stage of a ring slot, issued by one elected lane without a divergent branch.  commit_empty: last sub-stage of the slot (releases it to the producers);
// last: last sub-stage of the row (hands the accumulator to the epilogue); n_mma: 2 (0 / 1 only in timing experiments).
__device__ __forceinline__ void tc_stage_mma(uint32_t d, uint32_t lo_hi, uint32_t lo_lo, uint32_t lo_b2, uint32_t idesc_n1, uint32_t idesc_n2,
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

__device__ __forceinline__ void split_half(float x, __half &hi, __half &lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn((x - __half2float(hi)) * LO_SCALE);
}

// =====================================================================================================================
// The kernel.  Template parameters: DUMP (test hook), LDGSTS (producer flavour), NSTG_ ring slots of KS stages, NSA solver
// buffers in use, NP producer warps, NSW solver warps.
// 24 warps: epilogue 0-3 | solvers 4-15 | MMA 16 | producers 17-23.
// Warp roles are a template parameter set: NSW solver warps after the 4 epilogue warps, then the MMA issuer, then NP producers
// (the rest of the last warpgroup idles).  Register budgets (setmaxnreg, per warpgroup) must fit the launch allocation 768 x 80:
//   long rows : 12 solver warps x 104 + 4 x 72 + 8 x 40   (6 solver buffers in use: few systems, many stages)
//   short rows: 16 solver warps x  96 + 4 x 56 + 4 x 40   (the solvers pace this half-step: every warp that fits is one)
constexpr int V2_WARPS = 24, V2_W_SOLVE0 = 4;
constexpr int V2_REGS_AUX = 40;
using TL4 = cm::Tiles<4>;
constexpr int V2_BUF_FLOATS = TL4::FLOATS + 3 * D;                          // tiles | bz | invd | xs  per solver warp

// Cross-rank ordering of the half-steps when the solver warps store into every replica (one process per GPU, planes mapped over
// CUDA IPC): no host-side collective.  The LAST CTA of a half-step on side s publishes epoch e into slot [my_rank] of the flag
// words kept in the spare rows of EVERY replica's side-s plane buffer (system-scope release after all of the rank's remote
// stores); a half-step that reads side s first waits until all `n` slots of its LOCAL flag words reached the epoch it expects.
// Seeing every rank's flag also means every rank FINISHED that half-step, so the replicas may be overwritten by the next one.
struct FlagSync {
    int *ctr;                  // per-ctx counter of finished CTAs (last-block detection)
    const int *wait_flags;     // local flag words of the side being READ (partner), or nullptr: nothing to wait for
    int wait_epoch;
    int64_t signal_off;        // offset (in halves) of the flag words inside the plane buffer of the side being WRITTEN
    int signal_epoch, my_rank, n;
};
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

template <int NSTG_, int NSA, int KS>
struct Smem2 {
    static constexpr int ring = 0;
    static constexpr int bufs = ring + NSTG_ * KS * STAGE_BYTES;
    static constexpr int bars = bufs + NSA * V2_BUF_FLOATS * 4;
    static constexpr int n_bars = 2 * NSTG_ + 2 * NACC + 2 * NSA;
    static constexpr int tmem_ptr = bars + n_bars * 8;
    static constexpr int scale = tmem_ptr + 16;                // 2 x 3 floats: power-of-two scales of the system being unpacked
    static constexpr int total = tmem_ptr + 48;
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
__global__ void __launch_bounds__(V2_WARPS * 32, 1) k_als_rows_tc64(
    const __grid_constant__ CUtensorMap map_pl, const int64_t *__restrict__ ptr,
    const int32_t *__restrict__ stage_ptr, const int32_t *__restrict__ colp, const float *__restrict__ valp,
    const int32_t *__restrict__ cta_rows, float *__restrict__ T, __half *__restrict__ T_pl, float lambda,
    int64_t row_begin, int *__restrict__ info, float *__restrict__ dump, __half *const *__restrict__ peers, int n_peers,
    const __half *__restrict__ S_pl, int dbg, FlagSync fs, int n_act) {
    using SM = Smem2<NSTG_, NSA, KS>;
    constexpr int V2_W_MMA = V2_W_SOLVE0 + NSW, V2_W_PROD0 = V2_W_MMA + 1;
    constexpr int V2_REGS_EPI = NSW > 12 ? 56 : 72, V2_REGS_SOLVE = NSW > 12 ? 96 : 104;
    static_assert(NSW % 4 == 0 && V2_W_PROD0 + NP <= V2_WARPS, "warp roles");
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
    volatile float *scale_slot = reinterpret_cast<volatile float *>(sm + SM::scale);

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
    if (fs.wait_flags != nullptr && threadIdx.x == 0) {
        // every rank's rows of the side this half-step gathers from have landed in OUR replica (and every rank is done with it)
        const long long t0 = clock64();
        for (int r = 0; r < fs.n; ++r)
            while (ld_acquire_sys(fs.wait_flags + r) < fs.wait_epoch) {
                if (clock64() - t0 > 8000000000LL) { atomicAdd(info, 1 << 20); break; }      // ~4 s: a peer is gone; do not hang the GPU
            }
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
                    tc_stage_mma(d, e | LBO_T, (e + (TILE_BYTES >> 4)) | LBO_T, e | LBO_B2, idesc1, idesc2, (st > 0 || ks > 0) ? 1u : 0u,
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
            // fp16-packed solver (cm::solve_tiles_h): the matrix goes out scaled by the power of two that brings its first diagonal
            // entry into [1, 2); the right-hand side by that and a second power of two that brings its own magnitude to ~8 (x is
            // linear in b: the solver multiplies the inverse back in) -- an fp16 lo part resolves 6e-8 ABSOLUTE, so both operands
            // must sit near 1.  Warp 0 owns the entries; one named barrier among the four epilogue warps per system hands the
            // factors over (two slot sets: a slow warp still reads the previous one).
            float sc = 1.f, sb = 1.f;
            if (FMS_CHOL_F16 && !DUMP) {
                if (q == 0) {
                    float e1[4], e2[4], eb[4];
                    tc_ld16x256_x1(tmem_base + (uint32_t)acc * ACC_COLS, e1);
                    tc_ld16x256_x1(tmem_base + (uint32_t)acc * ACC_COLS + 64, e2);
                    tc_ld16x256_x1(tmem_base + (uint32_t)acc * ACC_COLS + 128, eb);
                    tc_wait_ld();
                    // largest |b| among the first 16 components (the lanes t = 0 of this warp hold them)
                    float m = t4 == 0 ? fmaxf(fabsf(fmaf(eb[1], LO_INV, eb[0])), fabsf(fmaf(eb[3], LO_INV, eb[2]))) : 0.f;
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
                    if (lane == 0) {
                        const uint32_t ex = (__float_as_uint(fmaf(LO_INV, e2[0], e1[0]) + reg) >> 23) & 0x1ffu;      // sign + exponent
                        const float s_a = (ex >= 1u && ex <= 253u) ? __uint_as_float((254u - ex) << 23) : 1.f;       // first diagonal entry -> [1, 2)
                        const uint32_t eb_ = (__float_as_uint(m * s_a) >> 23) & 0xffu;
                        const bool okb = eb_ >= 8u && eb_ <= 250u;
                        scale_slot[3 * (job & 1) + 0] = s_a;
                        scale_slot[3 * (job & 1) + 1] = okb ? __uint_as_float((257u - eb_) << 23) : 1.f;             // largest of those |b| -> [8, 16)
                        scale_slot[3 * (job & 1) + 2] = okb ? __uint_as_float((eb_ - 3u) << 23) : 1.f;               // its inverse, for x
                    }
                }
                asm volatile("bar.sync 2, 128;" ::: "memory");
                sc = scale_slot[3 * (job & 1)];
                sb = sc * scale_slot[3 * (job & 1) + 1];
            }
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
                        if (FMS_CHOL_F16 && !DUMP) { v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; }
                        tiles[32 * j] = v;
                    }
                }
            }
            float bb[4];
            tc_ld16x256_x1(tq + 128, bb);                          // column 128: sum r_hi y, column 129: the 2^-11 terms
            tc_wait_ld();
            if (t4 == 0) {
                float *bz = buf + TL4::FLOATS;
                bz[16 * q + g8] = -fmaf(bb[1], LO_INV, bb[0]) * sb;
                bz[16 * q + g8 + 8] = -fmaf(bb[3], LO_INV, bb[2]) * sb;
                if (q == 0 && g8 == 0 && FMS_CHOL_F16 && !DUMP) bz[D] = scale_slot[3 * (job & 1) + 2];   // 1 / (rhs-only scale), first scratch word
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar_acc_empty(acc)); mbar_arrive(bar_g_full(g)); }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(V2_REGS_SOLVE));
        // ================================================================== solvers: one system per warp on the HMMA path
        // n_act of the NSW solver warps take systems round robin (system `job` sits in buffer job % NSA).  Measured on cfg1's short-row
        // half-step (16 buffers): 16 / 15 / 14 / 13 active solvers -> 11.3 / 11.7 / 12.3 / 13.1 ms, and 17 or 18 buffers for 16
        // solvers (shallower ring) change nothing: the solvers, not the refill of their buffers, pace this half-step.
        const int sv = warp - V2_W_SOLVE0;
        if (sv < n_act) {
            for (uint32_t job = (uint32_t)sv; first + job < last; job += (uint32_t)n_act) {
                const int64_t row = first + job;
                const int gb = (int)(job % NSA);
                float *S = reinterpret_cast<float *>(sm + SM::bufs) + (size_t)gb * V2_BUF_FLOATS;
                float *bz = S + TL4::FLOATS, *invd = bz + D, *xs = invd + D;
                mbar_wait(bar_g_full(gb), (job / NSA) & 1);
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
                if (!(dbg & 16)) {
                    if (FMS_CHOL_F16 && !DUMP) {
                        const float inv_sb = invd[0];
                        cm::solve_tiles_h<4>(S, bz, invd, xs, lane, x);
                        x[0] *= inv_sb; x[1] *= inv_sb;
                    } else {
                        cm::solve_tiles<4>(S, bz, invd, xs, lane, x);
                    }
                }
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
                if (lane == 0) mbar_arrive(bar_g_empty(gb));          // tiles / scratch no longer needed
                T[row * D + lane] = b0;
                T[row * D + lane + 32] = b1;
                const int64_t poff = row * PLANE_PITCH + 4 * lane;
                if (n_peers == 0) {
                    *reinterpret_cast<uint2 *>(T_pl + poff) = prow;
                } else {
                    for (int p = 0; p < n_peers; ++p) *reinterpret_cast<uint2 *>(peers[p] + poff) = prow;
                }
                // NaN / inf: the factorisation broke down; beyond +-65504 the fp16 gather planes cannot carry the value to the next
                // half-step (the fp32 row in T is still exact): both are counted, fms_als_fit / fms_als_last_info report the count
                const bool bad = !(fabsf(b0) <= 65504.f) || !(fabsf(b1) <= 65504.f);
                if (__any_sync(0xffffffffu, bad) && lane == 0) atomicAdd(info, 1);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == V2_W_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
    if (fs.ctr != nullptr && threadIdx.x == 0) {
        __threadfence_system();                                    // this CTA's remote plane stores before its count
        if (atomicAdd(fs.ctr, 1) == (int)gridDim.x - 1) {          // last CTA of this rank's half-step
            *fs.ctr = 0;
            __threadfence_system();
            for (int p = 0; p < fs.n; ++p)
                st_release_sys(reinterpret_cast<int *>(peers[p] + fs.signal_off) + fs.my_rank, fs.signal_epoch);
        }
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
    constexpr size_t need = (size_t)(tc::Smem2<6, 6, 4>::total > tc::Smem2<7, 16, 1>::total ? tc::Smem2<6, 6, 4>::total : tc::Smem2<7, 16, 1>::total) + 1024;
    return rank == tc::D && ctx->smem_optin >= need && get_encode_fn() != nullptr;
}

// planes of side s: [hi: rows x 64 halves][lo: rows x 64 halves], rows = padded rows + 8 (row `padded` stays zero)
static int64_t plane_rows(const AlsState &A, int s) { return A.shard_rows[s] * A.shard_count + 8; }
// flag words of the in-kernel cross-rank ordering: spare row `padded + 4` of the plane buffer (row `padded` is the zero row the
// stage padding gathers; 64 ints = one slot per rank)
static int64_t flag_off_halves(const AlsState &A, int s) { return (A.shard_rows[s] * A.shard_count + 4) * (int64_t)tc::PLANE_PITCH; }

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
    // producer flavour per half-step (results are bitwise equal either way)
    const int64_t n_side = which == 0 ? A.n_users : A.n_items;
    const bool short_rows = n_side > 0 && A.nnz / n_side < 16 * tc::KROWS;
    static const int force_kernel = getenv("FMS_ALS_KERNEL") ? atoi(getenv("FMS_ALS_KERNEL")) : 0;     // experiments: 3 = TMA, 4 = LDGSTS producers
    const int kern = force_kernel ? force_kernel : A.kernel;
    // round-2 A/B with the HMMA solver (profiles/README.md): TMA gather4 now wins for short rows (users 11.21 vs 11.31 ms: fewer
    // producer instructions beside 16 busy solver warps), LDGSTS for long rows (items 4.07 vs 4.16 ms) -- the reverse of round 1
    const bool ldgsts = !dump && (kern == FMS_ALS_KERNEL_TC_LDGSTS || (kern != FMS_ALS_KERNEL_TC_TMA && !short_rows));
    {
        // ring slots of KS x 16 ratings (one producer / issuer handshake per slot), ring depth and active solver warps per
        // row-length regime: many short systems -> 16 ratings per slot, 7 slots, 3 producers, 16 solver warps; few long ones ->
        // 4 x 16 ratings per slot, 6 slots (144 KB of gathers in flight, 6 of the 7 producer warps active), 6 solver warps
        static const int nact_env = getenv("FMS_TC_NACT") ? atoi(getenv("FMS_TC_NACT")) : 0;   // experiments: active solver warps
        static const int dbg = getenv("FMS_TC_DBG") ? atoi(getenv("FMS_TC_DBG")) : 0;   // timing experiments (wrong results): see the kernel
        // in-kernel cross-rank ordering (flag words in the spare rows of the plane buffers), only with mapped replicas of BOTH sides
        tc::FlagSync fs = {nullptr, nullptr, 0, 0, 0, 0, 0};
        if (n_peers > 0 && A.peer_n[partner] == n_peers && A.flag_sync) {
            FMS_CUDA(ctx, A.sync_ctr.alloc(1));
            if (!A.sync_ctr_zeroed) { FMS_CUDA(ctx, cudaMemsetAsync(A.sync_ctr.p, 0, sizeof(int), ctx->stream)); A.sync_ctr_zeroed = true; }
            fs.ctr = A.sync_ctr.p;
            fs.wait_flags = reinterpret_cast<const int *>(p_hi + flag_off_halves(A, partner));
            fs.wait_epoch = A.plane_epoch[partner];
            fs.signal_off = flag_off_halves(A, which);
            fs.signal_epoch = ++A.plane_epoch[which];
            fs.my_rank = A.shard_index;
            fs.n = n_peers;
        }
        // (measured on cfg2: slots of 32 ratings raise the short-row half-step's no-solver floor from 8.8 to 10.7 ms -- every
        // system pays the padding to a multiple of the slot -- while slots of 64 cut the long-row half-step from 5.8 to 4.3 ms)
        const int ks = short_rows ? 1 : 4;
        FMS_TRY(ensure_stage_csr(ctx, which, rb, re, zero_row, grid, tc::KROWS * ks));
#define FMS_TC2_LAUNCH(DUMPF, LDG, NS, NA, KSV, NPV, NSWV, DUMP_PTR)                                                                    \
    do {                                                                                                                             \
        const size_t smem = (size_t)tc::Smem2<NS, NA, KSV>::total + 1024;                                                             \
        FMS_CUDA(ctx, cudaFuncSetAttribute(tc::k_als_rows_tc64<DUMPF, LDG, NS, NA, KSV, NPV, NSWV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        tc::k_als_rows_tc64<DUMPF, LDG, NS, NA, KSV, NPV, NSWV><<<grid, tc::V2_WARPS * 32, smem, ctx->stream>>>(                    \
            m_pl, A.ptr[which].p, A.stage_ptr[which].p, A.colp[which].p, A.valp[which].p, A.cta_rows[which].p, A.fac[which].p, t_hi,  \
            lambda, rb, A.info.p, DUMP_PTR, peers, n_peers, p_hi, dbg, fs, nact_env > 0 && nact_env <= NA && nact_env <= NSWV ? nact_env : (NA < NSWV ? NA : NSWV));                                                          \
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
    A.plane_epoch[s] = 0;
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
    // fresh flag words for the in-kernel ordering; the HOST must put one barrier across the ranks between this call and the
    // first half-step (a peer's first signal must not be overwritten by this memset)
    A.plane_epoch[s] = 0;
    FMS_CUDA(ctx, cudaMemsetAsync(reinterpret_cast<__half *>(A.plane[s].p) + flag_off_halves(A, s), 0, 64 * sizeof(int), ctx->stream));
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
