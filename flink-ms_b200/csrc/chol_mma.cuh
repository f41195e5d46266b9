// chol_mma.cuh -- one warp solves one SPD system (A + lambda n I) x = b of order D = 16 MT on the warp-level tensor path.
// Two product flavours with the same structure and the same accuracy (5e-6 worst case against an fp64 solve):
//   solve_tiles    mma.sync m16n8k8 TF32 (SASS HMMA.1688.F32.TF32), error-compensated 3xTF32: 3 instructions per product;
//   solve_tiles_h  mma.sync m16n8k16 f16 (SASS HMMA.16816.F32), hi and lo parts packed along K: 2 instructions per product,
//                  B fragments half the size; the caller scales the system by powers of two.  The ALS kernels use this one.
//
// Why: the round-1 solver (one matrix row per lane, trailing updates as FFMA2 against broadcast LDS.128) was bound by the
// shared-memory data pipe (ncu: l1tex data-pipe wavefronts 94 % of peak, 2 240 of them per 64x64 system) and by issue
// slots (7 000 warp instructions per system for ~750 useful FFMA2).  A broadcast LDS delivers two floats per wavefront; an
// MMA fragment holds its operands in registers, so the blocked factorisation below moves every tile through shared memory
// once per panel and issues ~3 HMMA where the old code issued 16 FFMA2 + 8 LDS.128.
//
// Data layout ("tile layout"): the NEGATED system S = -(A + lambda n I) is stored as the 16x8 tiles (I, j), j <= 2 I + 1,
// of its lower triangle (diagonal tiles complete), tile (I, j) at float offset 128 (I (I + 1) + j), each tile in
// mma-accumulator order: lane (g = lane / 4, t = lane % 4) owns the float4
//      ( S[16 I + g][8 j + 2 t], S[16 I + g][8 j + 2 t + 1], S[16 I + g + 8][8 j + 2 t], S[16 I + g + 8][8 j + 2 t + 1] ).
// That is exactly what tcgen05.ld.16x256b hands the epilogue warps (tools/micro/mma_sync_rate.cu prints the mapping), and an
// accumulator tile can be fed back as the A operand (rows) or the B operand (columns of a later tile) of the next MMA
// without moving data when the k slots are read as k = t -> column 2 t, k = t + 4 -> column 2 t + 1.
// The right-hand side travels as row D of the matrix (bz = -b): the panel loop then leaves z = L^-1 b in bz.
//
// Right-looking, panel width 8 (one n-tile):
//   1. the 8x8 diagonal block is read by every lane (broadcast LDS.64) and factored redundantly in registers;
//      every lane derives row g of W = Ld^-1, i.e. its B fragment of -W;
//   2. "trsm" as a product: L(I, J) = S(I, J) (-W)^T for every m-tile of the panel (3 HMMA each);
//   3. trailing update S(I, j) += L(I, J) L(j, J)^T for j > J (3 HMMA per tile, fragments straight from step 2);
//   4. the same two steps for the right-hand-side row.
// Back substitution is scalar: one column per lane and 32-column block, panel-wise x_J = W_J^T r_J.  (A variant on the HMMA path
// -- (1 x 8)(8 x 8) products, B fragments read transposed from the tiles -- cut the shared-memory wavefronts by 29 % and won 6 %
// stand-alone, but lost 3 % inside the ALS kernel, where the tensor datapath is shared with tcgen05: profiles/r02_backsub_variants.txt.)
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#ifndef FMS_CHOL_F16
#define FMS_CHOL_F16 1          // rank-64 kernel: 1 = fp16-packed HMMA.16816 products on a power-of-two scaled system, 0 = 3xTF32
#endif
#include <stdint.h>

namespace fms {
namespace cm {

// round to TF32 (10 explicit mantissa bits), ties away from zero: the same result as cvt.rna.tf32.f32 for finite inputs in two
// integer instructions (sm_100a emulates the cvt with FSETP + IMAD + LOP3; the tensor core ignores the low 13 bits anyway, the
// mask only keeps `x - hi` exact).  Inf / NaN stay non-finite.
__device__ __forceinline__ uint32_t tf32_rna(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
// x ~= hi + lo with hi, lo representable in TF32 (22 mantissa bits together)
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
    hi = tf32_rna(x);                                             // masked: x - hi below must be exact
    lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;      // + half an ulp; the tensor core's truncation completes the rounding
}
// A B fragment travels as ONE 64-bit value (b0 low, b1 high): ptxas then keeps it in the aligned register pair HMMA wants
// (as two separate 32-bit variables every use cost two MOVs: ~1 000 of the prototype's 7 300 instructions per system).
typedef unsigned long long BFrag;
__device__ __forceinline__ BFrag bfrag(uint32_t b0, uint32_t b1) {
    BFrag r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(b0), "r"(b1));
    return r;
}
__device__ __forceinline__ void mma_tf32(float4 &c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, BFrag b) {
    asm volatile(
        "{\n"
        ".reg .b32 b0, b1;\n"
        "mov.b64 {b0, b1}, %8;\n"
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {b0,b1}, {%0,%1,%2,%3};\n"
        "}\n"
        : "+f"(c.x), "+f"(c.y), "+f"(c.z), "+f"(c.w)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "l"(b));
}
// A operand from an accumulator-layout tile: rows g / g + 8, k slot t = column 2 t, k slot t + 4 = column 2 t + 1
struct AFrag {
    uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
    __device__ __forceinline__ void set(float x, float y, float z, float w) {   // (row g, col 2t), (g, 2t+1), (g+8, 2t), (g+8, 2t+1)
        split_tf32(x, h0, l0); split_tf32(z, h1, l1); split_tf32(y, h2, l2); split_tf32(w, h3, l3);
    }
};
// c += A B^T with 3xTF32: hi hi + hi lo + lo hi (the small terms first)
__device__ __forceinline__ void mma3(float4 &c, const AFrag &a, BFrag bh, BFrag bl) {
    mma_tf32(c, a.l0, a.l1, a.l2, a.l3, bh);
    mma_tf32(c, a.h0, a.h1, a.h2, a.h3, bl);
    mma_tf32(c, a.h0, a.h1, a.h2, a.h3, bh);
}

__device__ __forceinline__ float rsqrt_nr(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));      // bare MUFU.RSQ (rsqrtf() adds a denormal range fix-up)
    return r * fmaf(-0.5f * x * r, r, 1.5f);              // one Newton step: ~1 ulp; a non-positive pivot becomes NaN
}
__device__ __forceinline__ float sel4(int t, float a, float b, float c, float d) {
    const float lo = (t & 1) ? b : a, hi = (t & 1) ? d : c;
    return (t & 2) ? hi : lo;
}

template <int MT>
struct Tiles {
    static constexpr int D = 16 * MT, NT = 2 * MT;
    static constexpr int N_TILES = MT * (MT + 1);
    static constexpr int FLOATS = N_TILES * 128;
    __host__ __device__ static constexpr int off(int I, int j) { return (I * (I + 1) + j) * 128; }
    // float offset of element (k, i), k >= i
    __host__ __device__ static constexpr int at(int k, int i) {
        return off(k >> 4, i >> 3) + ((k & 7) * 4 + ((i & 7) >> 1)) * 4 + ((k >> 3) & 1) * 2 + (i & 1);
    }
};

// Solves one system in place.  S: tile layout (negated matrix) -- holds L afterwards.  bz: D floats, -b in, z out.
// invd: D floats scratch (unused by the solver since W replaces the diagonal blocks; kept for the caller's staging).  xs: D floats scratch (x, for the broadcasts of the back substitution).
// x_out[q] = x[lane + 32 q].  All four arrays are private to the calling warp; the caller provides the barrier that
// made S and bz visible (mbarrier wait or __syncwarp).
template <int MT>
__device__ __forceinline__ void solve_tiles(float *__restrict__ S, float *__restrict__ bz, float *__restrict__ invd, float *__restrict__ xs,
                                            const int lane, float (&x_out)[MT / 2]) {
    using TL = Tiles<MT>;
    constexpr int NT = TL::NT;
    const int g = lane >> 2, t = lane & 3;
    float4 *S4 = reinterpret_cast<float4 *>(S);
    BFrag bh[NT], bl[NT];                                   // B fragments (hi / lo) of the current panel's L, per n-tile
#pragma unroll
    for (int j = 0; j < NT; ++j) bh[j] = bl[j] = 0ull;

#pragma unroll 1
    for (int J = 0; J < NT; ++J) {
        const int Id = J >> 1, h = J & 1;
        // ---- 1. diagonal block, factored redundantly: a[r (r + 1) / 2 + c] = (A + lambda n I)[8 J + r][8 J + c], c <= r
        BFrag wb_h, wb_l;
        float2 w_own;
        {
            const float *Td = S + (Id * (Id + 1) + J) * 128 + 2 * h;
            float a[36];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int cp = 0; 2 * cp <= r; ++cp) {
                    const float2 v = *reinterpret_cast<const float2 *>(Td + (4 * r + cp) * 4);
                    a[r * (r + 1) / 2 + 2 * cp] = -v.x;
                    if (2 * cp + 1 <= r) a[r * (r + 1) / 2 + 2 * cp + 1] = -v.y;
                }
            float rinv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                rinv[k] = rsqrt_nr(a[k * (k + 1) / 2 + k]);
#pragma unroll
                for (int i = k + 1; i < 8; ++i) a[i * (i + 1) / 2 + k] *= rinv[k];
#pragma unroll
                for (int j = k + 1; j < 8; ++j)
#pragma unroll
                    for (int i = j; i < 8; ++i) a[i * (i + 1) / 2 + j] = fmaf(-a[i * (i + 1) / 2 + k], a[j * (j + 1) / 2 + k], a[i * (i + 1) / 2 + j]);
            }
            // row g of W = Ld^-1 (w Ld = e_g^T, solved from the right; entries right of the diagonal come out 0)
            float wv[8];
#pragma unroll
            for (int c = 7; c >= 0; --c) {
                float s = (c == g) ? 1.f : 0.f;
#pragma unroll
                for (int m = c + 1; m < 8; ++m) s = fmaf(-wv[m], a[m * (m + 1) / 2 + c], s);
                wv[c] = s * rinv[c];
            }
            // B fragment of -W: n = g (row of W), k slot t -> column 2 t, k slot t + 4 -> column 2 t + 1
            uint32_t h0, l0, h1, l1;
            w_own = make_float2(sel4(t, wv[0], wv[2], wv[4], wv[6]), sel4(t, wv[1], wv[3], wv[5], wv[7]));   // W[g][2t], W[g][2t+1]
            split_tf32(-w_own.x, h0, l0);
            split_tf32(-w_own.y, h1, l1);
            wb_h = bfrag(h0, h1);
            wb_l = bfrag(l0, l1);
        }
        // ---- 2 + 3. panel tiles: L(I, J) = S(I, J) (-W)^T, then S(I, j) += L(I, J) L(j, J)^T for J < j <= 2 I + 1
#pragma unroll
        for (int I = 0; I < MT; ++I) {
            if (I >= Id) {
                const float4 sv = S4[(I * (I + 1) + J) * 32 + lane];
                AFrag af;
                af.set(sv.x, sv.y, sv.z, sv.w);
                float4 L = make_float4(0.f, 0.f, 0.f, 0.f);
                mma3(L, af, wb_h, wb_l);
                if (I == Id) {
                    // rows of the diagonal block: keep the lower triangle (the product reproduces Ld up to rounding noise
                    // above the diagonal); rows of the m-tile above the block belong to finished panels: zero
                    const bool u0 = 2 * t > g, u1 = 2 * t + 1 > g;
                    if (h == 0) { L.x = u0 ? 0.f : L.x; L.y = u1 ? 0.f : L.y; }
                    else { L.x = 0.f; L.y = 0.f; L.z = u0 ? 0.f : L.z; L.w = u1 ? 0.f : L.w; }
                }
                S4[(I * (I + 1) + J) * 32 + lane] = L;
                af.set(L.x, L.y, L.z, L.w);
                bh[2 * I] = bfrag(af.h0, af.h2); bl[2 * I] = bfrag(af.l0, af.l2);                  // rows g     -> n-tile 2 I
                bh[2 * I + 1] = bfrag(af.h1, af.h3); bl[2 * I + 1] = bfrag(af.l1, af.l3);          // rows g + 8 -> n-tile 2 I + 1
                // descending j in PAIRS with an early exit (a real, uniform branch instead of predicated-off HMMA / LDS / STS):
                // the two tiles' 3-HMMA chains are independent, so interleaving them fills the ~30-cycle dependent-issue gaps
                // (one tile at a time, ptxas pads every HMMA of the chain with a NOP)
#pragma unroll
                for (int j = 2 * I + 1; j >= 1; j -= 2) {
                    if (j <= J) break;
                    const bool two = j - 1 > J;                         // the lower tile of the pair is live too
                    float4 c0 = S4[(I * (I + 1) + j) * 32 + lane];
                    float4 c1 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (two) c1 = S4[(I * (I + 1) + j - 1) * 32 + lane];
                    mma_tf32(c0, af.l0, af.l1, af.l2, af.l3, bh[j]);
                    mma_tf32(c1, af.l0, af.l1, af.l2, af.l3, bh[j - 1]);
                    mma_tf32(c0, af.h0, af.h1, af.h2, af.h3, bl[j]);
                    mma_tf32(c1, af.h0, af.h1, af.h2, af.h3, bl[j - 1]);
                    mma_tf32(c0, af.h0, af.h1, af.h2, af.h3, bh[j]);
                    mma_tf32(c1, af.h0, af.h1, af.h2, af.h3, bh[j - 1]);
                    S4[(I * (I + 1) + j) * 32 + lane] = c0;
                    if (two) S4[(I * (I + 1) + j - 1) * 32 + lane] = c1;
                }
            }
        }
        // the diagonal block of the factor is only ever used through its inverse again (back substitution): keep W = Ld^-1 there.
        // Lane (g, t) owns exactly the two entries W[g][2t], W[g][2t+1] of its fragment position.
        *reinterpret_cast<float2 *>(S + (Id * (Id + 1) + J) * 128 + lane * 4 + 2 * h) = w_own;
        // ---- 4. right-hand-side row (row D of the matrix): z_J = bz_J (-W)^T, bz_j += z_J L(j, J)^T
        {
            float2 bv = make_float2(0.f, 0.f);
            if (g == 0) bv = *reinterpret_cast<const float2 *>(bz + 8 * J + 2 * t);
            AFrag af;
            af.h1 = af.h3 = af.l1 = af.l3 = 0u;
            split_tf32(bv.x, af.h0, af.l0);
            split_tf32(bv.y, af.h2, af.l2);
            float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            mma3(z, af, wb_h, wb_l);
            if (g == 0) *reinterpret_cast<float2 *>(bz + 8 * J + 2 * t) = make_float2(z.x, z.y);
            split_tf32(z.x, af.h0, af.l0);
            split_tf32(z.y, af.h2, af.l2);
            // descending pairs like the trailing update (two independent 3-HMMA chains per branch)
#pragma unroll
            for (int j = NT - 1; j >= 1; j -= 2) {
                if (j <= J) break;
                const bool two = j - 1 > J;
                float2 cv0 = make_float2(0.f, 0.f), cv1 = make_float2(0.f, 0.f);
                if (g == 0) {
                    cv0 = *reinterpret_cast<const float2 *>(bz + 8 * j + 2 * t);
                    if (two) cv1 = *reinterpret_cast<const float2 *>(bz + 8 * j - 8 + 2 * t);
                }
                float4 c0 = make_float4(cv0.x, cv0.y, 0.f, 0.f), c1 = make_float4(cv1.x, cv1.y, 0.f, 0.f);
                mma_tf32(c0, af.l0, af.l1, af.l2, af.l3, bh[j]);
                mma_tf32(c1, af.l0, af.l1, af.l2, af.l3, bh[j - 1]);
                mma_tf32(c0, af.h0, af.h1, af.h2, af.h3, bl[j]);
                mma_tf32(c1, af.h0, af.h1, af.h2, af.h3, bl[j - 1]);
                mma_tf32(c0, af.h0, af.h1, af.h2, af.h3, bh[j]);
                mma_tf32(c1, af.h0, af.h1, af.h2, af.h3, bh[j - 1]);
                if (g == 0) {
                    *reinterpret_cast<float2 *>(bz + 8 * j + 2 * t) = make_float2(c0.x, c0.y);
                    if (two) *reinterpret_cast<float2 *>(bz + 8 * j - 8 + 2 * t) = make_float2(c1.x, c1.y);
                }
            }
        }
        __syncwarp();
    }

    // ---- back substitution L^T x = z: column i = lane + 32 q per lane, one panel (8 pivots) at a time, blocks of 32 columns
    // descending.  x_J = W_J^T r_J with W_J = Ld^-1 stored over the diagonal block: no sequential triangular solve.
    constexpr int NQ = MT / 2;
    float zq[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) zq[q] = bz[lane + 32 * q];
    const int lp = (lane >> 3) * 128 + ((lane & 7) >> 1) * 4 + (lane & 1);      // lane part of Tiles::at(k, 32 q + lane)
    const int sk = (lane >> 3) & 1, sk16 = 16 * sk;                             // row-order skew of the odd lane groups (bank conflicts)
#pragma unroll
    for (int q = NQ - 1; q >= 0; --q) {
        float x = 0.f;
#pragma unroll 1
        for (int jr = 24; jr >= 0; jr -= 8) {
            const int j0 = 32 * q + jr, J = j0 >> 3, I = J >> 1, hh = J & 1;
            const int rel = lane - jr;
            // pivot lane c = rel: x_c = sum_r W[r][c] r_r  (W lower triangular, zeros above the diagonal are stored)
            const float *wc = S + (I * (I + 1) + J) * 128 + 2 * hh + ((rel & 7) >> 1) * 4 + (rel & 1);
            float xc = 0.f;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float zr = __shfl_sync(0xffffffffu, zq[q], jr + r);
                xc = fmaf(wc[16 * r], zr, xc);
            }
            if (rel >= 0 && rel < 8) x = xc;
            // columns left of the panel: z_i -= sum_c L[j0 + c][i] x_c
            // rows j0 + c sit 16 floats apart, the four 8-column tiles of a lane group 128 floats: all four groups would hit the
            // same 8 banks (4-way conflict).  Odd groups walk the rows in the order 1, 0, 3, 2, ...: two bank sets, 2 wavefronts.
            const float *p = S + (I * (I + 1)) * 128 + 2 * hh + 4 * q * 128 + lp;
            float acc = zq[q];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float xv = __shfl_sync(0xffffffffu, xc, jr + (c ^ sk));
                acc = fmaf(-p[16 * c + ((c & 1) ? -sk16 : sk16)], xv, acc);
            }
            if (rel < 0) zq[q] = acc;
        }
        x_out[q] = x;
        if (q > 0) {
            xs[32 * q + lane] = x;
            __syncwarp();
            // earlier column blocks: z_i -= sum_{k in block q} L[k][i] x_k
#pragma unroll
            for (int q2 = 0; q2 < q; ++q2) {
                float acc = zq[q2];
#pragma unroll 1
                for (int k4 = 0; k4 < 32; k4 += 4) {
                    const int k = 32 * q + k4;
                    const int I = k >> 4;
                    const float *p = S + (I * (I + 1) + 4 * q2) * 128 + (k & 7) * 16 + ((k >> 3) & 1) * 2 + lp;
                    const float4 xv = *reinterpret_cast<const float4 *>(xs + k);
                    const float x0 = sk ? xv.y : xv.x, x1 = sk ? xv.x : xv.y, x2 = sk ? xv.w : xv.z, x3 = sk ? xv.z : xv.w;
                    acc = fmaf(-p[48 - sk16], x3, fmaf(-p[32 + sk16], x2, fmaf(-p[16 - sk16], x1, fmaf(-p[sk16], x0, acc))));
                }
                zq[q2] = acc;
            }
        }
    }
}

// ---- fp16-packed products (m16n8k16): see solve_tiles_h
__device__ __forceinline__ void split_h2(float x, float y, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
__device__ __forceinline__ void mma_f16(float4 &c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, BFrag b) {
    asm volatile(
        "{\n"
        ".reg .b32 b0, b1;\n"
        "mov.b64 {b0, b1}, %8;\n"
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {b0,b1}, {%0,%1,%2,%3};\n"
        "}\n"
        : "+f"(c.x), "+f"(c.y), "+f"(c.z), "+f"(c.w)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "l"(b));
}
// A operand from an accumulator-layout tile, hi and lo halves as f16x2: h0 / l0 = (row g, cols 2t, 2t+1), h1 / l1 = (row g + 8, ...).
// The instruction wants the operand in four consecutive registers [a | a]: the duplicates are made ONCE per tile (opaque movs, so
// ptxas keeps the quad instead of rebuilding it with two MOVs in front of every HMMA).
__device__ __forceinline__ uint32_t dup32(uint32_t v) {
    uint32_t r;
    asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
struct AFragH {
    uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
    __device__ __forceinline__ void set(float x, float y, float z, float w) {
        split_h2(x, y, h0, l0);
        split_h2(z, w, h1, l1);
        dup();
    }
    __device__ __forceinline__ void dup() { h2 = dup32(h0); h3 = dup32(h1); l2 = dup32(l0); l3 = dup32(l1); }
};
// c += A B^T, B = {hi, lo} packed along K: a_lo (b_hi + b_lo), then a_hi (b_hi + b_lo)
__device__ __forceinline__ void mma2h(float4 &c, const AFragH &a, BFrag b) {
    mma_f16(c, a.l0, a.l1, a.l2, a.l3, b);
    mma_f16(c, a.h0, a.h1, a.h2, a.h3, b);
}

// fp16-packed variant of solve_tiles (FMS_CHOL_F16 = 1): every product runs as TWO HMMA.16816.F32 with the hi and the lo parts
// of the B operand packed into the K dimension -- A = [a | a] (k slots 0..7 and 8..15 hold the same 8 columns), B = [b_hi ; b_lo]:
// a_lo (b_hi + b_lo) first, then a_hi (b_hi + b_lo): all four partial products, one third fewer tensor instructions than 3xTF32
// and half the B-fragment registers.  fp16 has 5 exponent bits: the CALLER scales the system (matrix and right-hand side by the
// same power of two, x is unchanged) so that the diagonal is O(1); entries below 2^-3 keep an absolute error of 3e-8.
// Solves one system in place.  S: tile layout (negated matrix) -- holds L afterwards.  bz: D floats, -b in, z out.
// invd: D floats scratch (unused by the solver since W replaces the diagonal blocks; kept for the caller's staging).  xs: D floats scratch (x, for the broadcasts of the back substitution).
// x_out[q] = x[lane + 32 q].  All four arrays are private to the calling warp; the caller provides the barrier that
// made S and bz visible (mbarrier wait or __syncwarp).
template <int MT, bool ROLL = false>
__device__ __forceinline__ void solve_tiles_h(float *__restrict__ S, float *__restrict__ bz, float *__restrict__ invd, float *__restrict__ xs,
                                            const int lane, float (&x_out)[MT / 2]) {
    using TL = Tiles<MT>;
    constexpr int NT = TL::NT;
    const int g = lane >> 2, t = lane & 3;
    float4 *S4 = reinterpret_cast<float4 *>(S);
    BFrag bf[NT];                                           // B fragments {hi, lo} of the current panel's L, per n-tile
#pragma unroll
    for (int j = 0; j < NT; ++j) bf[j] = 0ull;

#pragma unroll 1
    for (int J = 0; J < NT; ++J) {
        const int Id = J >> 1, h = J & 1;
        // ---- 1. diagonal block, factored redundantly: a[r (r + 1) / 2 + c] = (A + lambda n I)[8 J + r][8 J + c], c <= r
        BFrag wb;
        float2 w_own;
        {
            const float *Td = S + (Id * (Id + 1) + J) * 128 + 2 * h;
            float a[36];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int cp = 0; 2 * cp <= r; ++cp) {
                    const float2 v = *reinterpret_cast<const float2 *>(Td + (4 * r + cp) * 4);
                    a[r * (r + 1) / 2 + 2 * cp] = -v.x;
                    if (2 * cp + 1 <= r) a[r * (r + 1) / 2 + 2 * cp + 1] = -v.y;
                }
            float rinv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                rinv[k] = rsqrt_nr(a[k * (k + 1) / 2 + k]);
#pragma unroll
                for (int i = k + 1; i < 8; ++i) a[i * (i + 1) / 2 + k] *= rinv[k];
#pragma unroll
                for (int j = k + 1; j < 8; ++j)
#pragma unroll
                    for (int i = j; i < 8; ++i) a[i * (i + 1) / 2 + j] = fmaf(-a[i * (i + 1) / 2 + k], a[j * (j + 1) / 2 + k], a[i * (i + 1) / 2 + j]);
            }
            // row g of W = Ld^-1 (w Ld = e_g^T, solved from the right; entries right of the diagonal come out 0)
            float wv[8];
#pragma unroll
            for (int c = 7; c >= 0; --c) {
                float s = (c == g) ? 1.f : 0.f;
#pragma unroll
                for (int m = c + 1; m < 8; ++m) s = fmaf(-wv[m], a[m * (m + 1) / 2 + c], s);
                wv[c] = s * rinv[c];
            }
            // B fragment of -W: n = g (row of W), k slot t -> column 2 t, k slot t + 4 -> column 2 t + 1
            uint32_t h0, l0;
            w_own = make_float2(sel4(t, wv[0], wv[2], wv[4], wv[6]), sel4(t, wv[1], wv[3], wv[5], wv[7]));   // W[g][2t], W[g][2t+1]
            split_h2(-w_own.x, -w_own.y, h0, l0);
            wb = bfrag(h0, l0);
        }
        // ---- 2 + 3. panel tiles: L(I, J) = S(I, J) (-W)^T, then S(I, j) += L(I, J) L(j, J)^T for J < j <= 2 I + 1
        if constexpr (ROLL) {
            // Code-size variant for MT = 8 (als_wide.cu): the unrolled form above is 25 KB of SASS per panel iteration, beyond the
            // 32 KB instruction cache once ten warps run different parts of the kernel.  The row tiles are walked by two RUNTIME
            // loops (tiles 0 .. MT/2-1, which only meet n-tiles 0 .. MT-1, and tiles MT/2 .. MT-1, which meet all of them); the
            // B-fragment registers keep their static indices: the new pair is steered into bf[2 I], bf[2 I + 1] by a static
            // select, the pair loop is unrolled over all candidates and guarded by uniform branches.
            constexpr int HB = MT / 2;
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
#pragma unroll 1
                for (int Ir = 0; Ir < HB; ++Ir) {
                    const int I = hb * HB + Ir;
                    if (I < Id) continue;
                    const int row0 = I * (I + 1);
                    const float4 sv = S4[(row0 + J) * 32 + lane];
                    AFragH af;
                    af.set(sv.x, sv.y, sv.z, sv.w);
                    float4 L = make_float4(0.f, 0.f, 0.f, 0.f);
                    mma2h(L, af, wb);
                    if (I == Id) {
                        const bool u0 = 2 * t > g, u1 = 2 * t + 1 > g;
                        if (h == 0) { L.x = u0 ? 0.f : L.x; L.y = u1 ? 0.f : L.y; }
                        else { L.x = 0.f; L.y = 0.f; L.z = u0 ? 0.f : L.z; L.w = u1 ? 0.f : L.w; }
                    }
                    S4[(row0 + J) * 32 + lane] = L;
                    af.set(L.x, L.y, L.z, L.w);
                    const BFrag n0 = bfrag(af.h0, af.l0), n1 = bfrag(af.h1, af.l1);
#pragma unroll
                    for (int k = 0; k < HB; ++k)
                        if (Ir == k) { bf[2 * (hb * HB + k)] = n0; bf[2 * (hb * HB + k) + 1] = n1; }
#pragma unroll
                    for (int m = (hb + 1) * HB - 1; m >= 0; --m) {
                        const int j = 2 * m + 1;
                        if (m > I) continue;                                // n-tiles right of this row tile's diagonal
                        if (j <= J) break;
                        const bool two = j - 1 > J;
                        float4 c0 = S4[(row0 + j) * 32 + lane];
                        float4 c1 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (two) c1 = S4[(row0 + j - 1) * 32 + lane];
                        mma_f16(c0, af.l0, af.l1, af.l2, af.l3, bf[j]);
                        mma_f16(c1, af.l0, af.l1, af.l2, af.l3, bf[j - 1]);
                        mma_f16(c0, af.h0, af.h1, af.h2, af.h3, bf[j]);
                        mma_f16(c1, af.h0, af.h1, af.h2, af.h3, bf[j - 1]);
                        S4[(row0 + j) * 32 + lane] = c0;
                        if (two) S4[(row0 + j - 1) * 32 + lane] = c1;
                    }
                }
            }
        } else {
#pragma unroll
        for (int I = 0; I < MT; ++I) {
            if (I >= Id) {
                const float4 sv = S4[(I * (I + 1) + J) * 32 + lane];
                AFragH af;
                af.set(sv.x, sv.y, sv.z, sv.w);
                float4 L = make_float4(0.f, 0.f, 0.f, 0.f);
                mma2h(L, af, wb);
                if (I == Id) {
                    // rows of the diagonal block: keep the lower triangle (the product reproduces Ld up to rounding noise
                    // above the diagonal); rows of the m-tile above the block belong to finished panels: zero
                    const bool u0 = 2 * t > g, u1 = 2 * t + 1 > g;
                    if (h == 0) { L.x = u0 ? 0.f : L.x; L.y = u1 ? 0.f : L.y; }
                    else { L.x = 0.f; L.y = 0.f; L.z = u0 ? 0.f : L.z; L.w = u1 ? 0.f : L.w; }
                }
                S4[(I * (I + 1) + J) * 32 + lane] = L;
                af.set(L.x, L.y, L.z, L.w);
                bf[2 * I] = bfrag(af.h0, af.l0);                  // rows g     -> n-tile 2 I
                bf[2 * I + 1] = bfrag(af.h1, af.l1);              // rows g + 8 -> n-tile 2 I + 1
                // descending j in PAIRS with an early exit (a real, uniform branch instead of predicated-off HMMA / LDS / STS):
                // the two tiles' 3-HMMA chains are independent, so interleaving them fills the ~30-cycle dependent-issue gaps
                // (one tile at a time, ptxas pads every HMMA of the chain with a NOP)
#pragma unroll
                for (int j = 2 * I + 1; j >= 1; j -= 2) {
                    if (j <= J) break;
                    const bool two = j - 1 > J;                         // the lower tile of the pair is live too
                    float4 c0 = S4[(I * (I + 1) + j) * 32 + lane];
                    float4 c1 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (two) c1 = S4[(I * (I + 1) + j - 1) * 32 + lane];
                    mma_f16(c0, af.l0, af.l1, af.l2, af.l3, bf[j]);
                    mma_f16(c1, af.l0, af.l1, af.l2, af.l3, bf[j - 1]);
                    mma_f16(c0, af.h0, af.h1, af.h2, af.h3, bf[j]);
                    mma_f16(c1, af.h0, af.h1, af.h2, af.h3, bf[j - 1]);
                    S4[(I * (I + 1) + j) * 32 + lane] = c0;
                    if (two) S4[(I * (I + 1) + j - 1) * 32 + lane] = c1;
                }
            }
        }
        }
        // the diagonal block of the factor is only ever used through its inverse again (back substitution): keep W = Ld^-1 there.
        // Lane (g, t) owns exactly the two entries W[g][2t], W[g][2t+1] of its fragment position.
        *reinterpret_cast<float2 *>(S + (Id * (Id + 1) + J) * 128 + lane * 4 + 2 * h) = w_own;
        // ---- 4. right-hand-side row (row D of the matrix): z_J = bz_J (-W)^T, bz_j += z_J L(j, J)^T
        {
            float2 bv = make_float2(0.f, 0.f);
            if (g == 0) bv = *reinterpret_cast<const float2 *>(bz + 8 * J + 2 * t);
            AFragH af;
            af.h1 = af.l1 = 0u;
            split_h2(bv.x, bv.y, af.h0, af.l0);
            af.dup();
            float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            mma2h(z, af, wb);
            if (g == 0) *reinterpret_cast<float2 *>(bz + 8 * J + 2 * t) = make_float2(z.x, z.y);
            split_h2(z.x, z.y, af.h0, af.l0);
            af.dup();
            // descending pairs like the trailing update (two independent 3-HMMA chains per branch)
#pragma unroll
            for (int j = NT - 1; j >= 1; j -= 2) {
                if (j <= J) break;
                const bool two = j - 1 > J;
                float2 cv0 = make_float2(0.f, 0.f), cv1 = make_float2(0.f, 0.f);
                if (g == 0) {
                    cv0 = *reinterpret_cast<const float2 *>(bz + 8 * j + 2 * t);
                    if (two) cv1 = *reinterpret_cast<const float2 *>(bz + 8 * j - 8 + 2 * t);
                }
                float4 c0 = make_float4(cv0.x, cv0.y, 0.f, 0.f), c1 = make_float4(cv1.x, cv1.y, 0.f, 0.f);
                mma_f16(c0, af.l0, af.l1, af.l2, af.l3, bf[j]);
                mma_f16(c1, af.l0, af.l1, af.l2, af.l3, bf[j - 1]);
                mma_f16(c0, af.h0, af.h1, af.h2, af.h3, bf[j]);
                mma_f16(c1, af.h0, af.h1, af.h2, af.h3, bf[j - 1]);
                if (g == 0) {
                    *reinterpret_cast<float2 *>(bz + 8 * j + 2 * t) = make_float2(c0.x, c0.y);
                    if (two) *reinterpret_cast<float2 *>(bz + 8 * j - 8 + 2 * t) = make_float2(c1.x, c1.y);
                }
            }
        }
        __syncwarp();
    }

    // ---- back substitution L^T x = z: column i = lane + 32 q per lane, one panel (8 pivots) at a time, blocks of 32 columns
    // descending.  x_J = W_J^T r_J with W_J = Ld^-1 stored over the diagonal block: no sequential triangular solve.
    constexpr int NQ = MT / 2;
    float zq[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) zq[q] = bz[lane + 32 * q];
    const int lp = (lane >> 3) * 128 + ((lane & 7) >> 1) * 4 + (lane & 1);      // lane part of Tiles::at(k, 32 q + lane)
    const int sk = (lane >> 3) & 1, sk16 = 16 * sk;                             // row-order skew of the odd lane groups (bank conflicts)
#pragma unroll
    for (int q = NQ - 1; q >= 0; --q) {
        float x = 0.f;
#pragma unroll 1
        for (int jr = 24; jr >= 0; jr -= 8) {
            const int j0 = 32 * q + jr, J = j0 >> 3, I = J >> 1, hh = J & 1;
            const int rel = lane - jr;
            // pivot lane c = rel: x_c = sum_r W[r][c] r_r  (W lower triangular, zeros above the diagonal are stored)
            const float *wc = S + (I * (I + 1) + J) * 128 + 2 * hh + ((rel & 7) >> 1) * 4 + (rel & 1);
            float xc = 0.f;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float zr = __shfl_sync(0xffffffffu, zq[q], jr + r);
                xc = fmaf(wc[16 * r], zr, xc);
            }
            if (rel >= 0 && rel < 8) x = xc;
            // columns left of the panel: z_i -= sum_c L[j0 + c][i] x_c
            // rows j0 + c sit 16 floats apart, the four 8-column tiles of a lane group 128 floats: all four groups would hit the
            // same 8 banks (4-way conflict).  Odd groups walk the rows in the order 1, 0, 3, 2, ...: two bank sets, 2 wavefronts.
            const float *p = S + (I * (I + 1)) * 128 + 2 * hh + 4 * q * 128 + lp;
            float acc = zq[q];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float xv = __shfl_sync(0xffffffffu, xc, jr + (c ^ sk));
                acc = fmaf(-p[16 * c + ((c & 1) ? -sk16 : sk16)], xv, acc);
            }
            if (rel < 0) zq[q] = acc;
        }
        x_out[q] = x;
        if (q > 0) {
            xs[32 * q + lane] = x;
            __syncwarp();
            // earlier column blocks: z_i -= sum_{k in block q} L[k][i] x_k
#pragma unroll
            for (int q2 = 0; q2 < q; ++q2) {
                float acc = zq[q2];
#pragma unroll 1
                for (int k4 = 0; k4 < 32; k4 += 4) {
                    const int k = 32 * q + k4;
                    const int I = k >> 4;
                    const float *p = S + (I * (I + 1) + 4 * q2) * 128 + (k & 7) * 16 + ((k >> 3) & 1) * 2 + lp;
                    const float4 xv = *reinterpret_cast<const float4 *>(xs + k);
                    const float x0 = sk ? xv.y : xv.x, x1 = sk ? xv.x : xv.y, x2 = sk ? xv.w : xv.z, x3 = sk ? xv.z : xv.w;
                    acc = fmaf(-p[48 - sk16], x3, fmaf(-p[32 + sk16], x2, fmaf(-p[16 - sk16], x1, fmaf(-p[sk16], x0, acc))));
                }
                zq[q2] = acc;
            }
        }
    }
}

}  // namespace cm
}  // namespace fms
