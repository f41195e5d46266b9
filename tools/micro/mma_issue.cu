// Microbenchmark (development aid): cycles per stage of the tcgen05 issue loop of the ALS kernel, piece by piece.
// One CTA per SM, one issuing warp, operands = zeroed shared memory (no producers), accumulators in TMEM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__host__ __device__ constexpr uint32_t make_idesc(uint32_t m, uint32_t n) {
    return (1u << 4) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
constexpr int TILE = 2048, STAGE = 3 * TILE, NSTG = 13;
#ifndef SPIN_MODE
#define SPIN_MODE 0
#endif

template <int VARIANT, int SPIN>
__global__ void __launch_bounds__(896, 1) k_issue(int n_stages, long long *out) {
    extern __shared__ __align__(1024) unsigned char raw[];
    unsigned char *sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    const uint32_t base = smem_u32(sm);
    const uint32_t bars = base + NSTG * STAGE;                 // 2 * NSTG + 2 barriers
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + NSTG * STAGE + 512);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < NSTG * STAGE / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sm)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2 * NSTG + 3; ++s) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t d = *tmem_slot;
    const uint32_t bar_stop = bars + 8 * (2 * NSTG + 2);
    if (warp >= 4 && warp < 4 + SPIN) {                       // idle warps of the real kernel: hinted try_wait polling
        uint32_t ok;
        do {
            if (SPIN_MODE == 0)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar_stop), "r"(0), "r"(1000000u) : "memory");
            else { ok = mbar_test(bar_stop, 0); if (!ok) __nanosleep(1000); }
        } while (!ok);
    }
    if (warp == 0) {
        const uint32_t id1 = make_idesc(64, 136), id2 = make_idesc(64, 72);
        const uint32_t LBO_T = (uint32_t)(TILE >> 4) << 16, LBO_B2 = (uint32_t)((2 * TILE + 16) >> 4) << 16;
        const uint32_t enc0 = (base >> 4) & 0x3FFF;
        uint32_t slot = 0, ph = 0, enc = enc0, ready = 1;
        const long long t0 = clock64();
        for (int st = 0; st < n_stages; ++st) {
            if (VARIANT >= 4) {
                if (!ready) { while (!mbar_test(bars + 8 * slot, ph)) { } }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            uint32_t nslot = slot + 1, nph = ph, nenc = enc + (STAGE >> 4);
            if (nslot == NSTG) { nslot = 0; nph ^= 1u; nenc = enc0; }
            uint32_t nready = 1;
            if (VARIANT >= 4) nready = mbar_test(bars + 8 * (2 * NSTG), 1) | 1u;     // a barrier that never completes; result forced true
            const uint32_t lo_hi = enc | LBO_T, lo_lo = (enc + (TILE >> 4)) | LBO_T, lo_b2 = enc | LBO_B2;
            const uint32_t be = bars + 8 * (NSTG + slot);
            const uint32_t acc = st > 0 ? 1u : 0u;
            if (VARIANT == 0 || VARIANT == 4) {          // product: 2 MMAs + commit
                asm volatile("{\n.reg .pred pe, pa, pt;\n.reg .b32 d2;\n.reg .b64 ahi, alo, b2;\nelect.sync _|pe, 0xffffffff;\nsetp.ne.b32 pa, %6, 0;\nsetp.eq.u32 pt, %0, %0;\n"
                             "add.u32 d2, %0, 64;\nmov.b64 ahi, {%1, 0x40004040};\nmov.b64 alo, {%2, 0x40004040};\nmov.b64 b2, {%3, 0x40004040};\n"
                             "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], ahi, ahi, %4, pa;\n@pe tcgen05.mma.cta_group::1.kind::f16 [d2], alo, b2, %5, pt;\n"
                             "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n}\n" ::"r"(d), "r"(lo_hi), "r"(lo_lo), "r"(lo_b2),
                             "r"(id1), "r"(id2), "r"(acc), "r"(be) : "memory");
            } else if (VARIANT == 1) {                   // 2 MMAs, no commit
                asm volatile("{\n.reg .pred pe, pa, pt;\n.reg .b32 d2;\n.reg .b64 ahi, alo, b2;\nelect.sync _|pe, 0xffffffff;\nsetp.ne.b32 pa, %6, 0;\nsetp.eq.u32 pt, %0, %0;\n"
                             "add.u32 d2, %0, 64;\nmov.b64 ahi, {%1, 0x40004040};\nmov.b64 alo, {%2, 0x40004040};\nmov.b64 b2, {%3, 0x40004040};\n"
                             "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], ahi, ahi, %4, pa;\n@pe tcgen05.mma.cta_group::1.kind::f16 [d2], alo, b2, %5, pt;\n}\n" ::"r"(d),
                             "r"(lo_hi), "r"(lo_lo), "r"(lo_b2), "r"(id1), "r"(id2), "r"(acc), "r"(be) : "memory");
            } else if (VARIANT == 2) {                   // 1 MMA + commit
                asm volatile("{\n.reg .pred pe, pa;\n.reg .b64 ahi;\nelect.sync _|pe, 0xffffffff;\nsetp.ne.b32 pa, %3, 0;\nmov.b64 ahi, {%1, 0x40004040};\n"
                             "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], ahi, ahi, %2, pa;\n@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%4];\n}\n" ::"r"(d),
                             "r"(lo_hi), "r"(id1), "r"(acc), "r"(be) : "memory");
            } else if (VARIANT == 3) {                   // commit only
                asm volatile("{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(be) : "memory");
            } else if (VARIANT == 5) {                   // 1 MMA, no commit
                asm volatile("{\n.reg .pred pe, pa;\n.reg .b64 ahi;\nelect.sync _|pe, 0xffffffff;\nsetp.ne.b32 pa, %3, 0;\nmov.b64 ahi, {%1, 0x40004040};\n"
                             "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], ahi, ahi, %2, pa;\n}\n" ::"r"(d), "r"(lo_hi), "r"(id1), "r"(acc), "r"(be) : "memory");
            }
            slot = nslot; ph = nph; enc = nenc; ready = nready;
        }
        const long long t1 = clock64();
        // drain: commit to the last barrier and wait for it
        const uint32_t bl = bars + 8 * (2 * NSTG + 1);
        asm volatile("{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(bl) : "memory");
        while (!mbar_test(bl, 0)) { }
        const long long t2 = clock64();
        if (blockIdx.x == 0 && lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        if (lane == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar_stop) : "memory");   // release the spinners
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(d) : "memory");
}

template <int V, int SPIN>
void run(const char *name, long long *d_out) {
    const int n = 4000;
    const size_t smem = NSTG * STAGE + 2048 + 1024;
    cudaFuncSetAttribute(k_issue<V, SPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) {
        k_issue<V, SPIN><<<148, 896, smem>>>(n, d_out);
        long long h[2];
        cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
        if (rep) printf("%-44s spinners %2d: issue %.1f cyc/stage, incl. drain %.1f cyc/stage  (%s)\n", name, SPIN, (double)h[0] / n, (double)h[1] / n, cudaGetErrorString(cudaGetLastError()));
    }
}

int main() {
    long long *d_out;
    cudaMalloc(&d_out, 64);
    run<0, 0>("2 MMA (N=136, N=72) + commit", d_out);
    run<2, 0>("1 MMA (N=136) + commit", d_out);
    run<3, 0>("commit only", d_out);
    run<4, 0>("product loop: probe + fence + 2 MMA + commit", d_out);
    run<4, 4>("product loop", d_out);
    run<4, 16>("product loop", d_out);
    run<4, 24>("product loop", d_out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
