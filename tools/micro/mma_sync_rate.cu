// Microbenchmark (development aid, round 2): what the legacy warp-level tensor path (mma.sync, SASS HMMA) delivers on
// sm_100a, and the register layout of tcgen05.ld.16x256b -- the two facts the tensor-core Cholesky of als_tc.cu rests on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_rate mma_sync_rate.cu && ./mma_sync_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// KIND 0 tf32, 1 f16, 2 bf16; CHAINS independent accumulators per warp (dependent-issue latency vs throughput)
template <int KIND, int CHAINS>
__global__ void k_rate(int iters, float *out, long long *cyc) {
    uint32_t a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = 0x3c003c00u + threadIdx.x + i;
    for (int i = 0; i < 2; ++i) b[i] = 0x3c003c00u + 3 * threadIdx.x + i;
    float c[CHAINS][4];
    for (int k = 0; k < CHAINS; ++k) for (int i = 0; i < 4; ++i) c[k][i] = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) {
            if (KIND == 0) mma_tf32(c[k], a, b);
            else if (KIND == 1) mma_f16(c[k], a, b);
            else mma_bf16(c[k], a, b);
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int k = 0; k < CHAINS; ++k) for (int i = 0; i < 4; ++i) s += c[k][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND, int CHAINS>
static void run(const char *name, int warps, int macs_per_mma) {
    float *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    k_rate<KIND, CHAINS><<<148, warps * 32>>>(iters, out, cyc);
    k_rate<KIND, CHAINS><<<148, warps * 32>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double mma_per_clk_sm = (double)iters * CHAINS * warps / (double)mx;
    printf("%-5s warps/SM %2d chains %d : %.3f mma/clk/SM = %7.1f MAC/clk/SM  (%.2f clk per mma per SMSP)  [%s]\n", name, warps, CHAINS,
           mma_per_clk_sm, mma_per_clk_sm * macs_per_mma, 4.0 / mma_per_clk_sm, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

// ---- tcgen05.ld.16x256b register layout: fill TMEM through 32x32b stores (lane = row, register = column: value = 1000 row + col),
// read 16 columns back with 16x256b.x2 and print which (row, col) every thread's registers hold.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k_tmem_layout(float *out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    const uint32_t tq = base + ((uint32_t)(32 * warp) << 16);
    uint32_t v[16];
    for (int c = 0; c < 16; ++c) v[c] = __float_as_uint(1000.f * (32 * warp + lane) + c);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(tq),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                 "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tq)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[(warp * 32 + lane) * 16 + i] = __uint_as_float(r[i]);
    // the same with the lane field advanced by 16 (second half of the quadrant)
    uint32_t r2[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3]), "=r"(r2[4]), "=r"(r2[5]), "=r"(r2[6]), "=r"(r2[7])
                 : "r"(tq + (16u << 16))
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[(warp * 32 + lane) * 16 + 8 + i] = __uint_as_float(r2[i]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(base) : "memory");
}

int main() {
    for (int w : {4, 8, 16}) {
        run<0, 1>("tf32", w, 1024);
        run<0, 4>("tf32", w, 1024);
        run<0, 8>("tf32", w, 1024);
    }
    for (int w : {4, 8, 16}) {
        run<1, 1>("f16", w, 2048);
        run<1, 8>("f16", w, 2048);
    }
    for (int w : {4, 16}) run<2, 8>("bf16", w, 2048);

    float *out;
    cudaMalloc(&out, 128 * 16 * 4);
    cudaMemset(out, 0, 128 * 16 * 4);
    k_tmem_layout<<<1, 128>>>(out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("tmem layout kernel: %s\n", cudaGetErrorString(e));
    static float h[128 * 16];
    cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    for (int t = 0; t < 128; ++t) {
        if (t % 32 >= 8 && t % 32 != 31) continue;
        printf("thread %3d:", t);
        for (int i = 0; i < 16; ++i) printf(" %7.0f%s", h[t * 16 + i], i == 7 ? " |" : "");
        printf("\n");
    }
    return 0;
}
