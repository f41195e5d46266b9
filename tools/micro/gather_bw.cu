// Microbenchmark (development aid): DRAM throughput of random row gathers, two separate 128-byte planes vs one
// interleaved 256-byte row.  Decides whether interleaving the fp16 hi/lo gather planes of the ALS kernel can pay.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_sep(const uint4 *hi, const uint4 *lo, const int *idx, int64_t n, uint4 *out) {
    const int lane8 = threadIdx.x & 7;
    int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;       // one 8-lane group per row
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 3;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (; g + 3 * stride < n; g += 4 * stride) {
        uint4 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int64_t r = idx[g + u * stride]; a[u] = hi[r * 8 + lane8]; b[u] = lo[r * 8 + lane8]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc.x ^= a[u].x ^ b[u].x; acc.y ^= a[u].y ^ b[u].y; acc.z ^= a[u].z ^ b[u].z; acc.w ^= a[u].w ^ b[u].w; }
    }
    if (acc.x == 0x12345678u) out[0] = acc;
}
__global__ void k_int(const uint4 *pl, const int *idx, int64_t n, uint4 *out) {
    const int lane16 = threadIdx.x & 15;
    int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;      // one 16-lane group per 256-byte row
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 4;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (; g + 7 * stride < n; g += 8 * stride) {
        uint4 a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int64_t r = idx[g + u * stride]; a[u] = pl[r * 16 + lane16]; }
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc.x ^= a[u].x; acc.y ^= a[u].y; acc.z ^= a[u].z; acc.w ^= a[u].w; }
    }
    if (acc.x == 0x12345678u) out[0] = acc;
}
// Same gathers, but each warp walks a contiguous range of the index stream (like a CTA walking its CSR rows), and the
// stream is made of rows of 1000 ascending indices (like the item-major CSR of the ALS items half-step).
__global__ void k_sep_seq(const uint4 *hi, const uint4 *lo, const int *idx, int64_t n, uint4 *out) {
    const int lane = threadIdx.x & 31, lane8 = lane & 7, sub = lane >> 3;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t per = (n / nwarps) & ~15ll, b = warp * per, e = b + per;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int64_t g = b; g < e; g += 16) {
        uint4 a[4], c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int64_t r = idx[g + 4 * u + sub]; a[u] = hi[r * 8 + lane8]; c[u] = lo[r * 8 + lane8]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc.x ^= a[u].x ^ c[u].x; acc.y ^= a[u].y ^ c[u].y; acc.z ^= a[u].z ^ c[u].z; acc.w ^= a[u].w ^ c[u].w; }
    }
    if (acc.x == 0x12345678u) out[0] = acc;
}
// k_sep_seq plus an L2 prefetch PF entries ahead in the same stream (tests whether prefetch.global.L2 raises the
// effective memory-level parallelism of a latency-bound gather)
template <int PF>
__global__ void k_sep_seq_pf(const uint4 *hi, const uint4 *lo, const int *idx, int64_t n, uint4 *out) {
    const int lane = threadIdx.x & 31, lane8 = lane & 7, sub = lane >> 3;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t per = (n / nwarps) & ~15ll, b = warp * per, e = b + per;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int64_t g = b; g < e; g += 16) {
        if (g + PF + 16 <= e) {                                   // 16 rows x 2 planes = 32 lines, one per lane
            const int64_t r = idx[g + PF + (lane & 15)];
            const uint4 *p = (lane < 16 ? hi : lo) + r * 8;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        }
        uint4 a[4], c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int64_t r = idx[g + 4 * u + sub]; a[u] = hi[r * 8 + lane8]; c[u] = lo[r * 8 + lane8]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc.x ^= a[u].x ^ c[u].x; acc.y ^= a[u].y ^ c[u].y; acc.z ^= a[u].z ^ c[u].z; acc.w ^= a[u].w ^ c[u].w; }
    }
    if (acc.x == 0x12345678u) out[0] = acc;
}
__global__ void k_int_seq(const uint4 *pl, const int *idx, int64_t n, uint4 *out) {
    const int lane = threadIdx.x & 31, lane16 = lane & 15, sub = lane >> 4;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t per = (n / nwarps) & ~15ll, b = warp * per, e = b + per;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int64_t g = b; g < e; g += 16) {
        uint4 a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int64_t r = idx[g + 2 * u + sub]; a[u] = pl[r * 16 + lane16]; }
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc.x ^= a[u].x; acc.y ^= a[u].y; acc.z ^= a[u].z; acc.w ^= a[u].w; }
    }
    if (acc.x == 0x12345678u) out[0] = acc;
}
#include <algorithm>
int main() {
    const int64_t rows = 1000000, n = 100000000;
    uint4 *hi, *lo, *pl, *out; int *idx;
    cudaMalloc(&hi, rows * 128); cudaMalloc(&lo, rows * 128); cudaMalloc(&pl, rows * 256); cudaMalloc(&out, 64);
    cudaMemset(hi, 1, rows * 128); cudaMemset(lo, 2, rows * 128); cudaMemset(pl, 3, rows * 256);
    int *h = (int *)malloc(n * 4);
    uint64_t s = 88172645463325252ull;
    for (int64_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % rows); }
    for (int64_t r = 0; r < n / 1000; ++r) std::sort(h + r * 1000, h + (r + 1) * 1000);   // item-major CSR: ascending users per item
    cudaMalloc(&idx, n * 4); cudaMemcpy(idx, h, n * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int blocks : {148 * 8, 148 * 16}) {
        for (int rep = 0; rep < 3; ++rep) {
            float ms;
            cudaEventRecord(e0); k_sep<<<blocks, 256>>>(hi, lo, idx, n, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            printf("separate   blocks=%d  %.3f ms  %.0f GB/s (rows only)\n", blocks, ms, n * 256.0 / ms / 1e6);
            cudaEventRecord(e0); k_int<<<blocks, 256>>>(pl, idx, n, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            printf("interleave blocks=%d  %.3f ms  %.0f GB/s (rows only)\n", blocks, ms, n * 256.0 / ms / 1e6);
        }
    }
    for (int blocks : {74, 148, 148 * 2, 148 * 8}) {
        for (int rep = 0; rep < 2; ++rep) {
            float ms;
            cudaEventRecord(e0); k_sep_seq<<<blocks, 256>>>(hi, lo, idx, n, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            printf("separate-seq   blocks=%d  %.3f ms  %.0f GB/s\n", blocks, ms, n * 256.0 / ms / 1e6);
            cudaEventRecord(e0); k_int_seq<<<blocks, 256>>>(pl, idx, n, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            printf("interleave-seq blocks=%d  %.3f ms  %.0f GB/s\n", blocks, ms, n * 256.0 / ms / 1e6);
            cudaEventRecord(e0); k_sep_seq_pf<64><<<blocks, 256>>>(hi, lo, idx, n, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            printf("separate-seq+pf64  blocks=%d  %.3f ms  %.0f GB/s\n", blocks, ms, n * 256.0 / ms / 1e6);
            cudaEventRecord(e0); k_sep_seq_pf<256><<<blocks, 256>>>(hi, lo, idx, n, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            printf("separate-seq+pf256 blocks=%d  %.3f ms  %.0f GB/s\n", blocks, ms, n * 256.0 / ms / 1e6);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
