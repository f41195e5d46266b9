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
int main() {
    const int64_t rows = 1000000, n = 100000000;
    uint4 *hi, *lo, *pl, *out; int *idx;
    cudaMalloc(&hi, rows * 128); cudaMalloc(&lo, rows * 128); cudaMalloc(&pl, rows * 256); cudaMalloc(&out, 64);
    cudaMemset(hi, 1, rows * 128); cudaMemset(lo, 2, rows * 128); cudaMemset(pl, 3, rows * 256);
    int *h = (int *)malloc(n * 4);
    uint64_t s = 88172645463325252ull;
    for (int64_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % rows); }
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
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
