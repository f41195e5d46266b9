// Development harness for flink-ms_b200/csrc/chol_mma.cuh: ALS-like SPD systems (Y^T Y + lambda n I, b = Y^T r) -> tile
// layout -> warp solver -> compare with an fp64 Cholesky on the host; then solver throughput with W warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o chol_mma_test chol_mma_test.cu && ./chol_mma_test [MT]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../flink-ms_b200/csrc/chol_mma.cuh"

using namespace fms::cm;

template <int MT, bool H, bool R = false>
__global__ void __launch_bounds__(512, 1) k_solve(const float *__restrict__ A, const float *__restrict__ b, int n_sys, int reps, float *__restrict__ X,
                                                  long long *cyc) {
    using TL = Tiles<MT>;
    constexpr int D = TL::D;
    extern __shared__ __align__(16) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    float *S = sm + (size_t)warp * (TL::FLOATS + 3 * D);
    float *bz = S + TL::FLOATS, *invd = bz + D, *xs = invd + D;
    const long long t0 = clock64();
    // every warp solves `reps` systems (balanced: the cycle count then reads as per-warp latency x reps)
    for (int rep = 0; rep < reps; ++rep) {
        {
            const int s = ((blockIdx.x * nw + warp) * reps + rep) % n_sys;
            const float *As = A + (size_t)s * D * D;
            // fp16-packed variant: the caller scales the system by a power of two taken from the first diagonal entry
            const float sc = H ? __uint_as_float((254u - ((__float_as_uint(As[0]) >> 23) & 0xffu)) << 23) : 1.f;
            // what the epilogue warps do in the real kernel: negated system in tile layout, -b
#pragma unroll 1
            for (int I = 0; I < MT; ++I)
                for (int j = 0; j <= 2 * I + 1; ++j) {
                    const float2 r0 = *reinterpret_cast<const float2 *>(As + (16 * I + g) * D + 8 * j + 2 * t);
                    const float2 r1 = *reinterpret_cast<const float2 *>(As + (16 * I + g + 8) * D + 8 * j + 2 * t);
                    reinterpret_cast<float4 *>(S)[(I * (I + 1) + j) * 32 + lane] = make_float4(-r0.x * sc, -r0.y * sc, -r1.x * sc, -r1.y * sc);
                }
            for (int q = 0; q < D / 32; ++q) bz[lane + 32 * q] = -b[(size_t)s * D + lane + 32 * q] * sc;
            __syncwarp();
            float x[MT / 2];
            if (H) solve_tiles_h<MT, R>(S, bz, invd, xs, lane, x);
            else solve_tiles<MT>(S, bz, invd, xs, lane, x);
            for (int q = 0; q < MT / 2; ++q) X[(size_t)s * D + lane + 32 * q] = x[q];
            __syncwarp();
        }
    }
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

static void chol_solve64(std::vector<double> &a, std::vector<double> &b, int d) {
    for (int j = 0; j < d; ++j) {
        double s = a[j * d + j];
        for (int k = 0; k < j; ++k) s -= a[j * d + k] * a[j * d + k];
        const double l = std::sqrt(s);
        a[j * d + j] = l;
        for (int i = j + 1; i < d; ++i) {
            double v = a[i * d + j];
            for (int k = 0; k < j; ++k) v -= a[i * d + k] * a[j * d + k];
            a[i * d + j] = v / l;
        }
    }
    for (int i = 0; i < d; ++i) { double v = b[i]; for (int k = 0; k < i; ++k) v -= a[i * d + k] * b[k]; b[i] = v / a[i * d + i]; }
    for (int i = d - 1; i >= 0; --i) { double v = b[i]; for (int k = i + 1; k < d; ++k) v -= a[k * d + i] * b[k]; b[i] = v / a[i * d + i]; }
}

template <int MT, bool H, bool R = false>
static int run(int only_warps) {
    constexpr int D = 16 * MT;
    const int n_sys = 2048;
    std::vector<float> A((size_t)n_sys * D * D), b((size_t)n_sys * D);
    std::vector<double> xref((size_t)n_sys * D);
    srand(7);
    auto uni = []() { return (double)rand() / RAND_MAX; };
    for (int s = 0; s < n_sys; ++s) {
        const int n = s % 3 == 0 ? 4 + rand() % 12 : (s % 3 == 1 ? 60 + rand() % 80 : 600 + rand() % 800);   // ratings of the row
        std::vector<double> G((size_t)D * D, 0.0), rhs(D, 0.0), y(D);
        for (int k = 0; k < n; ++k) {
            for (int j = 0; j < D; ++j) y[j] = (s % 5 == 4) ? uni() * 2.0 - 1.0 : uni();
            const double r = 0.5 + 0.5 * (rand() % 10);
            for (int i = 0; i < D; ++i) { rhs[i] += r * y[i]; for (int j = 0; j < D; ++j) G[i * D + j] += y[i] * y[j]; }
        }
        for (int i = 0; i < D; ++i) G[i * D + i] += 0.9 * n;
        for (int i = 0; i < D * D; ++i) A[(size_t)s * D * D + i] = (float)G[i];
        for (int i = 0; i < D; ++i) b[(size_t)s * D + i] = (float)rhs[i];
        // the reference solves the fp32-rounded system in fp64 (identical inputs)
        std::vector<double> a64(D * D), b64(D);
        for (int i = 0; i < D * D; ++i) a64[i] = (double)A[(size_t)s * D * D + i];
        for (int i = 0; i < D; ++i) b64[i] = (double)b[(size_t)s * D + i];
        chol_solve64(a64, b64, D);
        for (int i = 0; i < D; ++i) xref[(size_t)s * D + i] = b64[i];
    }
    float *dA, *db, *dX;
    long long *dc;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dX, b.size() * 4); cudaMalloc(&dc, 148 * 8);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dX, 0, b.size() * 4);
    const size_t per_warp = (size_t)(Tiles<MT>::FLOATS + 3 * D) * 4;
    int rc = 0;
    for (int warps : {4, 8, 12, 16}) {
        const size_t smem = per_warp * warps;
        if (smem > 227 * 1024 || (only_warps && warps != only_warps)) continue;
        cudaFuncSetAttribute(k_solve<MT, H, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int reps = 16;
        k_solve<MT, H, R><<<148, warps * 32, smem>>>(dA, db, n_sys, 1, dX, dc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_solve<MT, H, R><<<148, warps * 32, smem>>>(dA, db, n_sys, reps, dX, dc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<float> X(b.size());
        cudaMemcpy(X.data(), dX, X.size() * 4, cudaMemcpyDeviceToHost);
        long long hc[148];
        cudaMemcpy(hc, dc, sizeof hc, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = hc[i] > mx ? hc[i] : mx;
        double worst = 0.0, sum = 0.0;
        int worst_s = -1, n_nan = 0;
        for (int s = 0; s < n_sys; ++s) {
            double num = 0.0, den = 0.0;
            for (int i = 0; i < D; ++i) {
                const double xr = xref[(size_t)s * D + i], xv = X[(size_t)s * D + i];
                if (!(xv == xv)) { ++n_nan; num = 1e30; }
                num = std::fmax(num, std::fabs(xv - xr));
                den = std::fmax(den, std::fabs(xr));
            }
            const double rel = num / den;
            sum += rel;
            if (rel > worst) { worst = rel; worst_s = s; }
        }
        const double total = 148.0 * warps * reps;
        printf("D=%d warps/SM %2d: max rel err %.3e (system %d), mean %.3e, NaN %d | %.3f ms for %.0f systems = %.1f M systems/s, %.0f cycles per system per SM (%.0f per warp)\n",
               D, warps, worst, worst_s, sum / n_sys, n_nan, ms, total, total / ms * 1e-3, (double)mx / (warps * reps),
               (double)mx / reps);
        if (!(worst < 1e-5)) rc = 2;
    }
    return rc;
}

int main(int argc, char **argv) {
    const int mt = argc > 1 ? atoi(argv[1]) : 4;
    const int only = argc > 2 ? atoi(argv[2]) : 0;
    const bool h = argc > 3 && (argv[3][0] == 'h' || argv[3][0] == 'r');   // fp16-packed products (solve_tiles_h); 'r': rolled row-tile loops
    if (argc > 3 && argv[3][0] == 'r') { printf("fp16-packed, rolled\n"); return mt == 8 ? run<8, true, true>(only) : run<4, true, true>(only); }
    if (h) printf("fp16-packed variant\n");
    if (mt == 4) return h ? run<4, true>(only) : run<4, false>(only);
    if (mt == 8) return h ? run<8, true>(only) : run<8, false>(only);
    if (mt == 2) return h ? run<2, true>(only) : run<2, false>(only);
    printf("MT must be 2, 4 or 8\n");
    return 1;
}
