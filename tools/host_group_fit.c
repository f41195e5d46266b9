/* A plain C host (no Python, no torch, no NCCL) driving libflinkms_b200.so: ALS + SVM fits on ONE GPU and on a GROUP of n GPUs
 * through the same entry points, results compared.  What a JVM host does through the JNI shim.
 *   gcc -O2 -I include tools/host_group_fit.c -o tools/host_group_fit -L flink-ms_b200 -lflinkms_b200 -Wl,-rpath,$PWD/flink-ms_b200 -lm
 *   tools/host_group_fit <n_devices> [users items nnz rank]                                                                   */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "flinkms.h"

static unsigned long long lcg = 88172645463325252ULL;
static unsigned int rnd(void) { lcg ^= lcg << 13; lcg ^= lcg >> 7; lcg ^= lcg << 17; return (unsigned int)(lcg >> 16); }
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
#define CHECK(ctx, call) do { int rc_ = (call); if (rc_ != FMS_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, fms_last_error(ctx)); return 1; } } while (0)

int main(int argc, char **argv) {
    const int n_dev = argc > 1 ? atoi(argv[1]) : 2;
    const long long U = argc > 2 ? atoll(argv[2]) : 20000, I = argc > 3 ? atoll(argv[3]) : 3000, nnz = argc > 4 ? atoll(argv[4]) : 1000000;
    const int rank = argc > 5 ? atoi(argv[5]) : 64, T = 3;
    int64_t *u = malloc(nnz * 8), *it = malloc(nnz * 8);
    double *r = malloc(nnz * 8);
    for (long long k = 0; k < nnz; ++k) { u[k] = rnd() % U; it[k] = rnd() % I; r[k] = 0.5 * (1 + rnd() % 10); }
    int ids[64];
    for (int g = 0; g < n_dev; ++g) ids[g] = g;
    double *F[2][2];
    int64_t nu[2], ni[2];
    for (int mode = 0; mode < 2; ++mode) {                        /* 0: one GPU, 1: the group */
        fms_ctx *ctx = NULL;
        const int rc = fms_create(&ctx, ids, mode ? n_dev : 1);
        if (rc != FMS_OK) { fprintf(stderr, "fms_create(%d devices) -> %d: %s\n", mode ? n_dev : 1, rc, fms_last_error(NULL)); return 1; }
        double t0 = now();
        CHECK(ctx, fms_als_fit(ctx, u, it, r, nnz, rank, T, 0.9, 42, &nu[mode], &ni[mode]));
        double t1 = now();
        F[mode][0] = malloc(nu[mode] * rank * 8); F[mode][1] = malloc(ni[mode] * rank * 8);
        CHECK(ctx, fms_als_get_factors(ctx, FMS_USERS, NULL, F[mode][0]));
        CHECK(ctx, fms_als_get_factors(ctx, FMS_ITEMS, NULL, F[mode][1]));
        printf("ALS rank %d, %lld ratings, %d device(s): fit %.3f s, %lld kernel launches\n", rank, nnz, mode ? n_dev : 1, t1 - t0, (long long)fms_launch_count(ctx));
        fms_destroy(ctx);
    }
    if (nu[0] != nu[1] || ni[0] != ni[1]) { fprintf(stderr, "row counts differ\n"); return 1; }
    const int same_u = memcmp(F[0][0], F[1][0], nu[0] * rank * 8) == 0, same_i = memcmp(F[0][1], F[1][1], ni[0] * rank * 8) == 0;
    printf("ALS group == single GPU, bitwise: users %s, items %s\n", same_u ? "yes" : "NO", same_i ? "yes" : "NO");
    /* SVM */
    const long long n = 8192; const int dim = 256, K = 64;
    float *X = malloc(n * dim * 4); double *y = malloc(n * 8); int32_t *bid = malloc(n * 4);
    double w[2][256];
    for (long long i = 0; i < n; ++i) {
        double s = 0;
        for (int j = 0; j < dim; ++j) { X[i * dim + j] = (float)((int)(rnd() % 2001) - 1000) / 1000.f; s += X[i * dim + j] * ((j % 7) - 3); }
        y[i] = s >= 0 ? 1.0 : -1.0; bid[i] = (int32_t)(i % K);
    }
    for (int mode = 0; mode < 2; ++mode) {
        fms_ctx *ctx = NULL;
        if (fms_create(&ctx, ids, mode ? n_dev : 1) != FMS_OK) { fprintf(stderr, "fms_create: %s\n", fms_last_error(NULL)); return 1; }
        CHECK(ctx, fms_svm_fit(ctx, X, FMS_X_DENSE_F32, NULL, NULL, y, n, dim, bid, K, 5, -1, 1.0, 1.0, 42, FMS_RNG_PER_BLOCK, w[mode]));
        fms_destroy(ctx);
    }
    double dmax = 0, wmax = 0;
    for (int j = 0; j < dim; ++j) { dmax = fmax(dmax, fabs(w[0][j] - w[1][j])); wmax = fmax(wmax, fabs(w[0][j])); }
    printf("SVM group vs single GPU: max |dw| / max |w| = %.3e\n", dmax / wmax);
    return (same_u && same_i && dmax / wmax < 1e-5) ? 0 : 2;
}
