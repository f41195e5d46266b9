/*
 * flinkml_oracle.c -- CPU restatement (IEEE fp64) of the FlinkML 1.3.1 hot paths that
 * mmziyad/flink-ms calls.  TEST INFRASTRUCTURE ONLY.
 *
 *   *** PARITY UNPINNED ***
 *   The arithmetic lives in org.apache.flink:flink-ml_2.11:1.3.1 (flink-als/pom.xml:54,
 *   :110-114; flink-svm/pom.xml:54, :110-114), which is NOT under /root/reference, and the
 *   reference holds no tests, fixtures or golden vectors for this path.  This file restates
 *   the published algorithm and is anchored on the reference's own
 *   call sites: ALSImpl.scala:35-57 (builder, lambda=0.9, seed=42, factorsOption order),
 *   SVMImpl.scala:21-35 (builder, weights -> 1-based ids), ALSPredictRandom.java:79-92 and
 *   SVMPredictRandom.java:68-88 (predict semantics).  What IS pinned: java.util.Random
 *   (public JDK spec; KATs in tests/golden/jrand_kat.json) and the linear-algebra identities
 *   checked against numpy/LAPACK in tests/test_oracle.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this library.  The product (libflinkms_b200.so) never links or calls it.
 *
 * Build: see oracle/Makefile  (gcc -O3 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Risk register: every point of SURVEY.md Appendix A that is RECOLLECTED (not readable in
 * /root/reference) and changes results is a named switch.  Defaults = the reading this oracle
 * (and the CUDA path) implements; tests/test_oracle.py checks what each switch does, and
 * tools/make_flink_goldens.sh produces the Flink 1.3.1 outputs that would pin them.
 *   als_lambda_times_n   1: diag += lambda * n_u (ALS-WR, Zhou et al.)     0: diag += lambda
 *   als_final_user_step  1: one extra users half-step after the loop       0: none
 *   svm_scaling_configured_blocks
 *                        1: dw scaled by stepsize / configured blocks      0: / non-empty blocks
 *   svm_alpha_scale_stepsize
 *                        1: alpha += stepsize * dAlpha                     0: alpha += (stepsize / blocks) * dAlpha
 * (SVM block membership -- hash vs source subtask -- and the RNG schedule are explicit inputs.)
 * ---------------------------------------------------------------------------------------- */
static struct { int als_lambda_times_n, als_final_user_step, svm_scaling_configured_blocks, svm_alpha_scale_stepsize; } g_sw = {1, 1, 1, 1};
static int *sw_slot(const char *name) {
    if (!strcmp(name, "als_lambda_times_n")) return &g_sw.als_lambda_times_n;
    if (!strcmp(name, "als_final_user_step")) return &g_sw.als_final_user_step;
    if (!strcmp(name, "svm_scaling_configured_blocks")) return &g_sw.svm_scaling_configured_blocks;
    if (!strcmp(name, "svm_alpha_scale_stepsize")) return &g_sw.svm_alpha_scale_stepsize;
    return NULL;
}
ORC_API int orc_set_switch(const char *name, int value) { int *p = sw_slot(name); if (!p) return -1; *p = value; return 0; }
ORC_API int orc_get_switch(const char *name) { int *p = sw_slot(name); return p ? *p : -1; }

/* ------------------------------------------------------------------------------------------
 * java.util.Random  (scala.util.Random(seed) wraps it)
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint64_t s; } jrand;
#define JR_MULT 0x5DEECE66DULL
#define JR_MASK ((1ULL << 48) - 1)

static inline void jr_seed(jrand *r, int64_t seed) { r->s = ((uint64_t)seed ^ JR_MULT) & JR_MASK; }
static inline int32_t jr_next(jrand *r, int bits) {
    r->s = (r->s * JR_MULT + 0xBULL) & JR_MASK;
    return (int32_t)(int64_t)(r->s >> (48 - bits)); /* (int)(state >>> (48-bits)) */
}
static inline int32_t jr_next_int(jrand *r) { return jr_next(r, 32); }
static inline int32_t jr_next_int_bound(jrand *r, int32_t bound) {
    int32_t v = jr_next(r, 31);
    int32_t m = bound - 1;
    if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)v) >> 31);
    /* Java int arithmetic wraps; do it in uint32 to stay defined in C */
    for (int32_t u = v;; u = jr_next(r, 31)) {
        v = u % bound;
        if ((int32_t)((uint32_t)u - (uint32_t)v + (uint32_t)m) >= 0) return v;
    }
}
static inline double jr_next_double(jrand *r) {
    int64_t hi = (int64_t)jr_next(r, 26);
    int64_t lo = (int64_t)jr_next(r, 27);
    return (double)((hi << 27) + lo) * 0x1.0p-53;
}

ORC_API void orc_jrand_ints(int64_t seed, int n, int32_t *out) {
    jrand r; jr_seed(&r, seed);
    for (int i = 0; i < n; ++i) out[i] = jr_next_int(&r);
}
ORC_API void orc_jrand_bounded(int64_t seed, int32_t bound, int n, int32_t *out) {
    jrand r; jr_seed(&r, seed);
    for (int i = 0; i < n; ++i) out[i] = jr_next_int_bound(&r, bound);
}
ORC_API void orc_jrand_doubles(int64_t seed, int n, double *out) {
    jrand r; jr_seed(&r, seed);
    for (int i = 0; i < n; ++i) out[i] = jr_next_double(&r);
}

/* ------------------------------------------------------------------------------------------
 * ALS  (FlinkML recommendation.ALS)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t nnz, n_users, n_items;
    int64_t *user_ids, *item_ids;          /* sorted ascending distinct (createBlockInformation) */
    /* by-user CSR and by-item CSR, partner index ascending inside a row */
    int64_t *uptr, *iptr;
    int32_t *ucol, *icol;
    double  *uval, *ival;
    int rank;
    double *U, *I;                          /* row-major N x rank */
} orc_als;

static int cmp_i64(const void *a, const void *b) {
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}
static int64_t sort_unique(int64_t *v, int64_t n) {
    qsort(v, (size_t)n, sizeof(int64_t), cmp_i64);
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) if (i == 0 || v[i] != v[i - 1]) v[m++] = v[i];
    return m;
}
static int64_t lower_bound(const int64_t *v, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (v[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}
typedef struct { int32_t r, c; double v; int64_t ord; } triple;
static int cmp_triple(const void *a, const void *b) {
    const triple *x = a, *y = b;
    if (x->r != y->r) return (x->r > y->r) - (x->r < y->r);
    if (x->c != y->c) return (x->c > y->c) - (x->c < y->c);
    return (x->ord > y->ord) - (x->ord < y->ord);       /* duplicates keep input order */
}
static void build_csr(int64_t nnz, int64_t nrows, const int32_t *r, const int32_t *c, const double *v,
                      int64_t **ptr_o, int32_t **col_o, double **val_o) {
    triple *t = malloc((size_t)nnz * sizeof(triple));
    for (int64_t k = 0; k < nnz; ++k) { t[k].r = r[k]; t[k].c = c[k]; t[k].v = v[k]; t[k].ord = k; }
    qsort(t, (size_t)nnz, sizeof(triple), cmp_triple);
    int64_t *ptr = calloc((size_t)nrows + 1, sizeof(int64_t));
    int32_t *col = malloc((size_t)(nnz ? nnz : 1) * sizeof(int32_t));
    double *val = malloc((size_t)(nnz ? nnz : 1) * sizeof(double));
    for (int64_t k = 0; k < nnz; ++k) { ptr[t[k].r + 1]++; col[k] = t[k].c; val[k] = t[k].v; }
    for (int64_t i = 0; i < nrows; ++i) ptr[i + 1] += ptr[i];
    free(t);
    *ptr_o = ptr; *col_o = col; *val_o = val;
}

/* ALS.fit's plan up to createBlockInformation: ids -> sorted distinct, ratings regrouped by
 * user and (transposed) by item.  Duplicated (u,i) pairs are kept as separate ratings, as a
 * DataSet[(Int,Int,Double)] would (ALSImpl.scala:29-32 does not dedupe). */
ORC_API orc_als *orc_als_create(const int64_t *user, const int64_t *item, const double *rating, int64_t nnz) {
    orc_als *h = calloc(1, sizeof(orc_als));
    h->nnz = nnz;
    h->user_ids = malloc((size_t)(nnz ? nnz : 1) * sizeof(int64_t));
    h->item_ids = malloc((size_t)(nnz ? nnz : 1) * sizeof(int64_t));
    memcpy(h->user_ids, user, (size_t)nnz * sizeof(int64_t));
    memcpy(h->item_ids, item, (size_t)nnz * sizeof(int64_t));
    h->n_users = sort_unique(h->user_ids, nnz);
    h->n_items = sort_unique(h->item_ids, nnz);
    int32_t *du = malloc((size_t)(nnz ? nnz : 1) * sizeof(int32_t));
    int32_t *di = malloc((size_t)(nnz ? nnz : 1) * sizeof(int32_t));
    for (int64_t k = 0; k < nnz; ++k) {
        du[k] = (int32_t)lower_bound(h->user_ids, h->n_users, user[k]);
        di[k] = (int32_t)lower_bound(h->item_ids, h->n_items, item[k]);
    }
    build_csr(nnz, h->n_users, du, di, rating, &h->uptr, &h->ucol, &h->uval);
    build_csr(nnz, h->n_items, di, du, rating, &h->iptr, &h->icol, &h->ival);
    free(du); free(di);
    return h;
}
ORC_API void orc_als_destroy(orc_als *h) {
    if (!h) return;
    free(h->user_ids); free(h->item_ids); free(h->uptr); free(h->iptr); free(h->ucol); free(h->icol);
    free(h->uval); free(h->ival); free(h->U); free(h->I); free(h);
}
ORC_API int64_t orc_als_n_users(const orc_als *h) { return h->n_users; }
ORC_API int64_t orc_als_n_items(const orc_als *h) { return h->n_items; }

/* initialItems: new scala.util.Random(id ^ seed); Array.fill(factors)(rnd.nextDouble())
 * (Appendix A.1 `randomFactors`).  Users are NOT initialised: the first half-step solves them. */
ORC_API void orc_als_init(orc_als *h, int rank, int64_t seed) {
    h->rank = rank;
    free(h->U); free(h->I);
    h->U = calloc((size_t)(h->n_users ? h->n_users : 1) * rank, sizeof(double));
    h->I = calloc((size_t)(h->n_items ? h->n_items : 1) * rank, sizeof(double));
    for (int64_t i = 0; i < h->n_items; ++i) {
        jrand r; jr_seed(&r, h->item_ids[i] ^ seed);
        for (int k = 0; k < rank; ++k) h->I[i * rank + k] = jr_next_double(&r);
    }
}

/* In-place Cholesky solve of the dense SPD system, upper storage -- LAPACK dposv("U").
 * a: d x d, only the upper triangle (column-major "U" == row-major lower of the mirror) is read.
 * Returns nonzero when a pivot is not positive (dposv info>0); FlinkML ignores info. */
static int chol_solve(double *a, double *b, int d) {
    /* factor A = R^T R, R upper, stored row-major in a[i*d+j], j>=i */
    for (int j = 0; j < d; ++j) {
        double s = a[j * d + j];
        for (int k = 0; k < j; ++k) s -= a[k * d + j] * a[k * d + j];
        if (!(s > 0.0)) return j + 1;
        double rjj = sqrt(s);
        a[j * d + j] = rjj;
        for (int i = j + 1; i < d; ++i) {
            double t = a[j * d + i];
            for (int k = 0; k < j; ++k) t -= a[k * d + j] * a[k * d + i];
            a[j * d + i] = t / rjj;
        }
    }
    for (int i = 0; i < d; ++i) {                 /* R^T z = b */
        double t = b[i];
        for (int k = 0; k < i; ++k) t -= a[k * d + i] * b[k];
        b[i] = t / a[i * d + i];
    }
    for (int i = d - 1; i >= 0; --i) {            /* R x = z */
        double t = b[i];
        for (int k = i + 1; k < d; ++k) t -= a[i * d + k] * b[k];
        b[i] = t / a[i * d + i];
    }
    return 0;
}

/* updateFactors coGroup body for rows [row_begin,row_end) of one side (Appendix A.1):
 *   per partner vector v: outerProduct(v) -> packed triangle; daxpy into XtX[u]; daxpy(r*v) into Xy[u]
 *   per row: generateFullMatrix; diag += lambda * n_u  (ALS-WR); dposv; out(u) = Xy[u]
 * which: 0 = update users from items, 1 = update items from users.  Returns #rows with info!=0. */
ORC_API int64_t orc_als_half_step_range(orc_als *h, int which, double lambda, int64_t row_begin, int64_t row_end,
                                        int nthreads) {
    const int d = h->rank;
    const int64_t *ptr = which == 0 ? h->uptr : h->iptr;
    const int32_t *col = which == 0 ? h->ucol : h->icol;
    const double *val = which == 0 ? h->uval : h->ival;
    const double *P = which == 0 ? h->I : h->U;
    double *T = which == 0 ? h->U : h->I;
    const int tri = d * (d + 1) / 2;
    int64_t bad = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel reduction(+ : bad)
    {
        double *xtx = malloc((size_t)tri * sizeof(double));
        double *vv = malloc((size_t)tri * sizeof(double));
        double *full = malloc((size_t)d * d * sizeof(double));
        double *xy = malloc((size_t)d * sizeof(double));
#pragma omp for schedule(dynamic, 64)
        for (int64_t u = row_begin; u < row_end; ++u) {
            memset(xtx, 0, (size_t)tri * sizeof(double));
            memset(xy, 0, (size_t)d * sizeof(double));
            int64_t n = 0;
            for (int64_t k = ptr[u]; k < ptr[u + 1]; ++k) {
                const double *v = P + (int64_t)col[k] * d;
                const double r = val[k];
                int pos = 0;                                   /* outerProduct */
                for (int row = 0; row < d; ++row)
                    for (int c = 0; c <= row; ++c) vv[pos++] = v[row] * v[c];
                for (int p = 0; p < tri; ++p) xtx[p] += vv[p];  /* blas.daxpy(tri, 1.0, vv, 1, xtx, 1) */
                for (int c = 0; c < d; ++c) xy[c] += r * v[c];  /* blas.daxpy(d, r, v, 1, xy, 1) */
                ++n;
            }
            int pos = 0;                                       /* generateFullMatrix */
            for (int row = 0; row < d; ++row)
                for (int c = 0; c <= row; ++c) { full[row * d + c] = xtx[pos]; full[c * d + row] = xtx[pos]; ++pos; }
            for (int f = 0; f < d; ++f) full[f * d + f] += g_sw.als_lambda_times_n ? lambda * (double)n : lambda;
            if (chol_solve(full, xy, d)) ++bad;
            memcpy(T + u * d, xy, (size_t)d * sizeof(double));
        }
        free(xtx); free(vv); free(full); free(xy);
    }
    return bad;
}
ORC_API int64_t orc_als_half_step(orc_als *h, int which, double lambda, int nthreads) {
    return orc_als_half_step_range(h, which, lambda, 0, which == 0 ? h->n_users : h->n_items, nthreads);
}

/* fitALS schedule: items = init; iterate(T){ users = f(items); items = f(users) }; users = f(items) */
ORC_API int64_t orc_als_fit(orc_als *h, int rank, int iterations, double lambda, int64_t seed, int nthreads) {
    int64_t bad = 0;
    orc_als_init(h, rank, seed);
    for (int t = 0; t < iterations; ++t) {
        bad += orc_als_half_step(h, 0, lambda, nthreads);
        bad += orc_als_half_step(h, 1, lambda, nthreads);
    }
    if (g_sw.als_final_user_step) bad += orc_als_half_step(h, 0, lambda, nthreads);
    return bad;
}
ORC_API void orc_als_get_factors(const orc_als *h, int which, int64_t *ids, double *factors) {
    int64_t n = which == 0 ? h->n_users : h->n_items;
    if (ids) memcpy(ids, which == 0 ? h->user_ids : h->item_ids, (size_t)n * sizeof(int64_t));
    if (factors) memcpy(factors, which == 0 ? h->U : h->I, (size_t)n * h->rank * sizeof(double));
}
ORC_API void orc_als_set_factors(orc_als *h, int which, int rank, const double *factors) {
    int64_t n = which == 0 ? h->n_users : h->n_items;
    if (h->rank != rank || !h->U) {
        h->rank = rank; free(h->U); free(h->I);
        h->U = calloc((size_t)(h->n_users ? h->n_users : 1) * rank, sizeof(double));
        h->I = calloc((size_t)(h->n_items ? h->n_items : 1) * rank, sizeof(double));
    }
    memcpy(which == 0 ? h->U : h->I, factors, (size_t)n * rank * sizeof(double));
}
/* ALS.predict: inner join on both factor sets, blas.ddot (sequential sum, index ascending);
 * serve-side twin ALSPredictRandom.java:79-92.  found[k]=0 => id missing from the model. */
ORC_API void orc_als_predict(const orc_als *h, const int64_t *u, const int64_t *i, int64_t n, double *out,
                             uint8_t *found) {
    const int d = h->rank;
    for (int64_t k = 0; k < n; ++k) {
        int64_t a = lower_bound(h->user_ids, h->n_users, u[k]);
        int64_t b = lower_bound(h->item_ids, h->n_items, i[k]);
        int ok = a < h->n_users && h->user_ids[a] == u[k] && b < h->n_items && h->item_ids[b] == i[k];
        double s = 0.0;
        if (ok) for (int j = 0; j < d; ++j) s += h->U[a * d + j] * h->I[b * d + j];
        out[k] = ok ? s : 0.0;
        if (found) found[k] = (uint8_t)ok;
    }
}
/* ALS.empiricalRisk (never called by the reference; used by tests as a monotonicity check):
 * sum residual^2 + lambda * (sum_u n_u |x_u|^2 + sum_i n_i |y_i|^2) */
ORC_API double orc_als_risk(const orc_als *h, double lambda) {
    const int d = h->rank;
    double loss = 0.0, reg = 0.0;
    for (int64_t u = 0; u < h->n_users; ++u) {
        double nu = (double)(h->uptr[u + 1] - h->uptr[u]), s2 = 0.0;
        for (int j = 0; j < d; ++j) s2 += h->U[u * d + j] * h->U[u * d + j];
        reg += nu * s2;
        for (int64_t k = h->uptr[u]; k < h->uptr[u + 1]; ++k) {
            double p = 0.0;
            for (int j = 0; j < d; ++j) p += h->U[u * d + j] * h->I[(int64_t)h->ucol[k] * d + j];
            loss += (h->uval[k] - p) * (h->uval[k] - p);
        }
    }
    for (int64_t i = 0; i < h->n_items; ++i) {
        double ni = (double)(h->iptr[i + 1] - h->iptr[i]), s2 = 0.0;
        for (int j = 0; j < d; ++j) s2 += h->I[i * d + j] * h->I[i * d + j];
        reg += ni * s2;
    }
    return loss + lambda * reg;
}

/* ------------------------------------------------------------------------------------------
 * SVM  (FlinkML classification.SVM, CoCoA + SDCA)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int x_dtype;               /* 0 = dense float32 row-major, 1 = dense float64 row-major, 2 = CSR (val float64) */
    const void *X;
    const int64_t *csr_ptr;
    const int32_t *csr_idx;
    int32_t dim;
} xmat;

static inline double x_dot(const xmat *m, int64_t row, const double *w) {
    double s = 0.0;
    if (m->x_dtype == 0) { const float *x = (const float *)m->X + row * (int64_t)m->dim; for (int j = 0; j < m->dim; ++j) s += (double)x[j] * w[j]; }
    else if (m->x_dtype == 1) { const double *x = (const double *)m->X + row * (int64_t)m->dim; for (int j = 0; j < m->dim; ++j) s += x[j] * w[j]; }
    else { const double *v = (const double *)m->X; for (int64_t p = m->csr_ptr[row]; p < m->csr_ptr[row + 1]; ++p) s += v[p] * w[m->csr_idx[p]]; }
    return s;
}
static inline double x_sq(const xmat *m, int64_t row) {
    double s = 0.0;
    if (m->x_dtype == 0) { const float *x = (const float *)m->X + row * (int64_t)m->dim; for (int j = 0; j < m->dim; ++j) s += (double)x[j] * (double)x[j]; }
    else if (m->x_dtype == 1) { const double *x = (const double *)m->X + row * (int64_t)m->dim; for (int j = 0; j < m->dim; ++j) s += x[j] * x[j]; }
    else { const double *v = (const double *)m->X; for (int64_t p = m->csr_ptr[row]; p < m->csr_ptr[row + 1]; ++p) s += v[p] * v[p]; }
    return s;
}
static inline void x_axpy(const xmat *m, int64_t row, double c, double *w, double *dw) {
    if (m->x_dtype == 0) { const float *x = (const float *)m->X + row * (int64_t)m->dim; for (int j = 0; j < m->dim; ++j) { double t = c * (double)x[j]; w[j] += t; dw[j] += t; } }
    else if (m->x_dtype == 1) { const double *x = (const double *)m->X + row * (int64_t)m->dim; for (int j = 0; j < m->dim; ++j) { double t = c * x[j]; w[j] += t; dw[j] += t; } }
    else { const double *v = (const double *)m->X; for (int64_t p = m->csr_ptr[row]; p < m->csr_ptr[row + 1]; ++p) { double t = c * v[p]; w[m->csr_idx[p]] += t; dw[m->csr_idx[p]] += t; } }
}

/* One localDualMethod.map call: H sequential SDCA steps (maximize) on a private copy of w.
 * rows: the block's member rows in arrival order.  alpha: the block's persistent alpha (len nk).
 * Updates alpha (alpha += stepsize * dAlpha) and writes dw (len dim). */
static void sdca_block_round(const xmat *m, const double *y, const int64_t *rows, int64_t nk, double *alpha,
                             jrand *rng, int64_t H, double lambda, double n_total, double alpha_scale,
                             const double *w0, double *w, double *dw, double *tmp_alpha, double *d_alpha) {
    const int D = m->dim;
    memcpy(w, w0, (size_t)D * sizeof(double));
    memset(dw, 0, (size_t)D * sizeof(double));
    memcpy(tmp_alpha, alpha, (size_t)nk * sizeof(double));
    memset(d_alpha, 0, (size_t)nk * sizeof(double));
    const double ln = lambda * n_total;
    for (int64_t it = 0; it < H; ++it) {
        int32_t idx = jr_next_int_bound(rng, (int32_t)nk);
        int64_t row = rows[idx];
        double yy = y[row], a = tmp_alpha[idx];
        /* maximize(x, y, lambda, a, w, n) */
        double g = (yy * x_dot(m, row, w) - 1.0) * ln;
        if (g != 0.0) {
            double q = x_sq(m, row);
            double an = (q != 0.0) ? fmin(fmax(a - g / q, 0.0), 1.0) : 1.0;
            double da = an - a;
            x_axpy(m, row, yy * da / ln, w, dw);
            tmp_alpha[idx] += da;
            d_alpha[idx] += da;
        }
    }
    for (int64_t i = 0; i < nk; ++i) alpha[i] += d_alpha[i] * alpha_scale;
}

/*
 * SVM.fit.  block_id[i] in [0,blocks) gives the CoCoA block of row i (FlinkMLTools.block assigns by
 * SOURCE SUBTASK index, which is an input-split property -- so it is an explicit input here).
 * local_iterations <= 0 means "H = n_k" (one epoch per round; bench configuration).
 * rng_mode 0: one java.util.Random(seed ^ k) per block k, persisting across rounds
 *             (one block per subtask, `r = new Random(seed ^ getIndexOfThisSubtask)`).
 * rng_mode 1: one shared Random(seed ^ 0) drawn by blocks in ascending order (all blocks on subtask 0).
 * scaling = stepsize / blocks with the CONFIGURED block count.  alpha_out (len n, optional) returns alpha
 * in row order.  Returns 0, or -1 on bad arguments.
 */
ORC_API int orc_svm_fit(const void *X, int x_dtype, const int64_t *csr_ptr, const int32_t *csr_idx, const double *y,
                        int64_t n, int32_t dim, const int32_t *block_id, int blocks, int iterations,
                        int64_t local_iterations, double regularization, double stepsize, int64_t seed, int rng_mode,
                        int nthreads, double *w_out, double *alpha_out) {
    if (n < 0 || dim <= 0 || blocks <= 0) return -1;
    xmat m = { x_dtype, X, csr_ptr, csr_idx, dim };
    /* group rows by block, arrival (input) order inside a block */
    int64_t *cnt = calloc((size_t)blocks + 1, sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) { if (block_id[i] < 0 || block_id[i] >= blocks) { free(cnt); return -1; } cnt[block_id[i] + 1]++; }
    for (int k = 0; k < blocks; ++k) cnt[k + 1] += cnt[k];
    int64_t *rows = malloc((size_t)(n ? n : 1) * sizeof(int64_t));
    int64_t *fill = malloc((size_t)blocks * sizeof(int64_t));
    memcpy(fill, cnt, (size_t)blocks * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) rows[fill[block_id[i]]++] = i;
    free(fill);
    double *alpha = calloc((size_t)(n ? n : 1), sizeof(double));      /* block-grouped order */
    double *tmpa = malloc((size_t)(n ? n : 1) * sizeof(double));
    double *dalp = malloc((size_t)(n ? n : 1) * sizeof(double));
    double *w = calloc((size_t)dim, sizeof(double));
    double *dws = malloc((size_t)blocks * dim * sizeof(double));
    jrand *rngs = malloc((size_t)blocks * sizeof(jrand));
    for (int k = 0; k < blocks; ++k) jr_seed(&rngs[k], seed ^ (int64_t)k);
    jrand shared; jr_seed(&shared, seed ^ 0);
    int nonempty = 0;
    for (int k = 0; k < blocks; ++k) nonempty += cnt[k + 1] > cnt[k];
    const double scaling = stepsize / (double)(g_sw.svm_scaling_configured_blocks ? blocks : (nonempty ? nonempty : 1));
    const double alpha_scale = g_sw.svm_alpha_scale_stepsize ? stepsize : stepsize / (double)blocks;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    for (int t = 0; t < iterations; ++t) {
        if (rng_mode == 0) {
#pragma omp parallel
            {
                double *wl = malloc((size_t)dim * sizeof(double));
#pragma omp for schedule(dynamic, 1)
                for (int k = 0; k < blocks; ++k) {
                    int64_t nk = cnt[k + 1] - cnt[k];
                    double *dw = dws + (int64_t)k * dim;
                    if (nk == 0) { memset(dw, 0, (size_t)dim * sizeof(double)); continue; }   /* empty block: no map call */
                    int64_t H = local_iterations > 0 ? local_iterations : nk;
                    sdca_block_round(&m, y, rows + cnt[k], nk, alpha + cnt[k], &rngs[k], H, regularization, (double)n,
                                     alpha_scale, w, wl, dw, tmpa + cnt[k], dalp + cnt[k]);
                }
                free(wl);
            }
        } else {
            double *wl = malloc((size_t)dim * sizeof(double));
            for (int k = 0; k < blocks; ++k) {
                int64_t nk = cnt[k + 1] - cnt[k];
                double *dw = dws + (int64_t)k * dim;
                if (nk == 0) { memset(dw, 0, (size_t)dim * sizeof(double)); continue; }
                int64_t H = local_iterations > 0 ? local_iterations : nk;
                sdca_block_round(&m, y, rows + cnt[k], nk, alpha + cnt[k], &shared, H, regularization, (double)n,
                                 alpha_scale, w, wl, dw, tmpa + cnt[k], dalp + cnt[k]);
            }
            free(wl);
        }
        /* w.union(dws.map(_ * scaling)).reduce(_ + _): summed in ascending block order */
        for (int k = 0; k < blocks; ++k)
            for (int j = 0; j < dim; ++j) w[j] += dws[(int64_t)k * dim + j] * scaling;
    }
    memcpy(w_out, w, (size_t)dim * sizeof(double));
    if (alpha_out) for (int64_t p = 0; p < n; ++p) alpha_out[rows[p]] = alpha[p];
    free(cnt); free(rows); free(alpha); free(tmpa); free(dalp); free(w); free(dws); free(rngs);
    return 0;
}

/* The sample-index stream one block draws: `r.nextInt(block.size)` H times per round (bit-exact KAT aid). */
ORC_API void orc_svm_index_stream(int64_t seed, int block, int32_t nk, int64_t count, int32_t *out) {
    jrand r; jr_seed(&r, seed ^ (int64_t)block);
    for (int64_t i = 0; i < count; ++i) out[i] = jr_next_int_bound(&r, nk);
}

/* SVM.predict: raw = x.w ; decision function or (raw > threshold ? 1.0 : -1.0)  -- STRICT '>'
 * (serve-side twin SVMPredictRandom.java:83-88). */
ORC_API void orc_svm_predict(const void *X, int x_dtype, const int64_t *csr_ptr, const int32_t *csr_idx, int64_t n,
                             int32_t dim, const double *w, double threshold, int output_decision_function,
                             double *out) {
    xmat m = { x_dtype, X, csr_ptr, csr_idx, dim };
    for (int64_t i = 0; i < n; ++i) {
        double raw = x_dot(&m, i, w);
        out[i] = output_decision_function ? raw : (raw > threshold ? 1.0 : -1.0);
    }
}

/* Dual objective D(alpha) = sum(alpha)/n - lambda/2 |w(alpha)|^2 with w = sum alpha_i y_i x_i/(lambda n),
 * and primal P(w) = lambda/2 |w|^2 + mean hinge -- for monotonicity / duality-gap tests. */
ORC_API void orc_svm_objectives(const void *X, int x_dtype, const int64_t *csr_ptr, const int32_t *csr_idx,
                                const double *y, int64_t n, int32_t dim, const double *w, const double *alpha,
                                double lambda, double *primal, double *dual) {
    xmat m = { x_dtype, X, csr_ptr, csr_idx, dim };
    double nw = 0.0, hinge = 0.0, sa = 0.0;
    for (int j = 0; j < dim; ++j) nw += w[j] * w[j];
    for (int64_t i = 0; i < n; ++i) {
        double mg = 1.0 - y[i] * x_dot(&m, i, w);
        hinge += mg > 0.0 ? mg : 0.0;
        if (alpha) sa += alpha[i];
    }
    *primal = 0.5 * lambda * nw + hinge / (double)n;
    if (dual) {
        double *wa = calloc((size_t)dim, sizeof(double)), *scratch = calloc((size_t)dim, sizeof(double));
        if (alpha) for (int64_t i = 0; i < n; ++i) x_axpy(&m, i, alpha[i] * y[i] / (lambda * (double)n), wa, scratch);
        double nwa = 0.0;
        for (int j = 0; j < dim; ++j) nwa += wa[j] * wa[j];
        *dual = sa / (double)n - 0.5 * lambda * nwa;
        free(wa); free(scratch);
    }
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
