// N4 -- GPU-side ingest of the two text formats the reference jobs read before training:
//   * ratings CSV  `user<delim>item<delim>rating`  (env.readCsvFile[(Int, Int, Double)], ALSImpl.scala:29-32, with
//     fieldDelimiter and ignoreFirstLine, :22-26)
//   * LibSVM       `label idx:val idx:val ...`      (env.readLibSVM, SVMImpl.scala:21; FlinkML 1.3.1 MLUtils.readLibSVM:
//     '#' starts a comment, blank lines are skipped, indices are 1-based in the file and 0-based in the vector,
//     dimension = largest index)
// One thread per line.  Integers follow Integer.parseInt (optional sign, digits, 32-bit range), doubles follow
// Double.parseDouble for decimal literals ([+-]digits[.digits][(e|E)[+-]digits], NaN, Infinity): correctly rounded
// with Clinger's exact fast path and the Eisel-Lemire algorithm (128-bit power-of-ten table, tools/gen_pow10_table.py).
// The rare literals Eisel-Lemire cannot decide (exact half-way cases, subnormal / overflowing magnitudes, hex floats)
// are reported as FMS_ERR_UNSUPPORTED with their line number -- there is no CPU fallback.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "common.cuh"
#include "pow10_table.cuh"

namespace fms {
namespace {

enum : int { ST_OK = 0, ST_SKIP = 1, ST_BAD = 2, ST_UNSUPPORTED = 3 };

__device__ __forceinline__ bool is_digit(char c) { return c >= '0' && c <= '9'; }

// Eisel-Lemire: w * 10^q -> nearest double, or false when the 128-bit approximation cannot decide.
__device__ bool eisel_lemire(uint64_t w, int q, double *out) {
    if (w == 0) { *out = 0.0; return true; }
    if (q < POW10_MIN || q > POW10_MAX) return false;
    const int lz = __clzll((long long)w);
    w <<= lz;
    uint64_t exp2 = (uint64_t)(((217706 * q) >> 16) + 64 + 1023) - (uint64_t)lz;
    const uint64_t t_hi = POW10_TABLE[q - POW10_MIN][0], t_lo = POW10_TABLE[q - POW10_MIN][1];
    uint64_t x_hi = __umul64hi(w, t_hi), x_lo = w * t_hi;
    if ((x_hi & 0x1FF) == 0x1FF && x_lo + w < w) {                      // wider approximation
        const uint64_t y_hi = __umul64hi(w, t_lo), y_lo = w * t_lo;
        uint64_t m_hi = x_hi, m_lo = x_lo + y_hi;
        if (m_lo < x_lo) ++m_hi;
        if ((m_hi & 0x1FF) == 0x1FF && m_lo + 1 == 0 && y_lo + w < w) return false;
        x_hi = m_hi; x_lo = m_lo;
    }
    const uint64_t msb = x_hi >> 63;
    uint64_t mant = x_hi >> (msb + 9);
    exp2 -= 1 ^ msb;
    // exact half-way case (round bit set, nothing below): decidable only where the product is exact (5^|q| fits the
    // table entry, q in [-4, 23]) -- then round to even, i.e. do not round up; elsewhere the 128 bits cannot tell
    const bool maybe_tie = x_lo == 0 && (x_hi & 0x1FF) == 0 && (mant & 3) == 1;
    if (maybe_tie && (q < -4 || q > 23)) return false;
    const bool tie = maybe_tie && (x_hi & ((1ull << (msb + 9)) - 1)) == 0;    // every bit below the round bit is zero
    if (!tie) mant += mant & 1;
    mant >>= 1;
    if (mant >> 53) { mant >>= 1; exp2 += 1; }
    if (exp2 - 1 >= 0x7FF - 1) return false;                             // subnormal or overflow: not decided here
    const uint64_t bits = (exp2 << 52) | (mant & 0x000FFFFFFFFFFFFFull);
    *out = __longlong_as_double((long long)bits);
    return true;
}

// [p, e) -> double.  Returns ST_OK / ST_BAD (not a literal) / ST_UNSUPPORTED (a literal this parser cannot round).
__device__ int parse_double(const char *p, const char *e, double *out) {
    if (p >= e) return ST_BAD;
    bool neg = false;
    if (*p == '+' || *p == '-') { neg = *p == '-'; ++p; }
    if (p >= e) return ST_BAD;
    const int len = (int)(e - p);
    if (len == 3 && p[0] == 'N' && p[1] == 'a' && p[2] == 'N') { *out = __longlong_as_double(0x7ff8000000000000ll); return ST_OK; }
    if (len == 8 && p[0] == 'I' && p[1] == 'n' && p[2] == 'f' && p[3] == 'i' && p[4] == 'n' && p[5] == 'i' && p[6] == 't' && p[7] == 'y') {
        *out = neg ? -__longlong_as_double(0x7ff0000000000000ll) : __longlong_as_double(0x7ff0000000000000ll);
        return ST_OK;
    }
    uint64_t mant = 0;
    int nd = 0, exp10 = 0, digits_seen = 0;
    bool dropped_nonzero = false;
    for (; p < e && is_digit(*p); ++p) {
        ++digits_seen;
        if (nd < 19) { mant = mant * 10 + (uint64_t)(*p - '0'); if (mant) ++nd; }
        else { ++exp10; dropped_nonzero |= *p != '0'; }
    }
    if (p < e && *p == '.') {
        ++p;
        for (; p < e && is_digit(*p); ++p) {
            ++digits_seen;
            if (nd < 19) { mant = mant * 10 + (uint64_t)(*p - '0'); if (mant) ++nd; --exp10; }
            else dropped_nonzero |= *p != '0';
        }
    }
    if (digits_seen == 0) return ST_BAD;
    if (p < e && (*p == 'e' || *p == 'E')) {
        ++p;
        bool eneg = false;
        if (p < e && (*p == '+' || *p == '-')) { eneg = *p == '-'; ++p; }
        if (p >= e || !is_digit(*p)) return ST_BAD;
        int ex = 0;
        for (; p < e && is_digit(*p); ++p) if (ex < 100000) ex = ex * 10 + (*p - '0');
        exp10 += eneg ? -ex : ex;
    }
    if (p < e && (*p == 'd' || *p == 'D' || *p == 'f' || *p == 'F')) ++p;   // Java's type suffixes
    if (p != e) return ST_BAD;
    double v;
    if (mant == 0) v = 0.0;
    else if (!dropped_nonzero && mant < (1ull << 53) && exp10 >= -22 && exp10 <= 22) {
        // Clinger: both operands exact in binary64, one correctly rounded operation
        const double p10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18,
                                1e19, 1e20, 1e21, 1e22};
        v = exp10 < 0 ? __ddiv_rn((double)mant, p10[-exp10]) : __dmul_rn((double)mant, p10[exp10]);
    } else {
        if (!eisel_lemire(mant, exp10, &v)) return ST_UNSUPPORTED;
        if (dropped_nonzero) {                                           // more than 19 digits: both neighbours must agree
            double v2;
            if (!eisel_lemire(mant + 1, exp10, &v2) || v2 != v) return ST_UNSUPPORTED;
        }
    }
    *out = neg ? -v : v;
    return ST_OK;
}

// Integer.parseInt
__device__ int parse_int32(const char *p, const char *e, int64_t *out) {
    if (p >= e) return ST_BAD;
    bool neg = false;
    if (*p == '+' || *p == '-') { neg = *p == '-'; ++p; }
    if (p >= e) return ST_BAD;
    int64_t v = 0;
    for (; p < e; ++p) {
        if (!is_digit(*p)) return ST_BAD;
        v = v * 10 + (*p - '0');
        if (v > (int64_t)1 << 31) return ST_BAD;
    }
    v = neg ? -v : v;
    if (v > 2147483647ll || v < -2147483648ll) return ST_BAD;
    *out = v;
    return ST_OK;
}

struct IsNewline {
    const char *t;
    __device__ bool operator()(int64_t i) const { return t[i] == '\n'; }
};

__device__ __forceinline__ void line_span(const int64_t *nl, int64_t n_nl, int64_t nbytes, int64_t k, int64_t *b, int64_t *e) {
    *b = k == 0 ? 0 : nl[k - 1] + 1;
    *e = k < n_nl ? nl[k] : nbytes;
}

__global__ void k_csv_lines(const char *__restrict__ t, int64_t nbytes, const int64_t *__restrict__ nl, int64_t n_nl, int64_t n_lines,
                            char delim, int skip_first, int64_t *__restrict__ u, int64_t *__restrict__ it, double *__restrict__ r,
                            int *__restrict__ status, unsigned long long *__restrict__ first_bad) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_lines) return;
    int64_t b, e;
    line_span(nl, n_nl, nbytes, k, &b, &e);
    if (e > b && t[e - 1] == '\r') --e;
    int st = ST_OK;
    if ((k == 0 && skip_first) || e == b) st = ST_SKIP;
    else {
        const char *p = t + b, *end = t + e;
        const char *f1 = p;
        while (f1 < end && *f1 != delim) ++f1;
        const char *f2 = f1 < end ? f1 + 1 : end;
        while (f2 < end && *f2 != delim) ++f2;
        if (f1 >= end || f2 >= end) st = ST_BAD;
        else {
            const char *f3 = f2 + 1;
            const char *f3e = f3;
            while (f3e < end && *f3e != delim) ++f3e;                     // further columns are ignored like unselected CSV fields
            int64_t a = 0, c = 0;
            double v = 0.0;
            st = parse_int32(p, f1, &a);
            if (st == ST_OK) st = parse_int32(f1 + 1, f2, &c);
            if (st == ST_OK) st = parse_double(f3, f3e, &v);
            u[k] = a; it[k] = c; r[k] = v;
        }
    }
    status[k] = st;
    if (st >= ST_BAD) atomicMin(first_bad, ((unsigned long long)k << 2) | (unsigned long long)st);
}

__global__ void k_scatter_csv(int64_t n_lines, const int *__restrict__ status, const int64_t *__restrict__ pos, const int64_t *__restrict__ u,
                              const int64_t *__restrict__ it, const double *__restrict__ r, int64_t *__restrict__ ou, int64_t *__restrict__ oi,
                              double *__restrict__ orr) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_lines || status[k] != ST_OK) return;
    const int64_t o = pos[k];
    ou[o] = u[k]; oi[o] = it[k]; orr[o] = r[k];
}

__global__ void k_ok_flags(int64_t n, const int *__restrict__ status, int64_t *__restrict__ flag) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) flag[k] = status[k] == ST_OK ? 1 : 0;
}

// LibSVM: trims, strips the comment, then either counts the features of the line (fill == 0) or parses them.
__global__ void k_libsvm_lines(const char *__restrict__ t, int64_t nbytes, const int64_t *__restrict__ nl, int64_t n_nl, int64_t n_lines, int fill,
                               int *__restrict__ status, int64_t *__restrict__ cnt, const int64_t *__restrict__ row_of, const int64_t *__restrict__ ptr,
                               double *__restrict__ label, int32_t *__restrict__ idx, double *__restrict__ val, int *__restrict__ max_idx,
                               unsigned long long *__restrict__ first_bad) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_lines) return;
    if (fill && status[k] != ST_OK) return;
    int64_t b, e;
    line_span(nl, n_nl, nbytes, k, &b, &e);
    const char *p = t + b, *end = t + e;
    for (const char *c = p; c < end; ++c) if (*c == '#') { end = c; break; }       // line.takeWhile(_ != '#')
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;                 // .trim
    while (end > p && (end[-1] == ' ' || end[-1] == '\t' || end[-1] == '\r')) --end;
    if (p == end) { if (!fill) { status[k] = ST_SKIP; cnt[k] = 0; } return; }
    int st = ST_OK;
    const char *q = p;
    while (q < end && *q != ' ') ++q;
    double lab = 0.0;
    st = parse_double(p, q, &lab);
    int64_t nfeat = 0;
    const int64_t base = fill ? ptr[row_of[k]] : 0;
    int mx = 0;
    while (st == ST_OK && q < end) {
        while (q < end && *q == ' ') ++q;
        if (q >= end) break;
        const char *tok = q;
        while (q < end && *q != ' ') ++q;
        const char *colon = tok;
        while (colon < q && *colon != ':') ++colon;
        if (colon >= q) { st = ST_BAD; break; }
        int64_t fi = 0;
        double fv = 0.0;
        st = parse_int32(tok, colon, &fi);
        if (st == ST_OK && fi < 1) st = ST_BAD;                          // 1-based in the file
        if (st == ST_OK) st = parse_double(colon + 1, q, &fv);
        if (st != ST_OK) break;
        if (fill) { idx[base + nfeat] = (int32_t)(fi - 1); val[base + nfeat] = fv; }
        mx = max(mx, (int)fi);
        ++nfeat;
    }
    if (!fill) {
        status[k] = st;
        cnt[k] = st == ST_OK ? nfeat : 0;
        if (st >= ST_BAD) atomicMin(first_bad, ((unsigned long long)k << 2) | (unsigned long long)st);
        else atomicMax(max_idx, mx);
    } else {
        label[row_of[k]] = lab;
    }
}

__global__ void k_row_ptr_from_cnt(int64_t n_lines, const int *__restrict__ status, const int64_t *__restrict__ row_of, const int64_t *__restrict__ off,
                                   int64_t *__restrict__ ptr) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n_lines && status[k] == ST_OK) ptr[row_of[k]] = off[k];
}

struct LineIndex {
    DevBuf<char> text;
    DevBuf<int64_t> nl;
    int64_t n_nl = 0, n_lines = 0;
};

int index_lines(fms_ctx *ctx, const char *h_text, int64_t nbytes, LineIndex &L) {
    cudaStream_t st = ctx->stream;
    FMS_CUDA(ctx, L.text.alloc((size_t)nbytes + 1));
    FMS_CUDA(ctx, cudaMemcpyAsync(L.text.p, h_text, (size_t)nbytes, cudaMemcpyDefault, st));
    FMS_CUDA(ctx, L.nl.alloc((size_t)nbytes + 1));
    DevBuf<int64_t> cnt;
    DevBuf<unsigned char> temp;
    FMS_CUDA(ctx, cnt.alloc(1));
    thrust::counting_iterator<int64_t> it(0);
    size_t tb = 0;
    FMS_CUDA(ctx, cub::DeviceSelect::If(nullptr, tb, it, L.nl.p, cnt.p, nbytes, IsNewline{L.text.p}, st));
    FMS_CUDA(ctx, temp.alloc(tb));
    FMS_CUDA(ctx, cub::DeviceSelect::If(temp.p, tb, it, L.nl.p, cnt.p, nbytes, IsNewline{L.text.p}, st));
    ctx->launches += 1;
    char last = '\n';
    FMS_CUDA(ctx, cudaMemcpyAsync(&L.n_nl, cnt.p, 8, cudaMemcpyDeviceToHost, st));
    if (nbytes > 0) FMS_CUDA(ctx, cudaMemcpyAsync(&last, L.text.p + nbytes - 1, 1, cudaMemcpyDeviceToHost, st));
    FMS_CUDA(ctx, cudaStreamSynchronize(st));
    L.n_lines = L.n_nl + (nbytes > 0 && last != '\n' ? 1 : 0);
    return FMS_OK;
}

int report_bad(fms_ctx *ctx, unsigned long long packed, const char *what) {
    const long long line = (long long)(packed >> 2) + 1;
    if ((packed & 3) == ST_UNSUPPORTED)
        return fail(ctx, FMS_ERR_UNSUPPORTED, "%s: line %lld holds a numeric literal this parser cannot round exactly", what, line);
    return fail(ctx, FMS_ERR_INVALID_ARG, "%s: line %lld is malformed", what, line);
}

int exclusive_scan(fms_ctx *ctx, const int64_t *in, int64_t *out, int64_t n, DevBuf<unsigned char> &temp) {
    size_t tb = 0;
    FMS_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, ctx->stream));
    FMS_CUDA(ctx, temp.alloc(tb));
    FMS_CUDA(ctx, cub::DeviceScan::ExclusiveSum(temp.p, tb, in, out, n, ctx->stream));
    ctx->launches += 1;
    return FMS_OK;
}

}  // namespace

int ingest_ratings_csv(fms_ctx *ctx, const char *text, int64_t nbytes, char delimiter, int ignore_first_line, int64_t *user, int64_t *item,
                       double *rating, int64_t capacity, int64_t *n_rows) {
    cudaStream_t st = ctx->stream;
    if (n_rows) *n_rows = 0;
    if (nbytes == 0) return FMS_OK;
    LineIndex L;
    FMS_TRY(index_lines(ctx, text, nbytes, L));
    const int64_t n = L.n_lines;
    if (n == 0) return FMS_OK;
    DevBuf<int64_t> u, it, flag, pos, ou, oi;
    DevBuf<double> r, orr;
    DevBuf<int> status;
    DevBuf<unsigned long long> bad;
    DevBuf<unsigned char> temp;
    FMS_CUDA(ctx, u.alloc((size_t)n)); FMS_CUDA(ctx, it.alloc((size_t)n)); FMS_CUDA(ctx, r.alloc((size_t)n));
    FMS_CUDA(ctx, status.alloc((size_t)n)); FMS_CUDA(ctx, flag.alloc((size_t)n + 1)); FMS_CUDA(ctx, pos.alloc((size_t)n + 1));
    FMS_CUDA(ctx, bad.alloc(1));
    FMS_CUDA(ctx, cudaMemsetAsync(bad.p, 0xff, 8, st));
    FMS_CUDA(ctx, cudaMemsetAsync(flag.p, 0, (size_t)(n + 1) * 8, st));
    const unsigned grid = (unsigned)((n + 127) / 128);
    k_csv_lines<<<grid, 128, 0, st>>>(L.text.p, nbytes, L.nl.p, L.n_nl, n, delimiter, ignore_first_line, u.p, it.p, r.p, status.p, bad.p);
    FMS_LAUNCH_CHECK(ctx);
    k_ok_flags<<<grid, 128, 0, st>>>(n, status.p, flag.p);
    FMS_LAUNCH_CHECK(ctx);
    FMS_TRY(exclusive_scan(ctx, flag.p, pos.p, n + 1, temp));
    unsigned long long h_bad = 0;
    int64_t total = 0;
    FMS_CUDA(ctx, cudaMemcpyAsync(&h_bad, bad.p, 8, cudaMemcpyDeviceToHost, st));
    FMS_CUDA(ctx, cudaMemcpyAsync(&total, pos.p + n, 8, cudaMemcpyDeviceToHost, st));
    FMS_CUDA(ctx, cudaStreamSynchronize(st));
    if (h_bad != ~0ull) return report_bad(ctx, h_bad, "fms_parse_ratings_csv");
    if (n_rows) *n_rows = total;
    if (!user && !item && !rating) return FMS_OK;                        // sizing call
    if (total > capacity) return fail(ctx, FMS_ERR_INVALID_ARG, "fms_parse_ratings_csv: %lld rows but capacity %lld", (long long)total, (long long)capacity);
    if (total == 0) return FMS_OK;
    FMS_CUDA(ctx, ou.alloc((size_t)total)); FMS_CUDA(ctx, oi.alloc((size_t)total)); FMS_CUDA(ctx, orr.alloc((size_t)total));
    k_scatter_csv<<<grid, 128, 0, st>>>(n, status.p, pos.p, u.p, it.p, r.p, ou.p, oi.p, orr.p);
    FMS_LAUNCH_CHECK(ctx);
    if (user) FMS_CUDA(ctx, cudaMemcpyAsync(user, ou.p, (size_t)total * 8, cudaMemcpyDefault, st));
    if (item) FMS_CUDA(ctx, cudaMemcpyAsync(item, oi.p, (size_t)total * 8, cudaMemcpyDefault, st));
    if (rating) FMS_CUDA(ctx, cudaMemcpyAsync(rating, orr.p, (size_t)total * 8, cudaMemcpyDefault, st));
    FMS_CUDA(ctx, cudaStreamSynchronize(st));
    return FMS_OK;
}

int ingest_libsvm(fms_ctx *ctx, const char *text, int64_t nbytes, double *label, int64_t row_capacity, int64_t *csr_ptr, int32_t *csr_idx,
                  double *csr_val, int64_t nnz_capacity, int64_t *n_rows, int64_t *nnz, int32_t *dim) {
    cudaStream_t st = ctx->stream;
    if (n_rows) *n_rows = 0;
    if (nnz) *nnz = 0;
    if (dim) *dim = 0;
    if (nbytes == 0) { if (csr_ptr && row_capacity >= 0) { int64_t z = 0; FMS_CUDA(ctx, cudaMemcpy(csr_ptr, &z, 8, cudaMemcpyDefault)); } return FMS_OK; }
    LineIndex L;
    FMS_TRY(index_lines(ctx, text, nbytes, L));
    const int64_t n = L.n_lines;
    DevBuf<int64_t> cnt, flag, row_of, off, ptr;
    DevBuf<int> status, mx;
    DevBuf<unsigned long long> bad;
    DevBuf<unsigned char> temp;
    DevBuf<double> lab, val;
    DevBuf<int32_t> idx;
    FMS_CUDA(ctx, cnt.alloc((size_t)n + 1)); FMS_CUDA(ctx, flag.alloc((size_t)n + 1)); FMS_CUDA(ctx, row_of.alloc((size_t)n + 1));
    FMS_CUDA(ctx, off.alloc((size_t)n + 1)); FMS_CUDA(ctx, status.alloc((size_t)n + 1)); FMS_CUDA(ctx, mx.alloc(1)); FMS_CUDA(ctx, bad.alloc(1));
    FMS_CUDA(ctx, cudaMemsetAsync(bad.p, 0xff, 8, st));
    FMS_CUDA(ctx, cudaMemsetAsync(mx.p, 0, 4, st));
    FMS_CUDA(ctx, cudaMemsetAsync(cnt.p, 0, (size_t)(n + 1) * 8, st));
    FMS_CUDA(ctx, cudaMemsetAsync(flag.p, 0, (size_t)(n + 1) * 8, st));
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (n > 0) {
        k_libsvm_lines<<<grid, 128, 0, st>>>(L.text.p, nbytes, L.nl.p, L.n_nl, n, 0, status.p, cnt.p, nullptr, nullptr, nullptr, nullptr, nullptr, mx.p, bad.p);
        FMS_LAUNCH_CHECK(ctx);
        k_ok_flags<<<grid, 128, 0, st>>>(n, status.p, flag.p);
        FMS_LAUNCH_CHECK(ctx);
    }
    FMS_TRY(exclusive_scan(ctx, flag.p, row_of.p, n + 1, temp));
    FMS_TRY(exclusive_scan(ctx, cnt.p, off.p, n + 1, temp));
    unsigned long long h_bad = 0;
    int64_t rows = 0, total = 0;
    int h_mx = 0;
    FMS_CUDA(ctx, cudaMemcpyAsync(&h_bad, bad.p, 8, cudaMemcpyDeviceToHost, st));
    FMS_CUDA(ctx, cudaMemcpyAsync(&rows, row_of.p + n, 8, cudaMemcpyDeviceToHost, st));
    FMS_CUDA(ctx, cudaMemcpyAsync(&total, off.p + n, 8, cudaMemcpyDeviceToHost, st));
    FMS_CUDA(ctx, cudaMemcpyAsync(&h_mx, mx.p, 4, cudaMemcpyDeviceToHost, st));
    FMS_CUDA(ctx, cudaStreamSynchronize(st));
    if (h_bad != ~0ull) return report_bad(ctx, h_bad, "fms_parse_libsvm");
    if (n_rows) *n_rows = rows;
    if (nnz) *nnz = total;
    if (dim) *dim = h_mx;
    if (!label && !csr_ptr && !csr_idx && !csr_val) return FMS_OK;      // sizing call
    if (rows > row_capacity || total > nnz_capacity)
        return fail(ctx, FMS_ERR_INVALID_ARG, "fms_parse_libsvm: %lld rows / %lld entries exceed the capacities", (long long)rows, (long long)total);
    FMS_CUDA(ctx, ptr.alloc((size_t)rows + 1)); FMS_CUDA(ctx, lab.alloc((size_t)rows + 1));
    FMS_CUDA(ctx, idx.alloc((size_t)total + 1)); FMS_CUDA(ctx, val.alloc((size_t)total + 1));
    if (n > 0) {
        k_row_ptr_from_cnt<<<grid, 128, 0, st>>>(n, status.p, row_of.p, off.p, ptr.p);
        FMS_LAUNCH_CHECK(ctx);
    }
    FMS_CUDA(ctx, cudaMemcpyAsync(ptr.p + rows, &total, 8, cudaMemcpyHostToDevice, st));
    if (n > 0 && rows > 0) {
        k_libsvm_lines<<<grid, 128, 0, st>>>(L.text.p, nbytes, L.nl.p, L.n_nl, n, 1, status.p, cnt.p, row_of.p, ptr.p, lab.p, idx.p, val.p, mx.p, bad.p);
        FMS_LAUNCH_CHECK(ctx);
    }
    if (label && rows) FMS_CUDA(ctx, cudaMemcpyAsync(label, lab.p, (size_t)rows * 8, cudaMemcpyDefault, st));
    if (csr_ptr) FMS_CUDA(ctx, cudaMemcpyAsync(csr_ptr, ptr.p, (size_t)(rows + 1) * 8, cudaMemcpyDefault, st));
    if (csr_idx && total) FMS_CUDA(ctx, cudaMemcpyAsync(csr_idx, idx.p, (size_t)total * 4, cudaMemcpyDefault, st));
    if (csr_val && total) FMS_CUDA(ctx, cudaMemcpyAsync(csr_val, val.p, (size_t)total * 8, cudaMemcpyDefault, st));
    FMS_CUDA(ctx, cudaStreamSynchronize(st));
    return FMS_OK;
}

}  // namespace fms
