// group.cu -- fms_create(device_ids, n > 1): ONE process (a JVM host, a C program) drives n GPUs of one NVLink node through the
// same whole-fit entry points.  SURVEY.md 8(b)/(e): the reference scales by Flink parallelism / setBlocks
// (flink-als/src/main/scala/de/tub/it4bi/ALSImpl.scala:39-41, flink-svm/src/main/scala/de/tub/it4bi/SVMImpl.scala:25); here the
// parent ctx owns one single-GPU member ctx per device and
//   ALS: every member pulls 1/n of the COO over PCIe, pulls the other slices from its peers over NVLink, indexes them and solves
//        its row shard; with the tensor-core kernel the solver warps store every new gather-plane row straight into all replicas
//        (peer pointers, no IPC needed inside one process), otherwise the fp32 shard is pushed with cudaMemcpyPeerAsync; the only
//        thing between half-steps is an event barrier across the members' streams -- no host synchronisation, no NCCL;
//   SVM: member g sweeps the CoCoA blocks k = g (mod n) on its own rows; one fused kernel per member then reads every member's
//        delta-w over NVLink, sums them in rank order (deterministic) and applies w += (stepsize / blocks) * sum:
//        reduce -> all-reduce -> apply in one launch.
// The one-process-per-GPU driver (flink-ms_b200/dist.py, torchrun) stays: bench.py's N > 1 contract launches one rank per GPU.
#include <algorithm>
#include <thread>

#include "common.cuh"

namespace fms {

int als_tc_set_peers_direct(fms_ctx *ctx, int s, void *const *ptrs, int n);

static int up(fms_ctx *P, fms_ctx *m, int rc) {          // surface a member's error on the parent
    if (rc != FMS_OK) P->err = "device " + std::to_string(m->device) + ": " + m->err;
    return rc;
}

#define GRP_CUDA(P, expr)                                                                                \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return ::fms::fail((P), e__ == cudaErrorMemoryAllocation ? FMS_ERR_OOM : FMS_ERR_CUDA,        \
                               "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));   \
    } while (0)

int group_create(fms_ctx *P, const int *device_ids, int n) {
    for (int g = 0; g < n; ++g) {
        for (int h = 0; h < g; ++h)
            if (device_ids[h] == device_ids[g]) return fail(P, FMS_ERR_INVALID_ARG, "fms_create: device %d listed twice", device_ids[g]);
        fms_ctx *m = nullptr;
        const int rc = fms_create(&m, device_ids + g, 1);
        if (rc != FMS_OK) { P->err = fms_last_error(nullptr); return rc; }
        P->members.push_back(m);
    }
    // peer access both ways between every pair (NVLink / NVSwitch node); "already enabled" is fine
    for (int g = 0; g < n; ++g) {
        GRP_CUDA(P, cudaSetDevice(device_ids[g]));
        for (int p = 0; p < n; ++p) {
            if (p == g) continue;
            int can = 0;
            GRP_CUDA(P, cudaDeviceCanAccessPeer(&can, device_ids[g], device_ids[p]));
            if (!can) return fail(P, FMS_ERR_UNSUPPORTED, "fms_create: device %d cannot access device %d (no peer-to-peer path)", device_ids[g], device_ids[p]);
            cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[p], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(P, FMS_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", device_ids[g], device_ids[p], cudaGetErrorString(e));
            cudaGetLastError();
        }
        cudaEvent_t ev;
        GRP_CUDA(P, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        P->member_ev.push_back(ev);
    }
    return FMS_OK;
}

void group_destroy(fms_ctx *P) {
    for (size_t g = 0; g < P->members.size(); ++g) {
        cudaSetDevice(P->members[g]->device);
        if (g < P->member_ev.size()) cudaEventDestroy(P->member_ev[g]);
        fms_destroy(P->members[g]);
    }
    P->members.clear();
    P->member_ev.clear();
}

// every member's stream waits for everything queued so far on every member's stream (device-side; the host does not block)
static int barrier_all(fms_ctx *P) {
    const int G = (int)P->members.size();
    for (int g = 0; g < G; ++g) {
        GRP_CUDA(P, cudaSetDevice(P->members[g]->device));
        GRP_CUDA(P, cudaEventRecord(P->member_ev[g], P->members[g]->stream));
    }
    for (int g = 0; g < G; ++g) {
        GRP_CUDA(P, cudaSetDevice(P->members[g]->device));
        for (int p = 0; p < G; ++p)
            if (p != g) GRP_CUDA(P, cudaStreamWaitEvent(P->members[g]->stream, P->member_ev[p], 0));
    }
    return FMS_OK;
}

static int sync_all(fms_ctx *P) {
    for (fms_ctx *m : P->members) {
        GRP_CUDA(P, cudaSetDevice(m->device));
        GRP_CUDA(P, cudaStreamSynchronize(m->stream));
    }
    return FMS_OK;
}

template <typename F>
static int for_members_parallel(fms_ctx *P, F fn) {        // fn(g) -> status; one host thread per member (loads synchronise their stream)
    const int G = (int)P->members.size();
    std::vector<int> rc((size_t)G, FMS_OK);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) th.emplace_back([&, g] { cudaSetDevice(P->members[g]->device); rc[(size_t)g] = fn(g); });
    for (auto &t : th) t.join();
    for (int g = 0; g < G; ++g)
        if (rc[(size_t)g] != FMS_OK) return up(P, P->members[g], rc[(size_t)g]);
    return FMS_OK;
}

// push this member's rows [S g, S (g + 1)) of an fp32 factor matrix into every replica
static int push_fp32_shard(fms_ctx *P, int g, int which) {
    fms_ctx *m = P->members[g];
    AlsState &A = m->als;
    const int64_t n = which == 0 ? A.n_users : A.n_items, S = A.shard_rows[which];
    const int64_t b = std::min<int64_t>(n, S * g), e = std::min<int64_t>(n, S * (g + 1));
    if (e <= b) return FMS_OK;
    GRP_CUDA(P, cudaSetDevice(m->device));
    for (size_t p = 0; p < P->members.size(); ++p) {
        if ((int)p == g) continue;
        fms_ctx *o = P->members[p];
        GRP_CUDA(P, cudaMemcpyPeerAsync(o->als.fac[which].p + b * A.rank, o->device, A.fac[which].p + b * A.rank, m->device,
                                        (size_t)(e - b) * A.rank * sizeof(float), m->stream));
    }
    return FMS_OK;
}

int group_als_fit(fms_ctx *P, const int64_t *user, const int64_t *item, const double *rating, int64_t nnz, int rank, int iterations,
                  double lambda, int64_t seed, int64_t *n_users, int64_t *n_items) {
    const int G = (int)P->members.size();
    if (nnz < 0 || (nnz > 0 && (!user || !item || !rating))) return fail(P, FMS_ERR_INVALID_ARG, "fms_als_fit: null input or nnz < 0");
    // ---- 1. the triples: 1/G per member over PCIe, the rest from the peers over NVLink
    std::vector<int64_t *> du((size_t)G, nullptr), di((size_t)G, nullptr);
    std::vector<double *> dr((size_t)G, nullptr);
    auto free_tmp = [&] {
        for (int g = 0; g < G; ++g) { cudaSetDevice(P->members[g]->device); cudaFree(du[(size_t)g]); cudaFree(di[(size_t)g]); cudaFree(dr[(size_t)g]); }
    };
    const int64_t sl = (nnz + G - 1) / G;
    for (int g = 0; g < G; ++g) {
        fms_ctx *m = P->members[g];
        GRP_CUDA(P, cudaSetDevice(m->device));
        const size_t bytes = (size_t)std::max<int64_t>(nnz, 1) * 8;
        cudaError_t e = cudaMalloc((void **)&du[(size_t)g], bytes);
        if (e == cudaSuccess) e = cudaMalloc((void **)&di[(size_t)g], bytes);
        if (e == cudaSuccess) e = cudaMalloc((void **)&dr[(size_t)g], bytes);
        if (e != cudaSuccess) { free_tmp(); return fail(P, FMS_ERR_OOM, "fms_als_fit: staging the triples on device %d: %s", m->device, cudaGetErrorString(e)); }
        const int64_t b = std::min<int64_t>(nnz, sl * g), c = std::min<int64_t>(nnz, sl * (g + 1)) - b;
        if (c > 0) {
            // cudaMemcpyDefault: the triples are host memory, or (chunked upload) already staged on the first device
            cudaMemcpyAsync(du[(size_t)g] + b, user + b, (size_t)c * 8, cudaMemcpyDefault, m->stream);
            cudaMemcpyAsync(di[(size_t)g] + b, item + b, (size_t)c * 8, cudaMemcpyDefault, m->stream);
            cudaMemcpyAsync(dr[(size_t)g] + b, rating + b, (size_t)c * 8, cudaMemcpyDefault, m->stream);
        }
    }
    int rc = barrier_all(P);
    for (int g = 0; g < G && rc == FMS_OK; ++g) {
        fms_ctx *m = P->members[g];
        cudaSetDevice(m->device);
        for (int p = 0; p < G; ++p) {
            if (p == g) continue;
            const int64_t b = std::min<int64_t>(nnz, sl * p), c = std::min<int64_t>(nnz, sl * (p + 1)) - b;
            if (c <= 0) continue;
            fms_ctx *o = P->members[p];
            cudaMemcpyPeerAsync(du[(size_t)g] + b, m->device, du[(size_t)p] + b, o->device, (size_t)c * 8, m->stream);
            cudaMemcpyPeerAsync(di[(size_t)g] + b, m->device, di[(size_t)p] + b, o->device, (size_t)c * 8, m->stream);
            cudaMemcpyPeerAsync(dr[(size_t)g] + b, m->device, dr[(size_t)p] + b, o->device, (size_t)c * 8, m->stream);
        }
    }
    if (rc == FMS_OK) rc = sync_all(P);                           // all pulls done before any member frees / reuses its staging buffers
    if (rc == FMS_OK && cudaGetLastError() != cudaSuccess) rc = fail(P, FMS_ERR_CUDA, "fms_als_fit: distributing the triples failed");
    // ---- 2. index (createBlockInformation) + init on every member, in parallel host threads
    if (rc == FMS_OK)
        rc = for_members_parallel(P, [&](int g) {
            fms_ctx *m = P->members[g];
            int r = fms_als_load(m, du[(size_t)g], di[(size_t)g], dr[(size_t)g], nnz, g, G, nullptr, nullptr);
            if (r == FMS_OK) r = fms_als_init(m, rank, seed);
            return r;
        });
    free_tmp();
    if (rc != FMS_OK) return rc;
    // ---- 3. replicas of the gather planes: direct peer pointers (tensor-core path only)
    const bool tc = P->members[0]->als.kernel != FMS_ALS_KERNEL_GENERIC && als_tc_supported(P->members[0], rank);
    if (tc) {
        for (int s = 0; s < 2; ++s) {
            std::vector<void *> tab((size_t)G, nullptr);
            for (int g = 0; g < G; ++g) {
                cudaSetDevice(P->members[g]->device);
                if ((rc = up(P, P->members[g], als_tc_planes(P->members[g], s, &tab[(size_t)g], nullptr, nullptr))) != FMS_OK) return rc;
            }
            for (int g = 0; g < G; ++g) {
                cudaSetDevice(P->members[g]->device);
                if ((rc = up(P, P->members[g], als_tc_set_peers_direct(P->members[g], s, tab.data(), G))) != FMS_OK) return rc;
            }
        }
        if ((rc = sync_all(P)) != FMS_OK) return rc;              // the zero fill of fresh plane buffers is complete everywhere
    }
    // ---- 4. iterate(T){ users; items }; users -- an event barrier across the members' streams after every half-step
    auto half = [&](int which) -> int {
        for (int g = 0; g < G; ++g) {
            fms_ctx *m = P->members[g];
            cudaSetDevice(m->device);
            int r = up(P, m, fms_als_half_step(m, which, lambda));
            if (r != FMS_OK) return r;
            if (!tc && (r = push_fp32_shard(P, g, which)) != FMS_OK) return r;
        }
        int r = barrier_all(P);
        if (r != FMS_OK) return r;
        for (int g = 0; g < G; ++g) {
            if (tc) fms_als_planes_updated(P->members[g], which);      // every replica of the planes is current behind the barrier
            else fms_als_factors_updated(P->members[g], which);
        }
        return FMS_OK;
    };
    for (int t = 0; t < iterations; ++t) {
        if ((rc = half(FMS_USERS)) != FMS_OK) return rc;
        if ((rc = half(FMS_ITEMS)) != FMS_OK) return rc;
    }
    if ((rc = half(FMS_USERS)) != FMS_OK) return rc;
    // ---- 5. the fp32 matrices (read-out, predict) were only written shard-wise on the tensor-core path: complete the replicas
    if (tc) {
        for (int g = 0; g < G; ++g)
            for (int s = 0; s < 2; ++s)
                if ((rc = push_fp32_shard(P, g, s)) != FMS_OK) return rc;
        if ((rc = barrier_all(P)) != FMS_OK) return rc;
    }
    if ((rc = sync_all(P)) != FMS_OK) return rc;
    for (fms_ctx *m : P->members) m->als.fac64_valid = false;
    int64_t bad = 0;
    for (fms_ctx *m : P->members) { int64_t b = 0; cudaSetDevice(m->device); fms_als_last_info(m, &b); bad += b; }
    if (bad) fail(P, FMS_OK, "fms_als_fit: %lld row(s) of the last half-step broke down (non-positive pivot, dposv info != 0)", (long long)bad);
    else P->err.clear();
    if (n_users) *n_users = P->members[0]->als.n_users;
    if (n_items) *n_items = P->members[0]->als.n_items;
    return FMS_OK;
}

// w += scaling * sum_g delta_g, the deltas read from every member over NVLink in rank order
__global__ void k_svm_allreduce_apply(float *__restrict__ w, const float *const *__restrict__ deltas, int G, int dim, float scaling) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= dim) return;
    float s = 0.f;
    for (int g = 0; g < G; ++g) s += deltas[g][j];
    w[j] = fmaf(s, scaling, w[j]);
}

int group_svm_fit(fms_ctx *P, const void *X, int x_dtype, const int64_t *csr_ptr, const int32_t *csr_idx, const double *y, int64_t n,
                  int32_t dim, const int32_t *block_id, int blocks, int iterations, int64_t local_iterations, double regularization,
                  double stepsize, int64_t seed, int rng_mode, double *w_out) {
    const int G = (int)P->members.size();
    if (rng_mode != FMS_RNG_PER_BLOCK) return fail(P, FMS_ERR_UNSUPPORTED, "fms_svm_fit on %d devices: the shared-subtask RNG schedule is one sequential stream and cannot be sharded", G);
    if (n < 0 || dim < 1 || blocks < 1 || (n > 0 && (!X || !y || !block_id))) return fail(P, FMS_ERR_INVALID_ARG, "fms_svm_fit: bad arguments");
    if (x_dtype < FMS_X_DENSE_F32 || x_dtype > FMS_X_CSR_F64) return fail(P, FMS_ERR_INVALID_ARG, "fms_svm_fit: x_dtype %d", x_dtype);
    if (x_dtype == FMS_X_CSR_F64 && n > 0 && (!csr_ptr || !csr_idx)) return fail(P, FMS_ERR_INVALID_ARG, "fms_svm_fit: CSR arrays missing");
    for (int64_t i = 0; i < n; ++i)
        if (block_id[i] < 0 || block_id[i] >= blocks) return fail(P, FMS_ERR_INVALID_ARG, "fms_svm_fit: block_id[%lld]=%d out of range", (long long)i, block_id[i]);
    // ---- member g gets the rows of the blocks k = g (mod G), compacted in arrival order, with their GLOBAL block ids
    int rc = for_members_parallel(P, [&](int g) {
        fms_ctx *m = P->members[g];
        std::vector<int64_t> rows;
        for (int64_t i = 0; i < n; ++i)
            if (block_id[i] % G == g) rows.push_back(i);
        const int64_t ng = (int64_t)rows.size();
        std::vector<double> yg((size_t)ng);
        std::vector<int32_t> bg((size_t)ng);
        for (int64_t r = 0; r < ng; ++r) { yg[(size_t)r] = y[rows[(size_t)r]]; bg[(size_t)r] = block_id[rows[(size_t)r]]; }
        int r = FMS_OK;
        if (x_dtype == FMS_X_CSR_F64) {
            std::vector<int64_t> pg((size_t)ng + 1, 0);
            for (int64_t q = 0; q < ng; ++q) pg[(size_t)q + 1] = pg[(size_t)q] + (csr_ptr[rows[(size_t)q] + 1] - csr_ptr[rows[(size_t)q]]);
            std::vector<int32_t> ig((size_t)pg[(size_t)ng]);
            std::vector<double> vg((size_t)pg[(size_t)ng]);
            const double *val = static_cast<const double *>(X);
            for (int64_t q = 0; q < ng; ++q) {
                const int64_t s0 = csr_ptr[rows[(size_t)q]], c = csr_ptr[rows[(size_t)q] + 1] - s0;
                std::copy(csr_idx + s0, csr_idx + s0 + c, ig.begin() + pg[(size_t)q]);
                std::copy(val + s0, val + s0 + c, vg.begin() + pg[(size_t)q]);
            }
            r = fms_svm_load(m, vg.data(), x_dtype, 0, pg.data(), ig.data(), yg.data(), ng, dim, bg.data(), blocks, g, G, regularization, stepsize, seed, rng_mode);
        } else {
            const size_t esz = x_dtype == FMS_X_DENSE_F32 ? 4 : 8, rb = (size_t)dim * esz;
            std::vector<unsigned char> xg((size_t)ng * rb);
            const unsigned char *xs = static_cast<const unsigned char *>(X);
            for (int64_t q = 0; q < ng; ++q) memcpy(xg.data() + (size_t)q * rb, xs + (size_t)rows[(size_t)q] * rb, rb);
            r = fms_svm_load(m, xg.data(), x_dtype, 0, nullptr, nullptr, yg.data(), ng, dim, bg.data(), blocks, g, G, regularization, stepsize, seed, rng_mode);
        }
        m->svm.n_total = n;                                        // lambda * n of `maximize` is the size of the WHOLE training set
        return r;
    });
    if (rc != FMS_OK) return rc;
    // ---- delta pointer table on every member
    std::vector<const float *> tab((size_t)G);
    for (int g = 0; g < G; ++g) tab[(size_t)g] = P->members[g]->svm.delta.p;
    for (int g = 0; g < G; ++g) {
        fms_ctx *m = P->members[g];
        GRP_CUDA(P, cudaSetDevice(m->device));
        GRP_CUDA(P, m->svm.peer_delta.alloc((size_t)G));
        GRP_CUDA(P, cudaMemcpyAsync(m->svm.peer_delta.p, tab.data(), sizeof(float *) * (size_t)G, cudaMemcpyHostToDevice, m->stream));
    }
    if ((rc = sync_all(P)) != FMS_OK) return rc;
    const float scaling = (float)(stepsize / (double)blocks);
    for (int t = 0; t < iterations; ++t) {
        for (int g = 0; g < G; ++g) {
            fms_ctx *m = P->members[g];
            cudaSetDevice(m->device);
            if ((rc = up(P, m, svm_launch_sweep(m, local_iterations))) != FMS_OK) return rc;     // sweep + local reduce into delta
        }
        if ((rc = barrier_all(P)) != FMS_OK) return rc;
        for (int g = 0; g < G; ++g) {
            fms_ctx *m = P->members[g];
            cudaSetDevice(m->device);
            k_svm_allreduce_apply<<<(dim + 255) / 256, 256, 0, m->stream>>>(m->svm.w.p, reinterpret_cast<const float *const *>(m->svm.peer_delta.p), G, dim, scaling);
            ++m->launches;
            GRP_CUDA(P, cudaGetLastError());
            m->svm.w64_valid = false;
        }
        if ((rc = barrier_all(P)) != FMS_OK) return rc;           // nobody overwrites its delta before every member has read it
    }
    if ((rc = sync_all(P)) != FMS_OK) return rc;
    cudaSetDevice(P->members[0]->device);
    return up(P, P->members[0], fms_svm_get_weights(P->members[0], w_out, dim));
}

}  // namespace fms
