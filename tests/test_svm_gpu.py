"""GPU parity tests for the CoCoA/SDCA SVM path against the fp64 oracle: sample-index streams bit-exact, weights
within 1e-4 relative (BASELINE.json north_star), labels exact on margin-safe queries."""
import importlib
import json
import os

import numpy as np
import pytest

import oracle
from conftest import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
W_TOL = 1e-4


@pytest.fixture(scope="module", autouse=True)
def _synth(fms):
    import importlib
    fms.synth = importlib.import_module("flink-ms_b200.synth")


def test_device_index_stream_is_bit_exact(fms, ctx):
    for seed, block, bound in [(42, 0, 1000), (42, 7, 33334), (-5, 3, (1 << 30) + 1), (123456789012345, 1, 3), (42, 3, 4096)]:
        got = ctx.debug_index_stream(seed, block, bound, 500)
        assert got.tolist() == oracle.svm_index_stream(seed, block, bound, 500).tolist()
    kat = json.load(open(os.path.join(GOLD, "jrand_kat.json")))
    for key, vals in kat["published"]["bounded"].items():
        seed, bound = map(int, key.split(":"))
        assert ctx.debug_index_stream(seed, 0, bound, len(vals)).tolist() == vals


def _kernels(fms, dim):
    ks = [fms._native.SVM_KERNEL_GENERIC]
    if dim % 128 == 0 and dim <= 2048:
        ks.append(fms._native.SVM_KERNEL_AUTO)
    return ks


@pytest.mark.parametrize("n,dim,K,H", [(600, 24, 4, 10), (600, 24, 4, -1), (3000, 128, 16, -1), (4000, 1024, 24, -1),
                                        (2000, 256, 7, 25), (500, 1024, 1, -1)])
def test_fit_parity_dense(fms, ctx, n, dim, K, H):
    X, y, _ = fms.synth.svm_data(n, dim, seed=n + dim)
    bid = (np.arange(n) % K).astype(np.int32)
    want = oracle.svm_fit(X, y, bid, K, 5, H, 1.0, 1.0, 42, 0)
    for kernel in _kernels(fms, dim):
        ctx.svm_set_kernel(kernel)
        w = ctx.svm_fit(X, y, bid, K, 5, H, 1.0, 1.0, 42)
        a = ctx.svm_get_alpha()
        assert rel_err(w, want) <= W_TOL, (kernel, rel_err(w, want))
        assert a.min() >= 0.0 and a.max() <= 1.0
    ctx.svm_set_kernel(fms._native.SVM_KERNEL_AUTO)


def test_fit_parity_f64_input_stepsize_and_regularization(fms, ctx):
    X, y, _ = fms.synth.svm_data(800, 40, seed=3)
    bid = (np.arange(800) * 5 // 800).astype(np.int32)                # contiguous splits, like a file source
    want, wa = oracle.svm_fit(X.astype(np.float64), y, bid, 5, 6, 40, 0.3, 0.6, 7, 0, return_alpha=True)
    w = ctx.svm_fit(X.astype(np.float64), y, bid, 5, 6, 40, 0.3, 0.6, 7)
    assert rel_err(w, want) <= W_TOL
    assert np.abs(ctx.svm_get_alpha() - wa).max() <= 1e-4


def test_fit_parity_csr_and_zero_rows(fms, ctx):
    csr, y = fms.synth.svm_sparse(500, 64, 0.2, seed=9)
    ptr, idx, val = csr
    bid = (np.arange(500) % 6).astype(np.int32)
    want = oracle.svm_fit(None, y, bid, 6, 4, -1, 1.0, 1.0, 42, 0, csr=csr, dim=64)
    w = ctx.svm_fit(None, y, bid, 6, 4, -1, 1.0, 1.0, 42, csr=csr, dim=64)
    assert rel_err(w, want) <= W_TOL
    X, y2, _ = fms.synth.svm_data(300, 16, seed=1)
    X[::5] = 0.0                                                        # q == 0 rows: alpha' = 1, no weight change
    bid = (np.arange(300) % 3).astype(np.int32)
    want, wa = oracle.svm_fit(X, y2, bid, 3, 3, -1, 1.0, 1.0, 42, 0, return_alpha=True)
    w = ctx.svm_fit(X, y2, bid, 3, 3, -1, 1.0, 1.0, 42)
    assert rel_err(w, want) <= W_TOL
    assert np.abs(ctx.svm_get_alpha() - wa).max() <= 1e-4


def test_shared_subtask_rng_and_empty_blocks(fms, ctx):
    X, y, _ = fms.synth.svm_data(400, 20, seed=4)
    bid = (np.arange(400) % 3).astype(np.int32) * 2                     # blocks 1, 3, 5 stay empty
    want = oracle.svm_fit(X, y, bid, 6, 4, 15, 1.0, 1.0, 42, 1)
    w = ctx.svm_fit(X, y, bid, 6, 4, 15, 1.0, 1.0, 42, rng_mode=fms._native.RNG_SHARED_SUBTASK)
    assert rel_err(w, want) <= W_TOL
    want0 = oracle.svm_fit(X, y, bid, 6, 4, 15, 1.0, 1.0, 42, 0)
    w0 = ctx.svm_fit(X, y, bid, 6, 4, 15, 1.0, 1.0, 42)
    assert rel_err(w0, want0) <= W_TOL


def test_reference_default_shape_one_partition(fms, ctx):
    """BASELINE configs[0]-style run (SVMImpl.scala:24-26 with FlinkML defaults H=10, lambda=1, stepsize=1) at parallelism 1:
    FlinkMLTools.block hashes the rows over the 10 blocks (flinkml_hash.py) and the single localDualMethod subtask draws all
    blocks' samples from ONE java.util.Random(seed ^ 0), ascending block order.  The round-1 reading (all rows in block 0, the
    block count only scaling the step) stays available as block_mode="subtask"."""
    import importlib
    H = importlib.import_module("flink-ms_b200.flinkml_hash")
    X, y, _ = fms.synth.svm_data(1000, 32, seed=8)
    svm = fms.SVM(ctx).setBlocks(10).setIterations(10).setSeed(42)
    svm.fit(X, y)
    bid = H.block_ids_by_hash(y, 10, X=X)
    assert len(set(bid.tolist())) == 10
    want = oracle.svm_fit(X, y, bid, 10, 10, 10, 1.0, 1.0, 42, 1)
    w = svm.weightsOption[0].data
    assert rel_err(w, want) <= W_TOL
    svm1 = fms.SVM(ctx).setBlocks(10).setIterations(10).setSeed(42)
    svm1.fit(X, y, block_mode="subtask")
    want1 = oracle.svm_fit(X, y, np.zeros(1000, np.int32), 10, 10, 10, 1.0, 1.0, 42, 0)
    assert rel_err(svm1.weightsOption[0].data, want1) <= W_TOL
    # SVMImpl.scala:33-35: weights -> (1-based feature id, value)
    model = [(j + 1, v) for j, v in enumerate(w)]
    assert model[0][0] == 1 and len(model) == 32


def test_golden_small_case(fms, ctx):
    g = json.load(open(os.path.join(GOLD, "oracle_small.json")))["svm"]
    X = np.array(g["X"], np.float32)
    w10 = ctx.svm_fit(X, g["y"], g["block_id"], g["blocks"], g["iterations"], 10, 1.0, 1.0, g["seed"])
    assert rel_err(w10, g["w_H10"]) <= W_TOL
    wep = ctx.svm_fit(X, g["y"], g["block_id"], g["blocks"], g["iterations"], -1, 1.0, 1.0, g["seed"])
    assert rel_err(wep, g["w_epoch"]) <= W_TOL


def test_predict_labels_and_decision_function(fms, ctx):
    rng = np.random.default_rng(2)
    w = rng.standard_normal(50)
    X = rng.standard_normal((2000, 50))
    X[0] = 0.0                                                          # raw == threshold == 0 -> -1 (strict >)
    got = ctx.svm_predict(X, w, 0.0, False)
    want = oracle.svm_predict(X, w, 0.0, False)
    raw = oracle.svm_predict(X, w, 0.0, True)
    safe = np.abs(raw) > 1e-9
    safe[0] = True
    assert np.array_equal(got[safe], want[safe]) and got[0] == -1.0
    assert np.allclose(ctx.svm_predict(X, w, 0.0, True), raw, rtol=1e-13, atol=1e-13)
    assert np.array_equal(ctx.svm_predict(X.astype(np.float32), w, 0.25, False)[safe],
                          oracle.svm_predict(X.astype(np.float32), w, 0.25, False)[safe])
    # sparse queries with 1-based -> 0-based ids already applied; ids beyond the model contribute nothing
    csr, _ = fms.synth.svm_sparse(300, 50, 0.3, seed=3)
    assert np.allclose(ctx.svm_predict(None, w, 0.0, True, csr=csr, dim=50), oracle.svm_predict(None, w, 0.0, True, csr=csr, dim=50),
                       rtol=1e-13, atol=1e-13)


def test_argument_errors(fms, ctx):
    X, y, _ = fms.synth.svm_data(50, 8)
    with pytest.raises(fms.FlinkMSError) as e:
        ctx.svm_fit(X, y, np.full(50, 3, np.int32), 2, 1, 1)            # block id out of range
    assert e.value.status == fms._native.ERR_INVALID_ARG
    with pytest.raises(fms.FlinkMSError) as e:
        ctx.svm_fit(X, y, np.zeros(50, np.int32), 0, 1, 1)
    assert e.value.status == fms._native.ERR_INVALID_ARG
    fresh = fms.Context(0)
    with pytest.raises(fms.FlinkMSError) as e:
        fresh.svm_predict(X, None)
    assert e.value.status == fms._native.ERR_STATE and "has not been trained" in str(e.value)
    fresh.close()


def test_sharded_rounds_match_unsharded(fms, ctx):
    """Two ctxs on one GPU play ranks 0/1 (blocks k % 2); deltas are summed by hand like the all-reduce."""
    import importlib

    import torch
    dv = importlib.import_module("flink-ms_b200.dist").device_view
    n, dim, K = 1200, 64, 6
    X, y, _ = fms.synth.svm_data(n, dim, seed=6)
    bid = (np.arange(n) % K).astype(np.int32)
    want = oracle.svm_fit(X, y, bid, K, 4, -1, 1.0, 1.0, 42, 0)
    ranks = [fms.Context(0), fms.Context(0)]
    for g, c in enumerate(ranks):
        c.svm_load(X, y, bid, K, g, 2, 1.0, 1.0, 42)
    for _ in range(4):
        deltas = []
        for c in ranks:
            c.svm_sweep(-1)
            c.synchronize()
            ptr, d = c.svm_delta_device()
            deltas.append(dv(ptr, d))
        total = deltas[0] + deltas[1]
        for t in deltas:
            t.copy_(total)
        torch.cuda.synchronize()
        for c in ranks:
            c.svm_apply_delta()
    for c in ranks:
        assert rel_err(c.svm_get_weights(dim), want) <= W_TOL
        c.close()


@pytest.mark.gpu
def test_chunked_upload_equals_one_shot_fit(fms, ctx):
    """fms_*_stream_*: what a JVM host whose direct ByteBuffers stay below 2 GiB calls; same results as the
    one-shot entry points, bit for bit."""
    synth = importlib.import_module("flink-ms_b200.synth")
    X, y, _ = synth.svm_data(3000, 128, seed=8)
    bid = (np.arange(3000) % 12).astype(np.int32)
    w1 = ctx.svm_fit(X, y, bid, 12, 3, -1, 1.0, 1.0, 5)
    w2 = ctx.svm_fit_stream(X, y, bid, 12, 3, -1, 1.0, 1.0, 5, chunk_rows=700)
    np.testing.assert_array_equal(w1, w2)
    w3 = ctx.svm_fit_stream(X.astype(np.float64), y, bid, 12, 3, -1, 1.0, 1.0, 5, chunk_rows=999)
    np.testing.assert_array_equal(w1, w3)
    u, i, r = synth.als_ratings(900, 300, 20000, seed=6)
    for rank in (64, 10):
        ctx.als_fit(u, i, r, rank, 2, 0.9, 42)
        a = [ctx.als_get_factors(w)[1].copy() for w in (0, 1)]
        ctx.als_fit_stream(u, i, r, rank, 2, 0.9, 42, chunk=3333)
        b = [ctx.als_get_factors(w)[1] for w in (0, 1)]
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])
    with pytest.raises(fms.FlinkMSError):
        ctx.als_fit_stream(u[:10], i[:10], r[:10], 0, 1, 0.9, 42)            # rank < 1
    ctx._check(ctx._lib.fms_als_stream_begin(ctx._h, 5))
    with pytest.raises(fms.FlinkMSError):                                      # fit before everything announced has arrived
        ctx._check(ctx._lib.fms_als_stream_fit(ctx._h, 8, 1, 0.9, 42, None, None))
