"""GPU parity tests for the ALS path: the CUDA kernels (through the C-ABI) against the fp64 oracle on the same
seeded inputs.  Tolerances: one half-step from identical inputs <= 1e-5 relative (max-norm); a whole fit
<= 1e-4 relative (BASELINE.json north_star); ids / found flags / init bit-exact."""
import json
import os

import numpy as np
import pytest

import oracle
from conftest import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
HALF_TOL = 1e-5
FIT_TOL = 1e-4


def _case(fms, U, I, nnz, seed=1234, stride=(1, 1), offset=(0, 0)):
    return fms.synth.als_ratings(U, I, nnz, seed, stride, offset)


@pytest.fixture(scope="module", autouse=True)
def _synth(fms):
    import importlib
    fms.synth = importlib.import_module("flink-ms_b200.synth")


def _kernels(fms, rank):
    ks = [fms._native.ALS_KERNEL_GENERIC]
    if rank in (64, 128):
        ks.append(fms._native.ALS_KERNEL_AUTO)
    return ks


@pytest.mark.parametrize("rank", [1, 8, 10, 16, 32, 64, 100, 128])
def test_half_step_parity_from_identical_inputs(fms, ctx, rank):
    u, i, r = _case(fms, 300, 200, 6000, seed=rank)
    orc = oracle.OracleALS(u, i, r)
    for kernel in _kernels(fms, rank):
        ctx.als_set_kernel(kernel)
        nu, ni = ctx.als_load(u, i, r)
        assert (nu, ni) == (orc.n_users, orc.n_items)
        ctx.als_init(rank, 42)
        ids, Y0 = ctx.als_get_factors(fms._native.ITEMS)
        orc.init(rank, 42)
        oid, OY = orc.factors(1)
        assert (ids == oid).all()
        assert np.array_equal(Y0, OY.astype(np.float32).astype(np.float64))      # init: bit-exact RNG, fp32 storage
        orc.set_factors(1, Y0)                                                    # identical inputs
        ctx.als_half_step(fms._native.USERS, 0.9)
        orc.half_step(0, 0.9)
        _, X = ctx.als_get_factors(fms._native.USERS)
        _, OX = orc.factors(0)
        assert rel_err(X, OX) <= HALF_TOL, (kernel, rel_err(X, OX))
        assert ctx.als_last_info() == 0
        orc.set_factors(0, X)
        ctx.als_half_step(fms._native.ITEMS, 0.9)
        orc.half_step(1, 0.9)
        _, Y = ctx.als_get_factors(fms._native.ITEMS)
        _, OY1 = orc.factors(1)
        assert rel_err(Y, OY1) <= HALF_TOL, (kernel, rel_err(Y, OY1))
    ctx.als_set_kernel(fms._native.ALS_KERNEL_AUTO)


def test_fit_parity_config1_rank8(fms, ctx):
    """BASELINE.json configs[0]: rank 8, 10k x 10k, 100k ratings, T=10, lambda 0.9, seed 42."""
    u, i, r = _case(fms, 10_000, 10_000, 100_000)
    nu, ni = ctx.als_fit(u, i, r, 8, 10, 0.9, 42)
    orc = oracle.OracleALS(u, i, r)
    assert orc.fit(8, 10, 0.9, 42) == 0
    for which in (0, 1):
        ids, F = ctx.als_get_factors(which)
        oid, OF = orc.factors(which)
        assert (ids == oid).all()
        assert rel_err(F, OF) <= FIT_TOL, rel_err(F, OF)


@pytest.mark.parametrize("rank,kernel", [(64, "auto"), (64, "generic"), (128, "auto")])
def test_fit_parity_rank64_128_long_rows(fms, ctx, rank, kernel):
    # items get ~400 ratings each, users ~40: both row-length regimes of the bench configuration, scaled down
    u, i, r = _case(fms, 4000, 400, 160_000, seed=5)
    ctx.als_set_kernel(fms._native.ALS_KERNEL_GENERIC if kernel == "generic" else fms._native.ALS_KERNEL_AUTO)
    ctx.als_fit(u, i, r, rank, 3, 0.9, 42)
    ctx.als_set_kernel(fms._native.ALS_KERNEL_AUTO)
    orc = oracle.OracleALS(u, i, r)
    orc.fit(rank, 3, 0.9, 42)
    for which in (0, 1):
        _, F = ctx.als_get_factors(which)
        _, OF = orc.factors(which)
        assert rel_err(F, OF) <= FIT_TOL, (which, rel_err(F, OF))


def test_golden_small_case_with_duplicates_and_sparse_ids(fms, ctx):
    g = json.load(open(os.path.join(GOLD, "oracle_small.json")))["als"]
    ctx.als_fit(g["user"], g["item"], g["rating"], g["rank"], g["iterations"], g["lambda"], g["seed"])
    uid, X = ctx.als_get_factors(0)
    iid, Y = ctx.als_get_factors(1)
    assert uid.tolist() == g["user_ids"] and iid.tolist() == g["item_ids"]
    assert rel_err(X, np.array(g["U"])) <= FIT_TOL and rel_err(Y, np.array(g["I"])) <= FIT_TOL


def test_negative_and_huge_ids_and_zero_iterations(fms, ctx):
    u, i, r = _case(fms, 50, 40, 600, seed=3, stride=(10 ** 12, 7), offset=(-5 * 10 ** 13, -100))
    ctx.als_fit(u, i, r, 5, 0, 0.9, 42)              # T=0: items = init, users = one solve
    orc = oracle.OracleALS(u, i, r)
    orc.fit(5, 0, 0.9, 42)
    iid, Y = ctx.als_get_factors(1)
    oid, OY = orc.factors(1)
    assert (iid == oid).all() and np.array_equal(Y, OY.astype(np.float32).astype(np.float64))
    _, X = ctx.als_get_factors(0)
    assert rel_err(X, orc.factors(0)[1]) <= HALF_TOL


def test_single_rating_rows_and_single_row(fms, ctx):
    u = np.array([5, 6, 7, 8], np.int64)
    i = np.array([1, 1, 2, 3], np.int64)
    r = np.array([4.0, 3.0, 5.0, 1.5])
    ctx.als_fit(u, i, r, 3, 2, 0.9, 42)
    orc = oracle.OracleALS(u, i, r)
    orc.fit(3, 2, 0.9, 42)
    for which in (0, 1):
        assert rel_err(ctx.als_get_factors(which)[1], orc.factors(which)[1]) <= FIT_TOL


def test_empty_input(fms, ctx):
    e = np.zeros(0, np.int64)
    assert ctx.als_fit(e, e, np.zeros(0), 4, 2, 0.9, 42) == (0, 0)
    ids, F = ctx.als_get_factors(0)
    assert ids.size == 0 and F.shape == (0, 4)
    out, found = ctx.als_predict([1], [2])
    assert found.tolist() == [False] and out.tolist() == [0.0]


def test_argument_errors(fms, ctx):
    u, i, r = _case(fms, 20, 20, 100)
    with pytest.raises(fms.FlinkMSError) as e:
        ctx.als_fit(u, i, r, 0, 1, 0.9, 42)
    assert e.value.status == fms._native.ERR_INVALID_ARG
    with pytest.raises(fms.FlinkMSError) as e:
        ctx.als_fit(u, i, r, 129, 1, 0.9, 42)
    assert e.value.status == fms._native.ERR_UNSUPPORTED
    with pytest.raises(fms.FlinkMSError) as e:
        ctx.als_fit(u, i, r, 4, -1, 0.9, 42)
    assert e.value.status == fms._native.ERR_INVALID_ARG
    fresh = fms.Context(0)
    with pytest.raises(fms.FlinkMSError) as e:
        fresh.als_predict([1], [1])
    assert e.value.status == fms._native.ERR_STATE and "has not been fitted" in str(e.value)
    fresh.close()


def test_predict_and_mse_match_oracle(fms, ctx):
    u, i, r = _case(fms, 500, 300, 8000, seed=11, stride=(3, 2), offset=(1, 5))
    ctx.als_fit(u, i, r, 16, 2, 0.9, 42)
    uid, X = ctx.als_get_factors(0)
    iid, Y = ctx.als_get_factors(1)
    orc = oracle.OracleALS(u, i, r)
    orc.set_factors(0, X)
    orc.set_factors(1, Y)
    rng = np.random.default_rng(0)
    qu = np.concatenate([rng.choice(u, 3000), [2, 10 ** 15, u[0]]])          # 2 and 1e15 are not user ids
    qi = np.concatenate([rng.choice(i, 3000), [i[0], i[0], 4]])              # 4 is not an item id
    out, found = ctx.als_predict(qu, qi)
    oout, ofound = orc.predict(qu, qi)
    assert np.array_equal(found, ofound)                                      # indexing: bit-exact
    assert found[-3:].tolist() == [False, False, False]
    assert np.allclose(out, oout, rtol=1e-13, atol=1e-15)                     # fp64 dot, order differs only
    mse, used = ctx.als_mse(qu, qi, np.full(len(qu), 3.0))
    assert used == int(found.sum())
    assert mse == pytest.approx(np.mean((3.0 - oout[ofound]) ** 2), rel=1e-12)


def test_external_model_lookup(fms, ctx):
    """Serve-side path (ALSPredictRandom.java:63-92): install `id,U|I,f1;..` rows, batched lookup-and-dot."""
    rng = np.random.default_rng(5)
    uid = np.sort(rng.choice(10 ** 6, 1000, replace=False)).astype(np.int64)
    iid = np.sort(rng.choice(10 ** 6, 800, replace=False)).astype(np.int64)
    U = rng.standard_normal((1000, 10))
    I = rng.standard_normal((800, 10))
    c = fms.Context(0)
    c.als_set_model(uid, U, iid, I)
    qu = rng.choice(uid, 5000)
    qi = rng.choice(iid, 5000)
    out, found = c.als_predict(qu, qi)
    want = np.einsum("kd,kd->k", U[np.searchsorted(uid, qu)], I[np.searchsorted(iid, qi)])
    assert found.all() and np.allclose(out, want, rtol=1e-13, atol=1e-14)
    with pytest.raises(fms.FlinkMSError):
        c.als_set_model(uid[::-1].copy(), U, iid, I)                          # ids must ascend
    c.close()


def test_operator_api_matches_reference_call_shape(fms, ctx):
    """ALSImpl.scala:35-57 written against the mirror."""
    u, i, r = _case(fms, 200, 100, 3000, seed=2)
    als = fms.ALS(ctx).setIterations(3).setNumFactors(6).setBlocks(4)
    parameters = fms.ParameterMap().add(fms.ALS.Lambda, 0.9).add(fms.ALS.Seed, 42)
    als.fit((u, i, r), parameters)
    users, items = als.factorsOption
    orc = oracle.OracleALS(u, i, r)
    orc.fit(6, 3, 0.9, 42)
    assert [f.id for f in users] == orc.factors(0)[0].tolist()
    assert rel_err(items.factors, orc.factors(1)[1]) <= FIT_TOL
    pu, pi, pred = als.predict(([u[0], 10 ** 9], [i[0], i[0]]))
    assert pu.tolist() == [u[0]] and len(pred) == 1                            # unknown id dropped (inner join)


def test_sharded_half_steps_equal_unsharded(fms, ctx):
    """Two ctxs on one GPU play ranks 0/1: each solves its row shard, shards are exchanged by hand."""
    import importlib

    import torch
    dv = importlib.import_module("flink-ms_b200.dist").device_view
    u, i, r = _case(fms, 301, 203, 5000, seed=8)
    ctx.als_load(u, i, r)
    ctx.als_init(8, 42)
    ctx.als_half_step(0, 0.9)
    ctx.als_half_step(1, 0.9)
    _, Yref = ctx.als_get_factors(1)
    _, Xref = ctx.als_get_factors(0)
    ranks = [fms.Context(0), fms.Context(0)]
    for g, c in enumerate(ranks):
        c.als_load(u, i, r, g, 2)
        c.als_init(8, 42)
    for which in (0, 1):
        views = []
        for c in ranks:
            c.als_half_step(which, 0.9)
            c.synchronize()
            ptr, rows, b, e = c.als_factors_device(which)
            views.append((dv(ptr, rows * 8).view(rows, 8), b, e))
        for src in range(2):                       # the all-gather, by hand
            t_src, b, e = views[src]
            views[1 - src][0][b:e].copy_(t_src[b:e])
        torch.cuda.synchronize()
        for c in ranks:
            c.als_factors_updated(which)
    for c in ranks:
        _, X = c.als_get_factors(0)
        _, Y = c.als_get_factors(1)
        assert np.array_equal(X, Xref) and np.array_equal(Y, Yref)             # same per-row math: bitwise equal
        c.close()


def test_tensor_core_gramian_is_fp32_grade(fms, ctx):
    """The tcgen05 kernel's system matrix (fp16 hi/lo split Gramian + lambda*n on the diagonal) and rhs, as the
    solver warps receive them, against fp64 numpy: the split must deliver fp32-level accuracy (plain fp16 is 5e-4)."""
    u, i, r = _case(fms, 400, 60, 12000, seed=3)
    ctx.als_set_kernel(fms._native.ALS_KERNEL_AUTO)
    ctx.als_load(u, i, r)
    ctx.als_init(64, 42)
    _, Y = ctx.als_get_factors(fms._native.ITEMS)
    uid, iid = np.unique(u), np.unique(i)
    G, b = ctx.debug_als_gramian(fms._native.USERS, 0.9, 0, 64)
    for row in range(64):
        sel = u == uid[row]
        V = Y[np.searchsorted(iid, i[sel])]
        Gr = V.T @ V + 0.9 * sel.sum() * np.eye(64)
        br = V.T @ r[sel]
        assert np.abs(G[row] - Gr).max() <= 2e-6 * np.abs(Gr).max()
        assert np.abs(b[row] - br).max() <= 2e-6 * np.abs(br).max()


def test_tensor_core_and_generic_kernels_agree_over_iterations(fms, ctx):
    u, i, r = _case(fms, 3000, 500, 90_000, seed=12)
    out = {}
    for name, k in (("tc", fms._native.ALS_KERNEL_TC), ("generic", fms._native.ALS_KERNEL_GENERIC)):
        ctx.als_set_kernel(k)
        ctx.als_fit(u, i, r, 64, 4, 0.9, 42)
        out[name] = (ctx.als_get_factors(0)[1], ctx.als_get_factors(1)[1])
    ctx.als_set_kernel(fms._native.ALS_KERNEL_AUTO)
    assert rel_err(out["tc"][0], out["generic"][0]) <= 5e-5 and rel_err(out["tc"][1], out["generic"][1]) <= 5e-5


def test_tensor_core_producer_flavours_are_bitwise_equal(fms, ctx):
    """The two producer flavours of the tensor-core kernel (TMA gather4 / 16-byte LDGSTS copies) feed the same bytes to the
    same MMAs: factors must be bitwise equal, whatever the automatic per-half-step choice picks."""
    u, i, r = _case(fms, 2000, 60, 70_000, seed=13)          # users: ~35 ratings per row, items: ~1170 per row
    out = {}
    for name, k in (("auto", fms._native.ALS_KERNEL_TC), ("tma", fms._native.ALS_KERNEL_TC_TMA), ("ldgsts", fms._native.ALS_KERNEL_TC_LDGSTS)):
        ctx.als_set_kernel(k)
        ctx.als_fit(u, i, r, 64, 3, 0.9, 42)
        out[name] = (ctx.als_get_factors(0)[1], ctx.als_get_factors(1)[1])
    ctx.als_set_kernel(fms._native.ALS_KERNEL_AUTO)
    for name in ("tma", "ldgsts"):
        assert np.array_equal(out[name][0], out["auto"][0]) and np.array_equal(out[name][1], out["auto"][1])


@pytest.mark.parametrize("rank", [8, 64])
def test_chunked_half_step_equals_full(fms, ctx, rank):
    """fms_als_half_step_chunk over all chunks == one fms_als_half_step (the pipelined exchange of dist.py relies on it)."""
    u, i, r = _case(fms, 1000, 150, 20000, seed=21)
    ctx.als_load(u, i, r)
    ctx.als_init(rank, 42)
    ctx.als_half_step(0, 0.9)
    ctx.als_half_step(1, 0.9)
    _, Xref = ctx.als_get_factors(0)
    _, Yref = ctx.als_get_factors(1)
    c2 = fms.Context(0)
    c2.als_load(u, i, r)
    c2.als_init(rank, 42)
    for which in (0, 1):
        for k in range(3):
            c2.als_half_step_chunk(which, 0.9, k, 3)
    _, X = c2.als_get_factors(0)
    _, Y = c2.als_get_factors(1)
    assert np.array_equal(X, Xref) and np.array_equal(Y, Yref)
    with pytest.raises(fms.FlinkMSError):
        c2.als_half_step_chunk(0, 0.9, 3, 3)
    c2.close()


def test_factor_rows_and_device_pointer_load(fms, ctx):
    """fms_als_get_factor_rows returns slices of fms_als_get_factors; fms_als_load accepts device pointers."""
    import torch
    u, i, r = _case(fms, 500, 120, 9000, seed=22)
    ctx.als_fit(u, i, r, 16, 2, 0.9, 42)
    ids, X = ctx.als_get_factors(0)
    for b, e in ((0, len(ids)), (17, 203), (5, 5)):
        ids_s, Xs = ctx.als_get_factor_rows(0, b, e)
        assert np.array_equal(ids_s, ids[b:e]) and np.array_equal(Xs, X[b:e])
    with pytest.raises(fms.FlinkMSError):
        ctx.als_get_factor_rows(0, 0, len(ids) + 1)
    du, di, dr = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (u.astype(np.int64), i.astype(np.int64), r.astype(np.float64)))
    torch.cuda.synchronize()
    c2 = fms.Context(0)
    assert c2.als_load_device(du.data_ptr(), di.data_ptr(), dr.data_ptr(), len(u)) == (len(ids), len(np.unique(i)))
    c2.als_init(16, 42)
    for _ in range(2):
        c2.als_half_step(0, 0.9)
        c2.als_half_step(1, 0.9)
    c2.als_half_step(0, 0.9)
    ids2, X2 = c2.als_get_factors(0)
    assert np.array_equal(ids2, ids) and np.array_equal(X2, X)
    c2.close()


def test_online_sgd_step_bit_exact_against_restatement(fms, ctx):
    """N2: fms_als_sgd_step == the Java arithmetic of SGD.java:160-217 (restated in oracle/np_restatement.sgd_step), bit for
    bit: snapshot reads, mean-vector cold start, last record wins on apply, unknown ids returned but not inserted."""
    from oracle import np_restatement as npr
    rng = np.random.default_rng(5)
    for rank in (5, 64, 100):
        uid = np.array([3, 7, 8, 20, 21, 1000], np.int64)
        iid = np.array([-4, 0, 5, 9], np.int64)
        U = rng.standard_normal((len(uid), rank))
        I = rng.standard_normal((len(iid), rank))
        um, im = U.mean(0), I.mean(0)
        ctx.als_set_model(uid, U, iid, I)
        users = np.array([7, 7, 3, 999, 20, 7, 1000, 8], np.int64)            # 999: unknown user; 7 three times
        items = np.array([0, 5, 5, 9, 123, 0, -4, 9], np.int64)               # 123: unknown item
        ratings = rng.uniform(1, 5, len(users))
        ou, oi = ctx.als_sgd_step(users, items, ratings, 0.05, 0.01, 0.02, um, im, apply=False)
        ru, ri, _, _ = npr.sgd_step(uid, U, iid, I, users, items, ratings, 0.05, 0.01, 0.02, um, im, apply=False)
        assert np.array_equal(ou, ru) and np.array_equal(oi, ri)
        _, U0 = ctx.als_get_factors(0)
        assert np.array_equal(U0, U)                                            # apply=False leaves the model alone
        ou, oi = ctx.als_sgd_step(users, items, ratings, 0.05, 0.01, 0.02, um, im, apply=True)
        ru, ri, U2, I2 = npr.sgd_step(uid, U, iid, I, users, items, ratings, 0.05, 0.01, 0.02, um, im, apply=True)
        assert np.array_equal(ou, ru) and np.array_equal(oi, ri)
        ids_u, Ug = ctx.als_get_factors(0)
        ids_i, Ig = ctx.als_get_factors(1)
        assert np.array_equal(ids_u, uid) and np.array_equal(ids_i, iid)        # nothing inserted
        assert np.array_equal(Ug, U2) and np.array_equal(Ig, I2)
    with pytest.raises(fms.FlinkMSError):
        ctx.als_sgd_step(users, items, ratings, 0.05, 0.0, 0.0, None, im)


def test_online_sgd_operator_emits_reference_records(fms, ctx):
    from oracle import np_restatement as npr
    formats = __import__("importlib").import_module("flink-ms_b200.formats")
    uid = np.array([1, 2], np.int64); iid = np.array([10, 11], np.int64)
    U = np.array([[0.5, -1.25], [2.0, 0.125]]); I = np.array([[1.0, 3.0], [-0.75, 0.2]])
    ctx.als_set_model(uid, U, iid, I)
    sgd = fms.OnlineSGD(ctx, {"learningRate": "0.1", "userMean": "0.1;0.2", "itemMean": "0.3;0.4"})
    recs = sgd.step([(1, 11, 4.0), (5, 10, 2.5)], apply=False)
    ru, ri, _, _ = npr.sgd_step(uid, U, iid, I, [1, 5], [11, 10], [4.0, 2.5], 0.1, 0.0, 0.0, [0.1, 0.2], [0.3, 0.4])
    assert recs == [formats.als_factor_row(1, "U", ru[0]), formats.als_factor_row(11, "I", ri[0]),
                    formats.als_factor_row(5, "U", ru[1]), formats.als_factor_row(10, "I", ri[1])]
    assert recs[0].startswith("1,U,") and recs[1].startswith("11,I,")
    with pytest.raises(RuntimeError):
        fms.OnlineSGD(ctx, {}).step([(1, 11, 4.0)])


def test_query_state_sees_online_sgd_updates(fms, ctx):
    """SGD.java:204-217 writes the updated factors back into the queryable state: a point lookup after an applied batch must
    return the NEW row (the client reads the device model, not the host copy it was loaded from)."""
    formats = __import__("importlib").import_module("flink-ms_b200.formats")
    uid = np.array([1, 2], np.int64); iid = np.array([10, 11], np.int64)
    U = np.array([[0.5, -1.25], [2.0, 0.125]]); I = np.array([[1.0, 3.0], [-0.75, 0.2]])
    client = formats.QueryableStateClient(ctx)
    client.load_als(uid, U, iid, I)
    before = client.queryState("ALS_MODEL", "1-U")
    assert before == ("1-U", "0.5;-1.25")
    recs = fms.OnlineSGD(ctx, {"learningRate": "0.1", "userMean": "0.1;0.2", "itemMean": "0.3;0.4"}).step([(1, 11, 4.0)], apply=True)
    after = client.queryState("ALS_MODEL", "1-U")
    assert after != before
    assert recs[0] == "1,U," + after[1]                     # the record the operator emitted is what the state now holds
    assert client.queryState("ALS_MODEL", "2-U") == ("2-U", "2.0;0.125")      # untouched row


@pytest.mark.parametrize("scale,lam", [(1e-3, 0.9), (1.0, 1e-3), (300.0, 0.9), (1.0, 30.0)])
def test_rank64_solver_scaling_over_rating_and_lambda_ranges(fms, ctx, scale, lam):
    """The rank-64 solver multiplies fp16 hi/lo operands on a system scaled by powers of two (matrix by its first diagonal entry,
    right-hand side by its own magnitude): ratings of very different size and regularisation from weak to dominant must stay
    inside the half-step tolerance -- or, where the system is so ill-conditioned that fp32 itself cannot (lambda = 1e-3 with
    fewer ratings than factors), within 3x of what the plain fp32 CUDA-core kernel reaches against the fp64 oracle."""
    u, i, r = _case(fms, 400, 120, 9000, seed=77)
    r = r * scale
    errs = {}
    for kernel in (fms._native.ALS_KERNEL_GENERIC, fms._native.ALS_KERNEL_AUTO):
        orc = oracle.OracleALS(u, i, r)
        ctx.als_set_kernel(kernel)
        ctx.als_load(u, i, r)
        ctx.als_init(64, 42)
        _, Y0 = ctx.als_get_factors(fms._native.ITEMS)
        orc.init(64, 42)
        orc.set_factors(1, Y0)
        ctx.als_half_step(fms._native.USERS, lam)
        orc.half_step(0, lam)
        _, X = ctx.als_get_factors(fms._native.USERS)
        _, OX = orc.factors(0)
        assert ctx.als_last_info() == 0
        e_users = rel_err(X, OX)
        orc.set_factors(0, X)
        ctx.als_half_step(fms._native.ITEMS, lam)      # now the Gramian is built from the (scaled) user factors
        orc.half_step(1, lam)
        _, Y = ctx.als_get_factors(fms._native.ITEMS)
        _, OY = orc.factors(1)
        assert ctx.als_last_info() == 0
        errs[kernel] = max(e_users, rel_err(Y, OY))
    ctx.als_set_kernel(fms._native.ALS_KERNEL_AUTO)
    bound = max(HALF_TOL, 3.0 * errs[fms._native.ALS_KERNEL_GENERIC])
    assert errs[fms._native.ALS_KERNEL_AUTO] <= bound, errs
