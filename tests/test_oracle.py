"""CPU tests of the oracle: java.util.Random KATs, C oracle vs the independent numpy restatement, linear-algebra
and optimisation identities that do not depend on recollecting FlinkML."""
import importlib
import json
import os

import numpy as np
import pytest

import oracle
from oracle import np_restatement as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_jrand_published_kats():
    kat = json.load(open(os.path.join(GOLD, "jrand_kat.json")))["published"]
    for seed, vals in kat["ints"].items():
        assert oracle.jrand_ints(int(seed), len(vals)).tolist() == vals
    for seed, vals in kat["doubles"].items():
        assert oracle.jrand_doubles(int(seed), len(vals)).tolist() == vals     # bit-exact doubles
    for key, vals in kat["bounded"].items():
        seed, bound = map(int, key.split(":"))
        assert oracle.jrand_bounded(seed, bound, len(vals)).tolist() == vals


def test_jrand_derived_streams_match_python_ints():
    kat = json.load(open(os.path.join(GOLD, "jrand_kat.json")))["derived"]
    for key, vals in kat["bounded"].items():
        seed, bound = map(int, key.split(":"))
        assert oracle.jrand_bounded(seed, bound, len(vals)).tolist() == vals
    for seed, vals in kat["doubles"].items():
        assert oracle.jrand_doubles(int(seed), len(vals)).tolist() == vals
    for seed, vals in kat["ints"].items():
        assert oracle.jrand_ints(int(seed), len(vals)).tolist() == vals


def test_jrand_rejection_path_is_exercised():
    # bound just above 2^30 rejects ~half of the draws: the C loop and the Python loop must agree
    bound = (1 << 30) + 1
    r = R.JavaRandom(99)
    want = [r.next_int(bound) for _ in range(200)]
    assert oracle.jrand_bounded(99, bound, 200).tolist() == want
    assert max(want) < bound and min(want) >= 0


def _als_case(seed=1, U=60, I=50, nnz=900):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, U, nnz) * 3 + 7
    i = rng.integers(0, I, nnz) * 5 - 20
    r = np.clip(np.round(rng.normal(3.5, 1, nnz) * 2) / 2, 0.5, 5)
    return u, i, r


def test_als_c_oracle_matches_numpy_restatement():
    u, i, r = _als_case()
    h = oracle.OracleALS(u, i, r)
    assert h.fit(8, 5, 0.9, 42) == 0
    uid, X = h.factors(0)
    iid, Y = h.factors(1)
    uid2, X2, iid2, Y2 = R.als_fit(u, i, r, 8, 5, 0.9, 42)
    assert (uid == uid2).all() and (iid == iid2).all()
    assert np.abs(X - X2).max() <= 1e-12 * np.abs(X2).max()
    assert np.abs(Y - Y2).max() <= 1e-12 * np.abs(Y2).max()


def test_als_rows_satisfy_normal_equations():
    u, i, r = _als_case(3)
    lam, d = 0.9, 7
    h = oracle.OracleALS(u, i, r)
    h.fit(d, 3, lam, 42)
    uid, X = h.factors(0)
    iid, Y = h.factors(1)
    # returned users are the exact ALS-WR solve against the returned items
    for row, user in enumerate(uid):
        sel = u == user
        V = Y[np.searchsorted(iid, i[sel])]
        A = V.T @ V + lam * sel.sum() * np.eye(d)
        assert np.allclose(A @ X[row], V.T @ r[sel], rtol=1e-10, atol=1e-10)


def test_als_init_is_random_of_id_xor_seed():
    u, i, r = _als_case(5)
    h = oracle.OracleALS(u, i, r)
    h.init(4, 42)
    iid, Y = h.factors(1)
    for k in (0, len(iid) // 2, len(iid) - 1):
        rr = R.JavaRandom(int(iid[k]) ^ 42)
        assert Y[k].tolist() == [rr.next_double() for _ in range(4)]


def test_als_risk_non_increasing():
    u, i, r = _als_case(9, nnz=1500)
    lam = 0.9
    h = oracle.OracleALS(u, i, r)
    h.init(6, 42)
    h.half_step(0, lam)
    prev = h.risk(lam)
    for _ in range(4):
        for which in (1, 0):
            h.half_step(which, lam)
            cur = h.risk(lam)
            assert cur <= prev * (1 + 1e-12)
            prev = cur


def test_als_duplicates_are_separate_ratings_and_golden():
    g = json.load(open(os.path.join(GOLD, "oracle_small.json")))["als"]
    h = oracle.OracleALS(g["user"], g["item"], g["rating"])
    h.fit(g["rank"], g["iterations"], g["lambda"], g["seed"])
    uid, X = h.factors(0)
    iid, Y = h.factors(1)
    assert uid.tolist() == g["user_ids"] and iid.tolist() == g["item_ids"]
    assert np.allclose(X, np.array(g["U"]), rtol=1e-12, atol=1e-14)
    assert np.allclose(Y, np.array(g["I"]), rtol=1e-12, atol=1e-14)
    # the case holds duplicated (u,i) pairs: n_u counts every rating
    pairs = list(zip(g["user"], g["item"]))
    assert len(set(pairs)) < len(pairs)


def test_als_predict_unknown_ids():
    u, i, r = _als_case(2)
    h = oracle.OracleALS(u, i, r)
    h.fit(4, 1, 0.9, 42)
    out, found = h.predict([u[0], 10 ** 9, u[1]], [i[0], i[0], -12345])
    assert found.tolist() == [True, False, False]
    uid, X = h.factors(0)
    iid, Y = h.factors(1)
    assert out[0] == pytest.approx(X[np.searchsorted(uid, u[0])] @ Y[np.searchsorted(iid, i[0])], rel=1e-14)


def _svm_case(seed=1, n=400, D=24):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, D)).astype(np.float32)
    ws = rng.standard_normal(D)
    y = np.sign(X @ ws + 0.1 * rng.standard_normal(n))
    y[y == 0] = 1
    return X, y


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("H", [10, -1])
def test_svm_c_oracle_matches_python_restatement(mode, H):
    X, y = _svm_case()
    K = 4
    bid = np.arange(len(y)) % K
    w, a = oracle.svm_fit(X, y, bid, K, 6, H, 1.0, 1.0, 42, mode, return_alpha=True)
    w2, a2 = R.svm_fit(X, y, bid, K, 6, H, 1.0, 1.0, 42, mode)
    assert np.abs(w - w2).max() <= 1e-13 * max(np.abs(w2).max(), 1e-300)
    assert np.abs(a - a2).max() <= 1e-13


def test_svm_alpha_box_and_dual_ascent_with_one_block():
    X, y = _svm_case(4)
    bid = np.zeros(len(y), np.int32)
    prev = -1.0
    for T in (1, 2, 4, 8):
        # K=1, stepsize=1: CoCoA == plain SDCA, whose dual objective never decreases and w == w(alpha)
        w, a = oracle.svm_fit(X, y, bid, 1, T, -1, 1.0, 1.0, 42, 0, return_alpha=True)
        assert a.min() >= 0.0 and a.max() <= 1.0
        p, d = oracle.svm_objectives(X, y, w, a, 1.0)
        assert d >= prev - 1e-12 and p >= d - 1e-12
        prev = d
        w_of_alpha = (X.astype(np.float64) * (a * y)[:, None]).sum(0) / (1.0 * len(y))
        assert np.allclose(w, w_of_alpha, rtol=1e-10, atol=1e-12)


def test_svm_dtypes_agree_and_csr():
    X, y = _svm_case(6, n=150, D=10)
    X[::7] = 0.0                                    # all-zero rows: q == 0 -> alpha' = 1
    K = 3
    bid = np.arange(len(y)) % K
    w32 = oracle.svm_fit(X, y, bid, K, 4, -1, 0.5, 0.7, 11, 0)
    w64 = oracle.svm_fit(X.astype(np.float64), y, bid, K, 4, -1, 0.5, 0.7, 11, 0)
    ptr = np.arange(0, X.size + 1, X.shape[1])
    idx = np.tile(np.arange(X.shape[1]), X.shape[0])
    wcsr = oracle.svm_fit(None, y, bid, K, 4, -1, 0.5, 0.7, 11, 0, csr=(ptr, idx, X.astype(np.float64).ravel()), dim=X.shape[1])
    assert np.array_equal(w32, w64) and np.allclose(w32, wcsr, rtol=1e-13, atol=1e-15)


def test_svm_golden_and_index_stream():
    g = json.load(open(os.path.join(GOLD, "oracle_small.json")))["svm"]
    X = np.array(g["X"], np.float32)
    w10 = oracle.svm_fit(X, g["y"], g["block_id"], g["blocks"], g["iterations"], 10, 1.0, 1.0, g["seed"], 0)
    wep = oracle.svm_fit(X, g["y"], g["block_id"], g["blocks"], g["iterations"], -1, 1.0, 1.0, g["seed"], 0)
    assert np.allclose(w10, g["w_H10"], rtol=1e-12, atol=1e-15) and np.allclose(wep, g["w_epoch"], rtol=1e-12, atol=1e-15)
    r = R.JavaRandom(42 ^ 3)
    assert oracle.svm_index_stream(42, 3, 50, 40).tolist() == [r.next_int(50) for _ in range(40)]


def test_svm_predict_strict_threshold():
    X = np.array([[1.0, 0.0], [0.0, 0.0], [-1.0, 0.0]])
    w = np.array([2.0, 5.0])
    assert oracle.svm_predict(X, w, 0.0, False).tolist() == [1.0, -1.0, -1.0]      # raw == threshold -> -1
    assert oracle.svm_predict(X, w, 0.0, True).tolist() == [2.0, 0.0, -2.0]


def test_sgd_restatement_properties():
    """oracle/np_restatement.sgd_step: zero learning rate is the identity, zero regularisation moves u along v, the last
    record of an id wins on apply, unknown ids use the mean rows and are not inserted."""
    from oracle import np_restatement as npr
    uid = np.array([1, 4], np.int64); iid = np.array([2], np.int64)
    U = np.array([[1.0, 2.0], [0.5, -1.0]]); I = np.array([[3.0, -1.0]])
    um, im = np.array([0.1, 0.1]), np.array([0.2, 0.2])
    ou, oi, U2, I2 = npr.sgd_step(uid, U, iid, I, [1], [2], [5.0], 0.0, 0.3, 0.3, um, im, apply=True)
    assert np.array_equal(ou[0], U[0]) and np.array_equal(oi[0], I[0]) and np.array_equal(U2, U)
    ou, oi, _, _ = npr.sgd_step(uid, U, iid, I, [1], [2], [5.0], 0.1, 0.0, 0.0, um, im)
    err = 5.0 - (1.0 * 3.0 + 2.0 * -1.0)
    assert np.allclose(ou[0], U[0] + 0.1 * err * I[0]) and np.allclose(oi[0], I[0] + 0.1 * err * U[0])
    ou, oi, U2, I2 = npr.sgd_step(uid, U, iid, I, [1, 9, 1], [2, 2, 7], [5.0, 1.0, 2.0], 0.1, 0.0, 0.0, um, im, apply=True)
    assert np.array_equal(U2[0], ou[2]) and np.array_equal(I2[0], oi[1])          # last record of user 1 / item 2 wins
    assert np.array_equal(U2[1], U[1]) and U2.shape == U.shape                    # user 9 not inserted
    e9 = 1.0 - float(um @ I[0])
    assert np.allclose(ou[1], um + 0.1 * e9 * I[0])                               # cold start from the mean row
    with pytest.raises(ValueError):
        npr.sgd_step(uid, U, iid, I, [1], [2], [5.0], 0.1, 0.0, 0.0, None, im)


def test_als_half_step_matches_lapack_dposv():
    """FlinkML solves every row with netlib LAPACK `dposv("U", ...)` on the upper triangle of X^T X + lambda n I (SURVEY.md
    Appendix A.1).  The same routine is reachable here through scipy: the oracle's rows must agree with it to a few ulps
    (identical algorithm, possibly different blocking inside the BLAS)."""
    scipy_lapack = pytest.importorskip("scipy.linalg.lapack")
    synth = __import__("importlib").import_module("flink-ms_b200.synth")
    u, i, r = synth.als_ratings(120, 40, 1500, seed=31)
    for rank in (4, 16, 64):
        orc = oracle.OracleALS(u, i, r)
        orc.init(rank, 7)
        ids_i, Y = orc.factors(1)
        orc.half_step(0, 0.3)
        ids_u, X = orc.factors(0)
        for row in range(0, orc.n_users, 7):
            sel = u == ids_u[row]
            V = Y[np.searchsorted(ids_i, i[sel])]
            A = V.T @ V + 0.3 * sel.sum() * np.eye(rank)
            b = V.T @ r[sel]
            _, x, info = scipy_lapack.dposv(A, b.reshape(-1, 1), lower=0)
            assert info == 0
            assert np.abs(x[:, 0] - X[row]).max() <= 1e-11 * max(1.0, np.abs(X[row]).max())


def test_svm_converges_to_the_hinge_svm_optimum_of_liblinear():
    """Run long enough with one block, CoCoA/SDCA reaches the optimum of  lambda/2 |w|^2 + 1/n sum hinge(y w.x)  -- the
    objective FlinkML states for its SVM.  liblinear (through scikit-learn, C = 1 / (lambda n), hinge loss, no
    intercept) solves the same problem with an unrelated implementation: primal objectives must agree to 1e-3."""
    sk = pytest.importorskip("sklearn.svm")
    synth = __import__("importlib").import_module("flink-ms_b200.synth")
    X, y, _ = synth.svm_data(400, 12, seed=17)
    X = np.asarray(X, np.float64); y = np.asarray(y, np.float64)
    lam = 0.05
    w, alpha = oracle.svm_fit(X, y, np.zeros(len(y), np.int32), 1, 60, len(y), regularization=lam, stepsize=1.0, seed=5, return_alpha=True)
    primal, dual = oracle.svm_objectives(X, y, w, alpha, lam)
    assert primal - dual <= 5e-4 * max(1.0, abs(primal))                     # duality gap: converged
    ref = sk.LinearSVC(loss="hinge", C=1.0 / (lam * len(y)), fit_intercept=False, dual=True, tol=1e-10, max_iter=200000).fit(X, y)
    wr = ref.coef_.ravel()
    pr = 0.5 * lam * float(wr @ wr) + float(np.maximum(0.0, 1.0 - y * (X @ wr)).mean())
    assert abs(primal - pr) <= 1e-3 * max(1.0, abs(pr))


# ---------------------------------------------------------------------------------------------- risk register (recollected details)
@pytest.fixture
def restore_switches():
    yield
    for name in oracle.SWITCHES:
        oracle.set_switch(name, 1)


def test_switch_names_and_defaults():
    assert all(oracle.get_switch(n) == 1 for n in oracle.SWITCHES)
    with pytest.raises(KeyError):
        oracle.set_switch("no_such_switch", 0)


def test_switch_als_lambda_scaling_and_final_user_step(restore_switches):
    synth = importlib.import_module("flink-ms_b200.synth")
    u, i, r = synth.als_ratings(60, 40, 900, seed=3)
    lam, d = 0.9, 6

    def row_solution(o, row, weight_by_n):
        ids_i, Y = o.factors(1)
        sel = np.flatnonzero(u == o.factors(0)[0][row])
        V = Y[np.searchsorted(ids_i, i[sel])]
        A = V.T @ V + (lam * len(sel) if weight_by_n else lam) * np.eye(d)
        return np.linalg.solve(A, V.T @ r[sel])

    for weight_by_n in (1, 0):
        oracle.set_switch("als_lambda_times_n", weight_by_n)
        o = oracle.OracleALS(u, i, r)
        o.fit(d, 2, lam, 42)
        X = o.factors(0)[1]
        for row in (0, 7, 31):
            np.testing.assert_allclose(X[row], row_solution(o, row, weight_by_n), rtol=1e-9, atol=1e-12)
    oracle.set_switch("als_lambda_times_n", 1)
    # without the extra users half-step the returned users were solved against the PREVIOUS items
    o1 = oracle.OracleALS(u, i, r); o1.fit(d, 2, lam, 42)
    oracle.set_switch("als_final_user_step", 0)
    o0 = oracle.OracleALS(u, i, r); o0.fit(d, 2, lam, 42)
    np.testing.assert_array_equal(o0.factors(1)[1], o1.factors(1)[1])          # items: identical
    assert np.abs(o0.factors(0)[1] - o1.factors(0)[1]).max() > 1e-6            # users: one half-step apart
    o0.half_step(0, lam)
    np.testing.assert_allclose(o0.factors(0)[1], o1.factors(0)[1], rtol=0, atol=0)


def test_switch_svm_scaling_and_alpha_scale(restore_switches):
    synth = importlib.import_module("flink-ms_b200.synth")
    X, y, _ = synth.svm_data(120, 16, seed=4)
    blocks = 6
    bid = (np.arange(120) % 4).astype(np.int32)                               # blocks 4 and 5 stay empty
    base = oracle.svm_fit(X, y, bid, blocks, 1, -1, 1.0, 1.0, 7, 0)
    oracle.set_switch("svm_scaling_configured_blocks", 0)
    eff = oracle.svm_fit(X, y, bid, blocks, 1, -1, 1.0, 1.0, 7, 0)
    np.testing.assert_allclose(eff, base * (6.0 / 4.0), rtol=1e-12)          # one round from w = 0: only the scale differs
    oracle.set_switch("svm_scaling_configured_blocks", 1)
    w1, a1 = oracle.svm_fit(X, y, bid, blocks, 1, -1, 1.0, 0.5, 7, 0, return_alpha=True)
    oracle.set_switch("svm_alpha_scale_stepsize", 0)
    w0, a0 = oracle.svm_fit(X, y, bid, blocks, 1, -1, 1.0, 0.5, 7, 0, return_alpha=True)
    np.testing.assert_allclose(w0, w1, rtol=1e-12)                            # the weights of round 1 do not see alpha's scale
    np.testing.assert_allclose(a0 * blocks, a1, rtol=1e-12)                   # alpha += (stepsize / blocks) dAlpha vs stepsize dAlpha


def test_flink_goldens_if_present():
    """Pinned parity: tests/golden/flink_*.json come from real Flink 1.3.1 runs (tools/make_flink_goldens.sh).  No JVM has
    been available so far, so the files do not exist and the oracle stays PARITY UNPINNED; the day they do, this test is
    what pins it."""
    import glob
    import json
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "flink_*.json")))
    if not files:
        pytest.skip("no Flink-generated goldens (no JVM / flink-ml 1.3.1 in this image): parity unpinned")
    for f in files:
        gold = json.load(open(f))
        if gold["kind"] == "als":
            o = oracle.OracleALS(np.asarray(gold["user"]), np.asarray(gold["item"]), np.asarray(gold["rating"]))
            o.fit(gold["rank"], gold["iterations"], gold["lambda"], gold["seed"])
            for which, key in ((0, "user_factors"), (1, "item_factors")):
                ids, F = o.factors(which)
                want = np.asarray([gold[key][str(int(k))] for k in ids])
                np.testing.assert_allclose(F, want, rtol=1e-9, atol=1e-12)
        elif gold["kind"] == "svm":
            w = oracle.svm_fit(np.asarray(gold["X"]), np.asarray(gold["y"]), np.asarray(gold["block_id"], np.int32), gold["blocks"],
                               gold["iterations"], gold["local_iterations"], gold["regularization"], gold["stepsize"], gold["seed"],
                               gold["rng_mode"])
            np.testing.assert_allclose(w, np.asarray(gold["weights"]), rtol=1e-9, atol=1e-12)
