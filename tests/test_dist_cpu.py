"""world_size-2 gloo tests of the sharding logic (flink-ms_b200/dist.py) on CPU.  The compute backend here is the
fp64 oracle restricted to the rank's shard (tests may use the oracle; the product backend is the CUDA ctx): the
sharded run must reproduce the unsharded oracle exactly (ALS) / to round-off (SVM all-reduce order)."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleALSBackend:
    def __init__(self, orc, rank_dim, world, me):
        self.o, self.d, self.world, self.me = orc, rank_dim, world, me
        D = importlib.import_module("flink-ms_b200.dist")
        self.shard = D.shard_range
        self.full = {}
        for which, n in ((0, orc.n_users), (1, orc.n_items)):
            s = (n + world - 1) // world
            self.full[which] = torch.zeros(s * world, rank_dim, dtype=torch.float64)
        self.full[1][:orc.n_items] = torch.from_numpy(orc.factors(1)[1])

    def half_step(self, which, lam):
        n = self.o.n_users if which == 0 else self.o.n_items
        b, e, _ = self.shard(n, self.world, self.me)
        self.o.half_step(which, lam, rows=(b, e))
        self.full[which][b:e] = torch.from_numpy(self.o.factors(which)[1][b:e])

    def factor_tensor(self, which, rank_dim):
        return self.full[which]

    def updated(self, which):
        n = self.o.n_users if which == 0 else self.o.n_items
        self.o.set_factors(which, self.full[which][:n].numpy())


def _als_worker(me, world, port, q):
    sys.path.insert(0, ROOT)
    import oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=me, world_size=world)
    D = importlib.import_module("flink-ms_b200.dist")
    synth = importlib.import_module("flink-ms_b200.synth")
    u, i, r = synth.als_ratings(101, 67, 1500, seed=3)
    orc = oracle.OracleALS(u, i, r)
    orc.init(5, 42)
    dals = D.DistributedALS(OracleALSBackend(orc, 5, world, me), 5)
    dals.fit(2, 0.9)
    q.put((me, orc.factors(0)[1], orc.factors(1)[1]))
    dist.destroy_process_group()


def test_sharded_als_equals_unsharded_over_gloo():
    import oracle
    synth = importlib.import_module("flink-ms_b200.synth")
    u, i, r = synth.als_ratings(101, 67, 1500, seed=3)
    ref = oracle.OracleALS(u, i, r)
    ref.fit(5, 2, 0.9, 42)
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctxm.Process(target=_als_worker, args=(g, 2, port, q)) for g in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(60) for p in procs]
    for me, X, Y in res:
        assert np.array_equal(X, ref.factors(0)[1]) and np.array_equal(Y, ref.factors(1)[1]), me


class NumpySVMBackend:
    """Rank-local CoCoA sweep in fp64 numpy (restatement), blocks k % world == me."""

    def __init__(self, X, y, bid, K, world, me, H, seed):
        R = importlib.import_module("oracle.np_restatement")
        self.R, self.X, self.y, self.K, self.H = R, X.astype(np.float64), y, K, H
        self.n = len(y)
        self.mine = [k for k in range(K) if k % world == me]
        self.members = {k: np.flatnonzero(bid == k) for k in self.mine}
        self.alpha = {k: np.zeros(len(self.members[k])) for k in self.mine}
        self.rng = {k: R.JavaRandom(seed ^ k) for k in self.mine}
        self.w = np.zeros(X.shape[1])
        self.delta = torch.zeros(X.shape[1], dtype=torch.float64)

    def sweep(self, local_iterations):
        total = np.zeros_like(self.w)
        ln = 1.0 * self.n
        for k in self.mine:
            rows = self.members[k]
            nk = len(rows)
            H = local_iterations if local_iterations > 0 else nk
            wl = self.w.copy()
            for _ in range(H):
                idx = self.rng[k].next_int(nk)
                x, yy = self.X[rows[idx]], self.y[rows[idx]]
                g = (yy * float(x @ wl) - 1.0) * ln
                if g != 0.0:
                    q = float(x @ x)
                    an = min(max(self.alpha[k][idx] - g / q, 0.0), 1.0) if q != 0.0 else 1.0
                    da = an - self.alpha[k][idx]
                    step = x * (yy * da / ln)
                    wl += step
                    total += step
                    self.alpha[k][idx] += da
        self.delta.copy_(torch.from_numpy(total))

    def delta_tensor(self):
        return self.delta

    def apply(self):
        self.w = self.w + self.delta.numpy() * (1.0 / self.K)


def _svm_worker(me, world, port, q):
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=me, world_size=world)
    D = importlib.import_module("flink-ms_b200.dist")
    synth = importlib.import_module("flink-ms_b200.synth")
    X, y, _ = synth.svm_data(240, 12, seed=2)
    bid = np.arange(240) % 4
    dsvm = D.DistributedSVM(NumpySVMBackend(X, y, bid, 4, world, me, -1, 42))
    dsvm.fit(3, -1)
    q.put((me, dsvm.b.w))
    dist.destroy_process_group()


def test_sharded_svm_matches_unsharded_over_gloo():
    import oracle
    synth = importlib.import_module("flink-ms_b200.synth")
    X, y, _ = synth.svm_data(240, 12, seed=2)
    bid = np.arange(240) % 4
    want = oracle.svm_fit(X, y, bid, 4, 3, -1, 1.0, 1.0, 42, 0)
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctxm.Process(target=_svm_worker, args=(g, 2, port, q)) for g in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(60) for p in procs]
    for me, w in res:
        assert np.allclose(w, want, rtol=1e-12, atol=1e-14), me


def test_shard_range_covers_rows_once():
    D = importlib.import_module("flink-ms_b200.dist")
    for n in (0, 1, 7, 8, 9, 1000003):
        for world in (1, 2, 3, 8):
            covered = []
            for g in range(world):
                b, e, s = D.shard_range(n, world, g)
                assert 0 <= b <= e <= n and e - b <= s
                covered.append((b, e))
            assert sum(e - b for b, e in covered) == n
            assert all(covered[k][1] == covered[k + 1][0] or covered[k + 1][0] == n for k in range(world - 1))


class ExchangeSVMBackend(NumpySVMBackend):
    """The fused exchange path of ContextSVMBackend restated on gloo: every rank's delta lands in slot [rank] of every rank's
    buffer (all_gather), the slots are summed IN RANK ORDER and applied -- what k_svm_exchange_apply does over NVLink.  sweep /
    delta_tensor / apply must not be called by DistributedSVM once p2p is on."""
    p2p = True

    def round_exchange(self, local_iterations):
        NumpySVMBackend.sweep(self, local_iterations)
        world = dist.get_world_size()
        slots = [torch.zeros_like(self.delta) for _ in range(world)]
        dist.all_gather(slots, self.delta)
        s = torch.zeros_like(self.delta)
        for r in range(world):
            s += slots[r]
        self.w = self.w + s.numpy() * (1.0 / self.K)

    def sweep(self, local_iterations):
        raise AssertionError("DistributedSVM.round must use round_exchange when the backend has p2p on")

    def apply(self):
        raise AssertionError("DistributedSVM.round must use round_exchange when the backend has p2p on")


def _svm_exchange_worker(me, world, port, q):
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=me, world_size=world)
    D = importlib.import_module("flink-ms_b200.dist")
    synth = importlib.import_module("flink-ms_b200.synth")
    X, y, _ = synth.svm_data(240, 12, seed=2)
    bid = np.arange(240) % 4
    dsvm = D.DistributedSVM(ExchangeSVMBackend(X, y, bid, 4, world, me, -1, 42))
    dsvm.fit(3, -1)
    q.put((me, dsvm.b.w))
    dist.destroy_process_group()


def test_sharded_svm_exchange_path_is_rank_ordered_and_identical_on_all_ranks():
    import oracle
    synth = importlib.import_module("flink-ms_b200.synth")
    X, y, _ = synth.svm_data(240, 12, seed=2)
    bid = np.arange(240) % 4
    want = oracle.svm_fit(X, y, bid, 4, 3, -1, 1.0, 1.0, 42, 0)
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctxm.Process(target=_svm_exchange_worker, args=(g, 2, port, q)) for g in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    assert np.array_equal(res[0], res[1])                       # same summation order on every rank: the same bits
    assert np.allclose(res[0], want, rtol=1e-12, atol=1e-14)
