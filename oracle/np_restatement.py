"""Independent numpy / pure-Python restatement of the same FlinkML algorithms,
written separately from flinkml_oracle.c so the two can be diffed.  Small cases only (Python loops).
TEST INFRASTRUCTURE ONLY -- parity unpinned (see flinkml_oracle.c header).
"""
from __future__ import annotations

import numpy as np

_MULT = 0x5DEECE66D
_MASK = (1 << 48) - 1


def _s32(v: int) -> int:
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


class JavaRandom:
    """java.util.Random per the JDK specification (Appendix B), in arbitrary-precision Python ints."""

    def __init__(self, seed: int):
        self.state = (seed ^ _MULT) & _MASK

    def next(self, bits: int) -> int:
        self.state = (self.state * _MULT + 0xB) & _MASK
        return _s32(self.state >> (48 - bits))

    def next_int(self, bound: int | None = None) -> int:
        if bound is None:
            return self.next(32)
        r = self.next(31)
        m = bound - 1
        if bound & m == 0:
            return _s32((bound * r) >> 31)
        u = r
        while True:
            r = u % bound
            if _s32(u - r + m) >= 0:
                return r
            u = self.next(31)

    def next_double(self) -> float:
        return ((self.next(26) << 27) + self.next(27)) * (1.0 / (1 << 53))


def als_fit(user, item, rating, rank, iterations, lam, seed):
    """ALS-WR with FlinkML's schedule: init items from Random(id ^ seed); T x (users, items); final users."""
    user = np.asarray(user, np.int64)
    item = np.asarray(item, np.int64)
    rating = np.asarray(rating, np.float64)
    uid, ui = np.unique(user, return_inverse=True)
    iid, ii = np.unique(item, return_inverse=True)
    Y = np.empty((len(iid), rank))
    for k, i in enumerate(iid):
        r = JavaRandom(int(i) ^ seed)
        Y[k] = [r.next_double() for _ in range(rank)]
    X = np.zeros((len(uid), rank))

    def half(target_idx, partner_idx, P, n_rows):
        out = np.zeros((n_rows, rank))
        order = np.argsort(target_idx, kind="stable")
        t_sorted = target_idx[order]
        starts = np.searchsorted(t_sorted, np.arange(n_rows + 1))
        for row in range(n_rows):
            sel = order[starts[row]:starts[row + 1]]
            V = P[partner_idx[sel]]
            A = V.T @ V + lam * len(sel) * np.eye(rank)
            b = V.T @ rating[sel]
            out[row] = np.linalg.solve(A, b)
        return out

    for _ in range(iterations):
        X = half(ui, ii, Y, len(uid))
        Y = half(ii, ui, X, len(iid))
    X = half(ui, ii, Y, len(uid))
    return uid, X, iid, Y


def svm_fit(X, y, block_id, blocks, iterations, local_iterations, lam=1.0, stepsize=1.0, seed=0, rng_mode=0):
    """CoCoA outer loop + SDCA local solver, pure Python."""
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64)
    n, D = X.shape
    members = [np.flatnonzero(np.asarray(block_id) == k) for k in range(blocks)]
    alphas = [np.zeros(len(m)) for m in members]
    rngs = [JavaRandom(seed ^ k) for k in range(blocks)]
    shared = JavaRandom(seed ^ 0)
    w = np.zeros(D)
    ln = lam * n
    for _ in range(iterations):
        dws = []
        for k in range(blocks):
            rows = members[k]
            nk = len(rows)
            if nk == 0:
                dws.append(np.zeros(D))
                continue
            rng = rngs[k] if rng_mode == 0 else shared
            H = local_iterations if local_iterations > 0 else nk
            wl = w.copy()
            dw = np.zeros(D)
            tmp = alphas[k].copy()
            dal = np.zeros(nk)
            for _ in range(H):
                idx = rng.next_int(nk)
                x = X[rows[idx]]
                yy = y[rows[idx]]
                g = (yy * float(x @ wl) - 1.0) * ln
                if g != 0.0:
                    q = float(x @ x)
                    an = min(max(tmp[idx] - g / q, 0.0), 1.0) if q != 0.0 else 1.0
                    da = an - tmp[idx]
                    step = x * (yy * da / ln)
                    wl += step
                    dw += step
                    tmp[idx] += da
                    dal[idx] += da
            alphas[k] += dal * stepsize
            dws.append(dw)
        for k in range(blocks):
            w = w + dws[k] * (stepsize / blocks)
    alpha = np.zeros(n)
    for k in range(blocks):
        alpha[members[k]] = alphas[k]
    return w, alpha


def sgd_step(user_ids, U, item_ids, I, users, items, ratings, learning_rate=0.1, user_reg=0.0, item_reg=0.0,
             user_mean=None, item_mean=None, apply=False):
    """als-ms online SGD for a micro-batch (als-ms/src/main/java/de/tub/it4bi/modelserving/qs/SGD.java:160-217), fp64,
    Java operation order (every operation rounded, no fused multiply-add: plain Python floats do exactly that).
    Snapshot semantics: every record reads the model as it was before the batch.  Returns (out_user, out_item, U', I');
    with apply the last record of a known id overwrites its row (ALSKafkaConsumer.java:78-92), unknown ids are not
    inserted.  PARITY UNPINNED like the rest of oracle/ (the reference has no tests for this path)."""
    user_ids = np.asarray(user_ids, np.int64); item_ids = np.asarray(item_ids, np.int64)
    U = np.asarray(U, np.float64); I = np.asarray(I, np.float64)
    if user_mean is None or item_mean is None:
        raise ValueError("Unable to load User mean or item mean factors.")       # SGD.java:150-152
    d = U.shape[1] if U.size else I.shape[1]
    n = len(users)
    ou = np.empty((n, d)); oi = np.empty((n, d))
    U2, I2 = U.copy(), I.copy()
    for k in range(n):
        a = int(np.searchsorted(user_ids, users[k])); a = a if a < len(user_ids) and user_ids[a] == users[k] else -1
        b = int(np.searchsorted(item_ids, items[k])); b = b if b < len(item_ids) and item_ids[b] == items[k] else -1
        x = [float(v) for v in (U[a] if a >= 0 else user_mean)]
        y = [float(v) for v in (I[b] if b >= 0 else item_mean)]
        pred = 0.0
        for j in range(d):
            pred += x[j] * y[j]
        err = float(ratings[k]) - pred
        for j in range(d):
            ou[k, j] = x[j] + learning_rate * (err * y[j] - user_reg * x[j])
            oi[k, j] = y[j] + learning_rate * (err * x[j] - item_reg * y[j])
        if apply:
            if a >= 0:
                U2[a] = ou[k]
            if b >= 0:
                I2[b] = oi[k]
    return ou, oi, U2, I2
