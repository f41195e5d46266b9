"""Regenerates tests/golden/*.json.

jrand_kat.json: java.util.Random known answers.  The `published` block holds widely published JDK outputs
(new Random(42).nextInt() = -1170105035, new Random(0).nextInt() = -1155484576, the nextInt(100) sequence
60 48 29 47 ... of seed 0, nextDouble() of seeds 0 and 42) typed in by hand -- NOT produced by our code; the
`derived` block is produced by the pure-Python JavaRandom (oracle/np_restatement.py) that reproduces every
published value, and pins longer streams incl. the rejection path of nextInt(bound).

oracle_small.json: small ALS / SVM cases solved by the fp64 oracle (self-goldens: they pin the oracle against
regressions; they are NOT outputs of the Flink reference, which cannot run here -- parity unpinned).

Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from oracle.np_restatement import JavaRandom  # noqa: E402

PUBLISHED = {
    "ints": {"42": [-1170105035, 234785527, -1360544799, 205897768],
             "0": [-1155484576, -723955400, 1033096058, -1690734402, -1557280266, 1327362106, -1930858313, 502539523,
                   -1728529858, -938301587]},
    "doubles": {"0": [0.730967787376657, 0.24053641567148587, 0.6374174253501083],
                "42": [0.7275636800328681, 0.6832234717598454, 0.30871945533265976]},
    "bounded": {"0:100": [60, 48, 29, 47, 15, 53, 91, 61, 19, 54], "42:10": [0, 3, 8, 4, 0, 5, 5, 8, 9, 3]},
}


def main():
    derived = {"bounded": {}, "doubles": {}, "ints": {}}
    for seed, bound, n in [(42, 1000, 64), (7 ^ 42, 33334, 64), (-5, (1 << 30) + 1, 64), (123456789012345, 3, 64),
                           (42 ^ 3, 1 << 12, 32)]:
        r = JavaRandom(seed)
        derived["bounded"][f"{seed}:{bound}"] = [r.next_int(bound) for _ in range(n)]
    for seed in (42 ^ 17, -1, 1 << 40):
        r = JavaRandom(seed)
        derived["doubles"][str(seed)] = [r.next_double() for _ in range(16)]
        r = JavaRandom(seed)
        derived["ints"][str(seed)] = [r.next_int() for _ in range(16)]
    json.dump({"published": PUBLISHED, "derived": derived}, open(os.path.join(HERE, "jrand_kat.json"), "w"), indent=1)

    # ---- oracle self-goldens
    rng = np.random.default_rng(7)
    U, I, nnz, d = 40, 30, 400, 6
    u = rng.integers(0, U, nnz) * 2 + 100
    i = rng.integers(0, I, nnz) * 3 - 10
    r = np.clip(np.round(rng.normal(3.5, 1, nnz) * 2) / 2, 0.5, 5)
    h = oracle.OracleALS(u, i, r)
    h.fit(d, 4, 0.9, 42)
    uid, X = h.factors(0)
    iid, Y = h.factors(1)
    n, D, K = 200, 12, 4
    Xs = rng.standard_normal((n, D)).astype(np.float32)
    ws = rng.standard_normal(D)
    y = np.sign(Xs @ ws + 0.1 * rng.standard_normal(n))
    y[y == 0] = 1
    bid = (np.arange(n) % K).astype(np.int32)
    w10 = oracle.svm_fit(Xs, y, bid, K, 5, 10, 1.0, 1.0, 42, 0)
    wep = oracle.svm_fit(Xs, y, bid, K, 5, -1, 1.0, 1.0, 42, 0)
    json.dump({
        "als": {"user": u.tolist(), "item": i.tolist(), "rating": r.tolist(), "rank": d, "iterations": 4, "lambda": 0.9,
                "seed": 42, "user_ids": uid.tolist(), "U": X.tolist(), "item_ids": iid.tolist(), "I": Y.tolist()},
        "svm": {"X": Xs.astype(np.float64).tolist(), "y": y.tolist(), "block_id": bid.tolist(), "blocks": K, "iterations": 5,
                "seed": 42, "w_H10": w10.tolist(), "w_epoch": wep.tolist()},
    }, open(os.path.join(HERE, "oracle_small.json"), "w"))


if __name__ == "__main__":
    main()
