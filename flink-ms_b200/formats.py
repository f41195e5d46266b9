"""Model text / lookup schemas of the reference pipeline, host side.

The reference moves trained models as TEXT: `ALSImpl` writes `id,U|I,f1;f2;...` rows (ALSImpl.scala:83-85), `SVMImpl` writes
`featureId,weight` rows with 1-BASED ids or range-bucket rows `bucket,id:w;id:w...` with bucket = id / range
(SVMImpl.scala:31-46, :63-71), `ALSMeanVector` writes `MEAN,U|I,...` (ALSMeanVector.scala:28-35); the Kafka consumers
turn them into queryable state keyed `"<id>-U"` / `"<id>-I"` -> `"f1;f2;..."` (ALSKafkaConsumer.java:78-92) and
`"<featureId>"` / `"<bucket>"` -> value string (SVMKafkaConsumer.java:79-92); clients look keys up and parse the
value with Double.parseDouble (ALSPredictRandom.java:63-92, SVMPredictRandom.java:68-88).

This module writes and parses exactly those strings (numbers formatted like Java's Double.toString) and offers
`QueryableStateClient`, a stand-in for QueryClientHelper.queryState (QueryClientHelper.java:105-139) backed by a
GPU-resident model: `queryState(name, key)` returns the (key, value) tuple or None (Optional.empty) on a miss; the
batched numeric path is `Context.als_predict` / `svm_predict`.
"""
from __future__ import annotations

import math

import numpy as np


def java_double_to_string(x: float) -> str:
    """java.lang.Double.toString: shortest digits that uniquely distinguish the double (JDK >= 19; older JDKs emit a
    few non-shortest outliers, which Double.parseDouble reads back to the same value), decimal notation for
    1e-3 <= |x| < 1e7, computerised scientific notation `d.dddE[-]n` otherwise, always at least one fraction digit.
    (Known deviation: Double.MIN_VALUE prints as 4.9E-324 in Java, 5.0E-324 here; both parse back to the same double.)"""
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    sign = "-" if x < 0 else ""
    r = repr(abs(x))                                   # shortest round-trip digits
    if "e" in r or "E" in r:
        mant, exp = r.lower().split("e")
        exp = int(exp)
    else:
        mant, exp = r, 0
    if "." in mant:
        ip, fp = mant.split(".")
    else:
        ip, fp = mant, ""
    digits = (ip + fp).lstrip("0")
    # decimal exponent of the first significant digit
    if ip.strip("0"):
        e10 = len(ip.lstrip("0")) - 1 + exp
    else:
        lead = len(fp) - len(fp.lstrip("0"))
        e10 = -(lead + 1) + exp
    digits = digits.rstrip("0") or "0"
    if -3 <= e10 < 7:
        if e10 >= 0:
            ipart = digits[:e10 + 1].ljust(e10 + 1, "0")
            fpart = digits[e10 + 1:] or "0"
        else:
            ipart = "0"
            fpart = "0" * (-e10 - 1) + digits
        return f"{sign}{ipart}.{fpart}"
    fpart = digits[1:] or "0"
    return f"{sign}{digits[0]}.{fpart}E{e10}"


# ------------------------------------------------------------------------------------------------ ALS rows
def als_factor_row(id_: int, kind: str, factors) -> str:
    """OutputFactor.toString: s"$id,$typeOfFeature,${factors.mkString(";")}"  (ALSImpl.scala:84)."""
    return f"{int(id_)},{kind},{';'.join(java_double_to_string(v) for v in factors)}"


def als_model_rows(ids, factors, kind: str):
    return [als_factor_row(i, kind, f) for i, f in zip(ids, factors)]


def als_mean_row(factors, kind: str) -> str:
    """ALSMeanVector.scala:28-35: column mean of all rows -> "MEAN,U|I,m1;m2;..."."""
    f = np.asarray(factors, np.float64)
    acc = np.zeros(f.shape[1])
    for row in f:                                      # reduce((l, r) => l zip r map (_ + _)): sequential sum
        acc = acc + row
    return "MEAN," + kind + "," + ";".join(java_double_to_string(v / len(f)) for v in acc)


def parse_als_rows(lines):
    """`id,U|I,f1;f2;...` rows (MEAN rows are returned separately) -> (user_ids, U, item_ids, I, means)."""
    users, items, means = {}, {}, {}
    for line in lines:
        line = line.strip()
        if not line:
            continue
        key, kind, vec = line.split(",")              # ALSKafkaConsumer.java:78-80 splits on ","
        v = np.array([float(t) for t in vec.split(";")], np.float64)
        if key == "MEAN":
            means[kind] = v
        elif kind == "U":
            users[int(key)] = v
        elif kind == "I":
            items[int(key)] = v
        else:
            raise ValueError(f"unknown factor type {kind!r}")
    uid = np.array(sorted(users), np.int64)
    iid = np.array(sorted(items), np.int64)
    rank = len(next(iter(users.values()))) if users else (len(next(iter(items.values()))) if items else 0)
    U = np.array([users[k] for k in uid], np.float64).reshape(len(uid), rank)
    I = np.array([items[k] for k in iid], np.float64).reshape(len(iid), rank)
    return uid, U, iid, I, means


# ------------------------------------------------------------------------------------------------ SVM rows
def svm_model_rows(weights, partition: bool = False, range_: int = 1000):
    """SVMImpl.scala:31-46: (1-based id, weight) rows, or range-bucket rows `bucket,id:w;id:w` with bucket = id / range."""
    w = np.asarray(weights, np.float64)
    if not partition:
        return [f"{j + 1},{java_double_to_string(v)}" for j, v in enumerate(w)]
    rows = {}
    for j, v in enumerate(w):
        fid = j + 1
        rows.setdefault(fid // range_, []).append(f"{fid}:{java_double_to_string(v)}")
    return [f"{b},{';'.join(parts)}" for b, parts in sorted(rows.items())]


def parse_svm_rows(lines, dim: int | None = None):
    """Either row form -> dense weight vector (0-based).  Missing features stay 0."""
    pairs = {}
    for line in lines:
        line = line.strip()
        if not line:
            continue
        key, val = line.split(",", 1)
        if ":" in val:
            for kv in val.split(";"):                 # RangePartitionSVMPredict.java:89-95
                k, v = kv.split(":")
                pairs[int(k)] = float(v)
        else:
            pairs[int(key)] = float(val)
    n = dim if dim is not None else (max(pairs) if pairs else 0)
    w = np.zeros(n, np.float64)
    for k, v in pairs.items():
        if 1 <= k <= n:
            w[k - 1] = v
    return w


# ------------------------------------------------------------------------------------------------ lookup client
class QueryableStateClient:
    """QueryClientHelper<String, Tuple2<String,String>>.queryState(name, key) over a GPU-resident model.

    ALS_MODEL keys: "<id>-U", "<id>-I", "MEAN-U", "MEAN-I" -> (key, "f1;f2;...")   (ALSKafkaConsumer.java:78-92, SGD.java:143-147)
    SVM_MODEL keys: "<featureId>" -> (key, "w")  or, for a range-partitioned model, "<bucket>" -> (key, "id:w;id:w")
    A missing key returns None (Optional.empty, QueryClientHelper.java:136-139); None arguments raise like the
    reference's NullPointerException.  With a Context attached, ALS point lookups read the live device model (one row D2H), so
    they see OnlineSGD updates; without one they are served from the host copy.  Batches go through Context.als_predict /
    svm_predict on the device.  Ids unknown to the model stay unknown (the reference's Kafka consumer would insert them).
    """

    def __init__(self, ctx=None):
        self._ctx = ctx
        self._als = None
        self._svm = None
        self._closed = False

    # ---- model installation
    def load_als(self, user_ids, U, item_ids, I, means=None):
        self._als = (np.asarray(user_ids, np.int64), np.asarray(U, np.float64), np.asarray(item_ids, np.int64),
                     np.asarray(I, np.float64), dict(means or {}))
        if self._ctx is not None:
            self._ctx.als_set_model(self._als[0], self._als[1], self._als[2], self._als[3])

    def load_svm(self, weights, partition: bool = False, range_: int = 1000):
        w = np.asarray(weights, np.float64)
        self._svm = (w, bool(partition), int(range_))
        if self._ctx is not None:
            self._ctx.svm_set_weights(w)

    # ---- QueryClientHelper.queryState
    def queryState(self, name: str, key: str):
        if name is None or key is None:
            raise TypeError("queryState: name and key must not be null")       # NullPointerException in the reference
        if self._closed:
            raise RuntimeError("client is closed")
        if name == "ALS_MODEL":
            if self._als is None:
                return None
            uid, U, iid, I, means = self._als
            ident, _, kind = key.rpartition("-")
            if ident == "MEAN":
                v = means.get(kind)
                return None if v is None else (key, ";".join(java_double_to_string(x) for x in v))
            try:
                k = int(ident)
            except ValueError:
                return None
            ids, F = (uid, U) if kind == "U" else (iid, I) if kind == "I" else (None, None)
            if ids is None:
                return None
            pos = int(np.searchsorted(ids, k))
            if pos >= len(ids) or ids[pos] != k:
                return None
            row = F[pos]
            if self._ctx is not None:
                # the device model is the state: OnlineSGD.step(apply=True) updates it in place (SGD.java:204-217 writes the
                # new factors back into the same queryable state), a host copy taken at load time would go stale
                row = self._ctx.als_get_factor_rows(0 if kind == "U" else 1, pos, pos + 1)[1][0]
            return key, ";".join(java_double_to_string(x) for x in row)
        if name == "SVM_MODEL":
            if self._svm is None:
                return None
            w, partition, range_ = self._svm
            try:
                k = int(key)
            except ValueError:
                return None
            if not partition:
                return (key, java_double_to_string(w[k - 1])) if 1 <= k <= len(w) else None
            ids = [j for j in range(max(k * range_, 1), min((k + 1) * range_, len(w) + 1))]
            if not ids:
                return None
            return key, ";".join(f"{j}:{java_double_to_string(w[j - 1])}" for j in ids)
        return None                                                             # unknown state name

    def close(self):
        self._closed = True

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
