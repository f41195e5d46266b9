"""Block assignment of FlinkML 1.3.1's `FlinkMLTools.block` (what `SVM.fit` calls before the CoCoA loop; entered from
flink-svm/src/main/scala/de/tub/it4bi/SVMImpl.scala:29):

    blockID = element.hashCode() % numBlocks ; if (blockID < 0) blockID += numBlocks ; groupBy(blockID)

`element` is a `LabeledVector(label: Double, vector: Vector)` -- a Scala 2.11 case class with a primitive field, so the
compiler-generated hashCode is

    acc = 0xcafebabe; acc = mix(acc, doubleHash(label)); acc = mix(acc, vector.hashCode); finalizeHash(acc, 2)

(scala.runtime.Statics; mix / finalizeHash are MurmurHash3's), with FlinkML's vector hashes

    SparseVector.hashCode = ((3 * 41 + size) * 41 + Arrays.hashCode(indices)) * 41 + Arrays.hashCode(data)
    DenseVector.hashCode  = Arrays.hashCode(data)

`env.readLibSVM` (SVMImpl.scala:21) yields SparseVectors of dimension = largest index in the file.

RECOLLECTED like the rest of SURVEY.md Appendix A (the jar is not under /root/reference): the round-1 advisor and this
author recall the same code independently; round 1 had "block by source subtask" instead, kept here as `by_subtask` so the
two readings stay switchable until a JVM can pin one (tools/make_flink_goldens.sh).  Everything below is the public JDK /
Scala-library arithmetic (java.util.Arrays.hashCode, Double.doubleToLongBits, MurmurHash3), vectorised with numpy int32
wrap-around.
"""
from __future__ import annotations

import numpy as np

_U32 = np.uint32


def _rotl(x, r):
    x = x.astype(_U32)
    return ((x << _U32(r)) | (x >> _U32(32 - r))).astype(_U32)


def mix_last(h, k):
    k = (k.astype(_U32) * _U32(0xCC9E2D51)).astype(_U32)
    k = _rotl(k, 15)
    k = (k * _U32(0x1B873593)).astype(_U32)
    return (h.astype(_U32) ^ k).astype(_U32)


def mix(h, k):
    """scala.runtime.Statics.mix == MurmurHash3.mix (one 32-bit block)."""
    h = mix_last(h, k)
    h = _rotl(h, 13)
    return (h * _U32(5) + _U32(0xE6546B64)).astype(_U32)


def avalanche(h):
    h = h.astype(_U32)
    h ^= h >> _U32(16)
    h = (h * _U32(0x85EBCA6B)).astype(_U32)
    h ^= h >> _U32(13)
    h = (h * _U32(0xC2B2AE35)).astype(_U32)
    h ^= h >> _U32(16)
    return h


def finalize_hash(h, length):
    return avalanche(h.astype(_U32) ^ _U32(length))


def double_hash(d):
    """scala.runtime.Statics.doubleHash: int value if integral and in int range, else float bits if exactly a float, else
    low 32 bits of the long if integral, else hi ^ lo of the IEEE bits."""
    d = np.atleast_1d(np.asarray(d, np.float64))
    bits = d.view(np.uint64)
    with np.errstate(invalid="ignore", over="ignore"):
        is_int = np.isfinite(d) & (np.abs(d) < 2147483648.0) & (np.floor(d) == d)
        iv = np.where(is_int, d, 0.0).astype(np.int64)
        f = d.astype(np.float32)
        is_f32 = (f.astype(np.float64) == d)                               # NaN fails, +-inf pass (as in Java)
        is_long = np.isfinite(d) & (np.abs(d) < 9.2e18) & (np.floor(d) == d)
        lv = np.where(is_long, d, 0.0).astype(np.int64)
    out = (bits ^ (bits >> np.uint64(32))).astype(_U32)
    out = np.where(is_long, lv.astype(np.uint64).astype(_U32), out)
    out = np.where(is_f32, f.view(_U32), out)
    # Java's (int) cast of -0.0 is 0 and 0 == -0.0 holds, so -0.0 hashes like 0
    out = np.where(is_int, iv.astype(np.uint64).astype(_U32), out)
    return out.astype(_U32)


def _double_bits_hash(v):
    """(int)(bits ^ (bits >>> 32)) of Double.doubleToLongBits(v) -- one element of Arrays.hashCode(double[])."""
    v = np.asarray(v, np.float64).copy()
    v[np.isnan(v)] = np.nan                                                 # doubleToLongBits canonicalises NaN
    bits = v.view(np.uint64)
    return (bits ^ (bits >> np.uint64(32))).astype(_U32)


def _segmented_arrays_hash(elem_hash, ptr):
    """java.util.Arrays.hashCode per CSR row: result = 1; for e: result = 31 * result + hash(e)."""
    ptr = np.asarray(ptr, np.int64)
    n = len(ptr) - 1
    res = np.ones(n, _U32)
    lens = ptr[1:] - ptr[:-1]
    for k in range(int(lens.max()) if n else 0):
        act = np.flatnonzero(lens > k)
        res[act] = (res[act] * _U32(31) + elem_hash[ptr[act] + k]).astype(_U32)
    return res


def sparse_vector_hash(ptr, idx, val, dim):
    hi = _segmented_arrays_hash(np.asarray(idx).astype(np.int64).astype(np.uint64).astype(_U32), ptr)
    hd = _segmented_arrays_hash(_double_bits_hash(val), ptr)
    h = np.full(len(hi), 3 * 41 + int(dim), np.uint64).astype(_U32)
    h = (h * _U32(41) + hi).astype(_U32)
    return (h * _U32(41) + hd).astype(_U32)


def dense_vector_hash(X):
    X = np.asarray(X, np.float64)
    n, d = X.shape
    eh = _double_bits_hash(X.reshape(-1))
    return _segmented_arrays_hash(eh, np.arange(n + 1, dtype=np.int64) * d)


def labeled_vector_hash(label, vector_hash):
    acc = np.full(len(vector_hash), 0xCAFEBABE, _U32)
    acc = mix(acc, double_hash(label))
    acc = mix(acc, np.asarray(vector_hash, _U32))
    return finalize_hash(acc, 2).view(np.int32)                            # Java int


def block_ids_by_hash(y, blocks, X=None, csr=None, dim=None) -> np.ndarray:
    """FlinkMLTools.block: LabeledVector.hashCode() % blocks, made non-negative.  csr=(ptr, idx, val) rows are hashed as
    SparseVector(dim, ...) (what readLibSVM builds), dense X rows as DenseVector."""
    if csr is not None:
        ptr, idx, val = csr
        vh = sparse_vector_hash(ptr, idx, val, int(dim))
    else:
        vh = dense_vector_hash(X)
    h = labeled_vector_hash(np.asarray(y, np.float64), vh).astype(np.int64)
    b = np.fmod(h, blocks)                                                   # Java %: sign of the dividend
    b[b < 0] += blocks
    return b.astype(np.int32)


def block_ids_by_subtask(n: int, blocks: int, parallelism: int = 1) -> np.ndarray:
    """The round-1 reading: blockID = sourceSubtaskIndex % blocks, a file source handing each of the
    `parallelism` subtasks one contiguous split.  Kept as a named alternative of the risk register."""
    split = (np.arange(n, dtype=np.int64) * parallelism) // max(n, 1)
    return (split % blocks).astype(np.int32)
