"""CPU tests of the host-side mirror of the FlinkML operator surface (no GPU work)."""
import numpy as np
import pytest


def test_als_builder_is_fluent_and_defaults_match_flinkml(fms):
    als = fms.ALS.__new__(fms.ALS)
    fms.ALS.__init__(als, ctx=object())
    assert als.setIterations(7).setNumFactors(5).setBlocks(3).setTemporaryPath("/tmp/x") is als
    assert als._param(fms.ALS.Iterations, None) == 7 and als._param(fms.ALS.NumFactors, None) == 5
    assert als._param(fms.ALS.Lambda, None) == 1.0 and als._param(fms.ALS.Seed, None) == 0     # FlinkML defaults
    pm = fms.ParameterMap().add(fms.ALS.Lambda, 0.9).add(fms.ALS.Seed, 42)                     # ALSImpl.scala:47-49
    assert als._param(fms.ALS.Lambda, pm) == 0.9 and als._param(fms.ALS.Seed, pm) == 42
    assert als.factorsOption is None
    with pytest.raises(RuntimeError, match="has not been fitted"):
        als.predict(([1], [2]))


def test_svm_builder_and_block_assignment(fms):
    svm = fms.SVM(ctx=object())
    assert svm.setBlocks(10).setIterations(10) is svm                                            # SVMImpl.scala:24-26
    assert svm._p["local_iterations"] == 10 and svm._p["regularization"] == 1.0 and svm._p["stepsize"] == 1.0
    assert svm._p["threshold"] == 0.0 and svm._p["seed"] is None
    with pytest.raises(RuntimeError, match="has not been trained"):
        svm.predict(np.zeros((1, 2)))
    # parallelism 1: every row in block 0 whatever `blocks` says (FlinkMLTools.block keys by source subtask)
    assert set(fms.SVM.default_block_ids(100, 10, 1).tolist()) == {0}
    b = fms.SVM.default_block_ids(100, 4, 8)
    assert b.tolist() == sorted(b.tolist(), key=lambda v: 0) and set(b.tolist()) == {0, 1, 2, 3}
    chunks = np.split(b, 8)[:8] if len(b) % 8 == 0 else None
    splits = (np.arange(100) * 8) // 100
    assert ((splits % 4) == b).all()


def test_dense_vector_surface(fms):
    v = fms.DenseVector([1.0, 2.0, 3.0])
    assert v.size() == 3 and v.apply(1) == 2.0 and v.dot(fms.DenseVector([1, 1, 1])) == 6.0
    v.update(0, 5.0)
    assert v.data[0] == 5.0


# ---------------------------------------------------------------------------------------------- FlinkMLTools.block (hash reading)
def _murmur3_32(data: bytes, seed: int) -> int:
    """Textbook MurmurHash3_x86_32 for whole 4-byte blocks (independent scalar implementation)."""
    def rotl(x, r):
        return ((x << r) | (x >> (32 - r))) & 0xFFFFFFFF
    h = seed & 0xFFFFFFFF
    for i in range(0, len(data), 4):
        k = int.from_bytes(data[i:i + 4], "little")
        k = (k * 0xCC9E2D51) & 0xFFFFFFFF
        k = rotl(k, 15)
        k = (k * 0x1B873593) & 0xFFFFFFFF
        h ^= k
        h = rotl(h, 13)
        h = (h * 5 + 0xE6546B64) & 0xFFFFFFFF
    h ^= len(data)
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & 0xFFFFFFFF
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & 0xFFFFFFFF
    h ^= h >> 16
    return h


def test_murmur_mix_matches_published_murmur3_vectors():
    import importlib
    H = importlib.import_module("flink-ms_b200.flinkml_hash")
    # SMHasher verification values of MurmurHash3_x86_32 (public): (bytes, seed) -> hash
    kats = [(b"\xff\xff\xff\xff", 0, 0x76293B50), (b"\x21\x43\x65\x87", 0, 0xF55B516B), (b"\x21\x43\x65\x87", 0x5082EDEE, 0x2362F9DE)]
    for data, seed, want in kats:
        assert _murmur3_32(data, seed) == want
        k = np.array([int.from_bytes(data, "little")], np.uint32)
        h = H.mix(np.array([seed], np.uint32), k)
        assert int(H.finalize_hash(h, 4)[0]) == want          # one block: finalizeHash(mix(seed, k), length 4)


def _java_labeled_sparse_hash(label, dim, idx, val):
    """LabeledVector(label, SparseVector(dim, idx, val)).hashCode(), scalar restatement with Python ints."""
    import struct
    def i32(x):
        x &= 0xFFFFFFFF
        return x
    def arrays_hash_int(a):
        r = 1
        for e in a:
            r = i32(31 * r + (e & 0xFFFFFFFF))
        return r
    def arrays_hash_double(a):
        r = 1
        for e in a:
            bits = struct.unpack("<Q", struct.pack("<d", e))[0]
            r = i32(31 * r + ((bits ^ (bits >> 32)) & 0xFFFFFFFF))
        return r
    vh = 3
    for part in (dim, arrays_hash_int(idx), arrays_hash_double(val)):
        vh = i32(vh * 41 + part)
    lab = int(label) & 0xFFFFFFFF                        # doubleHash of an integral label
    h = 0xCAFEBABE
    for k in (lab, vh):
        # one MurmurHash3 block step == what _murmur3_32 does per 4 bytes, without the finaliser
        kk = (k * 0xCC9E2D51) & 0xFFFFFFFF
        kk = ((kk << 15) | (kk >> 17)) & 0xFFFFFFFF
        kk = (kk * 0x1B873593) & 0xFFFFFFFF
        h ^= kk
        h = ((h << 13) | (h >> 19)) & 0xFFFFFFFF
        h = (h * 5 + 0xE6546B64) & 0xFFFFFFFF
    h ^= 2
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & 0xFFFFFFFF
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & 0xFFFFFFFF
    h ^= h >> 16
    return h - (1 << 32) if h & 0x80000000 else h


def test_block_ids_by_hash_match_the_scalar_restatement():
    import importlib
    H = importlib.import_module("flink-ms_b200.flinkml_hash")
    rng = np.random.default_rng(5)
    n, dim, blocks = 200, 50, 7
    ptr = [0]; idx = []; val = []
    for _ in range(n):
        k = int(rng.integers(0, 9))
        cols = np.sort(rng.choice(dim, size=k, replace=False))
        idx += cols.tolist(); val += rng.standard_normal(k).tolist(); ptr.append(ptr[-1] + k)
    y = rng.choice([-1.0, 1.0], size=n)
    ptr = np.asarray(ptr, np.int64); idx = np.asarray(idx, np.int32); val = np.asarray(val, np.float64)
    got = H.block_ids_by_hash(y, blocks, csr=(ptr, idx, val), dim=dim)
    for r in range(n):
        h = _java_labeled_sparse_hash(y[r], dim, idx[ptr[r]:ptr[r + 1]].tolist(), val[ptr[r]:ptr[r + 1]].tolist())
        b = abs(h) % blocks * (1 if h >= 0 else -1)          # Java %: sign of the dividend
        if b < 0:
            b += blocks
        assert got[r] == b
    assert got.min() >= 0 and got.max() < blocks and len(set(got.tolist())) == blocks     # rows spread over all blocks
    # doubleHash branches: integral, exactly-a-float, integral beyond int, generic
    dh = H.double_hash(np.array([1.0, -1.0, 0.5, 2.0 ** 40, 0.1, -0.0]))
    import struct
    assert int(dh[0]) == 1 and int(dh[1]) == 0xFFFFFFFF and int(dh[5]) == 0
    assert int(dh[2]) == struct.unpack("<I", struct.pack("<f", 0.5))[0]
    assert int(dh[3]) == struct.unpack("<I", struct.pack("<f", 2.0 ** 40))[0]          # float check comes before the long check
    bits = struct.unpack("<Q", struct.pack("<d", 0.1))[0]
    assert int(dh[4]) == ((bits ^ (bits >> 32)) & 0xFFFFFFFF)
