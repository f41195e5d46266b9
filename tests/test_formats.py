"""CPU tests of the model text / lookup schemas (flink-ms_b200/formats.py) against the reference's formats:
ALSImpl.scala:84, SVMImpl.scala:35/:42-45/:63-71, ALSMeanVector.scala:28-35, ALSKafkaConsumer.java:78-92,
QueryClientHelper.java:105-139."""
import importlib
import os

import numpy as np
import pytest

F = importlib.import_module("flink-ms_b200.formats")


@pytest.mark.parametrize("x,s", [
    (0.0, "0.0"), (-0.0, "-0.0"), (1.0, "1.0"), (-2.5, "-2.5"), (100.0, "100.0"), (0.001, "0.001"), (0.0001, "1.0E-4"),
    (1234567.0, "1234567.0"), (12345678.0, "1.2345678E7"), (1e7, "1.0E7"), (9999999.999, "9999999.999"), (1e-3, "0.001"),
    (9.999e-4, "9.999E-4"), (0.1, "0.1"), (1.0 / 3.0, "0.3333333333333333"), (3.5, "3.5"), (1e21, "1.0E21"), (1e-10, "1.0E-10"),
    (123456789012.0, "1.23456789012E11"), (0.7275636800328681, "0.7275636800328681"), (float("inf"), "Infinity"),
    (float("-inf"), "-Infinity"), (1.7976931348623157e308, "1.7976931348623157E308"),
])
def test_java_double_to_string_known_values(x, s):
    # known outputs of java.lang.Double.toString (JDK specification examples and widely published values)
    assert F.java_double_to_string(x) == s
    if s not in ("Infinity", "-Infinity"):
        assert float(s.replace("E", "e")) == x                  # Double.parseDouble round trip


def test_java_double_to_string_round_trips_random():
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.standard_normal(2000), rng.standard_normal(500) * 1e9, rng.standard_normal(500) * 1e-7])
    for v in vals:
        s = F.java_double_to_string(float(v))
        assert float(s.replace("E", "e")) == float(v)
        assert ("E" in s) == (not (1e-3 <= abs(v) < 1e7))


def test_als_rows_match_outputfactor_tostring_and_parse_back():
    ids = np.array([7, 42, 100000], np.int64)
    fac = np.array([[0.5, -1.25, 3.0], [1e-5, 2.0, 1 / 3], [1e8, 0.0, -0.1]])
    rows = F.als_model_rows(ids, fac, "U")
    assert rows[0] == "7,U,0.5;-1.25;3.0"                       # s"$id,$typeOfFeature,${factors.mkString(";")}"
    assert rows[1] == "42,U,1.0E-5;2.0;0.3333333333333333"
    assert rows[2] == "100000,U,1.0E8;0.0;-0.1"
    irows = F.als_model_rows([3, 1], [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], "I")
    mean = F.als_mean_row(fac, "U")
    assert mean.startswith("MEAN,U,")
    uid, U, iid, I, means = F.parse_als_rows(rows + irows + [mean])
    assert uid.tolist() == [7, 42, 100000] and np.array_equal(U, fac)
    assert iid.tolist() == [1, 3] and np.array_equal(I, [[4.0, 5.0, 6.0], [1.0, 2.0, 3.0]])
    assert np.allclose(means["U"], fac.mean(0), rtol=1e-15)


def test_svm_rows_one_based_and_range_buckets():
    w = np.arange(1, 8, dtype=np.float64) / 4
    plain = F.svm_model_rows(w)
    assert plain[0] == "1,0.25" and plain[-1] == "7,1.75"        # SVMImpl.scala:33-35: (index + 1, value)
    bucket = F.svm_model_rows(w, partition=True, range_=3)       # bucket = id / range (integer division)
    assert bucket == ["0,1:0.25;2:0.5", "1,3:0.75;4:1.0;5:1.25", "2,6:1.5;7:1.75"]
    assert np.array_equal(F.parse_svm_rows(plain), w) and np.array_equal(F.parse_svm_rows(bucket), w)
    assert np.array_equal(F.parse_svm_rows(plain[:3], dim=7), [0.25, 0.5, 0.75, 0, 0, 0, 0])


def test_queryable_state_client_schema():
    c = F.QueryableStateClient()
    uid, iid = np.array([5, 9], np.int64), np.array([2], np.int64)
    U, I = np.array([[1.0, 2.0], [0.5, 1e-4]]), np.array([[3.0, 4.0]])
    c.load_als(uid, U, iid, I, {"U": np.array([0.75, 1.00005])})
    assert c.queryState("ALS_MODEL", "9-U") == ("9-U", "0.5;1.0E-4")           # ALSKafkaConsumer.java:78-80
    assert c.queryState("ALS_MODEL", "2-I") == ("2-I", "3.0;4.0")
    assert c.queryState("ALS_MODEL", "MEAN-U") == ("MEAN-U", "0.75;1.00005")    # SGD.java:143-147
    assert c.queryState("ALS_MODEL", "6-U") is None and c.queryState("ALS_MODEL", "2-U") is None   # Optional.empty
    assert c.queryState("OTHER", "x") is None
    with pytest.raises(TypeError):
        c.queryState("ALS_MODEL", None)
    # the client-side prediction of ALSPredictRandom.java:79-92 on the served strings
    u = [float(t) for t in c.queryState("ALS_MODEL", "5-U")[1].split(";")]
    i = [float(t) for t in c.queryState("ALS_MODEL", "2-I")[1].split(";")]
    assert sum(a * b for a, b in zip(u, i)) == 11.0
    w = np.array([0.5, -1.0, 2.0, 4.0])
    c.load_svm(w)
    assert c.queryState("SVM_MODEL", "3") == ("3", "2.0") and c.queryState("SVM_MODEL", "5") is None
    c.load_svm(w, partition=True, range_=2)
    assert c.queryState("SVM_MODEL", "1") == ("1", "2:-1.0;3:2.0") and c.queryState("SVM_MODEL", "0") == ("0", "1:0.5")
    c.close()
    with pytest.raises(RuntimeError):
        c.queryState("SVM_MODEL", "1")


def _pow10_table():
    import re
    path = os.path.join(os.path.dirname(__file__), "..", "flink-ms_b200", "csrc", "pow10_table.cuh")
    rows = re.findall(r"\{0x([0-9a-f]{16})ULL, 0x([0-9a-f]{16})ULL\}", open(path).read())
    assert len(rows) == 651
    return [(int(h, 16), int(l, 16)) for h, l in rows]


def _eisel_lemire(w, q, table):
    """Python transcription of csrc/ingest.cu::eisel_lemire (same table, same steps) -- pins the generated table and the
    algorithm on the CPU; the device code itself is checked bit-exactly in tests/test_ingest_gpu.py."""
    import struct
    M = (1 << 64) - 1
    if w == 0:
        return 0.0
    if q < -342 or q > 308:
        return None
    lz = 64 - w.bit_length()
    w <<= lz
    exp2 = (((217706 * q) >> 16) + 64 + 1023) - lz
    t_hi, t_lo = table[q + 342]
    x = w * t_hi
    x_hi, x_lo = x >> 64, x & M
    if (x_hi & 0x1FF) == 0x1FF and ((x_lo + w) & M) < w:
        y = w * t_lo
        y_hi, y_lo = y >> 64, y & M
        m_hi, m_lo = x_hi, (x_lo + y_hi) & M
        if m_lo < x_lo:
            m_hi += 1
        if (m_hi & 0x1FF) == 0x1FF and ((m_lo + 1) & M) == 0 and ((y_lo + w) & M) < w:
            return None
        x_hi, x_lo = m_hi, m_lo
    msb = x_hi >> 63
    mant = x_hi >> (msb + 9)
    exp2 -= 1 ^ msb
    maybe_tie = x_lo == 0 and (x_hi & 0x1FF) == 0 and (mant & 3) == 1
    if maybe_tie and (q < -4 or q > 23):
        return None
    tie = maybe_tie and (x_hi & ((1 << (msb + 9)) - 1)) == 0
    if not tie:
        mant += mant & 1
    mant >>= 1
    if mant >> 53:
        mant >>= 1
        exp2 += 1
    if not (1 <= exp2 <= 0x7FE):
        return None
    return struct.unpack("<d", struct.pack("<Q", (exp2 << 52) | (mant & 0x000FFFFFFFFFFFFF)))[0]


def test_eisel_lemire_table_and_algorithm_match_python_float():
    table = _pow10_table()
    rng = np.random.default_rng(11)
    undecided = 0
    for it in range(40000):
        digits = int(rng.integers(1, 20))
        w = int(rng.integers(1, 10 ** digits)) if digits < 19 else int(rng.integers(10 ** 18, 2 ** 63))
        q = int(rng.integers(-300, 290)) if it % 2 else int(rng.integers(-6, 26))   # half of them where ties are decidable
        got = _eisel_lemire(w, q, table)
        if got is None:
            undecided += 1
            continue
        assert got == float(f"{w}e{q}"), (w, q)
    assert undecided < 400                                                     # the fallback-free parser rejects only rare cases
    assert _eisel_lemire(9007199254740993, 0, table) == 9007199254740992.0     # exact tie -> even
    assert _eisel_lemire(9007199254740995, 0, table) == 9007199254740996.0
    assert _eisel_lemire(17976931348623157, 292, table) == 1.7976931348623157e308
