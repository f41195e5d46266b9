"""CPU tests of the model text / lookup schemas (flink-ms_b200/formats.py) against the reference's formats:
ALSImpl.scala:84, SVMImpl.scala:35/:42-45/:63-71, ALSMeanVector.scala:28-35, ALSKafkaConsumer.java:78-92,
QueryClientHelper.java:105-139."""
import importlib

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
