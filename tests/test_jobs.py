"""The job drivers (flink-ms_b200/jobs.py) against the reference mains they mirror (ALSImpl.scala:15-75, SVMImpl.scala:13-56):
flag parsing on the CPU; the jobs themselves (GPU ingest -> fit -> Java-formatted rows) on the GPU."""
import importlib
import io

import numpy as np
import pytest

import oracle

jobs = importlib.import_module("flink-ms_b200.jobs")
formats = importlib.import_module("flink-ms_b200.formats")


def test_parameter_tool_from_args():
    p = jobs.ParameterTool.fromArgs(["--input", "a.csv", "-iterations", "7", "--lambda", "-0.5", "--flag", "--ignoreFirstLine", "FALSE"])
    assert p.get("input") == "a.csv" and p.getInt("iterations") == 7 and p.getDouble("lambda") == -0.5
    assert p.has("flag") and p.get("flag") == jobs.ParameterTool.NO_VALUE_KEY
    assert p.getBoolean("ignoreFirstLine", True) is False and p.getBoolean("missing", True) is True
    with pytest.raises(RuntimeError):
        p.getBoolean("missing")
    assert p.getInt("missing", 3) == 3
    with pytest.raises(RuntimeError):
        p.getRequired("training")
    with pytest.raises(ValueError):
        jobs.ParameterTool.fromArgs(["input", "x"])


def test_als_job_without_input_prints_the_reference_hint():
    out = io.StringIO()
    assert jobs.als_impl(["--iterations", "3"], ctx=object(), out=out) is None
    assert out.getvalue().strip() == "Use --input to specify file input."
    with pytest.raises(ValueError):
        jobs.als_impl(["--input", "x", "--fieldDelimiter", "semicolon"], ctx=object())


@pytest.mark.gpu
def test_als_job_end_to_end_matches_oracle_and_reference_text(fms, ctx, tmp_path):
    synth = importlib.import_module("flink-ms_b200.synth")
    u, i, r = synth.als_ratings(150, 40, 2500, seed=41)
    src = tmp_path / "ratings.tsv"
    src.write_text("user\titem\trating\n" + "".join(f"{a}\t{b}\t{float(c)!r}\n" for a, b, c in zip(u, i, r)))
    uf, itf = tmp_path / "U.txt", tmp_path / "I.txt"
    jobs.als_impl(["--input", str(src), "--fieldDelimiter", "tab", "--iterations", "3", "--numFactors", "8", "--lambda", "0.7", "--seed", "5",
                   "--blocks", "4", "--userFactors", str(uf), "--itemFactors", str(itf)], ctx=ctx)
    uid, U, iid, I, _ = formats.parse_als_rows(uf.read_text().splitlines() + itf.read_text().splitlines())
    orc = oracle.OracleALS(u, i, r)
    orc.fit(8, 3, 0.7, 5)
    ids_u, Xo = orc.factors(0)
    ids_i, Yo = orc.factors(1)
    assert np.array_equal(uid, ids_u) and np.array_equal(iid, ids_i)
    assert np.abs(U - Xo).max() <= 1e-4 * np.abs(Xo).max() and np.abs(I - Yo).max() <= 1e-4 * np.abs(Yo).max()
    first = uf.read_text().splitlines()[0].split(",")
    assert first[1] == "U" and len(first[2].split(";")) == 8                      # OutputFactor.toString
    out = io.StringIO()
    jobs.als_impl(["--input", str(src), "--fieldDelimiter", "tab", "--iterations", "1", "--numFactors", "2"], ctx=ctx, out=out)
    lines = out.getvalue().splitlines()
    assert lines[0].startswith("Printing results to stdout.") and lines[1] == "==== USER FACTORS ====" and "==== ITEM FACTORS ====" in lines


@pytest.mark.gpu
def test_svm_job_end_to_end(fms, ctx, tmp_path):
    synth = importlib.import_module("flink-ms_b200.synth")
    X, y, _ = synth.svm_data(250, 15, seed=42)
    X = np.asarray(X, np.float64)
    src = tmp_path / "train.libsvm"
    src.write_text("".join(("%+d" % int(lab)) + "".join(f" {j + 1}:{float(v)!r}" for j, v in enumerate(row) if v != 0.0) + "\n" for lab, row in zip(y, X)))
    dst = tmp_path / "model.txt"
    jobs.svm_impl(["--training", str(src), "--blocks", "3", "--iteration", "4", "--seed", "9", "--partition", "false", "--output", str(dst)], ctx=ctx)
    w_job = formats.parse_svm_rows(dst.read_text().splitlines(), dim=X.shape[1])
    # FlinkMLTools.block: LabeledVector.hashCode() % blocks over the SparseVectors readLibSVM builds (dim = largest index seen);
    # parallelism 1: the three blocks share the one java.util.Random(seed ^ 0) of the single localDualMethod subtask
    H = importlib.import_module("flink-ms_b200.flinkml_hash")
    ptr = np.arange(len(y) + 1, dtype=np.int64) * X.shape[1]                        # no exact zeros in N(0,1) samples: all 15 stored
    bid = H.block_ids_by_hash(y, 3, csr=(ptr, np.tile(np.arange(X.shape[1], dtype=np.int32), len(y)), X.reshape(-1)), dim=X.shape[1])
    assert len(set(bid.tolist())) == 3
    w_orc = oracle.svm_fit(X, y, bid, 3, 4, 10, 1.0, 1.0, 9, 1)
    assert np.abs(w_job - w_orc).max() <= 1e-4 * max(1.0, np.abs(w_orc).max())
    assert dst.read_text().splitlines()[0].startswith("1,")                          # 1-based feature ids
    out = io.StringIO()
    jobs.svm_impl(["--training", str(src), "--blocks", "3", "--iteration", "2", "--seed", "9", "--partition", "true", "--range", "4"], ctx=ctx, out=out)
    lines = out.getvalue().splitlines()
    assert lines[0].startswith("Printing result to stdout.") and lines[1].startswith("0,1:") and lines[2].startswith("1,4:")
    with pytest.raises(RuntimeError):
        jobs.svm_impl(["--blocks", "3"], ctx=ctx)
    with pytest.raises(RuntimeError) as ei:                                         # --partition is required (getBoolean without default)
        jobs.svm_impl(["--training", str(src), "--seed", "9"], ctx=ctx)
    assert "partition" in str(ei.value)


def test_predict_consoles_follow_the_reference_clients():
    """ALSPredict.java:58-92 / SVMPredict.java:58-90 over the host-side QueryableStateClient (no GPU involved)."""
    client = formats.QueryableStateClient(None)
    U = np.array([[0.5, -1.25, 2.0], [1.0, 0.0, 3.0]]); I = np.array([[2.0, 4.0, -0.5]])
    client.load_als([7, 9], U, [3], I)
    out = io.StringIO()
    jobs.als_predict_console(["7,3\n", " 9,3 ", "8,3", "7", "x,y"], client, out)
    text = out.getvalue()
    assert text.startswith("Enter <User,Item> to predict.\n[info] Querying the model for <user,item> pair '7,3'\n")
    assert "ALS Prediction =  %f \n" % (0.5 * 2.0 + -1.25 * 4.0 + 2.0 * -0.5) in text
    assert "ALS Prediction =  %f \n" % (1.0 * 2.0 + 3.0 * -0.5) in text
    assert text.count("User or Item Factors do not exist in the model for the query: ") == 2       # unknown user 8, non-numeric ids
    assert text.count("Query failed because of the following Exception:") == 1                   # "7": tokens[1] missing
    w = np.array([0.25, -1.0, 0.0, 2.0])
    client.load_svm(w)
    out = io.StringIO()
    jobs.svm_predict_console(["1:2.0 4:0.5", "2:1.0", "1:1.0 9:3.0", "1:0.0", "bad"], client, False, 0.0, out)
    lines = out.getvalue().splitlines()
    assert lines[0] == "Enter Vector data to predict."
    preds = [l for l in lines if l.startswith("SVM Prediction")]
    assert preds == ["SVM Prediction =  1.000000 ", "SVM Prediction =  -1.000000 ", "SVM Prediction =  1.000000 ", "SVM Prediction =  -1.000000 "]
    assert "Could not find the value for feature ID: 9 " in lines and out.getvalue().count("Query failed") == 1
    out = io.StringIO()
    jobs.svm_predict_console(["1:2.0 4:0.5"], client, True, 0.0, out)
    assert "SVM Prediction =  %f " % (0.25 * 2.0 + 2.0 * 0.5) in out.getvalue()


def test_als_mean_vector_job(tmp_path):
    """ALSMeanVector.scala:12-48: left-to-right fp64 sum / count, formatted like Scala's mkString of doubles."""
    rows = ["3,U,0.1;0.2;-1.5", "9,U,0.3;0.4;2.5", "11,U,1.0E-4;5.0;0.25"]
    src = tmp_path / "U.txt"
    src.write_text("\n".join(rows) + "\n")
    out = io.StringIO()
    row = jobs.als_mean_vector(["--type", "user", "--input", str(src)], out=out)
    m = [((0.1 + 0.3) + 1.0e-4) / 3, ((0.2 + 0.4) + 5.0) / 3, ((-1.5 + 2.5) + 0.25) / 3]
    assert row == "MEAN,U," + ";".join(formats.java_double_to_string(v) for v in m)
    assert out.getvalue().splitlines() == ["Printing results to stdout. Use --output to specify output location", row]
    dst = tmp_path / "mean.txt"
    jobs.als_mean_vector(["--type", "item", "--input", str(src), "--output", str(dst)])
    assert dst.read_text() == row.replace("MEAN,U,", "MEAN,I,") + "\n"
    with pytest.raises(ValueError):
        jobs.als_mean_vector(["--type", "both", "--input", str(src)])
    with pytest.raises(RuntimeError):
        jobs.als_mean_vector(["--input", str(src)])
