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
    jobs.svm_impl(["--training", str(src), "--blocks", "3", "--iteration", "4", "--seed", "9", "--output", str(dst)], ctx=ctx)
    w_job = formats.parse_svm_rows(dst.read_text().splitlines(), dim=X.shape[1])
    w_orc = oracle.svm_fit(X, y, np.zeros(len(y), np.int32), 3, 4, 10, 1.0, 1.0, 9)    # parallelism 1: every row in block 0
    assert np.abs(w_job - w_orc).max() <= 1e-4 * max(1.0, np.abs(w_orc).max())
    assert dst.read_text().splitlines()[0].startswith("1,")                          # 1-based feature ids
    out = io.StringIO()
    jobs.svm_impl(["--training", str(src), "--blocks", "3", "--iteration", "2", "--seed", "9", "--partition", "true", "--range", "4"], ctx=ctx, out=out)
    lines = out.getvalue().splitlines()
    assert lines[0].startswith("Printing result to stdout.") and lines[1].startswith("0,1:") and lines[2].startswith("1,4:")
    with pytest.raises(RuntimeError):
        jobs.svm_impl(["--blocks", "3"], ctx=ctx)
