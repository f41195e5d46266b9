"""N4: GPU-side ingest (fms_parse_ratings_csv / fms_parse_libsvm) against Python's own parsers -- float() is correctly
rounded like Java's Double.parseDouble for decimal literals, int() like Integer.parseInt inside the 32-bit range.
Bit-exact: every parsed double must equal float(token)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _literals(rng, n):
    out = []
    for k in range(n):
        kind = k % 8
        if kind == 0:
            out.append(repr(float(rng.uniform(0, 5))))                       # 17 significant digits: Eisel-Lemire territory
        elif kind == 1:
            out.append("%.1f" % rng.uniform(0, 5))                            # typical rating
        elif kind == 2:
            out.append(str(int(rng.integers(0, 6))))                          # integer literal
        elif kind == 3:
            if k % 16 == 3:                                                       # integers times small powers of ten: exact products, ties decidable
                out.append("%de%d" % (int(rng.integers(1, 2 ** 62)), int(rng.integers(-6, 26))))
            else:
                out.append("%.6e" % float(rng.standard_normal() * 10.0 ** int(rng.integers(-200, 200))))
        elif kind == 4:
            out.append(("-" if rng.random() < 0.5 else "+") + "%.12f" % rng.uniform(0, 1e6))
        elif kind == 5:
            out.append(repr(float(np.float32(rng.standard_normal()))))
        elif kind == 6:
            out.append("%d.%017d" % (int(rng.integers(1, 100)), int(rng.integers(0, 10 ** 17))))    # 18-19 significant digits
        else:
            out.append(["0", "0.0", ".5", "5.", "1e3", "1E-3", "2.5e+10", "NaN", "Infinity", "-Infinity", "1.0d", "3f", "0.30000000000000004",
                        "9007199254740993", "293710372636606e3", "0.1000000000000000055511151231257827021181583404541015625", "1.7976931348623157e308", "2.2250738585072014e-308", "123456789012345678"][k // 8 % 19])
    return out


def test_ratings_csv_matches_python_parsers(fms, ctx):
    rng = np.random.default_rng(1)
    n = 4000
    lits = _literals(rng, n)
    us = rng.integers(-5, 2_000_000_000, n)
    its = rng.integers(0, 100000, n)
    for delim, eol, header in ((",", "\n", False), ("\t", "\r\n", True), (";", "\n", True)):
        lines = [f"{u}{delim}{i}{delim}{v}" for u, i, v in zip(us, its, lits)]
        lines.insert(7, "")                                                     # empty lines are skipped
        text = (("user%sitem%srating%s" % (delim, delim, eol)) if header else "") + eol.join(lines) + (eol if delim != ";" else "")
        u, i, r = ctx.parse_ratings_csv(text.encode(), delim, header)
        assert len(u) == n
        assert np.array_equal(u, us) and np.array_equal(i, its)
        ref = np.array([float(v.rstrip("dDfF")) for v in lits])
        assert np.array_equal(r.view(np.uint64), ref.view(np.uint64))           # bit-exact, NaN included
    u, i, r = ctx.parse_ratings_csv(b"", ",", False)
    assert len(u) == 0
    u, i, r = ctx.parse_ratings_csv(b"1,2,3.5,extra,columns\n", ",", False)
    assert u.tolist() == [1] and i.tolist() == [2] and r.tolist() == [3.5]
    for bad in (b"1,2\n", b"1,x,3.0\n", b"1,2,3.0.1\n", b"99999999999,2,3.0\n", b"1,2,\n", b"1,2,0x1p3\n"):
        with pytest.raises(fms.FlinkMSError) as ei:
            ctx.parse_ratings_csv(b"5,6,7\n" + bad, ",", False)
        assert "line 2" in str(ei.value)


def test_ratings_csv_feeds_fit_like_the_reference_job(fms, ctx):
    import importlib
    synth = importlib.import_module("flink-ms_b200.synth")
    u, i, r = synth.als_ratings(200, 50, 3000, seed=9)
    text = "user,item,rating\n" + "".join(f"{a},{b},{float(c)!r}\n" for a, b, c in zip(u, i, r))
    pu, pi, pr = ctx.parse_ratings_csv(text.encode(), ",", True)
    assert np.array_equal(pu, u) and np.array_equal(pi, i) and np.array_equal(pr, r)
    ctx.als_fit(pu, pi, pr, 8, 2, 0.9, 42)
    _, X = ctx.als_get_factors(0)
    c2 = fms.Context(0)
    c2.als_fit(u, i, r, 8, 2, 0.9, 42)
    assert np.array_equal(c2.als_get_factors(0)[1], X)
    c2.close()


def test_libsvm_matches_python_parsers(fms, ctx):
    rng = np.random.default_rng(2)
    rows, text = [], []
    for k in range(1500):
        nf = int(rng.integers(0, 12))
        idx = np.sort(rng.choice(np.arange(1, 400), nf, replace=False))
        vals = _literals(rng, nf)
        vals = [v for v in vals]
        lab = "+1" if rng.random() < 0.5 else "-1"
        rows.append((float(lab), idx - 1, [float(v.rstrip("dDfF")) for v in vals]))
        line = lab + "".join(f" {j}:{v}" for j, v in zip(idx, vals))
        if k % 97 == 0:
            line += "   # a trailing comment 1:2"
        text.append(line)
        if k % 211 == 0:
            text.append("# a comment line")
            text.append("   ")
    blob = ("\n".join(text) + "\n").encode()
    y, (ptr, idx, val), dim = ctx.parse_libsvm(blob)
    assert len(y) == len(rows) and ptr[0] == 0 and ptr[-1] == len(idx)
    assert dim == max((int(r[1].max()) + 1 for r in rows if len(r[1])), default=0)
    for k, (lab, ji, jv) in enumerate(rows):
        assert y[k] == lab
        assert np.array_equal(idx[ptr[k]:ptr[k + 1]], ji)
        assert np.array_equal(val[ptr[k]:ptr[k + 1]].view(np.uint64), np.array(jv, np.float64).view(np.uint64))
    y0, (p0, i0, v0), d0 = ctx.parse_libsvm(b"")
    assert len(y0) == 0 and p0.tolist() == [0] and d0 == 0
    for bad in (b"1 3\n", b"1 0:2.0\n", b"x 1:2\n", b"1 2:abc\n"):
        with pytest.raises(fms.FlinkMSError) as ei:
            ctx.parse_libsvm(b"-1 1:0.5\n" + bad)
        assert "line 2" in str(ei.value)


def test_libsvm_feeds_fit_like_the_reference_job(fms, ctx):
    """readLibSVM -> SVM.fit: the parsed CSR trains the same model as the dense array the file was written from."""
    import importlib
    synth = importlib.import_module("flink-ms_b200.synth")
    X, y, _ = synth.svm_data(300, 20, seed=4)
    X = np.asarray(X, np.float64)
    text = "".join(("%+d" % int(lab)) + "".join(f" {j + 1}:{float(v)!r}" for j, v in enumerate(row) if v != 0.0) + "\n" for lab, row in zip(y, X))
    py, (ptr, idx, val), dim = ctx.parse_libsvm(text.encode())
    assert np.array_equal(py, np.asarray(y, np.float64)) and dim == X.shape[1]
    bid = (np.arange(len(y)) % 4).astype(np.int32)
    w_csr = ctx.svm_fit(None, py, bid, 4, 3, 40, csr=(ptr, idx, val), dim=dim, seed=3)
    w_dense = ctx.svm_fit(X, y, bid, 4, 3, 40, seed=3)
    assert np.isfinite(w_csr).all()
    assert np.abs(w_csr - w_dense).max() <= 1e-5 * max(1.0, np.abs(w_dense).max())
