"""fms_create(device_ids, n > 1): one process drives several GPUs through the whole-fit entry points (csrc/group.cu).
Needs >= 2 GPUs (gpurun --gpus 2); on a 1-GPU box these tests skip.  The reference scales the same fits by Flink parallelism /
setBlocks (flink-als/src/main/scala/de/tub/it4bi/ALSImpl.scala:39-41, flink-svm/src/main/scala/de/tub/it4bi/SVMImpl.scala:25)."""
import importlib
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def group(fms):
    try:
        g = fms.Context(devices=[0, 1])
    except fms.FlinkMSError as e:
        if "out of range" in str(e) or "peer" in str(e):
            pytest.skip(f"needs two peer-connected GPUs: {e}")
        raise
    yield g
    g.close()


@pytest.mark.parametrize("rank,shape", [(64, (3000, 500, 90000)), (8, (700, 300, 9000)), (10, (257, 129, 4000))])
def test_group_als_fit_equals_single_gpu_bitwise(fms, ctx, group, rank, shape):
    synth = importlib.import_module("flink-ms_b200.synth")
    u, i, r = synth.als_ratings(*shape, seed=11)
    ctx.als_fit(u, i, r, rank, 3, 0.9, 42)
    group.als_fit(u, i, r, rank, 3, 0.9, 42)
    for which in (0, 1):
        ids1, f1 = ctx.als_get_factors(which)
        ids2, f2 = group.als_get_factors(which)
        np.testing.assert_array_equal(ids1, ids2)
        np.testing.assert_array_equal(f1, f2)                 # per-row arithmetic does not depend on the sharding
    q_u, q_i = u[:1000], i[:1000]
    np.testing.assert_array_equal(ctx.als_predict(q_u, q_i)[0], group.als_predict(q_u, q_i)[0])


def test_group_svm_fit_matches_single_gpu(fms, ctx, group):
    synth = importlib.import_module("flink-ms_b200.synth")
    X, y, _ = synth.svm_data(6000, 256, seed=3)
    K = 48
    bid = (np.arange(len(y)) % K).astype(np.int32)
    w1 = ctx.svm_fit(X, y, bid, K, 4, -1, 1.0, 1.0, 42)
    w2 = group.svm_fit(X, y, bid, K, 4, -1, 1.0, 1.0, 42)
    assert np.abs(w1 - w2).max() <= 1e-6 * np.abs(w1).max()       # only the order of the delta-w sum differs
    (ptr, idx, val), ys = synth.svm_sparse(900, 40, seed=4)
    bid = (np.arange(900) % 6).astype(np.int32)
    w1 = ctx.svm_fit(None, ys, bid, 6, 3, 10, 1.0, 1.0, 7, csr=(ptr, idx, val), dim=40)
    w2 = group.svm_fit(None, ys, bid, 6, 3, 10, 1.0, 1.0, 7, csr=(ptr, idx, val), dim=40)
    assert np.abs(w1 - w2).max() <= 1e-6 * max(np.abs(w1).max(), 1e-30)


def test_group_chunked_upload(fms, ctx, group):
    synth = importlib.import_module("flink-ms_b200.synth")
    u, i, r = synth.als_ratings(1500, 400, 40000, seed=12)
    ctx.als_fit(u, i, r, 64, 2, 0.9, 42)
    group.als_fit_stream(u, i, r, 64, 2, 0.9, 42, chunk=7001)
    for which in (0, 1):
        np.testing.assert_array_equal(ctx.als_get_factors(which)[1], group.als_get_factors(which)[1])
    X, y, _ = synth.svm_data(2000, 128, seed=9)
    bid = (np.arange(2000) % 10).astype(np.int32)
    w1 = group.svm_fit(X, y, bid, 10, 3, -1, 1.0, 1.0, 3)
    w2 = group.svm_fit_stream(X, y, bid, 10, 3, -1, 1.0, 1.0, 3, chunk_rows=333)
    np.testing.assert_array_equal(w1, w2)


def test_group_refuses_step_level_calls(fms, group):
    with pytest.raises(fms.FlinkMSError) as e:
        group.als_half_step(0, 0.9)
    assert e.value.status == fms._native.ERR_UNSUPPORTED
    with pytest.raises(fms.FlinkMSError):
        fms.Context(devices=[0, 0])                             # a device listed twice


def test_plain_c_host_drives_the_group(group):
    """tools/host_group_fit.c: no Python, no torch, no NCCL in the process that trains."""
    exe = os.path.join(ROOT, "tools", "host_group_fit")
    subprocess.check_call(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "host_group_fit.c"), "-o", exe,
                           "-L", os.path.join(ROOT, "flink-ms_b200"), "-lflinkms_b200", "-Wl,-rpath," + os.path.join(ROOT, "flink-ms_b200"), "-lm"])
    out = subprocess.run([exe, "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "users yes, items yes" in out.stdout
