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
