import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def fms():
    import flinkms_b200
    return flinkms_b200


@pytest.fixture(scope="session")
def ctx(fms):
    """One shared fms_ctx on cuda:0.  Creation fails loudly without a GPU (no CPU fallback)."""
    c = fms.Context(0)
    yield c
    c.close()


def rel_err(a, b):
    import numpy as np
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.abs(b).max() if b.size else 1.0
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0)) if a.size else 0.0
