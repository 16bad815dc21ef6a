import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-scale-pointcloud-registration_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def armadillo_pair(oracle):
    """Config-1 clouds after the README pipeline at np.random.seed(0)."""
    np.random.seed(0)
    src = oracle.load_scan("ArmadilloBack_330")
    tgt = oracle.load_scan("ArmadilloBack_0")
    ps = oracle.Preprocessor([oracle.RandomDownsampler(5000), oracle.SOR()])
    pt = oracle.Preprocessor([oracle.RandomDownsampler(5000), oracle.SOR()])
    s = ps.preprocess(src)
    t = pt.preprocess(tgt)
    return s, t, np.random.get_state()
