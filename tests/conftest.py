import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _seed():
    # same session seed as the reference's python tests (libgdf/python/tests/conftest.py:14-20)
    np.random.seed(0xabcdef % (2 ** 32))
    yield


@pytest.fixture(scope="session")
def gdf():
    """(ffi, libgdf) of the product library; RMM in pool mode like the reference's conftest (:7-9)."""
    from libgdf_b200.librmm_cffi import librmm, librmm_config
    if _cuda_available():
        librmm_config.use_pool_allocator = True
        librmm.finalize()
        librmm.initialize()
    from libgdf_b200.libgdf_cffi import ffi, libgdf
    return ffi, libgdf
