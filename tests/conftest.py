import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _have_gpu():
    try:
        from mola_fe_lidar_b200 import capi
        return capi.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a GPU must FAIL loudly (no silent skip): only
    # deselect gpu tests when the user did not ask for them.
    pass


@pytest.fixture(scope="session")
def oracle():
    import oracle_api
    oracle_api.build()
    return oracle_api


@pytest.fixture(scope="session")
def capi():
    from mola_fe_lidar_b200 import capi as m
    return m


@pytest.fixture(scope="session")
def icp(capi):
    """One ICP object with the shipped regular settings on cuda:0."""
    obj = capi.ICP(capi.default_params(), device=0)
    yield obj
    obj.close()


@pytest.fixture(scope="session")
def rng():
    return np.random.default_rng(12345)
