import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import orc
    return orc.Oracle()


@pytest.fixture(scope="session")
def refshim():
    from oracle import orc
    if not orc.RefShim.available():
        pytest.skip("oracle/_ref/libcrixus_refshim.so not built (needs /root/reference; run oracle/build_ref.sh)")
    return orc.RefShim()


@pytest.fixture(scope="session")
def cxlib():
    from crixus_b200 import api, build
    if not os.path.exists(api.LIB_PATH):
        build.build()
    return api.lib()


@pytest.fixture(scope="session")
def cuda_backend(cxlib):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from backends import CudaBackend
    return CudaBackend()
