import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        from vierkant_b200 import capi
        return capi.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def port_oracle():
    """This repo's C restatement of the reference path (oracle/bc7_oracle.c ...), built on demand with gcc."""
    from oracle import pyoracle
    pyoracle.build("port")
    return pyoracle.PortOracle()


@pytest.fixture(scope="session")
def ref_oracle():
    """The unmodified reference compiled in place (oracle/_ref/libvkt_ref.so); skipped where it was never built."""
    from oracle import pyoracle
    if os.path.isdir("/root/reference"):
        pyoracle.build("ref")
    if not pyoracle.RefOracle.available():
        pytest.skip("oracle/_ref/libvkt_ref.so not present")
    return pyoracle.RefOracle()


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library; built here if nvcc is available and the .so is stale or missing."""
    from vierkant_b200 import build, capi
    if build.find_nvcc() is not None:
        build.build_cuda()
    return capi.load_library()


@pytest.fixture(scope="session")
def ctx(cuda_lib):
    from vierkant_b200 import capi
    c = capi.BcnContext([0])
    yield c
    c.close()
