import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def backend():
    """Binds learningbycheating_b200._lib to the CUDA library on a GPU box, else (CPU container) to the
    host-emulation build of the same sources -- test infrastructure for the host-side logic only."""
    import torch
    from learningbycheating_b200 import _lib, build
    if torch.cuda.is_available():
        build.build_cuda()
        _lib.lib()
        assert _lib.lib().lbc_device_kind() == 1
        return "cuda"
    path = build.build_hostemu()
    _lib.use_library_for_tests(path)
    return "cpu"


@pytest.fixture(scope="session")
def device(backend):
    return backend


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
