import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def emul_lib():
    from tests.emul import EmulLib
    return EmulLib()


@pytest.fixture(scope="session")
def cuda_lib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from generalizableracing_b200 import build, _lib
    build.build()
    return _lib.load()


def backend_params():
    return [pytest.param("emul", id="emul"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture
def backend(request, ):
    """('cpu', EmulLib) for the CPU emulation of the kernel sources, ('cuda:0', None) for the real library."""
    kind = request.param
    if kind == "emul":
        return "cpu", request.getfixturevalue("emul_lib")
    request.getfixturevalue("cuda_lib")
    return "cuda:0", None
