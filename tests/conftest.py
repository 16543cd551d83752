import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _ab_library():
    """Test-infrastructure hook for A/B builds (python -m tcdiff_b200.build --define ... --out libX.so): run the GPU suite
    against another build of the library with TCDIFF_TEST_LIB=path.  The product never reads this variable."""
    path = os.environ.get("TCDIFF_TEST_LIB")
    if path:
        from tcdiff_b200 import _lib
        _lib.LIB_PATH = os.path.abspath(path)


def pytest_configure(config):
    _ab_library()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    # np.sqrt(torch tensor) is kept on purpose (bit-identical schedule buffers, model/diffusion.py:155,159)
    config.addinivalue_line("filterwarnings", "ignore:__array_wrap__:DeprecationWarning")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


@pytest.fixture(scope="session")
def dev():
    return torch.device("cuda:0")
