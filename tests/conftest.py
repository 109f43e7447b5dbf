import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


REF_SO = os.path.join(ROOT, "oracle", "_ref", "libtacs_ref.so")


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref), when it has been built."""
    from tacs_b200 import binding

    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libtacs_ref.so not built (needs /root/reference)")
    from tests import ref_binding
    return ref_binding.load_reference(REF_SO)


@pytest.fixture(scope="session")
def lib():
    """The product library on a GPU; fails (not skips) when the extension cannot drive a device."""
    import tacs_b200

    L = tacs_b200.load()   # a missing extension is an error, not a skip: there is no fallback to test instead
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = True    # let init decide
    if not have_gpu:
        pytest.skip("no CUDA device visible: the -m gpu tests need a B200")
    assert L.init(0) == 0, "libtacs_b200.so could not initialise the CUDA device"
    return L
