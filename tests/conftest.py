import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import orcbind
    return orcbind.load()


@pytest.fixture(scope="session")
def ref_scalar():
    import refbind
    if not refbind.available("scalar"):
        pytest.skip("oracle/_ref/libdfpsr_ref_scalar.so not built (needs /root/reference; see oracle/Makefile)")
    return refbind.Ref("scalar")


@pytest.fixture(scope="session")
def ref_sse():
    import refbind
    if not refbind.available("sse"):
        pytest.skip("oracle/_ref/libdfpsr_ref_sse.so not built (needs /root/reference; see oracle/Makefile)")
    return refbind.Ref("sse")


@pytest.fixture(scope="session")
def cuda():
    """The product library on cuda:0. GPU tests fail (not skip) when the extension is missing."""
    import torch
    from dfpsr_b200 import lib
    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    handle = lib.load()
    lib.check(handle.dfpsr_init(0))
    torch.cuda.set_device(0)
    return handle
