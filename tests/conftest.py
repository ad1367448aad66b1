import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the oracle's C restatement is test infrastructure: build it on demand
    import oracle
    if not os.path.exists(oracle.PORT_SO) or (
            os.path.exists(os.path.join(oracle.REFERENCE_ROOT, "src", "dct.c")) and not oracle.have_reference()):
        oracle.build()
    # the product library: build on demand when nvcc is around; never fall back
    lib = os.path.join(ROOT, "jpeg_gpu_b200", "libjpeg_gpu_b200.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "jpeg_gpu_b200", "csrc")], check=True)


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.port()


@pytest.fixture(scope="session")
def checker():
    """The compiled reference when oracle/_ref exists, else our C port."""
    import oracle
    return oracle.best()


@pytest.fixture(scope="session")
def reference():
    import oracle
    if not oracle.have_reference():
        pytest.skip("oracle/_ref not built (reference tree not present)")
    return oracle.reference()


@pytest.fixture(scope="session")
def gpu_ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import jpeg_gpu_b200 as J
    ctx = J.Context(0)
    yield ctx
    ctx.close()
