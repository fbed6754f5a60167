import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref/libzultra_ref.so) - the checker, never the product."""
    import refharness
    if not os.path.exists(refharness.REF_SO):
        pytest.skip("oracle/_ref not built (needs /root/reference; run `make -C oracle ref`)")
    return refharness.Ref()


@pytest.fixture(scope="session")
def emu():
    import refharness
    if not os.path.exists(refharness.EMU_SO):
        import subprocess
        subprocess.check_call(["make", "-C", ROOT, "emu"])
    return refharness.Emu()


@pytest.fixture(scope="session")
def ctx():
    import zultra_b200 as z
    c = z.CudaCtx()
    yield c
    c.close()
