"""pytest configuration: `-m gpu` tests need a real B200 (they call the sm_100a library through the
C ABI); everything else runs on CPU.  The CPU suite may execute the kernel *sources* through the
g++-built emulation under tests/emu/ -- test infrastructure only, never loaded by the product."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def emu_lib():
    from emu import build_emu
    return build_emu.build()


@pytest.fixture(scope="session")
def emu_ctx(emu_lib):
    from funscript_flow_b200 import _native
    ctx = _native.FlowContext(0, emu_lib)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def gpu_ctx():
    from funscript_flow_b200 import _native, build
    build.build()
    if _native.device_count() < 1:
        pytest.fail("gpu-marked test started without a CUDA device")
    ctx = _native.FlowContext(0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
