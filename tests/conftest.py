import os
import sys
from pathlib import Path

import pytest

os.environ.setdefault("BCI_TUNING", "1")   # the kernel A/B switches some GPU tests flip (BCI_SCAN, BCI_SORT, BCI_MATMUL) are tuning-mode only

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by `pytest -m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Builds libbci.so and liboracle.so in-tree when missing (no-op when up to date)."""
    from bee_code_interpreter_b200 import build

    build.build_lib()
    build.build_oracle()
    build.build_server()
    return True


@pytest.fixture(scope="session")
def dev(built):
    """Binds the test process to GPU 0 through the C ABI; fails loudly without one."""
    from bee_code_interpreter_b200 import device

    device.init(int(os.environ.get("BCI_DEVICE", "0")))
    yield device
    device.sync()
