import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the reference build under oracle/_ref")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build test infrastructure (oracle, emulator) and the product library once per session."""
    from grsr_b200 import build
    build.build_oracle()
    build.build_cuda()
    yield
