import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with `-m gpu`)")


@pytest.fixture(scope="session")
def engine():
    """One evg_ctx for the whole GPU session; fails loudly without a device."""
    from evergreen_b200 import scheduler
    eng = scheduler.Engine(0)
    yield eng
    eng.close()
