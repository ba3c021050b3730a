import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def engine_lib():
    import mpidopenmmplugin_b200
    from mpidopenmmplugin_b200 import build
    build.build()
    return mpidopenmmplugin_b200.load_library()
