import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The C library and the CPU checkers must exist before anything runs."""
    import __graft_entry__ as entry

    entry.build()


@pytest.fixture(scope="session")
def scenes_dir():
    return os.path.join(ROOT, "tests", "golden", "scenes")


EXAMPLES = ["scene", "scene2", "scene3", "scene4"]
