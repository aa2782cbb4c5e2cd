import os
import sys

import pytest

os.environ.setdefault("KB200_RANDOM_VGG", "1")      # no network: tests run the VGG trunk with seeded weights (explicit opt-in)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_1thread():
    import oracle
    oracle.set_threads(1)
    return oracle
