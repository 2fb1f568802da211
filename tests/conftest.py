import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle9():
    import harness
    return harness.Oracle(planes=9, max_stages=6)


@pytest.fixture(scope="session")
def oracle15():
    import harness
    return harness.Oracle(planes=15, max_stages=7)


@pytest.fixture(scope="session")
def ref9():
    import harness
    if not harness.have_ref():
        pytest.skip("oracle/_ref not built")
    return harness.Ref(False)


@pytest.fixture(scope="session")
def ref15():
    import harness
    if not harness.have_ref():
        pytest.skip("oracle/_ref not built")
    return harness.Ref(True)
