import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")
    config.addinivalue_line("markers", "ref: needs the compiled reference under oracle/_ref (built here from /root/reference)")


@pytest.fixture(scope="session")
def gen():
    from harness import load_gen
    return load_gen()
