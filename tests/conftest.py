import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """libbwq.so built in-tree (nvcc cross-compiles without a GPU)."""
    from ml_qem_b200 import build, engine

    build.build_library()
    return engine.load_library()


@pytest.fixture(scope="session")
def engine_gpu(lib):
    from ml_qem_b200.engine import Engine

    return Engine(0)
