import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_present():
    """True when bwq_create(0) finds a device (no torch import: the library itself decides)."""
    try:
        from ml_qem_b200 import build, engine

        build.build_library()
        eng = engine.Engine(0)
    except Exception:  # noqa: BLE001 - no nvcc, no library, no device: all mean "skip the gpu tests"
        return False
    eng.close()
    return True


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a GPU skips the gpu-marked tests instead of erroring
    (the engine has no CPU fallback: Engine(0) raises there)."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the engine has no CPU fallback (run with -m gpu on the B200 box)")
    for it in gpu_items:
        it.add_marker(skip)


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
