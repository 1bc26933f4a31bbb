import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DATA = os.path.join(ROOT, "paos_b200", "lens_data")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _ensure_library():
    """The CUDA library is a build artefact (git-ignored): build it in-tree when a fresh checkout lacks it, so that the
    CPU suite can at least load it and check its exports (nvcc cross-compiles sm_100a without a GPU)."""
    lib = os.path.join(ROOT, "paos_b200", "libpaos_b200.so")
    if os.path.exists(lib):
        return
    import importlib.util

    spec = importlib.util.spec_from_file_location("_paos_b200_build", os.path.join(ROOT, "paos_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    _ensure_library()


@pytest.fixture(scope="session")
def data_dir():
    return DATA


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
