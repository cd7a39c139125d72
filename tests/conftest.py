import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session", autouse=True)
def native_artefacts():
    """libcrf.so / libcrf_tools.so (nvcc, sm_100a -- cross-compiles without a GPU) and the C oracle are built once per session
    if missing or older than their sources, so the suite does not depend on a prior build step."""
    pkg = os.path.join(ROOT, "colab-repeat-finder_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from crf_b200 import build as crf_build
    from oracle import oracle, ref
    oracle.build()
    ref.build()                                  # copies the Python reference to oracle/_ref where /root/reference exists
    if crf_build.have_nvcc():
        crf_build.build()
    # without nvcc the CPU-only suites (oracle, partition, host logic) still run; tests that need libcrf.so fail
    # loudly in _cabi.lib() if no prebuilt library is there
