import os
import sys

import pytest

# The loopback multi-rank tests run N ranks on ONE device, each on its own stream, and a rank's tiny wait kernels spin until
# its peers' kernels have run.  With the default 8 hardware work queues two such streams can share a queue, and the
# peer's kernels would then sit behind the spinning one.  Must be set before CUDA is initialised.  (N real GPUs have one
# rank per device: no such coupling.)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

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
