import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# NVRTC output of tape systems is cached on disk (csrc/runtime.cpp); keep the cache inside the repository for test runs
os.environ.setdefault("HB_JIT_CACHE_DIR", os.path.join(ROOT, ".jit_cache", "gpu" if os.path.exists("/dev/nvidiactl") else "cpu"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O
    O.lib()
    return O
