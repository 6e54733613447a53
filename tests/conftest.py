import os
import sys

# several slab handles share one device inside this process (tests/test_gpu_slab.py): a kernel that spins on a peer must never
# wait for another kernel's lazy module load, which cannot happen while a kernel is running (CUDA lazy-loading caveat)
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle
    oracle.build()
    return oracle
