import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# Several engine handles on ONE device (tests/test_gpu_shards.py) wait for each other's kernels: give every stream its own
# hardware queue so that a spinning kernel never sits in front of the one it waits for (must be set before CUDA starts).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# a lost hand-shake should fail a test in seconds, not in the production limit of 20 s (LOGREG: 30x)
os.environ.setdefault("PGN_TIMEOUT_S", "4")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running statistical check")


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU oracle (test infrastructure) behind the same marshalling code as the product."""
    from oracle_adapter import load_oracle
    return load_oracle()


@pytest.fixture(scope="session")
def gpu_lib():
    import pigeons_jl_b200 as pg
    return pg.EngineLib()
