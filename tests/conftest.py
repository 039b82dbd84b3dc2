import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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
