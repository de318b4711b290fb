import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import edo
    edo.build()
    return edo


@pytest.fixture()
def ed():
    """The product's host API, initialised on cuda:0 as a single rank; finalised afterwards."""
    from cdmft_lanc_ed_b200 import ed_hamiltonian as E
    E.ed_init(0)
    yield E
    E.ed_finalize()
