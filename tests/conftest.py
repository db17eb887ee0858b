import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure); builds oracle/libphsh_oracle.so on first use."""
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def re_blocks():
    from oracle import fortran_io as fio
    return fio.read_mufftin_rel(os.path.join(GOLDEN, "Re0001", "mufftin_Re_slab.d"))
